"""GLUE fine-tuning / scoring of a BERT checkpoint (capability parity with ``BERT/bert/compute_glue_scores.py``:
task processors ``:170-550``, ``convert_examples_to_features :553-628``, metrics ``:648-695``, evaluation loop ``:49-136``).

    python -m oktopk_b200.train.glue --task mrpc --data-dir /data/glue/MRPC --vocab /data/vocab.txt \
        --checkpoint ./checkpoints/bert_base-rank0-epoch0.pth --epochs 3

Processors read the standard GLUE ``train.tsv`` / ``dev.tsv`` files; ``--synthetic`` generates a small learnable task of
the same shape when no data is on disk (this image has no network).  Fine-tuning runs through the same
``DistributedOptimizer`` / ``BertAdam`` path as pre-training (dense by default; ``--compressor oktopk`` works too).
"""
from __future__ import annotations

import argparse
import csv
import json
import math
import os
import sys
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from ..utils.tokenization import BertTokenizer
from .bert_data import truncate_seq_pair


@dataclass
class InputExample:
    guid: str
    text_a: str
    text_b: Optional[str] = None
    label: Optional[str] = None


# task -> (text_a column, text_b column, label column, labels, skip header, metric)
TASKS: Dict[str, Tuple] = {
    "cola": (3, None, 1, ["0", "1"], False, "mcc"),
    "sst-2": (0, None, 1, ["0", "1"], True, "acc"),
    "mrpc": (3, 4, 0, ["0", "1"], True, "acc_and_f1"),
    "sts-b": (7, 8, -1, [None], True, "pearson_and_spearman"),
    "qqp": (3, 4, 5, ["0", "1"], True, "acc_and_f1"),
    "mnli": (8, 9, -1, ["contradiction", "entailment", "neutral"], True, "acc"),
    "mnli-mm": (8, 9, -1, ["contradiction", "entailment", "neutral"], True, "acc"),
    "qnli": (1, 2, -1, ["entailment", "not_entailment"], True, "acc"),
    "rte": (1, 2, -1, ["entailment", "not_entailment"], True, "acc"),
    "wnli": (1, 2, -1, ["0", "1"], True, "acc"),
}
DEV_FILE = {"mnli": "dev_matched.tsv", "mnli-mm": "dev_mismatched.tsv"}


class DataProcessor:
    """One class parameterised by the task table instead of the reference's ten near-identical processor classes."""

    def __init__(self, task: str):
        task = task.lower()
        if task not in TASKS:
            raise KeyError("unknown GLUE task %r (have %s)" % (task, sorted(TASKS)))
        self.task = task
        self.col_a, self.col_b, self.col_y, self.labels, self.header, self.metric = TASKS[task]

    def get_labels(self) -> List:
        return list(self.labels)

    @property
    def output_mode(self) -> str:
        return "regression" if self.labels == [None] else "classification"

    def _read(self, path: str, set_type: str) -> List[InputExample]:
        out = []
        with open(path, "r", encoding="utf-8-sig") as f:
            for i, row in enumerate(csv.reader(f, delimiter="\t", quotechar=None)):
                if i == 0 and self.header:
                    continue
                try:
                    a = row[self.col_a]
                    b = row[self.col_b] if self.col_b is not None else None
                    y = row[self.col_y]
                except IndexError:
                    continue
                out.append(InputExample("%s-%d" % (set_type, i), a, b, y))
        return out

    def get_train_examples(self, data_dir: str) -> List[InputExample]:
        return self._read(os.path.join(data_dir, "train.tsv"), "train")

    def get_dev_examples(self, data_dir: str) -> List[InputExample]:
        return self._read(os.path.join(data_dir, DEV_FILE.get(self.task, "dev.tsv")), "dev")


def convert_examples_to_features(examples: Sequence[InputExample], label_list: Sequence, max_seq_length: int,
                                 tokenizer: BertTokenizer, output_mode: str = "classification"):
    """``[CLS] a [SEP] (b [SEP])`` ids / mask / segments / label tensors."""
    label_map = {l: i for i, l in enumerate(label_list)}
    ids, masks, segs, ys = [], [], [], []
    for ex in examples:
        ta = tokenizer.tokenize(ex.text_a)
        tb = tokenizer.tokenize(ex.text_b) if ex.text_b else None
        if tb is not None:
            truncate_seq_pair(ta, tb, max_seq_length - 3)
        else:
            ta = ta[:max_seq_length - 2]
        toks = ["[CLS]"] + ta + ["[SEP]"]
        seg = [0] * len(toks)
        if tb is not None:
            toks += tb + ["[SEP]"]
            seg += [1] * (len(tb) + 1)
        i = tokenizer.convert_tokens_to_ids(toks)
        pad = max_seq_length - len(i)
        ids.append(i + [0] * pad)
        masks.append([1] * len(i) + [0] * pad)
        segs.append(seg + [0] * pad)
        ys.append(float(ex.label) if output_mode == "regression" else label_map[ex.label])
    y = torch.tensor(ys, dtype=torch.float32 if output_mode == "regression" else torch.long)
    return torch.tensor(ids), torch.tensor(segs), torch.tensor(masks), y


# ----------------------------------------------------------------------------------------------- metrics
def simple_accuracy(preds: torch.Tensor, labels: torch.Tensor) -> float:
    return float((preds == labels).float().mean())


def f1_score(preds: torch.Tensor, labels: torch.Tensor) -> float:
    tp = float(((preds == 1) & (labels == 1)).sum())
    fp = float(((preds == 1) & (labels == 0)).sum())
    fn = float(((preds == 0) & (labels == 1)).sum())
    return 0.0 if tp == 0 else 2 * tp / (2 * tp + fp + fn)


def matthews_corrcoef(preds: torch.Tensor, labels: torch.Tensor) -> float:
    tp = float(((preds == 1) & (labels == 1)).sum())
    tn = float(((preds == 0) & (labels == 0)).sum())
    fp = float(((preds == 1) & (labels == 0)).sum())
    fn = float(((preds == 0) & (labels == 1)).sum())
    den = math.sqrt((tp + fp) * (tp + fn) * (tn + fp) * (tn + fn))
    return 0.0 if den == 0 else (tp * tn - fp * fn) / den


def pearson(x: torch.Tensor, y: torch.Tensor) -> float:
    x, y = x.double() - x.double().mean(), y.double() - y.double().mean()
    den = float(x.norm() * y.norm())
    return 0.0 if den == 0 else float((x * y).sum()) / den


def spearman(x: torch.Tensor, y: torch.Tensor) -> float:
    def rank(t):
        order = t.argsort()
        r = torch.empty_like(order, dtype=torch.double)
        r[order] = torch.arange(t.numel(), dtype=torch.double)
        # average ranks over ties
        vals, inv, cnt = torch.unique(t, return_inverse=True, return_counts=True)
        sums = torch.zeros(vals.numel(), dtype=torch.double).index_add_(0, inv, r)
        return (sums / cnt.double())[inv]
    return pearson(rank(x), rank(y))


def compute_metrics(task: str, preds: torch.Tensor, labels: torch.Tensor) -> Dict[str, float]:
    """``compute_glue_scores.py:672-695``."""
    kind = TASKS[task.lower()][5]
    if kind == "mcc":
        return {"mcc": matthews_corrcoef(preds, labels)}
    if kind == "acc":
        return {"acc": simple_accuracy(preds, labels)}
    if kind == "acc_and_f1":
        a, f = simple_accuracy(preds, labels), f1_score(preds, labels)
        return {"acc": a, "f1": f, "acc_and_f1": (a + f) / 2}
    p, s = pearson(preds, labels), spearman(preds, labels)
    return {"pearson": p, "spearmanr": s, "corr": (p + s) / 2}


# ----------------------------------------------------------------------------------------------- run
def synthetic_examples(n: int, task: str = "mrpc", seed: int = 0) -> List[InputExample]:
    """Learnable stand-in of a GLUE-shaped task: the label is decided by the first word of sentence A."""
    g = torch.Generator().manual_seed(seed)
    words = ["kato", "mire", "sola", "nevu", "dipo", "aner", "inon", "stth", "toka", "remi"]
    labels = TASKS[task][3]
    out = []
    for i in range(n):
        a = [words[int(j)] for j in torch.randint(len(words), (6,), generator=g)]
        b = [words[int(j)] for j in torch.randint(len(words), (6,), generator=g)]
        cls = int(words.index(a[0]) < len(words) // 2)
        y = labels[cls % len(labels)] if labels != [None] else str(float(cls) * 5.0)
        out.append(InputExample("syn-%d" % i, " ".join(a), " ".join(b) if TASKS[task][1] is not None else None, y))
    return out


@torch.no_grad()
def run_evaluation(model, features, task: str, output_mode: str, batch_size: int = 32, device=None) -> Dict[str, float]:
    model.eval()
    ids, seg, mask, y = features
    preds = []
    for s in range(0, ids.size(0), batch_size):
        sl = slice(s, s + batch_size)
        b = [t[sl].to(device) if device is not None else t[sl] for t in (ids, seg, mask)]
        logits = model(b[0], b[1], b[2])
        preds.append(logits.float().cpu())
    logits = torch.cat(preds)
    p = logits.view(-1) if output_mode == "regression" else logits.argmax(-1)
    res = compute_metrics(task, p, y)
    model.train()
    return res


def finetune_and_score(task: str, train, dev, config, tokenizer, max_seq_length: int = 128, epochs: int = 3, lr: float = 2e-5,
                       batch_size: int = 32, device=None, checkpoint: Optional[str] = None, compressor: str = "none",
                       density: float = 1.0, seed: int = 0) -> Dict[str, float]:
    from ..models.bert_heads import BertForSequenceClassification, load_pretraining_encoder
    from ..optimizer import BertAdam
    torch.manual_seed(seed)
    proc = DataProcessor(task)
    labels, mode = proc.get_labels(), proc.output_mode
    model = BertForSequenceClassification(config, num_labels=1 if mode == "regression" else len(labels))
    if checkpoint:
        ck = torch.load(checkpoint, map_location="cpu", weights_only=False)
        load_pretraining_encoder(model, ck.get("state", ck))
    device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
    model.to(device)
    ftr = convert_examples_to_features(train, labels, max_seq_length, tokenizer, mode)
    fdv = convert_examples_to_features(dev, labels, max_seq_length, tokenizer, mode)
    steps = max(1, epochs * math.ceil(ftr[0].size(0) / batch_size))
    opt = BertAdam(model.parameters(), lr=lr, warmup=0.1, t_total=steps, named_parameters=list(model.named_parameters()),
                   compressor=compressor, density=density)
    g = torch.Generator().manual_seed(seed)
    for _ in range(epochs):
        perm = torch.randperm(ftr[0].size(0), generator=g)
        for s in range(0, perm.numel(), batch_size):
            idx = perm[s:s + batch_size]
            ids, seg, mask, y = (t[idx].to(device) for t in ftr)
            opt.zero_grad()
            model(ids, seg, mask, y).backward()
            opt.step()
    res = run_evaluation(model, fdv, task, mode, batch_size, device)
    opt.close()
    return res


def main(argv=None) -> int:
    from ..models.bert import BertConfig
    p = argparse.ArgumentParser()
    p.add_argument("--task", default="mrpc", choices=sorted(TASKS))
    p.add_argument("--data-dir", default=None)
    p.add_argument("--vocab", default=None)
    p.add_argument("--bert-config", default=None)
    p.add_argument("--checkpoint", default=None)
    p.add_argument("--max-seq-length", type=int, default=128)
    p.add_argument("--epochs", type=int, default=3)
    p.add_argument("--lr", type=float, default=2e-5)
    p.add_argument("--batch-size", type=int, default=32)
    p.add_argument("--compressor", default="none")
    p.add_argument("--density", type=float, default=1.0)
    p.add_argument("--synthetic", action="store_true")
    a = p.parse_args(argv)
    proc = DataProcessor(a.task)
    tok = BertTokenizer(a.vocab) if a.vocab else BertTokenizer.synthetic()
    cfg = BertConfig.from_json_file(a.bert_config) if a.bert_config else BertConfig.bert_base()
    if a.synthetic or not a.data_dir:
        train, dev = synthetic_examples(512, a.task, 0), synthetic_examples(128, a.task, 1)
    else:
        train, dev = proc.get_train_examples(a.data_dir), proc.get_dev_examples(a.data_dir)
    res = finetune_and_score(a.task, train, dev, cfg, tok, a.max_seq_length, a.epochs, a.lr, a.batch_size,
                             checkpoint=a.checkpoint, compressor=a.compressor, density=a.density)
    print(json.dumps({"task": a.task, **res}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
