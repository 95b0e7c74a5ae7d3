"""BERT pre-training data: sentence-pair sampling with on-the-fly masking from a text corpus.

Capability parity with ``BERTDataset`` / ``convert_example_to_features`` / ``random_word``
(``BERT/bert/main_bert.py:257-639``) and the partitioned variant (``BERT/bert/dataset.py:93-227``):

* corpus format: one sentence per line, documents separated by blank lines;
* a sample = (sentence i of a document, its successor with p=0.5 else a random sentence of another document) +
  ``is_next`` label;
* masking: each WordPiece is chosen with p=0.15; a chosen token becomes ``[MASK]`` (80 %), a random vocabulary
  token (10 %) or stays (10 %); ``lm_label_ids`` holds the original id there and -1 elsewhere;
* features: ``[CLS] A [SEP] B [SEP]`` truncated pair-wise to ``seq_len``, ``segment_ids``, ``input_mask``, zero padding.

Different by design: the corpus is indexed once into flat (document, line) arrays and tokenised lazily per sample
with a per-sample ``torch.Generator`` (reproducible under any sharding / worker count; the reference keeps file
cursors and global ``random`` state), and shards are a ``DistributedSampler`` over the same index instead of
pre-split files.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
from torch.utils.data import Dataset

from ..utils.tokenization import BertTokenizer


@dataclass
class InputExample:
    guid: int
    tokens_a: List[str]
    tokens_b: Optional[List[str]] = None
    is_next: Optional[int] = None
    lm_labels: Optional[List[int]] = None


@dataclass
class InputFeatures:
    input_ids: List[int]
    input_mask: List[int]
    segment_ids: List[int]
    is_next: int
    lm_label_ids: List[int]


def truncate_seq_pair(tokens_a: List[str], tokens_b: List[str], max_length: int) -> None:
    """Trim the longer sequence one token at a time until the pair fits (``main_bert.py:463-478``)."""
    while len(tokens_a) + len(tokens_b) > max_length:
        (tokens_a if len(tokens_a) > len(tokens_b) else tokens_b).pop()


def random_word(tokens: List[str], tokenizer: BertTokenizer, gen: torch.Generator, mask_prob: float = 0.15
                ) -> Tuple[List[str], List[int]]:
    """Mask ``tokens`` in place; returns (tokens, labels) with -1 for untouched positions (``main_bert.py:481-520``)."""
    labels = []
    n = len(tokens)
    if n == 0:
        return tokens, labels
    r = torch.rand(n, generator=gen).tolist()
    vocab_items = None
    unk = tokenizer.vocab.get("[UNK]", 0)
    for i, tok in enumerate(tokens):
        if r[i] < mask_prob:
            q = r[i] / mask_prob
            if q < 0.8:
                tokens[i] = "[MASK]"
            elif q < 0.9:
                if vocab_items is None:
                    vocab_items = list(tokenizer.vocab.keys())
                tokens[i] = vocab_items[int(torch.randint(len(vocab_items), (1,), generator=gen))]
            labels.append(tokenizer.vocab.get(tok, unk))
        else:
            labels.append(-1)
    return tokens, labels


def convert_example_to_features(example: InputExample, max_seq_length: int, tokenizer: BertTokenizer,
                                gen: torch.Generator) -> InputFeatures:
    """``main_bert.py:535-614``."""
    a, b = list(example.tokens_a), list(example.tokens_b or [])
    truncate_seq_pair(a, b, max_seq_length - 3)
    a, la = random_word(a, tokenizer, gen)
    b, lb = random_word(b, tokenizer, gen)
    tokens = ["[CLS]"] + a + ["[SEP]"] + b + ["[SEP]"]
    seg = [0] * (len(a) + 2) + [1] * (len(b) + 1)
    lm = [-1] + la + [-1] + lb + [-1]
    ids = tokenizer.convert_tokens_to_ids(tokens)
    mask = [1] * len(ids)
    pad = max_seq_length - len(ids)
    ids += [0] * pad
    mask += [0] * pad
    seg += [0] * pad
    lm += [-1] * pad
    assert len(ids) == len(mask) == len(seg) == len(lm) == max_seq_length
    return InputFeatures(ids, mask, seg, int(example.is_next), lm)


class BERTDataset(Dataset):
    """``BERTDataset(corpus_path, tokenizer, seq_len)`` -> ``(input_ids, segment_ids, input_mask, lm_label_ids, is_next)``
    (the tensor order our ``BertForPreTraining.forward`` / ``Trainer`` expect)."""

    def __init__(self, corpus_path: Optional[str], tokenizer: BertTokenizer, seq_len: int = 128, encoding: str = "utf-8",
                 corpus_lines: Optional[int] = None, on_memory: bool = True, seed: int = 0,
                 lines: Optional[Sequence[str]] = None):
        self.tokenizer, self.seq_len, self.seed = tokenizer, seq_len, seed
        self.vocab = tokenizer.vocab
        self.docs: List[List[str]] = []
        doc: List[str] = []
        if lines is None:
            if corpus_path is None or not os.path.isfile(corpus_path):
                raise FileNotFoundError("corpus %r not found" % (corpus_path,))
            with open(corpus_path, "r", encoding=encoding) as f:
                lines = f.read().splitlines()
        for line in lines:
            line = line.strip()
            if line == "":
                if doc:
                    self.docs.append(doc)
                doc = []
            else:
                doc.append(line)
        if doc:
            self.docs.append(doc)
        self.docs = [d for d in self.docs if len(d) >= 2]
        if len(self.docs) < 2:
            raise ValueError("need at least two documents of >= 2 sentences each")
        # sample index: every line that has a successor inside its document (main_bert.py:323-325)
        self.sample_to_doc: List[Tuple[int, int]] = [(di, li) for di, d in enumerate(self.docs) for li in range(len(d) - 1)]
        self.num_docs = len(self.docs)
        self.corpus_lines = sum(len(d) for d in self.docs)

    def __len__(self) -> int:
        return len(self.sample_to_doc)

    def random_sent(self, index: int, gen: torch.Generator) -> Tuple[str, str, int]:
        di, li = self.sample_to_doc[index]
        t1 = self.docs[di][li]
        if float(torch.rand(1, generator=gen)) > 0.5:
            return t1, self.docs[di][li + 1], 1
        for _ in range(10):                                   # a sentence of a *different* document
            rd = int(torch.randint(self.num_docs, (1,), generator=gen))
            if rd != di:
                break
        rl = int(torch.randint(len(self.docs[rd]), (1,), generator=gen))
        return t1, self.docs[rd][rl], 0

    def __getitem__(self, item: int):
        gen = torch.Generator().manual_seed(self.seed * 1_000_003 + item)
        t1, t2, is_next = self.random_sent(item, gen)
        ex = InputExample(item, self.tokenizer.tokenize(t1), self.tokenizer.tokenize(t2), is_next)
        f = convert_example_to_features(ex, self.seq_len, self.tokenizer, gen)
        return (torch.tensor(f.input_ids), torch.tensor(f.segment_ids), torch.tensor(f.input_mask),
                torch.tensor(f.lm_label_ids), torch.tensor(f.is_next))


def synthetic_corpus(n_docs: int = 64, sents: int = 8, words: int = 12, seed: int = 0, vocab_words: int = 2000) -> List[str]:
    """A small deterministic text corpus in the expected format (for tests and shape-faithful dry runs)."""
    g = torch.Generator().manual_seed(seed)
    syl = ["ka", "to", "mi", "re", "so", "la", "ne", "vu", "di", "po", "an", "er", "in", "on", "st", "th"]
    lines: List[str] = []
    for _ in range(n_docs):
        for _ in range(sents):
            ws = torch.randint(vocab_words, (words,), generator=g).tolist()
            lines.append(" ".join(syl[w % 16] + syl[(w // 16) % 16] for w in ws) + " .")
        lines.append("")
    return lines


def extended_attention_mask(input_mask: torch.Tensor) -> torch.Tensor:
    """``(1 - mask) * -10000`` broadcast to ``[B,1,1,S]`` (``main_bert.py:616-639``)."""
    return (1.0 - input_mask[:, None, None, :].to(torch.float32)) * -10000.0


# ------------------------------------------------------------------------------------------------------------------
# Pre-created training instances, stored as partitions (capability parity with ``BERT/bert/sources.py:30-255``
# ``TokenInstance`` / ``PretrainingDataCreator`` / ``GenericPretrainingDataCreator`` / ``WikiPretrainingDataCreator`` and
# ``BERT/bert/dataset.py:93-227`` ``BERTDatasetPartitioned``).  Documents are tokenised ONCE, packed into sentence-pair
# instances of ~``max_seq_length`` tokens (BERT's original recipe: segments A/B split at a sentence boundary, B replaced by
# a random document's text half of the time), duplicated ``dupe_factor`` times with different packings, and saved as
# shards; the dataset then only masks on the fly.  Storage is ``torch.save`` of plain lists (not pickled Python objects).
# ------------------------------------------------------------------------------------------------------------------
@dataclass
class TokenInstance:
    tokens_a: List[str]
    tokens_b: List[str]
    is_next: int                      # 0 = B continues A, 1 = B is random (the reference's convention, sources.py:34)

    def get_values(self):
        return self.tokens_a, self.tokens_b, self.is_next


class PretrainingDataCreator:
    """``documents``: list of documents, each a list of sentences (strings)."""

    def __init__(self, documents: Sequence[Sequence[str]], tokenizer: BertTokenizer, max_seq_length: int = 128,
                 dupe_factor: int = 5, small_seq_prob: float = 0.1, seed: int = 0):
        self.max_seq_length, self.dupe_factor, self.small_seq_prob = max_seq_length, dupe_factor, small_seq_prob
        gen = torch.Generator().manual_seed(seed)
        docs = [[tokenizer.tokenize(s) for s in d if s.strip()] for d in documents]
        docs = [[s for s in d if s] for d in docs]
        self._docs = [d for d in docs if len(d) >= 2]
        self.instances: List[TokenInstance] = []
        for _ in range(dupe_factor):
            for di in range(len(self._docs)):
                self.instances.extend(self._create(di, gen))
        perm = torch.randperm(len(self.instances), generator=gen).tolist()
        self.instances = [self.instances[i] for i in perm]
        self._docs = None

    @classmethod
    def from_corpus(cls, path: str, tokenizer: BertTokenizer, sep: Optional[str] = None, **kw) -> "PretrainingDataCreator":
        """``sep=None``: blank-line separated documents, one sentence per line (the ``BERTDataset`` format);
        ``sep='<sep>'``: one document per line with ``<sep>``-joined sentences (``sources.py:55-63``)."""
        docs, cur = [], []
        with open(path, encoding="utf-8") as f:
            for line in f:
                line = line.rstrip("\n")
                if sep is not None:
                    parts = [p for p in line.split(sep) if p.strip()]
                    if len(parts) > 3:
                        docs.append(parts)
                elif line.strip() == "":
                    if cur:
                        docs.append(cur)
                    cur = []
                else:
                    cur.append(line)
        if cur:
            docs.append(cur)
        return cls(docs, tokenizer, **kw)

    def __len__(self) -> int:
        return len(self.instances)

    def _rand(self, gen, lo: int, hi: int) -> int:
        return int(torch.randint(lo, hi + 1, (1,), generator=gen))

    def _create(self, di: int, gen) -> List[TokenInstance]:
        doc = self._docs[di]
        max_tokens = self.max_seq_length - 3                       # [CLS] + 2 x [SEP]
        target = max_tokens
        if float(torch.rand(1, generator=gen)) < self.small_seq_prob:
            target = self._rand(gen, 5, max_tokens)
        out, chunk, length, i = [], [], 0, 0
        while i < len(doc):
            chunk.append(doc[i])
            length += len(doc[i])
            if i == len(doc) - 1 or length >= target:
                if chunk:
                    a_end = 1 if len(chunk) < 2 else self._rand(gen, 1, len(chunk) - 1)
                    a = [t for s in chunk[:a_end] for t in s]
                    is_random = len(chunk) == 1 or float(torch.rand(1, generator=gen)) < 0.5
                    if is_random and len(self._docs) > 1:
                        rd = di
                        for _ in range(10):
                            rd = self._rand(gen, 0, len(self._docs) - 1)
                            if rd != di:
                                break
                        rdoc = self._docs[rd]
                        start = self._rand(gen, 0, len(rdoc) - 1)
                        b, budget = [], target - len(a)
                        for s in rdoc[start:]:
                            b.extend(s)
                            if len(b) >= budget:
                                break
                        i -= len(chunk) - a_end                    # the unused sentences go back (no text is wasted)
                        nxt = 1
                    else:
                        b = [t for s in chunk[a_end:] for t in s]
                        nxt = 0
                    if a and b:
                        truncate_seq_pair(a, b, max_tokens)
                        out.append(TokenInstance(a, b, nxt))
                chunk, length = [], 0
            i += 1
        return out

    # ---- partitions -------------------------------------------------------------------------------------------
    def save(self, filename: str) -> None:
        torch.save({"max_seq_length": self.max_seq_length,
                    "instances": [(x.tokens_a, x.tokens_b, x.is_next) for x in self.instances]}, filename)

    def save_partitions(self, folder: str, n: int) -> List[str]:
        os.makedirs(folder, exist_ok=True)
        paths = []
        for p in range(n):
            part = self.instances[p::n]
            path = os.path.join(folder, "part-%05d.pt" % p)
            torch.save({"max_seq_length": self.max_seq_length, "instances": [(x.tokens_a, x.tokens_b, x.is_next) for x in part]},
                       path)
            paths.append(path)
        return paths

    @staticmethod
    def load_instances(filename: str) -> List[TokenInstance]:
        d = torch.load(filename, weights_only=False)
        return [TokenInstance(list(a), list(b), int(n)) for a, b, n in d["instances"]]


GenericPretrainingDataCreator = PretrainingDataCreator
WikiPretrainingDataCreator = PretrainingDataCreator


def get_random_partition(data_directory: str, index: int) -> str:
    parts = sorted(os.path.join(data_directory, x) for x in os.listdir(data_directory))
    return parts[index % len(parts)]


class BERTDatasetPartitioned(Dataset):
    """Instances from (up to ``num_partitions``) shards of a folder, masked on the fly (``dataset.py:93-227``:
    ``masked_lm_prob`` 0.15 capped at ``max_predictions_per_seq``, 80/10/10 replacement).  Returns the same tensor tuple
    as ``BERTDataset``; ``is_next`` follows OUR loss convention (1 = B really follows A)."""

    def __init__(self, tokenizer: BertTokenizer, folder: str, max_seq_length: int = 128, max_predictions_per_seq: int = 20,
                 masked_lm_prob: float = 0.15, num_partitions: int = 8, seed: int = 0):
        self.tokenizer, self.max_seq_length, self.seed = tokenizer, max_seq_length, seed
        self.max_pred, self.p = max_predictions_per_seq, masked_lm_prob
        self.vocab_words = list(tokenizer.vocab.keys())
        seen, self.instances = set(), []
        for i in range(num_partitions):
            path = get_random_partition(folder, i)
            if path not in seen:
                seen.add(path)
                self.instances.extend(PretrainingDataCreator.load_instances(path))

    def __len__(self) -> int:
        return len(self.instances)

    def __getitem__(self, index: int):
        inst = self.instances[index % len(self.instances)]
        gen = torch.Generator().manual_seed(self.seed * 1_000_003 + index)
        a, b = list(inst.tokens_a), list(inst.tokens_b)
        truncate_seq_pair(a, b, self.max_seq_length - 3)
        tokens = ["[CLS]"] + a + ["[SEP]"] + b + ["[SEP]"]
        seg = [0] * (len(a) + 2) + [1] * (len(b) + 1)
        cand = [i for i, t in enumerate(tokens) if t not in ("[CLS]", "[SEP]")]
        n_pred = min(self.max_pred, max(1, int(round(len(tokens) * self.p))))
        order = torch.randperm(len(cand), generator=gen).tolist()[:n_pred]
        lm = [-1] * len(tokens)
        unk = self.tokenizer.vocab.get("[UNK]", 0)
        for j in order:
            pos = cand[j]
            lm[pos] = self.tokenizer.vocab.get(tokens[pos], unk)
            r = float(torch.rand(1, generator=gen))
            if r < 0.8:
                tokens[pos] = "[MASK]"
            elif r < 0.9:
                tokens[pos] = self.vocab_words[int(torch.randint(len(self.vocab_words), (1,), generator=gen))]
        ids = self.tokenizer.convert_tokens_to_ids(tokens)
        pad = self.max_seq_length - len(ids)
        mask = [1] * len(ids) + [0] * pad
        ids, seg, lm = ids + [0] * pad, seg + [0] * pad, lm + [-1] * pad
        return (torch.tensor(ids), torch.tensor(seg), torch.tensor(mask), torch.tensor(lm), torch.tensor(1 - inst.is_next))
