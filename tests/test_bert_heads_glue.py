"""BERT fine-tuning heads and GLUE scoring (``modeling.py:828-1210``, ``compute_glue_scores.py``) -- CPU only."""
import math

import pytest
import torch

from oktopk_b200.models.bert import BertConfig, BertForPreTraining
from oktopk_b200.models.bert_heads import (BertForMaskedLM, BertForMultipleChoice, BertForNextSentencePrediction,
                                           BertForQuestionAnswering, BertForSequenceClassification,
                                           BertForTokenClassification, BertModel, load_pretraining_encoder)
from oktopk_b200.train import glue
from oktopk_b200.utils.tokenization import BertTokenizer

CFG = BertConfig(vocab_size=3000, hidden_size=32, num_hidden_layers=4, num_attention_heads=2, intermediate_size=64,
                 max_position_embeddings=64, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)


def _inputs(b=3, s=12):
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(5, 3000, (b, s), generator=g)
    seg = (torch.arange(s)[None] >= s // 2).long().expand(b, s).contiguous()
    mask = torch.ones(b, s, dtype=torch.long)
    mask[:, -2:] = 0
    return ids, seg, mask


def test_heads_shapes_and_losses():
    ids, seg, mask = _inputs()
    seq, pooled = BertModel(CFG)(ids, seg, mask)
    assert seq.shape == (3, 12, 32) and pooled.shape == (3, 32)
    m = BertForSequenceClassification(CFG, 3)
    assert m(ids, seg, mask).shape == (3, 3)
    assert m(ids, seg, mask, torch.tensor([0, 1, 2])).dim() == 0
    assert BertForSequenceClassification(CFG, 1)(ids, seg, mask, torch.tensor([0.5, 1.0, 4.0])).dim() == 0   # regression
    tc = BertForTokenClassification(CFG, 5)
    assert tc(ids, seg, mask).shape == (3, 12, 5)
    assert torch.isfinite(tc(ids, seg, mask, torch.randint(0, 5, (3, 12))))
    qa = BertForQuestionAnswering(CFG)
    s, e = qa(ids, seg, mask)
    assert s.shape == e.shape == (3, 12)
    assert torch.isfinite(qa(ids, seg, mask, torch.tensor([1, 2, 40]), torch.tensor([3, 4, 50])))    # out-of-window ignored
    mc = BertForMultipleChoice(CFG, 2)
    ids2 = torch.stack([ids, ids.flip(1)], 1)
    assert mc(ids2, torch.stack([seg, seg], 1), torch.stack([mask, mask], 1)).shape == (3, 2)
    lm = BertForMaskedLM(CFG)
    assert lm(ids, seg, mask).shape == (3, 12, 3000)
    labels = torch.full((3, 12), -1)
    labels[:, 3] = ids[:, 3]
    assert torch.isfinite(lm(ids, seg, mask, labels))
    assert BertForNextSentencePrediction(CFG)(ids, seg, mask).shape == (3, 2)


def test_pretraining_checkpoint_loads_into_a_task_head():
    torch.manual_seed(0)
    pre = BertForPreTraining(CFG, depth=2)
    head = BertForSequenceClassification(CFG, 2)
    n = load_pretraining_encoder(head, pre.state_dict())
    own = head.bert.state_dict()
    assert n == len(own)                                            # embeddings + all 4 layers + pooler
    ids, seg, mask = _inputs()
    pre.eval(); head.eval()
    from oktopk_b200.models.bert import extended_attention_mask
    x = pre.stages[0](ids, seg, extended_attention_mask(mask))
    x = pre.stages[1].layers[0](x, extended_attention_mask(mask))
    x = pre.stages[1].layers[1](x, extended_attention_mask(mask))
    seq, _ = head.bert(ids, seg, mask)
    torch.testing.assert_close(seq, x)


def test_glue_metrics():
    p = torch.tensor([1, 0, 1, 1, 0, 0, 1, 0])
    y = torch.tensor([1, 0, 0, 1, 0, 1, 1, 0])
    assert glue.simple_accuracy(p, y) == pytest.approx(0.75)
    assert glue.f1_score(p, y) == pytest.approx(0.75)
    assert glue.matthews_corrcoef(p, y) == pytest.approx(0.5)
    x = torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0])
    assert glue.pearson(x, 2 * x + 1) == pytest.approx(1.0)
    assert glue.spearman(x, x ** 3) == pytest.approx(1.0)
    assert glue.spearman(x, -x) == pytest.approx(-1.0)
    assert set(glue.compute_metrics("mrpc", p, y)) == {"acc", "f1", "acc_and_f1"}
    assert set(glue.compute_metrics("sts-b", x, x)) == {"pearson", "spearmanr", "corr"}
    assert "mcc" in glue.compute_metrics("cola", p, y)
    from scipy import stats
    a, b = torch.randn(50), torch.randn(50)
    assert glue.spearman(a, b) == pytest.approx(stats.spearmanr(a.numpy(), b.numpy())[0], abs=1e-6)


def test_processors_read_tsv(tmp_path):
    (tmp_path / "train.tsv").write_text("Quality\t#1 ID\t#2 ID\t#1 String\t#2 String\n1\t1\t2\tthe cat sat\ta cat sat\n"
                                        "0\t3\t4\tdogs bark\tthe sky is blue\n")
    (tmp_path / "dev.tsv").write_text("Quality\t#1 ID\t#2 ID\t#1 String\t#2 String\n1\t1\t2\tx y\tx z\n")
    proc = glue.DataProcessor("mrpc")
    tr, dv = proc.get_train_examples(str(tmp_path)), proc.get_dev_examples(str(tmp_path))
    assert [e.label for e in tr] == ["1", "0"] and tr[0].text_b == "a cat sat" and len(dv) == 1
    assert proc.get_labels() == ["0", "1"] and proc.output_mode == "classification"
    assert glue.DataProcessor("sts-b").output_mode == "regression"
    tok = BertTokenizer.synthetic(3000)
    ids, seg, mask, y = glue.convert_examples_to_features(tr, proc.get_labels(), 16, tok)
    assert ids.shape == (2, 16) and y.tolist() == [1, 0] and int(mask[0].sum()) <= 16 and int(seg.max()) == 1
    with pytest.raises(KeyError):
        glue.DataProcessor("nope")


def test_finetune_on_synthetic_task_beats_chance():
    tok = BertTokenizer.synthetic(3000)
    train, dev = glue.synthetic_examples(256, "mrpc", 0), glue.synthetic_examples(96, "mrpc", 1)
    import dataclasses
    cfg = dataclasses.replace(CFG, num_hidden_layers=2)
    res = glue.finetune_and_score("mrpc", train, dev, cfg, tok, max_seq_length=24, epochs=16, lr=5e-4, batch_size=32,
                                  device=torch.device("cpu"))
    assert set(res) == {"acc", "f1", "acc_and_f1"} and all(math.isfinite(v) for v in res.values())
    assert res["acc"] > 0.6, res
