"""WordPiece tokenizer + BERT pre-training dataset (``BERT/bert/transformers/tokenization.py``,
``BERT/bert/main_bert.py:257-639``) -- CPU only."""
import os

import pytest
import torch

from oktopk_b200.train.bert_data import (BERTDataset, InputExample, convert_example_to_features, extended_attention_mask,
                                         random_word, synthetic_corpus, truncate_seq_pair)
from oktopk_b200.utils.tokenization import BasicTokenizer, BertTokenizer, WordpieceTokenizer, load_vocab


def test_basic_and_wordpiece_tokenizers():
    assert BasicTokenizer().tokenize("Hello, World!  naïve\tcafé") == ["hello", ",", "world", "!", "naive", "cafe"]
    assert BasicTokenizer(do_lower_case=False).tokenize("Hello [MASK] x") == ["Hello", "[MASK]", "x"]
    assert BasicTokenizer().tokenize("ab你好cd") == ["ab", "你", "好", "cd"]
    wp = WordpieceTokenizer({"un": 0, "##aff": 1, "##able": 2, "[UNK]": 3})
    assert wp.tokenize("unaffable") == ["un", "##aff", "##able"]
    assert wp.tokenize("unaffablex") == ["[UNK]"]
    assert wp.tokenize("x" * 200) == ["[UNK]"]


def test_bert_tokenizer_roundtrip(tmp_path):
    vf = tmp_path / "vocab.txt"
    toks = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "the", "quick", "brown", "fox", "##es", ",", "jump", "##s"]
    vf.write_text("\n".join(toks) + "\n")
    assert list(load_vocab(str(vf))) == toks
    tok = BertTokenizer(str(vf))
    pieces = tok.tokenize("The quick brown foxes, jumps zebra")
    assert pieces == ["the", "quick", "brown", "fox", "##es", ",", "jump", "##s", "[UNK]"]
    ids = tok.convert_tokens_to_ids(pieces)
    assert tok.convert_ids_to_tokens(ids) == pieces
    assert BertTokenizer.from_pretrained(str(tmp_path)).vocab == tok.vocab
    with pytest.raises(ValueError):
        BertTokenizer(str(tmp_path / "missing.txt"))
    with pytest.raises(ValueError):
        BertTokenizer(str(vf), max_len=3).convert_tokens_to_ids(pieces)
    assert len(BertTokenizer.synthetic()) == 30522


def test_masking_statistics_and_features():
    tok = BertTokenizer.synthetic(2000)
    g = torch.Generator().manual_seed(0)
    words = [w for w in tok.vocab if w.isalpha()][:200]
    tokens = (words * 50)[:10000]
    masked, labels = random_word(list(tokens), tok, g)
    chosen = [i for i, l in enumerate(labels) if l != -1]
    assert 0.13 < len(chosen) / len(tokens) < 0.17
    as_mask = sum(masked[i] == "[MASK]" for i in chosen) / len(chosen)
    kept = sum(masked[i] == tokens[i] for i in chosen) / len(chosen)
    assert 0.75 < as_mask < 0.85 and 0.06 < kept < 0.15
    assert all(labels[i] == tok.vocab[tokens[i]] for i in chosen)
    a, b = list("abcdefghij"), list("klm")
    truncate_seq_pair(a, b, 8)
    assert len(a) + len(b) == 8 and len(b) == 3
    f = convert_example_to_features(InputExample(0, words[:5], words[5:8], 1), 16, tok, g)
    assert len(f.input_ids) == 16 and f.input_ids[0] == tok.vocab["[CLS]"]
    assert f.segment_ids[:7] == [0] * 7 and f.segment_ids[7:11] == [1] * 4 and f.input_mask == [1] * 11 + [0] * 5
    assert f.lm_label_ids[0] == -1 and f.lm_label_ids[11:] == [-1] * 5 and f.is_next == 1


def test_bert_dataset_is_deterministic_and_feeds_the_model(tmp_path):
    lines = synthetic_corpus(n_docs=12, sents=5)
    (tmp_path / "train.txt").write_text("\n".join(lines))
    tok = BertTokenizer.synthetic(3000)
    ds = BERTDataset(str(tmp_path / "train.txt"), tok, seq_len=32)
    assert len(ds) == 12 * 4 and ds.num_docs == 12
    x, y = ds[5], ds[5]
    assert all(torch.equal(a, b) for a, b in zip(x, y))
    nexts = [int(ds[i][4]) for i in range(len(ds))]
    assert 0.25 < sum(nexts) / len(nexts) < 0.75
    from oktopk_b200.models.bert import BertConfig, BertForPreTraining
    from oktopk_b200.train.data import build_dataset
    ds2 = build_dataset("wikipedia", str(tmp_path), seq=64)
    assert isinstance(ds2, BERTDataset)
    loader = torch.utils.data.DataLoader(ds2, batch_size=4)
    ids, seg, mask, lm, nxt = next(iter(loader))
    cfg = BertConfig(vocab_size=30522, hidden_size=32, num_hidden_layers=2, num_attention_heads=2, intermediate_size=64,
                     max_position_embeddings=64)
    loss = BertForPreTraining(cfg, depth=2)(ids, seg, mask, lm, nxt)
    assert torch.isfinite(loss)
    m = extended_attention_mask(mask)
    assert m.shape == (4, 1, 1, 64) and float(m.max()) == 0.0 and float(m.min()) == -10000.0


def test_stage_runtime_training_loop_with_flushes_and_recompute():
    """``StageRuntime`` API parity (BERT/runtime.py:842-900): forward x update_interval, backward x update_interval,
    ``optimizer.step()``; ``--recompute_step`` activations; named-tensor stage wiring."""
    from oktopk_b200.models.bert import BertConfig, synthetic_batch
    from oktopk_b200.optimizer import BertAdam
    from oktopk_b200.train.stage_runtime import InputSource, StageRuntime, bert_stage_model
    torch.manual_seed(0)
    cfg = BertConfig(vocab_size=2000, hidden_size=32, num_hidden_layers=4, num_attention_heads=2, intermediate_size=64,
                     max_position_embeddings=64, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = bert_stage_model(cfg, 4)
    assert len(model) == 5                                   # 4 stages + criterion
    batches = [synthetic_batch(4, 16, vocab=2000, generator=torch.Generator().manual_seed(i)) for i in range(4)]

    def run(recompute):
        torch.manual_seed(1)
        m = bert_stage_model(cfg, 4)
        r = StageRuntime(m, device=torch.device("cpu"), enable_recompute=recompute)
        r.set_input_source(InputSource(list(batches)))
        opt = BertAdam(list(r.parameters()), lr=1e-3, warmup=0.1, t_total=100, named_parameters=list(r.named_parameters()),
                       compressor="oktopk", density=0.05)
        r.run_training_loop_with_flushes(4, opt, update_interval=2)
        loss = float(r.run_forward().detach())
        flat = torch.cat([p.detach().view(-1) for p in r.parameters()])
        opt.close()
        return loss, flat

    l0, f0 = run(False)
    l1, f1 = run(True)
    assert l0 == l0 and abs(l0 - l1) < 1e-4                   # finite, and recompute does not change the math
    torch.testing.assert_close(f0, f1, rtol=1e-4, atol=1e-6)


def test_flops_counter_matches_known_models():
    from oktopk_b200.models import create_net
    from oktopk_b200.utils.flops import count_flops, get_model_complexity_info
    net, _ = create_net(10, "vgg16")
    macs, params, per = count_flops(net, torch.zeros(1, 3, 32, 32))
    assert params == 14_728_266 and 3.1e8 < macs < 3.2e8 and per["conv"] > 0.99 * 3.13e8
    s_macs, s_params = get_model_complexity_info(net, (3, 32, 32))
    assert s_macs.endswith("MMac") and s_params == "14.73 M"


def test_ptb_and_an4_file_loaders(tmp_path):
    """Real-data loaders (used when the files are on disk; synthetic otherwise): PTB text and an AN4-style manifest."""
    import numpy as np
    from scipy.io import wavfile
    from oktopk_b200.models.deepspeech import AN4_LABELS
    from oktopk_b200.train.data import AN4Manifest, PTBText, an4_collate, build_dataset
    (tmp_path / "ptb.train.txt").write_text(" the cat sat on the mat \n the dog sat \n" * 30)
    (tmp_path / "ptb.valid.txt").write_text(" the cat sat \n" * 10)
    ds = build_dataset("ptb", str(tmp_path), train=True, num_steps=5)
    assert isinstance(ds, PTBText) and ds.vocab["the"] == 0 and "<eos>" in ds.vocab
    x, y = ds[0]
    assert x.shape == (5,) and torch.equal(x[1:], y[:-1])
    sr = 16000
    for i, text in enumerate(["HELLO WORLD", "GO"]):
        t = np.arange(int(sr * (0.5 + 0.3 * i))) / sr
        wavfile.write(str(tmp_path / ("u%d.wav" % i)), sr, (0.3 * np.sin(2 * np.pi * 440 * (i + 1) * t) * 32767).astype(np.int16))
        (tmp_path / ("u%d.txt" % i)).write_text(text)
    (tmp_path / "an4_train_manifest.csv").write_text("u0.wav,u0.txt\nu1.wav,u1.txt\n")
    an4 = build_dataset("an4", str(tmp_path), train=True)
    assert isinstance(an4, AN4Manifest) and len(an4) == 2
    spec, tgt = an4[0]
    assert spec.shape[0] == 161 and spec.shape[1] == 51 and tgt.numel() == len("HELLO WORLD")
    assert [AN4_LABELS[int(i)] for i in tgt] == list("HELLO WORLD")
    inputs, targets, pct, sizes = an4_collate([an4[0], an4[1]])
    assert inputs.shape[:3] == (2, 1, 161) and sizes.tolist() == [2, 11] or sizes.tolist() == [11, 2]


def test_pretraining_data_creator_partitions_and_partitioned_dataset(tmp_path):
    """Pre-created instance pipeline (``sources.py`` / ``dataset.py:93-227``): create -> shard -> load -> mask on the fly."""
    from oktopk_b200.train.bert_data import (BERTDatasetPartitioned, PretrainingDataCreator, TokenInstance,
                                             get_random_partition, synthetic_corpus)
    tok = BertTokenizer.synthetic(3000)
    lines = synthetic_corpus(n_docs=20, sents=6, words=9)
    (tmp_path / "corpus.txt").write_text("\n".join(lines))
    pc = PretrainingDataCreator.from_corpus(str(tmp_path / "corpus.txt"), tok, max_seq_length=48, dupe_factor=2, seed=1)
    assert len(pc) >= 40 and all(isinstance(x, TokenInstance) for x in pc.instances)
    assert all(len(x.tokens_a) + len(x.tokens_b) <= 45 and x.tokens_a and x.tokens_b for x in pc.instances)
    frac_random = sum(x.is_next for x in pc.instances) / len(pc)
    assert 0.2 < frac_random < 0.8
    paths = pc.save_partitions(str(tmp_path / "parts"), 3)
    assert len(paths) == 3 and get_random_partition(str(tmp_path / "parts"), 4) == paths[1]
    ds = BERTDatasetPartitioned(tok, str(tmp_path / "parts"), max_seq_length=48, max_predictions_per_seq=5)
    assert len(ds) == len(pc)
    ids, seg, mask, lm, nxt = ds[7]
    assert ids.shape == seg.shape == mask.shape == lm.shape == (48,) and int(nxt) in (0, 1)
    n_masked = int((lm != -1).sum())
    assert 1 <= n_masked <= 5 and int(ids[0]) == tok.vocab["[CLS]"]
    again = ds[7]
    assert all(torch.equal(x, y) for x, y in zip(ds[7], again))
    # <sep>-joined single-line documents (the Wikipedia creator's input format)
    (tmp_path / "wiki.txt").write_text("\n".join("<sep>".join(lines[i * 7:i * 7 + 6]) for i in range(10)))
    pw = PretrainingDataCreator.from_corpus(str(tmp_path / "wiki.txt"), tok, sep="<sep>", max_seq_length=32, dupe_factor=1)
    assert len(pw) > 0
