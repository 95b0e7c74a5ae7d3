"""BERT for pre-training (MLM + NSP) with the reference's *untied* decoder.

Architecture parity with ``BERT/bert/transformers/modeling.py:59-522`` (``BertConfig``, embeddings,
self-attention, ``BertLayer``, pooler, ``BertPreTrainingHeads``) and with the stage-split module lists
``BERT/bert/models/bert/depth=N`` (``StartingStage`` = embeddings + L/N layers, ``IntermediateStage``,
``EndingStage`` = layers + pooler + heads, whose decoder matrix is a *fresh* embedding-shaped
parameter, ``depth=4/__init__.py:17``) => BERT-base has 133,547,324 parameters, the paper's 133.5 M.

Fresh implementation: fused QKV projection, ``F.scaled_dot_product_attention`` (the reference does
matmul -> softmax -> matmul on the full [B,12,S,S] score tensor), ``F.layer_norm`` instead of apex.
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class BertConfig:
    vocab_size: int = 30522
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    hidden_act: str = "gelu"
    hidden_dropout_prob: float = 0.1
    attention_probs_dropout_prob: float = 0.1
    max_position_embeddings: int = 512
    type_vocab_size: int = 2
    initializer_range: float = 0.02
    layer_norm_eps: float = 1e-12

    @classmethod
    def from_json_file(cls, path: str) -> "BertConfig":
        with open(path) as f:
            d = json.load(f)
        return cls(**{k: v for k, v in d.items() if k in cls.__dataclass_fields__})

    @classmethod
    def bert_base(cls) -> "BertConfig":
        return cls()

    @classmethod
    def bert_large(cls) -> "BertConfig":
        return cls(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096)


def _act(name: str):
    return {"gelu": F.gelu, "relu": F.relu, "tanh": torch.tanh}[name]


class BertEmbeddings(nn.Module):
    def __init__(self, c: BertConfig):
        super().__init__()
        self.word_embeddings = nn.Embedding(c.vocab_size, c.hidden_size)
        self.position_embeddings = nn.Embedding(c.max_position_embeddings, c.hidden_size)
        self.token_type_embeddings = nn.Embedding(c.type_vocab_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)

    def forward(self, input_ids: torch.Tensor, token_type_ids: torch.Tensor) -> torch.Tensor:
        pos = torch.arange(input_ids.size(1), device=input_ids.device).unsqueeze(0)
        e = self.word_embeddings(input_ids) + self.position_embeddings(pos) + self.token_type_embeddings(token_type_ids)
        return self.dropout(self.LayerNorm(e))


class BertSelfAttention(nn.Module):
    def __init__(self, c: BertConfig):
        super().__init__()
        self.h, self.dh = c.num_attention_heads, c.hidden_size // c.num_attention_heads
        self.qkv = nn.Linear(c.hidden_size, 3 * c.hidden_size)          # same parameter count as 3 separate projections
        self.p_drop = c.attention_probs_dropout_prob

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
        b, s, _ = x.shape
        q, k, v = self.qkv(x).view(b, s, 3, self.h, self.dh).permute(2, 0, 3, 1, 4)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=self.p_drop if self.training else 0.0)
        return o.transpose(1, 2).reshape(b, s, self.h * self.dh)


class BertLayer(nn.Module):
    def __init__(self, c: BertConfig):
        super().__init__()
        self.attention = BertSelfAttention(c)
        self.attn_out = nn.Linear(c.hidden_size, c.hidden_size)
        self.attn_norm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.intermediate = nn.Linear(c.hidden_size, c.intermediate_size)
        self.output = nn.Linear(c.intermediate_size, c.hidden_size)
        self.out_norm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)
        self.act = _act(c.hidden_act)

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
        a = self.dropout(self.attn_out(self.attention(x, mask)))
        x = self.attn_norm(x + a)
        f = self.dropout(self.output(self.act(self.intermediate(x))))
        return self.out_norm(x + f)


class BertPooler(nn.Module):
    def __init__(self, c: BertConfig):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.hidden_size)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return torch.tanh(self.dense(x[:, 0]))


class BertPreTrainingHeads(nn.Module):
    """MLM head (dense + act + LN + decoder) and NSP head.  ``decoder_weight`` is an independent
    ``[vocab, hidden]`` parameter (untied, see module docstring) plus a vocab-sized bias."""

    def __init__(self, c: BertConfig):
        super().__init__()
        self.transform = nn.Linear(c.hidden_size, c.hidden_size)
        self.transform_norm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.act = _act(c.hidden_act)
        self.decoder_weight = nn.Parameter(torch.empty(c.vocab_size, c.hidden_size).normal_(std=c.initializer_range))
        self.decoder_bias = nn.Parameter(torch.zeros(c.vocab_size))
        self.seq_relationship = nn.Linear(c.hidden_size, 2)

    def forward(self, seq: torch.Tensor, pooled: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        h = self.transform_norm(self.act(self.transform(seq)))
        return F.linear(h, self.decoder_weight, self.decoder_bias), self.seq_relationship(pooled)


def extended_attention_mask(input_mask: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """``(1 - m) * -10000`` additive mask ``[B,1,1,S]`` (``BERT/bert/main_bert.py:616-639``)."""
    return ((1.0 - input_mask[:, None, None, :].to(dtype)) * -10000.0)


# ---- stage-split module lists (depth=N) --------------------------------------------------------
class StartingStage(nn.Module):
    def __init__(self, c: BertConfig, n_layers: int):
        super().__init__()
        self.embeddings = BertEmbeddings(c)
        self.layers = nn.ModuleList(BertLayer(c) for _ in range(n_layers))

    def forward(self, input_ids, token_type_ids, mask):
        x = self.embeddings(input_ids, token_type_ids)
        for l in self.layers:
            x = l(x, mask)
        return x


class IntermediateStage(nn.Module):
    def __init__(self, c: BertConfig, n_layers: int):
        super().__init__()
        self.layers = nn.ModuleList(BertLayer(c) for _ in range(n_layers))

    def forward(self, x, mask):
        for l in self.layers:
            x = l(x, mask)
        return x


class EndingStage(nn.Module):
    def __init__(self, c: BertConfig, n_layers: int):
        super().__init__()
        self.layers = nn.ModuleList(BertLayer(c) for _ in range(n_layers))
        self.pooler = BertPooler(c)
        self.heads = BertPreTrainingHeads(c)

    def forward(self, x, mask):
        for l in self.layers:
            x = l(x, mask)
        return self.heads(x, self.pooler(x))


def build_stages(c: BertConfig, depth: int = 4) -> List[nn.Module]:
    """``models.bert<L>.depth=<N>``: N sub-modules run back to back on one GPU (the reference's
    data-parallel ``StageRuntime`` instantiates every stage on every rank, ``BERT/runtime.py:128-151``)."""
    L = c.num_hidden_layers
    assert depth >= 2 and L % depth == 0, "depth must divide the layer count"
    per = L // depth
    return [StartingStage(c, per)] + [IntermediateStage(c, per) for _ in range(depth - 2)] + [EndingStage(c, per)]


class PretrainingCriterion(nn.Module):
    """CE(MLM, ignore_index=-1) + CE(NSP) (``BERT/runtime.py:585-596``)."""

    def __init__(self, vocab_size: int):
        super().__init__()
        self.vocab_size = vocab_size

    def forward(self, prediction_scores, seq_relationship_score, masked_lm_labels, next_sentence_labels):
        mlm = F.cross_entropy(prediction_scores.view(-1, self.vocab_size), masked_lm_labels.view(-1), ignore_index=-1)
        nsp = F.cross_entropy(seq_relationship_score.view(-1, 2), next_sentence_labels.view(-1))
        return mlm + nsp


class BertForPreTraining(nn.Module):
    def __init__(self, config: Optional[BertConfig] = None, depth: int = 4, recompute: bool = False):
        super().__init__()
        self.config = config or BertConfig()
        self.recompute = recompute           # ``--recompute_step`` (BERT/runtime.py:546-557, modeling.py:414-431)
        self.stages = nn.ModuleList(build_stages(self.config, depth))
        self.criterion = PretrainingCriterion(self.config.vocab_size)
        self.apply(self._init)
        nn.init.normal_(self.stages[-1].heads.decoder_weight, std=self.config.initializer_range)

    def _init(self, m: nn.Module) -> None:
        if isinstance(m, (nn.Linear, nn.Embedding)):
            nn.init.normal_(m.weight, std=self.config.initializer_range)
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, masked_lm_labels=None,
                next_sentence_label=None):
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        mask = None if attention_mask is None else extended_attention_mask(attention_mask)
        x = self.stages[0](input_ids, token_type_ids, mask)
        for st in self.stages[1:-1]:
            if self.recompute and self.training:
                from torch.utils.checkpoint import checkpoint
                x = checkpoint(st, x, mask, use_reentrant=False)
            else:
                x = st(x, mask)
        scores, nsp = self.stages[-1](x, mask)
        if masked_lm_labels is not None and next_sentence_label is not None:
            return self.criterion(scores, nsp, masked_lm_labels, next_sentence_label)
        return scores, nsp


def bert_base(depth: int = 4) -> BertForPreTraining:
    return BertForPreTraining(BertConfig.bert_base(), depth)


def synthetic_batch(batch: int, seq: int, vocab: int = 30522, device="cpu", generator=None, mask_prob: float = 0.15):
    """Wikipedia-shaped synthetic pre-training batch: ``input_ids, segment_ids, input_mask, lm_label_ids`` of
    ``[B,S]`` int64 (-1 = not masked) and ``is_next [B]`` (``BERT/bert/main_bert.py:535-614`` feature layout)."""
    g = generator
    ids = torch.randint(1000, vocab, (batch, seq), generator=g)
    seg = (torch.arange(seq).unsqueeze(0) >= torch.randint(seq // 4, 3 * seq // 4, (batch, 1), generator=g)).long()
    lens = torch.randint(seq // 2, seq + 1, (batch, 1), generator=g)
    mask = (torch.arange(seq).unsqueeze(0) < lens).long()
    sel = (torch.rand(batch, seq, generator=g) < mask_prob) & mask.bool()
    labels = torch.where(sel, ids, torch.full_like(ids, -1))
    ids = torch.where(sel, torch.full_like(ids, 103), ids) * mask
    is_next = torch.randint(0, 2, (batch,), generator=g)
    out = (ids, seg, mask, labels, is_next)
    return tuple(t.to(device) for t in out)
