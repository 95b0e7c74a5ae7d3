"""DeepSpeech-style acoustic model "lstman4" (the reference's LSTM workload).

Architecture parity with ``VGG/models/lstm_models.py:148-239`` and the factory defaults of
``VGG/models/lstman4.py:8-33`` (hidden 800, 5 uni-directional LSTM layers, look-ahead context 20,
29 labels): two masked Conv2d+BN+Hardtanh blocks over the (freq, time) spectrogram, a stack of
BatchNorm+LSTM layers on packed sequences, a look-ahead convolution, BatchNorm + bias-free Linear.
27,569,568 parameters in 40 tensors.  Input ``(B, 1, 161, T)`` + lengths; output ``(B, T', classes)``
logits (softmax only in eval mode) + output lengths; trained with CTC.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

AN4_LABELS = "_'ABCDEFGHIJKLMNOPQRSTUVWXYZ "     # 29 symbols, index 0 = CTC blank


class _TimeMaskedConv(nn.Module):
    """Conv stack that re-zeroes the padded time steps after every layer (``MaskConv``, :43-70)."""

    def __init__(self, seq: nn.Sequential):
        super().__init__()
        self.seq_module = seq

    def forward(self, x: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
        t = torch.arange(x.size(3), device=x.device).view(1, 1, 1, -1)
        for m in self.seq_module:
            x = m(x)
            if t.size(3) != x.size(3):
                t = torch.arange(x.size(3), device=x.device).view(1, 1, 1, -1)
            x = x.masked_fill(t >= lengths.to(x.device).view(-1, 1, 1, 1), 0)
        return x


class _SeqBN(nn.Module):
    """BatchNorm1d applied on (T*N, H) (``SequenceWise``, :21-40)."""

    def __init__(self, module: nn.Module):
        super().__init__()
        self.module = module

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        t, n = x.size(0), x.size(1)
        return self.module(x.reshape(t * n, -1)).view(t, n, -1)


class BatchRNN(nn.Module):
    def __init__(self, input_size: int, hidden_size: int, rnn_type=nn.LSTM, bidirectional: bool = False,
                 batch_norm: bool = True):
        super().__init__()
        self.bidirectional = bidirectional
        self.batch_norm = _SeqBN(nn.BatchNorm1d(input_size)) if batch_norm else None
        self.rnn = rnn_type(input_size=input_size, hidden_size=hidden_size, bidirectional=bidirectional, bias=True)

    def forward(self, x: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
        if self.batch_norm is not None:
            x = self.batch_norm(x)
        total = x.size(0)
        x = nn.utils.rnn.pack_padded_sequence(x, lengths.cpu(), enforce_sorted=False)
        x, _ = self.rnn(x)
        x, _ = nn.utils.rnn.pad_packed_sequence(x, total_length=total)
        if self.bidirectional:
            x = x.view(x.size(0), x.size(1), 2, -1).sum(2)
        return x


class Lookahead(nn.Module):
    """Look-ahead convolution (Wang et al. 2016): per-feature weighted sum over the next ``context``
    frames -- a depthwise 1-D convolution (the reference materialises a TxNxHx(context+1) tensor, :119-132)."""

    def __init__(self, n_features: int, context: int):
        super().__init__()
        assert context > 0
        self.n_features, self.context = n_features, context
        self.weight = nn.Parameter(torch.empty(n_features, context + 1))
        stdv = 1.0 / math.sqrt(context + 1)
        nn.init.uniform_(self.weight, -stdv, stdv)

    def forward(self, x: torch.Tensor) -> torch.Tensor:          # T x N x H
        y = F.pad(x.permute(1, 2, 0), (0, self.context))          # N x H x (T + context)
        y = F.conv1d(y, self.weight.unsqueeze(1), groups=self.n_features)
        return y.permute(2, 0, 1).contiguous()


class DeepSpeech(nn.Module):
    def __init__(self, rnn_hidden_size: int = 800, nb_layers: int = 5, labels: str = AN4_LABELS,
                 rnn_type=nn.LSTM, bidirectional: bool = False, context: int = 20, sample_rate: int = 16000,
                 window_size: float = 0.02):
        super().__init__()
        self._labels = labels
        self._bidirectional = bidirectional
        num_classes = len(labels)
        self.conv = _TimeMaskedConv(nn.Sequential(
            nn.Conv2d(1, 32, kernel_size=(41, 11), stride=(2, 2), padding=(20, 5)),
            nn.BatchNorm2d(32), nn.Hardtanh(0, 20, inplace=True),
            nn.Conv2d(32, 32, kernel_size=(21, 11), stride=(2, 1), padding=(10, 5)),
            nn.BatchNorm2d(32), nn.Hardtanh(0, 20, inplace=True)))
        f = int(math.floor(sample_rate * window_size / 2) + 1)          # 161 frequency bins
        f = int(math.floor(f + 2 * 20 - 41) / 2 + 1)
        f = int(math.floor(f + 2 * 10 - 21) / 2 + 1)
        rnn_in = f * 32
        rnns = [BatchRNN(rnn_in, rnn_hidden_size, rnn_type, bidirectional, batch_norm=False)]
        for _ in range(nb_layers - 1):
            rnns.append(BatchRNN(rnn_hidden_size, rnn_hidden_size, rnn_type, bidirectional))
        self.rnns = nn.ModuleList(rnns)
        self.lookahead = None if bidirectional else nn.Sequential(
            Lookahead(rnn_hidden_size, context=context), nn.Hardtanh(0, 20, inplace=True))
        self.fc = _SeqBN(nn.Sequential(nn.BatchNorm1d(rnn_hidden_size),
                                       nn.Linear(rnn_hidden_size, num_classes, bias=False)))

    def get_seq_lens(self, input_length: torch.Tensor) -> torch.Tensor:
        seq = input_length
        for m in self.conv.modules():
            if isinstance(m, nn.Conv2d):
                seq = (seq + 2 * m.padding[1] - m.dilation[1] * (m.kernel_size[1] - 1) - 1) // m.stride[1] + 1
        return seq.int()

    def forward(self, x: torch.Tensor, lengths: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        out_lens = self.get_seq_lens(lengths.cpu().int())
        x = self.conv(x, out_lens)
        b, c, d, t = x.size()
        x = x.view(b, c * d, t).permute(2, 0, 1).contiguous()        # T x N x H
        for rnn in self.rnns:
            x = rnn(x, out_lens)
        if self.lookahead is not None:
            x = self.lookahead(x)
        x = self.fc(x).transpose(0, 1)                                # N x T x classes
        if not self.training:
            x = F.softmax(x, dim=-1)
        return x, out_lens


def lstman4(hidden_size: int = 800, hidden_layers: int = 5, bidirectional: bool = False) -> DeepSpeech:
    """``VGG/models/lstman4.py:8`` defaults."""
    return DeepSpeech(rnn_hidden_size=hidden_size, nb_layers=hidden_layers, bidirectional=bidirectional)


class PTBLSTM(nn.Module):
    """2-layer 1500-hidden word-level LSTM language model (``VGG/models/lstm.py:5-40``), vocab 10k."""

    def __init__(self, vocab_size: int = 10000, embedding_dim: int = 1500, num_steps: int = 35, batch_size: int = 20,
                 num_layers: int = 2, dp_keep_prob: float = 0.35):
        super().__init__()
        self.embedding_dim, self.num_layers = embedding_dim, num_layers
        self.dropout = nn.Dropout(1 - dp_keep_prob)
        self.word_embeddings = nn.Embedding(vocab_size, embedding_dim)
        self.lstm = nn.LSTM(embedding_dim, embedding_dim, num_layers=num_layers, dropout=1 - dp_keep_prob)
        self.sm_fc = nn.Linear(embedding_dim, vocab_size)
        for w in (self.word_embeddings.weight, self.sm_fc.weight):
            nn.init.uniform_(w, -0.1, 0.1)
        nn.init.zeros_(self.sm_fc.bias)

    def init_hidden(self, batch_size: int, device=None):
        z = torch.zeros(self.num_layers, batch_size, self.embedding_dim, device=device)
        return (z, z.clone())

    def forward(self, inputs: torch.Tensor, hidden):
        emb = self.dropout(self.word_embeddings(inputs))
        out, hidden = self.lstm(emb, hidden)
        out = self.dropout(out)
        logits = self.sm_fc(out.view(-1, self.embedding_dim))
        return logits.view(inputs.size(0), inputs.size(1), -1), hidden
