"""Greedy CTC decoding and WER/CER (``VGG/decoder.py``: ``GreedyDecoder`` :105-197, ``wer`` :53-74,
``cer`` :76-86).  Pure-Python Levenshtein (the reference needs the ``python-Levenshtein`` C module)."""
from __future__ import annotations

from typing import List, Sequence

import torch


def levenshtein(a: Sequence, b: Sequence) -> int:
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def wer(s1: str, s2: str) -> int:
    """Word-level edit distance."""
    return levenshtein(s1.split(), s2.split())


def cer(s1: str, s2: str) -> int:
    """Character-level edit distance (spaces removed)."""
    return levenshtein(s1.replace(" ", ""), s2.replace(" ", ""))


class GreedyDecoder:
    def __init__(self, labels: str, blank_index: int = 0):
        self.labels = labels
        self.blank = blank_index

    def convert_targets(self, targets: torch.Tensor, sizes: torch.Tensor) -> List[str]:
        out, off = [], 0
        for s in sizes.tolist():
            out.append("".join(self.labels[int(i)] for i in targets[off:off + s]))
            off += s
        return out

    def decode(self, probs: torch.Tensor, sizes: torch.Tensor) -> List[str]:
        """``probs``: N x T x C; collapse repeats, drop blanks."""
        best = probs.argmax(2)
        res = []
        for n in range(best.size(0)):
            seq, prev = [], None
            for t in range(int(sizes[n])):
                c = int(best[n, t])
                if c != self.blank and c != prev:
                    seq.append(self.labels[c])
                prev = c
            res.append("".join(seq))
        return res
