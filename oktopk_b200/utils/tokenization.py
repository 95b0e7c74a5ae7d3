"""WordPiece tokenisation for the BERT workload (capability parity with
``BERT/bert/transformers/tokenization.py:51-392``: ``BertTokenizer`` = basic tokenizer + greedy
longest-match WordPiece, ``load_vocab``, ``convert_tokens_to_ids`` / ``convert_ids_to_tokens``).

Fresh implementation: the basic tokenizer is one regular-expression pass over NFD-normalised text instead
of the reference's character-by-character state machines, and WordPiece matching walks a prefix set.
There is no network here, so ``from_pretrained`` only accepts local paths; ``synthetic_vocab`` builds a
deterministic vocabulary of the bert-base-uncased size (30522) for shape-faithful synthetic runs.
"""
from __future__ import annotations

import collections
import os
import re
import unicodedata
from typing import Dict, Iterable, List, Optional

SPECIAL = ("[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]")

_CJK = ("一-鿿㐀-䶿\U00020000-\U0002a6df\U0002a700-\U0002b73f\U0002b740-\U0002b81f"
        "\U0002b820-\U0002ceaf豈-﫿\U0002f800-\U0002fa1f")


def _is_punct(ch: str) -> bool:
    cp = ord(ch)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(ch).startswith("P")


def load_vocab(vocab_file: str) -> "collections.OrderedDict[str, int]":
    vocab = collections.OrderedDict()
    with open(vocab_file, "r", encoding="utf-8") as f:
        for i, line in enumerate(f):
            tok = line.rstrip("\n")
            if tok == "" and i > 0 and not line.strip("\n"):
                tok = line.strip()
            vocab[tok.strip()] = i
    return vocab


def synthetic_vocab(size: int = 30522) -> "collections.OrderedDict[str, int]":
    """Deterministic stand-in vocabulary: specials, single characters, ``##`` continuations, then filler words."""
    vocab = collections.OrderedDict()
    for t in SPECIAL:
        vocab[t] = len(vocab)
    for i in range(99 - len(vocab)):
        vocab["[unused%d]" % i] = len(vocab)
    chars = "abcdefghijklmnopqrstuvwxyz0123456789"
    for c in chars + "!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~":
        vocab.setdefault(c, len(vocab))
    for c in chars:
        vocab.setdefault("##" + c, len(vocab))
    syl = ["ka", "to", "mi", "re", "so", "la", "ne", "vu", "di", "po", "an", "er", "in", "on", "st", "th"]
    i = 0
    while len(vocab) < size:
        w = syl[i % 16] + syl[(i // 16) % 16] + (syl[(i // 256) % 16] if i >= 256 else "") + (str(i // 4096) if i >= 4096 else "")
        vocab.setdefault(w if i % 3 else "##" + w, len(vocab))
        i += 1
    return vocab


def whitespace_tokenize(text: str) -> List[str]:
    text = text.strip()
    return text.split() if text else []


class BasicTokenizer:
    """Clean, (optionally) lower-case + strip accents, isolate punctuation and CJK characters."""

    def __init__(self, do_lower_case: bool = True, never_split: Iterable[str] = SPECIAL):
        self.do_lower_case = do_lower_case
        self.never_split = set(never_split)
        self._cjk = re.compile("([%s])" % _CJK)

    def _clean(self, text: str) -> str:
        out = []
        for ch in text:
            cp = ord(ch)
            if cp == 0 or cp == 0xFFFD:
                continue
            cat = unicodedata.category(ch)
            if ch in "\t\n\r" or cat == "Zs":
                out.append(" ")
            elif cat.startswith("C"):
                continue
            else:
                out.append(ch)
        return "".join(out)

    def _split_punct(self, word: str) -> List[str]:
        pieces, cur = [], []
        for ch in word:
            if _is_punct(ch):
                if cur:
                    pieces.append("".join(cur))
                    cur = []
                pieces.append(ch)
            else:
                cur.append(ch)
        if cur:
            pieces.append("".join(cur))
        return pieces

    def tokenize(self, text: str) -> List[str]:
        text = self._cjk.sub(r" \1 ", self._clean(text))
        out: List[str] = []
        for w in whitespace_tokenize(text):
            if w in self.never_split:
                out.append(w)
                continue
            if self.do_lower_case:
                w = "".join(c for c in unicodedata.normalize("NFD", w.lower()) if unicodedata.category(c) != "Mn")
            out.extend(self._split_punct(w))
        return out


class WordpieceTokenizer:
    """Greedy longest-match-first: ``unaffable`` -> ``un ##aff ##able``."""

    def __init__(self, vocab: Dict[str, int], unk_token: str = "[UNK]", max_input_chars_per_word: int = 100):
        self.vocab, self.unk, self.maxc = vocab, unk_token, max_input_chars_per_word

    def tokenize(self, text: str) -> List[str]:
        out: List[str] = []
        for word in whitespace_tokenize(text):
            if len(word) > self.maxc:
                out.append(self.unk)
                continue
            pieces, start, bad = [], 0, False
            while start < len(word):
                end, cur = len(word), None
                while start < end:
                    sub = word[start:end] if start == 0 else "##" + word[start:end]
                    if sub in self.vocab:
                        cur = sub
                        break
                    end -= 1
                if cur is None:
                    bad = True
                    break
                pieces.append(cur)
                start = end
            out.extend([self.unk] if bad else pieces)
        return out


class BertTokenizer:
    def __init__(self, vocab_file: Optional[str] = None, do_lower_case: bool = True, max_len: Optional[int] = None,
                 never_split: Iterable[str] = SPECIAL, vocab: Optional[Dict[str, int]] = None):
        if vocab is None:
            if vocab_file is None or not os.path.isfile(vocab_file):
                raise ValueError("Can't find a vocabulary file at path %r (there is no download here; pass a local "
                                 "vocab.txt or vocab=synthetic_vocab())" % (vocab_file,))
            vocab = load_vocab(vocab_file)
        self.vocab = vocab
        self.ids_to_tokens = collections.OrderedDict((i, t) for t, i in vocab.items())
        self.basic_tokenizer = BasicTokenizer(do_lower_case, never_split)
        self.wordpiece_tokenizer = WordpieceTokenizer(self.vocab)
        self.max_len = max_len if max_len is not None else int(1e12)

    @classmethod
    def from_pretrained(cls, path: str, *a, **kw) -> "BertTokenizer":
        vf = os.path.join(path, "vocab.txt") if os.path.isdir(path) else path
        return cls(vf, *a, **kw)

    @classmethod
    def synthetic(cls, size: int = 30522, **kw) -> "BertTokenizer":
        return cls(vocab=synthetic_vocab(size), **kw)

    def tokenize(self, text: str) -> List[str]:
        out = []
        for tok in self.basic_tokenizer.tokenize(text):
            out.extend([tok] if tok in self.basic_tokenizer.never_split else self.wordpiece_tokenizer.tokenize(tok))
        return out

    def convert_tokens_to_ids(self, tokens: Iterable[str]) -> List[int]:
        unk = self.vocab.get("[UNK]", 0)
        ids = [self.vocab.get(t, unk) for t in tokens]
        if len(ids) > self.max_len:
            raise ValueError("Token indices sequence length is longer than the specified maximum sequence length "
                             "for this BERT model (%d > %d)" % (len(ids), self.max_len))
        return ids

    def convert_ids_to_tokens(self, ids: Iterable[int]) -> List[str]:
        return [self.ids_to_tokens[int(i)] for i in ids]

    def __len__(self) -> int:
        return len(self.vocab)
