"""Gradient compressors: selection + error-feedback residual bookkeeping.

Behavioural parity with ``VGG/compression.py`` (``NoneCompressor`` :11-21,
``TopKCompressor`` :24-165, ``GaussianCompressor`` :167-482, the nine name-only
subclasses :484-509 and the ``compressors`` registry :512-523) with two deliberate
differences: state is *per instance* (the reference keeps it in class-level
dicts, so two optimizers in one process collide), and the threshold math also exists as
fused sm_100a kernels (``oktopk_b200/csrc``) which the CUDA engine uses instead of these
torch-op formulations.  These methods are device agnostic (CPU and CUDA) and are what
the oracle, the gloo path and the NCCL baseline path call.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

from .config import OkTopkConfig


def gen_threshold_from_normal_distribution(p_value: float, mu: float, sigma: float) -> Tuple[float, float]:
    """``VGG/utils.py:136-138``: two-sided interval of N(mu, sigma) holding mass ``p_value``.

    Closed form via ``torch.special.ndtri`` (the reference calls ``scipy.stats.norm.ppf``).
    """
    q = (1.0 - p_value) / 2.0
    z = float(torch.special.ndtri(torch.tensor(q, dtype=torch.float64)))
    return mu + z * sigma, mu - z * sigma


class Compressor:
    """Base: holds the per-bucket residuals and last selection."""

    name = "base"
    family = "none"

    def __init__(self, cfg: Optional[OkTopkConfig] = None):
        self.cfg = cfg or OkTopkConfig()
        self.residuals: Dict[str, torch.Tensor] = {}
        self.values: Dict[str, torch.Tensor] = {}
        self.indexes: Dict[str, torch.Tensor] = {}

    # -- state -------------------------------------------------------------
    def clear(self) -> None:
        self.residuals.clear()
        self.values.clear()
        self.indexes.clear()

    def residual(self, name: str, like: torch.Tensor) -> torch.Tensor:
        r = self.residuals.get(name)
        if r is None or r.shape != like.shape or r.device != like.device:
            r = torch.zeros_like(like)
            self.residuals[name] = r
        return r

    def get_residuals(self, name: str, like: torch.Tensor) -> torch.Tensor:
        return self.residual(name, like)

    def state_dict(self) -> Dict:
        return {"residuals": {k: v.detach().cpu().clone() for k, v in self.residuals.items()}}

    def load_state_dict(self, sd: Dict, device=None) -> None:
        self.residuals = {k: (v.to(device) if device is not None else v.clone())
                          for k, v in sd.get("residuals", {}).items()}

    # -- shared threshold helpers (VGG/compression.py:324-356) --------------
    @staticmethod
    def compressbythreshold(tensor: torch.Tensor, thres: float = 0.0):
        """Strict ``|x| > thres`` select; int32 indices + values (``:324-333``)."""
        with torch.no_grad():
            idx = (tensor.abs() > thres).nonzero(as_tuple=False).view(-1)
            val = tensor[idx]
            return idx.to(torch.int32), val

    @staticmethod
    def compressbythresholdlong(tensor: torch.Tensor, thres: float = 0.0):
        """Indices only, int64 (``:350-356``)."""
        with torch.no_grad():
            return (tensor.abs() > thres).nonzero(as_tuple=False).view(-1)

    @staticmethod
    def k2globalthreshold(tensor: torch.Tensor, k: int = 0):
        """Exact top-``min(len,k)`` by magnitude (``:407-415``) -> (values, positions, kth |value|)."""
        kk = min(tensor.numel(), k)
        with torch.no_grad():
            if kk <= 0:
                e = tensor.new_zeros(0)
                return e, torch.zeros(0, dtype=torch.long, device=tensor.device), 0.0
            vals, pos = torch.topk(tensor.abs(), k=kk)
            thr = float(vals[-1].item())
            return tensor[pos], pos, thr

    def update_residuals(self, involved_indexes: torch.Tensor, name: str) -> None:
        """``residual[involved] = 0`` (``:467-471``)."""
        with torch.no_grad():
            self.residuals[name][involved_indexes.long()] = 0.0


class NoneCompressor(Compressor):
    """Dense: identity (``VGG/compression.py:11-21``)."""

    name = "none"
    family = "none"

    @staticmethod
    def compress(tensor, name=None, ratio=None, **kw):
        return tensor, None

    @staticmethod
    def decompress(tensor, ctc=None, name=None):
        return tensor


class TopKCompressor(Compressor):
    """Exact top-k family (``VGG/compression.py:24-165``)."""

    name = "topk"
    family = "topk"

    def compress_org(self, tensor: torch.Tensor, name: str, ratio: float = 0.05, **kw):
        """``:37-62``: ``t += res``; keep exact top-k in ``t``; rest goes to the residual."""
        with torch.no_grad():
            res = self.residual(name, tensor)
            k = max(int(tensor.numel() * ratio), 1)
            tensor.add_(res)
            _, idx = torch.topk(tensor.abs(), k=k)
            vals = tensor[idx].clone()
            res.copy_(tensor)
            res[idx] = 0.0
            tensor.zero_()
            tensor[idx] = vals
            self.values[name] = vals
            self.indexes[name] = idx
            return tensor, idx

    def compress(self, tensor: torch.Tensor, name: str, ratio: float = 0.05, **kw):
        """``:65-83``: tensor stays dense, residual zeroed at the (unsorted) top-k."""
        with torch.no_grad():
            res = self.residual(name, tensor)
            k = max(int(tensor.numel() * ratio), 1)
            tensor.add_(res)
            vals, idx = torch.topk(tensor.abs(), k=k, sorted=False)
            vals = tensor[idx]
            res.copy_(tensor)
            res[idx] = 0.0
            self.values[name] = vals
            self.indexes[name] = idx
            return tensor, idx

    def ratio2threshold(self, tensor: torch.Tensor, name: str, ratio: float = 0.05) -> float:
        """``:86-106``: ``t += res``; residual = acc with exact top-k zeroed; returns k-th |value|."""
        with torch.no_grad():
            res = self.residual(name, tensor)
            k = max(int(tensor.numel() * ratio), 1)
            tensor.add_(res)
            vals, idx = torch.topk(tensor.abs(), k=k)
            res.copy_(tensor)
            res[idx] = 0.0
            return float(vals[-1].item())

    def add_residuals(self, included_indexes: Optional[torch.Tensor], name: str) -> None:
        """``:151-160``: put back local picks that did not survive globally."""
        with torch.no_grad():
            vals = self.values[name]
            if included_indexes is not None and included_indexes.numel() > 0:
                vals = vals.clone()
                vals[included_indexes.long()] = 0.0
            self.residuals[name][self.indexes[name]] += vals

    @staticmethod
    def decompress(tensor, ctc=None, name=None):
        return tensor


class GaussianCompressor(Compressor):
    """Threshold family (``VGG/compression.py:167-482``)."""

    name = "gaussiank"
    family = "gaussian"

    # -- Gaussiank proper ----------------------------------------------------
    def compress(self, tensor: torch.Tensor, name: str, ratio: float = 0.05, **kw):
        """``:220-266`` (+ LSTM/BERT correction flavours).

        ``t += res``; thr from a normal fit; bounded multiplicative search so the selected
        count lands near k; residual = acc with the selected entries zeroed.
        """
        cfg = self.cfg
        with torch.no_grad():
            res = self.residual(name, tensor)
            n = tensor.numel()
            k = max(int(n * ratio), 1)
            tensor.add_(res)
            std = float(torch.std(tensor)) if n > 1 else 0.0
            mean = float(torch.mean(tensor))
            _, thr = gen_threshold_from_normal_distribution(1.0 - ratio, mean, std)
            absx = tensor.abs()
            thr = gaussian_correct_threshold(absx, thr, k, cfg)
            idx = (absx > thr).nonzero(as_tuple=False).view(-1)
            vals = tensor[idx]
            res.copy_(tensor)
            res[idx] = 0.0
            return idx.to(torch.int32), vals

    def predictratio2threshold(self, tensor: torch.Tensor, name: str = None, ratio: float = 0.05):
        """``:307-320``: Gaussian-predicted threshold and the count it would select."""
        with torch.no_grad():
            std = float(torch.std(tensor))
            mean = float(torch.mean(tensor))
            _, thr = gen_threshold_from_normal_distribution(1.0 - ratio, mean, std)
            return thr, int((tensor.abs() > thr).sum())

    # -- Ok-Topk / topkAopt / gaussiankSA helpers ------------------------------
    def ratio2threshold(self, tensor: torch.Tensor, name: str, ratio: float = 0.05) -> float:
        """``:370-381``: ``t += res``; ``res = t`` (nothing zeroed); exact k-th |value|."""
        with torch.no_grad():
            res = self.residual(name, tensor)
            k = max(int(tensor.numel() * ratio), 1)
            tensor.add_(res)
            vals, _ = torch.topk(tensor.abs(), k=k)
            res.copy_(tensor)
            return float(vals[-1].item())

    def add2residual(self, tensor: torch.Tensor, name: str, thrd: float, tk: int) -> float:
        """``:384-404``: ``t += res; res = t``; raise the stale threshold while it over-selects."""
        cfg = self.cfg
        with torch.no_grad():
            res = self.residual(name, tensor)
            tensor.add_(res)
            res.copy_(tensor)
            if cfg.overselect_guard_loops <= 0:
                return thrd
            absx = tensor.abs()
            thres = thrd
            limit = cfg.overselect_guard_num * tk // cfg.overselect_guard_den
            for _ in range(cfg.overselect_guard_loops):
                if int((absx > thres).sum()) > limit:
                    thres *= cfg.overselect_guard_factor
                else:
                    break
            return thres

    def compressbythreshold_residual(self, tensor: torch.Tensor, name: str, thres: float = 0.0):
        """``:336-347``: threshold select that also zeroes the residual at the selection."""
        with torch.no_grad():
            idx = (tensor.abs() > thres).nonzero(as_tuple=False).view(-1)
            self.residuals[name][idx] = 0.0
            return idx.to(torch.int32), tensor[idx]


def gaussian_correct_threshold(absx: torch.Tensor, thr: float, k: int, cfg: OkTopkConfig) -> float:
    """The per-workload bounded search of Gaussiank (SURVEY A.1 'Gaussiank count correction')."""
    count = lambda t: int((absx > t).sum())  # noqa: E731
    init = count(thr)
    mode = cfg.gaussian_mode
    if mode == "vgg":
        lo, hi = 3 * k // 4, 5 * k // 4
        if init < lo:
            for _ in range(cfg.gaussian_loops):
                if count(thr) < lo:
                    thr /= cfg.gaussian_factor
                else:
                    break
        elif init > hi:
            for _ in range(cfg.gaussian_loops):
                if count(thr) > hi:
                    thr *= cfg.gaussian_factor
                else:
                    break
    elif mode == "lstm":
        lo = 3 * k // 4
        for _ in range(cfg.gaussian_loops):
            if count(thr) < lo:
                thr /= cfg.gaussian_factor
            else:
                break
    else:  # bert
        if init < 3 * k // 4:
            tgt = 5 * k // 6
            for _ in range(cfg.gaussian_loops):
                if count(thr) < tgt:
                    thr /= cfg.gaussian_factor
                else:
                    break
    return thr


# ---- name-only subclasses: the ``name`` is what the engine dispatches on -------
class TopKACompressor(TopKCompressor):
    name = "topkA"


class TopKACompressor2(TopKCompressor):
    name = "topkA2"


class TopKSACompressor(TopKCompressor):
    name = "topkSA"


class gTopKCompressor(TopKCompressor):
    name = "gtopk"


class TopKAoptCompressor(GaussianCompressor):
    name = "topkAopt"


class GaussianKCompressor(GaussianCompressor):
    name = "gaussiank"


class GaussianKConcatCompressor(GaussianCompressor):
    name = "gaussiankconcat"


class GaussianKSACompressor(GaussianCompressor):
    name = "gaussiankSA"


class OkTopKCompressor(GaussianCompressor):
    name = "oktopk"


compressors = {
    "topkA": TopKACompressor,
    "topkAopt": TopKAoptCompressor,
    "topkA2": TopKACompressor2,
    "topkSA": TopKSACompressor,
    "topkDSA": TopKSACompressor,      # the launch scripts' spelling (VGG/vgg16_topkDSA.sh:23)
    "gtopk": gTopKCompressor,
    "gaussiank": GaussianKCompressor,
    "gaussiankconcat": GaussianKConcatCompressor,
    "gaussiankSA": GaussianKSACompressor,
    "oktopk": OkTopKCompressor,
    "none": NoneCompressor,
    None: NoneCompressor,
}


def resolve_compressor(c, cfg: Optional[OkTopkConfig] = None) -> Compressor:
    """Accept a registry key, a compressor class (reference style) or an instance."""
    if isinstance(c, Compressor):
        if cfg is not None:
            c.cfg = cfg
        return c
    if isinstance(c, type) and issubclass(c, Compressor):
        return c(cfg)
    if c in compressors:
        return compressors[c](cfg)
    raise KeyError("unknown compressor %r (have %s)" % (c, sorted(k for k in compressors if k)))
