"""Analytic cost models: the reference's alpha-beta tables for Ethernet clusters
(``VGG/utils.py:62-134``) next to the B200 / NVLink-5 numbers this library is built for, and the
per-scheme communication-volume formulas of the paper's Table 1."""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

# (alpha [s], beta [s/byte]) per P -- hard-coded in the reference for GbE / 10GbE
GBE = {2: (1.6e-3, 1.0e-8), 4: (2.7e-3, 1.3e-8), 8: (4.0e-3, 1.5e-8), 16: (1.7e-3, 1.7e-8)}
TEN_GBE = {2: (1.5e-5, 5.7e-11), 4: (3.6e-5, 1.1e-10), 8: (8.5e-5, 1.4e-10), 16: (1.4e-4, 2.0e-10)}
# B200 HGX: NVLink 5 through NVSwitch (measured peer copy 770 GB/s/dir, flag round trip ~2-4 us)
NVLINK5 = {p: (3.0e-6, 1.0 / 770e9) for p in (2, 4, 8)}
HBM_BYTES_PER_S = 6.48e12


def topk(tensor: np.ndarray, k: int):
    """NumPy top-k by magnitude (``VGG/utils.py:19-24``)."""
    idx = np.argpartition(np.abs(tensor), -k)[-k:]
    return idx, tensor[idx]


def volume_elems(scheme: str, n: int, k: int, P: int) -> float:
    """Per-rank scalars sent+received (paper Table 1; Ok-Topk is the upper bound 6k(P-1)/P)."""
    s = scheme.lower()
    if s in ("dense", "none"):
        return 2.0 * n * (P - 1) / P
    if s in ("topka", "topka2", "topkaopt", "gaussiank", "gaussiankconcat"):
        return 2.0 * k * (P - 1)
    if s in ("topksa", "topkdsa", "gaussianksa"):
        return 4.0 * k * (P - 1) / P
    if s == "gtopk":
        return 4.0 * k * math.log2(max(P, 2))
    if s == "oktopk":
        return 6.0 * k * (P - 1) / P
    raise KeyError(scheme)


def allreduce_time(scheme: str, n: int, k: int, P: int, table: Dict = NVLINK5, bytes_per_elem: int = 4) -> float:
    alpha, beta = table.get(P, table[max(table)])
    rounds = {"gtopk": 2 * math.log2(max(P, 2))}.get(scheme.lower(), 2.0)
    return rounds * alpha + volume_elems(scheme, n, k, P) * bytes_per_elem * beta


def oktopk_roofline(n: int, k: int, P: int) -> Dict[str, float]:
    """one streaming pass (16 B/element) vs 6k * 8 B over NVLink."""
    hbm = 16.0 * n / HBM_BYTES_PER_S
    link = 6.0 * k * 8.0 * (P - 1) / P / 770e9
    return {"hbm_s": hbm, "link_s": link, "floor_s": max(hbm, link)}
