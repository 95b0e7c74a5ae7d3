"""Observability: phase timers with CUDA events instead of ``time.time()``
(``VGG/allreducer.py:256-261,379-443``), a scalar writer (tensorboard if installed, JSONL always), NVTX
ranges, and the algorithm-quality probes of ``settings.PROFILING_NORM`` (relative error of the sparse
result against the true dense top-k, ``VGG/allreducer.py:1072-1080``)."""
from __future__ import annotations

import contextlib
import json
import os
import time
from collections import defaultdict
from typing import Dict, List, Optional

import torch


def force_insert_item(d: Dict[str, list], key: str, item) -> None:
    """``VGG/utils.py`` helper kept for API parity."""
    d.setdefault(key, []).append(item)


class PhaseTimers:
    """Named timers; ``cuda_range`` measures on the device with events (no host sync until ``summary``)."""

    def __init__(self, window: int = 50):
        self.window = window
        self.host: Dict[str, List[float]] = defaultdict(list)
        self._events: Dict[str, list] = defaultdict(list)

    def add(self, name: str, seconds: float) -> None:
        self.host[name].append(seconds)

    @contextlib.contextmanager
    def cuda_range(self, name: str, stream: Optional["torch.cuda.Stream"] = None):
        if not torch.cuda.is_available():
            t0 = time.perf_counter()
            yield
            self.add(name, time.perf_counter() - t0)
            return
        s = stream or torch.cuda.current_stream()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.nvtx.range_push(name)
        a.record(s)
        try:
            yield
        finally:
            b.record(s)
            torch.cuda.nvtx.range_pop()
            self._events[name].append((a, b))

    def summary(self, reset: bool = True) -> Dict[str, float]:
        out = {}
        if self._events:
            torch.cuda.synchronize()
        for k, evs in self._events.items():
            ts = [a.elapsed_time(b) * 1e-3 for a, b in evs]
            if ts:
                out[k] = sum(ts) / len(ts)
        for k, ts in self.host.items():
            if ts:
                out[k] = sum(ts) / len(ts)
        if reset:
            self._events.clear()
            self.host.clear()
        return out


class MetricsWriter:
    """Rank-0 scalar sink: ``<log_dir>/metrics.jsonl`` always, tensorboard ``SummaryWriter`` when importable
    (the reference uses tensorboardX on rank 0, ``VGG/main_trainer.py:170-172``)."""

    def __init__(self, log_dir: Optional[str], rank: int = 0):
        self.f = None
        self.tb = None
        if log_dir and rank == 0:
            os.makedirs(log_dir, exist_ok=True)
            self.f = open(os.path.join(log_dir, "metrics.jsonl"), "a")
            from . import settings
            if settings.TENSORBOARD:                  # OKTOPK_TENSORBOARD=1 (VGG/settings.py TENSORBOARD)
                try:
                    from torch.utils.tensorboard import SummaryWriter  # noqa: WPS433
                    self.tb = SummaryWriter(log_dir)
                except Exception:  # noqa: BLE001
                    self.tb = None

    def add_scalars(self, tag: str, scalars: Dict[str, float], step: int) -> None:
        if self.f is not None:
            self.f.write(json.dumps({"tag": tag, "step": step, **{k: float(v) for k, v in scalars.items()}}) + "\n")
            self.f.flush()
        if self.tb is not None:
            for k, v in scalars.items():
                self.tb.add_scalar("%s/%s" % (tag, k), float(v), step)

    def close(self) -> None:
        if self.f is not None:
            self.f.close()
            self.f = None
        if self.tb is not None:
            self.tb.close()


@torch.no_grad()
def sparsification_error(dense_mean: torch.Tensor, sparse_result: torch.Tensor, k: int) -> Dict[str, float]:
    """EPS of ``PROFILING_NORM``: ``|| topk(dense_mean) - sparse_result || / || dense_mean ||`` plus norms."""
    k = max(min(k, dense_mean.numel()), 1)
    idx = torch.topk(dense_mean.abs(), k).indices
    ideal = torch.zeros_like(dense_mean)
    ideal[idx] = dense_mean[idx]
    gn = float(dense_mean.norm())
    return {"eps": float((ideal - sparse_result).norm()) / max(gn, 1e-30), "grad_norm": gn,
            "topk_norm": float(ideal.norm()), "result_norm": float(sparse_result.norm()),
            "nnz": int((sparse_result != 0).sum())}
