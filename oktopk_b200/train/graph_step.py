"""Whole-step CUDA graphs: forward + backward + fused sparse allreduce + fused optimizer update
captured once per *iteration type* and replayed.

Small-batch workloads (VGG-16 at 16 images/GPU, the reference's configuration) are launch bound: a
step is several hundred kernels of a few microseconds each.  Everything on our path is capturable
because nothing depends on host-visible values: thresholds, region edges, slot cursors, flag epochs
and the learning rate live in device memory, and the persistent cooperative kernel is a normal graph
node.  The only host-side variation is *which* flavour of the kernel a step needs (exact threshold
re-computation / region re-partition iterations), so one graph is captured per flavour
the first time it occurs and the host picks the graph from the iteration counter.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

from ..ops import ext


class GraphedTrainStep:
    def __init__(self, trainer, warmup_eager: int = 3):
        self.tr = trainer
        self.opt = trainer.optimizer
        self.graphs: Dict[Tuple, torch.cuda.CUDAGraph] = {}
        self.launches: Dict[Tuple, int] = {}
        self.static_in: Optional[Tuple[torch.Tensor, ...]] = None
        self.static_loss: Optional[torch.Tensor] = None
        self.eager_left = warmup_eager
        self.enabled = True
        self.pool = None
        self.why_disabled = ""
        self._precaptured = False
        self._cap_counters, self._cap_opt_counter = [], None
        self._loss_of: Dict[Tuple, torch.Tensor] = {}

    # ------------------------------------------------------------------ iteration flavour
    def _engines(self):
        return [self.opt._allreducer._engines.get(b.name) for b in self.opt._buckets]

    def _key(self, counters=None) -> Tuple:
        """The flavour of the step the engines are about to run (or would run at the given iteration counters)."""
        cfg = self.opt._cfg
        key = [("density", self.opt.get_current_density())]     # k and the guard limits are baked into the launch
        engines = self._engines()
        if counters is None:
            counters = [None if e is None else e.host.counter for e in engines]
        for eng, c in zip(engines, counters):
            if eng is None:
                return ("nograph",)
            if (not cfg.sparse) or c < cfg.warmup_iters:
                key.append(("dense",))
                continue
            it = c - cfg.warmup_iters
            name = self.opt._allreducer.compressor.name
            if name == "oktopk":
                key.append((it % cfg.local_recompute_interval == 0, it % cfg.global_recompute_interval == 0,
                            it % cfg.repartition_interval == 0))
            elif name == "topkAopt":
                key.append((it % cfg.topkaopt_recompute_interval == 0,))
            else:
                key.append(("every",))
        return tuple(key)

    def _sparse_flavours(self):
        """Every (key, representative sparse-iteration index) the schedule can produce -- a handful: for Ok-Topk the
        common threshold-reuse step, the exact-threshold step (1 in tau) and the exact + re-partition step."""
        cfg = self.opt._cfg
        name = self.opt._allreducer.compressor.name
        if name == "oktopk":
            period = 1
            for v in (cfg.local_recompute_interval, cfg.global_recompute_interval, cfg.repartition_interval):
                period = period * v // math.gcd(period, v)
        elif name == "topkAopt":
            period = cfg.topkaopt_recompute_interval
        else:
            period = 1
        seen = {}
        n_eng = len(self._engines())
        for it in range(min(period, 1 << 16)):
            k = self._key([cfg.warmup_iters + it] * n_eng)
            if k not in seen:
                seen[k] = it
        return seen

    def precapture_sparse(self) -> int:
        """Capture the graph of EVERY sparse-phase flavour now (capturing records launches, it executes nothing), so that
        the rare exact-threshold / re-partition iterations are replayed like the common one instead of running eagerly
        inside somebody's timed region.  Engine iteration counters are faked for the capture and restored."""
        if not self.enabled or self.static_in is None:
            return 0
        cfg = self.opt._cfg
        if not cfg.sparse:
            return 0
        engines = self._engines()
        if any(e is None for e in engines):
            return 0
        real = [e.host.counter for e in engines]
        real_opt = getattr(self.opt, "counter", None)
        made = 0
        for key, it in self._sparse_flavours().items():
            if key in self.graphs or key == ("nograph",):
                continue
            for e in engines:
                e.host.counter = cfg.warmup_iters + it
            ok = self._capture(key) is not None
            for e, c in zip(engines, real):
                e.host.counter = c
            if real_opt is not None:
                self.opt.counter = real_opt
            if not ok:
                break
            made += 1
        self._precaptured = True
        # every long-lived object of the training process exists now (model, optimizer state, engines, graphs): move them
        # to the permanent generation so that the cyclic collector's full passes stop walking them -- a generation-2
        # collection otherwise stalls a 1.2 ms step loop for tens of milliseconds
        import gc
        gc.collect()
        gc.freeze()
        return made

    # ------------------------------------------------------------------ one step
    def _eager(self, batch) -> torch.Tensor:
        tr = self.tr
        self.opt.zero_grad()
        loss, _ = tr._forward_loss(batch)
        loss.backward()
        tr.update_model()
        return loss.detach()

    def step(self, batch) -> torch.Tensor:
        """Run one optimizer step on ``batch`` (device tensors); returns the (device) loss."""
        if not self.enabled or self.eager_left > 0:
            self.eager_left -= 1
            return self._eager(batch)
        key = self._key()
        if key == ("nograph",):
            return self._eager(batch)
        if self.static_in is None:
            self.static_in = tuple(t.clone() if torch.is_tensor(t) else t for t in batch)
        for s, t in zip(self.static_in, batch):
            if torch.is_tensor(t):
                if s.shape != t.shape:
                    self.enabled, self.why_disabled = False, "batch shape changed"
                    return self._eager(batch)
                s.copy_(t, non_blocking=True)
        self.opt.refresh_lr()
        if not self._precaptured and all(k != ("dense",) for k in key[1:]):
            self.precapture_sparse()             # first sparse step: capture every flavour of the schedule at once
        g = self.graphs.get(key)
        if g is None:
            g = self._capture(key)
            if g is None:
                return self._eager(batch)
            for eng, c in zip(self._engines(), self._cap_counters):
                eng.host.counter = c             # the capture ran the Python side effects; the replay below is the real step
            if hasattr(self.opt, "counter") and self._cap_opt_counter is not None:
                self.opt.counter = self._cap_opt_counter
        for eng in self._engines():
            eng.host.counter += 1
        if hasattr(self.opt, "counter"):
            self.opt.counter += 1
        g.replay()
        self.opt._allreducer.poll_faults()       # pinned host mirror of the device fault words: a plain load per bucket
        ext.LAUNCH_COUNT["total"] += self.launches.get(key, 0)
        self.static_loss = self._loss_of[key]
        return self.static_loss

    def _capture(self, key) -> Optional[torch.cuda.CUDAGraph]:
        tr = self.tr
        counters = [eng.host.counter for eng in self._engines()]
        opt_counter = getattr(self.opt, "counter", None)
        self._cap_counters, self._cap_opt_counter = counters, opt_counter
        l0 = ext.LAUNCH_COUNT["total"]
        g = torch.cuda.CUDAGraph()
        try:
            torch.cuda.synchronize()
            with torch.cuda.graph(g, pool=self.pool):
                self.opt.zero_grad()
                loss, _ = tr._forward_loss(self.static_in)
                loss.backward()
                tr.update_model()
                out = loss.detach()
            if self.static_loss is None:
                self.static_loss = out
            else:
                # every graph must write the same loss buffer: re-point through a copy node is not possible after
                # capture, so keep one buffer per graph and expose the latest
                self.static_loss = out
            self._loss_of[key] = out
            if self.pool is None:
                self.pool = g.pool()
        except Exception as e:  # noqa: BLE001 - fall back to eager for good
            self.enabled, self.why_disabled = False, "capture failed: %r" % (e,)
            # capture executed the Python side effects (counters) but no kernels: undo them
            for eng, c in zip(self._engines(), counters):
                eng.host.counter = c
            if opt_counter is not None:
                self.opt.counter = opt_counter
            try:
                torch.cuda.synchronize()
            except Exception:  # noqa: BLE001
                pass
            self.opt._after_step()
            return None
        self.launches[key] = ext.LAUNCH_COUNT["total"] - l0
        ext.LAUNCH_COUNT["total"] = l0
        self.graphs[key] = g
        return g

    def loss_tensor(self, key=None) -> torch.Tensor:
        return self.static_loss
