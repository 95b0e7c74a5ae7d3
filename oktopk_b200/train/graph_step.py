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

    # ------------------------------------------------------------------ iteration flavour
    def _engines(self):
        return [self.opt._allreducer._engines.get(b.name) for b in self.opt._buckets]

    def _key(self) -> Tuple:
        cfg = self.opt._cfg
        key = [("density", self.opt.get_current_density())]     # k and the guard limits are baked into the launch
        for eng in self._engines():
            if eng is None:
                return ("nograph",)
            c = eng.host.counter
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
            elif name in ("topkA2", "gtopk"):
                return ("nograph",)          # host-driven tree / re-selection: not capturable
            else:
                key.append(("every",))
        return tuple(key)

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
        if key == ("nograph",) or any(any(f is True for f in k) for k in key[1:]):
            # rare flavours (exact-threshold / re-partition iterations, 1 in 32..128) stay eager: capturing them
            # costs more than they save, and the common flavour is what the step time is made of
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
        g = self.graphs.get(key)
        if g is None:
            g = self._capture(key)
            if g is None:
                return self._eager(batch)
        else:
            for eng in self._engines():
                eng.host.counter += 1
            if hasattr(self.opt, "counter"):
                self.opt.counter += 1
        g.replay()
        ext.LAUNCH_COUNT["total"] += self.launches.get(key, 0)
        self.static_loss = self._loss_of[key]
        return self.static_loss

    def _capture(self, key) -> Optional[torch.cuda.CUDAGraph]:
        tr = self.tr
        counters = [eng.host.counter for eng in self._engines()]
        opt_counter = getattr(self.opt, "counter", None)
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
            self._loss_of = getattr(self, "_loss_of", {})
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
