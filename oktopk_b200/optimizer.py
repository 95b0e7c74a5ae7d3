"""User-facing optimizer wrappers (L3): ``DistributedOptimizer`` and ``BertAdam``.

API parity with the reference:
``DistributedOptimizer(optimizer, named_parameters=None, compression=NoneCompressor, is_sparse=False,
err_handler=None, layerwise_times=None, sigma_scale=2.5, density=0.1, norm_clip=None, writer=None)``
(``VGG/distributed_optimizer.py:203-207``) with ``step() / synchronize() / zero_grad() / stop() /
add_train_epoch() / get_current_density()`` and the ``local`` gradient-accumulation gate (:78,186);
``BertAdam(params, lr, warmup, t_total, ..., density, compressor, rank)``
(``BERT/bert/transformers/optimization.py:68-227``).

What is different by design (B200-first):
  * no consumer thread, no queues, no per-hook ``torch.cuda.synchronize()`` (:58-59,90,93): a
    bucket's reduction is enqueued on a side CUDA stream from the post-accumulate hook of its
    last gradient, in a fixed bucket order on every rank, and ``synchronize()`` is a stream wait;
  * gradients / parameters / momentum are views into flat buffers, the update is one fused kernel
    per (bucket, param group) that also zeroes the gradient bucket (``zero_grad()`` is then free);
  * any torch optimizer can be wrapped (the reference re-implements SGD only, A.4-7): SGD and
    BertAdam take the fused kernels, everything else falls through to its own ``step()``;
  * ``state_dict()`` carries residuals, thresholds, region boundaries and counters.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional

import torch

from .compression import NoneCompressor, compressors, resolve_compressor
from .config import OkTopkConfig
from .ops import ext
from .parallel.allreducer import AllReducer
from .parallel.buckets import Bucket, attach, build_buckets
from .parallel.world import World, world as _world


# ====================================================================================== comm mixin
class _BucketedComm:
    """Bucket bookkeeping, autograd hooks, stream choreography.  Mixed into optimizer classes."""

    def _okt_setup(self, named_parameters, allreducer: AllReducer, flatten_params: bool = True) -> None:
        if named_parameters is not None:
            named_parameters = list(named_parameters)
            if any(not isinstance(p, tuple) for p in named_parameters):
                raise ValueError("named_parameters should be a sequence of (name, parameter) tuples, "
                                 "usually produced by model.named_parameters().")
            names = {v: k for k, v in named_parameters}
        else:
            names = {}
        i = 0
        for g in self.param_groups:
            for p in g["params"]:
                if p not in names:
                    names[p] = "allreduce.noname.%d" % i
                i += 1
        self._parameter_names = names
        self._allreducer = allreducer
        self._cfg: OkTopkConfig = allreducer.cfg
        self.local = False
        self._synced = False
        self.momentum_correction = False
        self._buckets: List[Bucket] = build_buckets(self.param_groups, names, self._cfg.bucket_elems)
        self._bucket_of: Dict[torch.nn.Parameter, Bucket] = {}
        self._flat_state: Dict[int, Dict[str, torch.Tensor]] = {}
        self._next_launch = 0
        self._hook_handles = []
        self._comm_stream = None
        self._use_streams = False
        # Gradient landing: autograd produces every gradient in a fresh tensor; ONE multi-tensor kernel per bucket copies
        # them into the flat symmetric bucket when the bucket's last gradient is ready (instead of autograd accumulating
        # into pre-existing bucket views = one elementwise add kernel per parameter per step).
        self._land = False
        for b in self._buckets:
            dev = b.params[0].device
            grad = allreducer.register_bucket(b.name, b.numel, dev)
            attach(b, grad, flatten_params)
            if dev.type == "cuda" and self._cfg.land_grads and ext.available():
                self._land = True
            b.pending = len(b.params)
            b.dirty = False                       # freshly zeroed
            if dev.type == "cuda":
                self._use_streams = self._cfg.overlap
                b.event = torch.cuda.Event()
            for p in b.params:
                self._bucket_of[p] = b
                self._hook_handles.append(p.register_post_accumulate_grad_hook(self._make_hook(p)))
        if self._use_streams:
            # high priority: a bucket's (SM-partitioned) communication kernel should get its SMs as soon as backward
            # kernels retire CTAs, not after the whole backward queue
            self._comm_stream = torch.cuda.Stream(priority=-1)
        allreducer.add_resync_hook(self._resync_replicas)
        # device-resident learning rates (one float per param group): the fused update kernels read lr from
        # memory so that a captured CUDA graph of the whole step stays valid when the schedule moves
        self._lr_dev = None
        self._lr_pin, self._lr_ev, self._lr_ring, self._lr_last = [], [], 0, None
        dev0 = self._buckets[0].params[0].device if self._buckets else torch.device("cpu")
        if dev0.type == "cuda" and ext.available():
            G = len(self.param_groups)
            self._lr_dev = torch.zeros(max(G, 1), dtype=torch.float32, device=dev0)
            self._lr_pin = [torch.zeros(max(G, 1), dtype=torch.float32).pin_memory() for _ in range(8)]
            self._lr_ev = [None] * len(self._lr_pin)

    def _resync_replicas(self) -> None:
        """After a handled fault: every replica takes rank 0's parameters (and momentum) again."""
        w = self._allreducer.world
        if w.size == 1:
            return
        for b in self._buckets:
            if b.flat_param is not None:
                w.broadcast(b.flat_param, 0)
            else:
                for p in b.params:
                    w.broadcast(p.data, 0)
            for t in self._flat_state.get(b.index, {}).values():
                if torch.is_tensor(t):
                    w.broadcast(t, 0)

    def _group_lr(self, gi: int) -> float:
        return float(self.param_groups[gi]["lr"])

    def refresh_lr(self) -> None:
        """Push the current per-group learning rates to the device (one tiny async H2D, only when changed).
        Never called while a stream is capturing: graph replays call it right before ``replay()``."""
        if self._lr_dev is None:
            return
        vals = [self._group_lr(gi) for gi in range(len(self.param_groups))]
        if vals == self._lr_last:
            return
        self._lr_ring = (self._lr_ring + 1) % len(self._lr_pin)
        pin = self._lr_pin[self._lr_ring]
        if self._lr_ev[self._lr_ring] is not None:          # the host may run many (graph-replayed) steps ahead:
            self._lr_ev[self._lr_ring].synchronize()        # never overwrite a staging slot whose copy is pending
        for i, v in enumerate(vals):
            pin[i] = v
        self._lr_dev.copy_(pin, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._lr_ev[self._lr_ring] = ev
        self._lr_last = vals

    def _lr_ptr(self, gi: int) -> int:
        return 0 if self._lr_dev is None else self._lr_dev.data_ptr() + 4 * gi

    def _maybe_refresh_lr(self) -> None:
        if self._lr_dev is not None and not torch.cuda.is_current_stream_capturing():
            self.refresh_lr()

    # ------------------------------------------------------------------ hooks
    def _make_hook(self, p):
        def hook(param):
            b = self._bucket_of[p]
            b.dirty = True
            if self.local:                        # accumulation micro-step: no communication (:78)
                return
            b.pending -= 1
            if b.pending == 0:
                self._launch_ready()
        return hook

    def _launch_ready(self) -> None:
        # strictly in bucket order on every rank: the fused kernels spin on peers' flags, so two
        # ranks must never enqueue two buckets in opposite orders.
        while self._next_launch < len(self._buckets) and self._buckets[self._next_launch].pending <= 0:
            self._launch(self._buckets[self._next_launch])
            self._next_launch += 1

    def _land_bucket(self, b: Bucket) -> None:
        """Copy the autograd-produced gradients of bucket ``b`` into the flat bucket with one kernel launch and re-point
        ``p.grad`` at the bucket views (what the user sees after ``synchronize()`` is the reduced gradient)."""
        srcs, offs, numels = [], [], []
        for p, o, v in zip(b.params, b.offsets, b.grad_views):
            g = p.grad
            if g is None:                         # no gradient in this step (unused parameter, or a conv bias folded
                srcs.append(0)                    # into a fused batch-norm): zero-filled by the same launch
                offs.append(o)
                numels.append(p.numel())
            elif g.data_ptr() == v.data_ptr():
                continue                          # already accumulated in place (gradients were never detached)
            elif g.dtype == torch.float32 and g.is_cuda and g.stride() == v.stride():
                srcs.append(g.data_ptr())
                offs.append(o)
                numels.append(g.numel())
            else:
                v.copy_(g)
            p.grad = v
        if srcs:
            ext.require().land_grads(srcs, offs, numels, b.grad.data_ptr(), torch.cuda.current_stream().cuda_stream)

    def _launch(self, b: Bucket) -> None:
        if b.launched:
            return
        b.launched = True
        if self._land:
            self._land_bucket(b)
        if self.momentum_correction:
            self._apply_momentum_correction(b)
        if self._comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._comm_stream.wait_event(ev)
            with torch.cuda.stream(self._comm_stream):
                self._allreducer.reduce_bucket(b.name, b.grad, stream=self._comm_stream)
                b.event.record(self._comm_stream)
        else:
            self._allreducer.reduce_bucket(b.name, b.grad)

    def _apply_momentum_correction(self, b: Bucket) -> None:
        """``VGG/distributed_optimizer.py:81-88``: communicate the momentum-accumulated gradient."""
        fs = self._flat_state.setdefault(b.index, {})
        buf = fs.get("mc_buf")
        if buf is None:
            buf = fs["mc_buf"] = torch.zeros_like(b.grad)
        if b.grad.is_cuda and ext.available():
            ext.require().momentum_correct(b.grad.data_ptr(), buf.data_ptr(), b.numel, 0.9,
                                           torch.cuda.current_stream().cuda_stream)
        else:
            buf.mul_(0.9).add_(b.grad)
            b.grad.copy_(buf)

    # ------------------------------------------------------------------ public API
    def synchronize(self) -> None:
        """Block (stream-wise) until every bucket holds its reduced gradient."""
        if self._synced:
            return
        for b in self._buckets:                   # flush buckets whose hooks did not all fire (unused params)
            b.pending = 0
        self._launch_ready()
        if self._comm_stream is not None:
            cur = torch.cuda.current_stream()
            for b in self._buckets:
                cur.wait_event(b.event)
        self._synced = True

    def _after_step(self) -> None:
        for b in self._buckets:
            b.pending = len(b.params)
            b.launched = False
        self._next_launch = 0
        self._synced = False
        self._allreducer.poll_faults()           # pinned host flag, no sync: a timed-out peer wait surfaces at once

    def zero_grad(self, set_to_none: bool = False) -> None:  # noqa: ARG002 - views must survive
        if self._land:
            # detach the gradients from the bucket: the next backward hands every parameter a fresh tensor, which the
            # landing kernel copies in (the bucket itself is overwritten, never accumulated into: no memset either)
            for b in self._buckets:
                for p in b.params:
                    p.grad = None
                b.dirty = False
            return
        for b in self._buckets:
            if b.dirty:
                b.grad.zero_()
                b.dirty = False

    def stop(self) -> None:
        self._allreducer.stop()

    def add_train_epoch(self) -> None:
        self._allreducer.train_epoch += 1

    def get_current_density(self) -> float:
        return self._allreducer.get_current_density()

    def comm_stats(self) -> Dict:
        return self._allreducer.stats()

    def check_faults(self) -> None:
        """Raise / call ``err_handler`` if a peer timed out inside a communication kernel (synchronous read)."""
        self._allreducer.check_faults()

    def close(self) -> None:
        for h in self._hook_handles:
            h.remove()
        self._hook_handles = []
        self._allreducer.close()

    # ------------------------------------------------------------------ fused updates
    def _fused_sgd(self, b: Bucket) -> None:
        fs = self._flat_state.setdefault(b.index, {})
        first = "momentum" not in fs
        mom = fs.get("momentum")
        if mom is None:
            mom = fs["momentum"] = torch.zeros_like(b.grad)
            for p, v in zip(b.params, b.views(mom)):
                self.state[p]["momentum_buffer"] = v
        use_kernel = b.grad.is_cuda and b.flat_param is not None and ext.available()
        for gi, s, e in b.group_slices:
            g = self.param_groups[gi]
            lr, m, damp, wd, nest = g["lr"], g.get("momentum", 0.0), g.get("dampening", 0.0), \
                g.get("weight_decay", 0.0), bool(g.get("nesterov", False))
            if self.momentum_correction:
                m = 0.0                            # momentum already applied before communication
            if use_kernel:
                ext.require().fused_sgd(b.flat_param.data_ptr() + 4 * s, b.grad.data_ptr() + 4 * s,
                                        mom.data_ptr() + 4 * s, e - s, lr, m, damp, wd, int(nest), int(first),
                                        0 if self._land else 1, 1.0,
                                        torch.cuda.current_stream().cuda_stream, self._lr_ptr(gi),
                                        self._allreducer.fault_ptr(b.name))
            else:
                gs, ms = b.grad[s:e], mom[s:e]
                if b.flat_param is not None:
                    self._sgd_math(b.flat_param[s:e], gs, ms, lr, m, damp, wd, nest, first)
                else:
                    for p, o in zip(b.params, b.offsets):
                        if s <= o < e:
                            self._sgd_math(p.data.view(-1), gs[o - s:o - s + p.numel()], ms[o - s:o - s + p.numel()],
                                           lr, m, damp, wd, nest, first)
                gs.zero_()
        b.dirty = False

    @staticmethod
    def _sgd_math(p, g, mom, lr, m, damp, wd, nest, first) -> None:
        d = g.add(p, alpha=wd) if wd != 0 else g.clone()
        if m != 0:
            if first:
                mom.copy_(d)
            else:
                mom.mul_(m).add_(d, alpha=1 - damp)
            d = d.add(mom, alpha=m) if nest else mom
        p.add_(d, alpha=-lr)


# ====================================================================================== DistributedOptimizer
class _DistributedOptimizerMixin(_BucketedComm):
    def step(self, closure=None):
        """``synchronize()`` + parameter update (``VGG/distributed_optimizer.py:185-190``)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if not self.local:
            self.synchronize()
        if self._okt_is_sgd:
            self._maybe_refresh_lr()
            with torch.no_grad():
                for b in self._buckets:
                    self._fused_sgd(b)
        else:
            super().step()
            for b in self._buckets:
                b.dirty = True
        self._after_step()
        return loss

    def state_dict(self):
        sd = super().state_dict()
        sd["oktopk"] = self._allreducer.state_dict()
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        okt = state_dict.pop("oktopk", None)
        super().load_state_dict(state_dict)
        if self._okt_is_sgd:                      # re-alias momentum into the flat buffers
            for b in self._buckets:
                fs = self._flat_state.setdefault(b.index, {})
                have = [("momentum_buffer" in self.state.get(p, {})) and self.state[p]["momentum_buffer"] is not None
                        for p in b.params]
                if any(have):
                    mom = fs.get("momentum")
                    if mom is None:
                        mom = fs["momentum"] = torch.zeros_like(b.grad)
                    for p, v, h in zip(b.params, b.views(mom), have):
                        if h:
                            v.copy_(self.state[p]["momentum_buffer"])
                        self.state[p]["momentum_buffer"] = v
        if okt is not None:
            self._allreducer.load_state_dict(okt)


def DistributedOptimizer(optimizer: torch.optim.Optimizer, named_parameters=None, compression=NoneCompressor,
                         is_sparse: bool = False, err_handler=None, layerwise_times=None, sigma_scale: float = 2.5,
                         density: float = 0.1, norm_clip: Optional[float] = None, writer=None,
                         cfg: Optional[OkTopkConfig] = None, world: Optional[World] = None,
                         backend: Optional[str] = None, flatten_params: bool = True):
    """Wrap ``optimizer`` so that ``step()`` first allreduces the gradients with the chosen scheme.

    Horovod-style dynamic subclass of the user's optimizer class, as in the reference
    (``VGG/distributed_optimizer.py:203-207``).  ``compression`` may be a registry key, a compressor
    class (``compressors['oktopk']``) or an instance; ``cfg`` overrides the scalar arguments.
    """
    base_cls = optimizer.__class__
    cls = type(base_cls.__name__, (_DistributedOptimizerMixin, base_cls), {})
    obj = cls.__new__(cls)
    base_cls.__init__(obj, optimizer.param_groups)
    obj.state.update(optimizer.state)
    ar = AllReducer(compression=compression, sparse=is_sparse, density=density, cfg=cfg, world=world,
                    backend=backend, err_callback=err_handler, layerwise_times=layerwise_times,
                    sigma_scale=sigma_scale, norm_clip=norm_clip, writer=writer)
    obj._okt_is_sgd = isinstance(optimizer, torch.optim.SGD)
    obj._okt_setup(named_parameters, ar, flatten_params=flatten_params)
    return obj


def rank() -> int:
    return _world().rank


def size() -> int:
    return _world().size


def broadcast_parameters(model_or_params, root_rank: int = 0, world: Optional[World] = None) -> None:
    """One-time parameter sync (the reference pickles the whole ``state_dict`` through
    ``comm.bcast``, ``VGG/main_trainer.py:52-55``)."""
    w = world or _world()
    if w.size == 1:
        return
    if isinstance(model_or_params, torch.nn.Module):
        tensors = list(model_or_params.state_dict().values())
    else:
        tensors = [p.data if isinstance(p, torch.nn.Parameter) else p for p in model_or_params]
    for t in tensors:
        if torch.is_tensor(t):
            w.broadcast(t, root_rank)
    w.barrier()


# ====================================================================================== BertAdam
def warmup_cosine(x, warmup=0.002):
    if x < warmup:
        return x / warmup
    return 0.5 * (1.0 + math.cos(math.pi * x))


def warmup_constant(x, warmup=0.002):
    if x < warmup:
        return x / warmup
    return 1.0


def warmup_linear(x, warmup=0.002):
    if x < warmup:
        return x / warmup
    return max((x - 1.0) / (warmup - 1.0), 0.0)


def warmup_poly(x, warmup=0.002, degree=0.5):
    if x < warmup:
        return x / warmup
    return (1.0 - x) ** degree


SCHEDULES = {
    "warmup_cosine": warmup_cosine,
    "warmup_constant": warmup_constant,
    "warmup_linear": warmup_linear,
    "warmup_poly": warmup_poly,
}


class BertAdam(_BucketedComm, torch.optim.Optimizer):
    """BERT's Adam (no bias correction, decoupled weight decay, warm-up schedules) with the sparse
    allreducer embedded -- ``BERT/bert/transformers/optimization.py:68-227``.

    ``max_grad_norm``: the reference calls ``clip_grad_norm_(p, ...)`` on the *local* ``p.grad`` and
    then applies the *reduced* gradient, so its clipping never affects the update (A.4-6).  Here
    ``clip_reduced=True`` clips the reduced per-parameter gradient for real; the default reproduces the
    reference's effective behaviour (no clipping).
    """

    def __init__(self, params, lr=1e-3, warmup=-1, t_total=-1, schedule="warmup_linear", b1=0.9, b2=0.999, e=1e-6,
                 weight_decay=0.01, max_grad_norm=1.0, density=1.0, compressor="none", rank=-1, named_parameters=None,
                 cfg: Optional[OkTopkConfig] = None, world: Optional[World] = None, backend: Optional[str] = None,
                 clip_reduced: bool = False, flatten_params: bool = True, **_ignored):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if schedule not in SCHEDULES:
            raise ValueError("Invalid schedule parameter: {}".format(schedule))
        if not 0.0 <= warmup < 1.0 and not warmup == -1:
            raise ValueError("Invalid warmup: {} - should be in [0.0, 1.0[ or -1".format(warmup))
        if not 0.0 <= b1 < 1.0 or not 0.0 <= b2 < 1.0 or not e >= 0.0:
            raise ValueError("Invalid Adam hyper-parameters")
        defaults = dict(lr=lr, schedule=schedule, warmup=warmup, t_total=t_total, b1=b1, b2=b2, e=e,
                        weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        torch.optim.Optimizer.__init__(self, params, defaults)
        self.rank = rank
        self.counter = 0
        self.clip_reduced = clip_reduced
        base = cfg if cfg is not None else OkTopkConfig(
            warmup_iters=0, local_recompute_interval=128, global_recompute_interval=128, overselect_guard_loops=0,
            local_adapt_low=4 / 5, local_adapt_high=5 / 4, local_adapt_factor=1.025, global_adapt_low=4 / 5,
            global_adapt_high=5 / 4, global_adapt_inc=1.036, global_adapt_dec=1.025, balanced_allgather=True)
        ar = AllReducer(compression=compressor, sparse=(compressor != "none"), density=density,
                        cfg=base.replace(density=density), world=world, backend=backend)
        self._okt_setup(named_parameters, ar, flatten_params=flatten_params)

    def _group_lr(self, gi: int) -> float:
        return float(self._scheduled_lr(self.param_groups[gi], self.counter))

    def get_lr(self) -> List[float]:
        out = []
        for g in self.param_groups:
            out.append(self._scheduled_lr(g, self.counter))
        return out if self.counter > 0 else [0]

    @staticmethod
    def _scheduled_lr(g: dict, step: int) -> float:
        if g["t_total"] != -1:
            return g["lr"] * SCHEDULES[g["schedule"]](step / g["t_total"], g["warmup"])
        return g["lr"]

    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if not self.local:
            self.synchronize()
        self._maybe_refresh_lr()
        with torch.no_grad():
            for b in self._buckets:
                self._fused_adam(b)
        self.counter += 1
        self._after_step()
        return loss

    def _fused_adam(self, b: Bucket) -> None:
        fs = self._flat_state.setdefault(b.index, {})
        if "next_m" not in fs:
            fs["next_m"] = torch.zeros_like(b.grad)
            fs["next_v"] = torch.zeros_like(b.grad)
            for p, vm, vv in zip(b.params, b.views(fs["next_m"]), b.views(fs["next_v"])):
                self.state[p]["next_m"], self.state[p]["next_v"], self.state[p]["step"] = vm, vv, 0
        m, v = fs["next_m"], fs["next_v"]
        use_kernel = b.grad.is_cuda and b.flat_param is not None and ext.available()
        for gi, s, e in b.group_slices:
            g = self.param_groups[gi]
            lr = self._scheduled_lr(g, self.counter)
            if self.clip_reduced and g["max_grad_norm"] > 0:
                for p, o in zip(b.params, b.offsets):
                    if s <= o < e:
                        gv = b.grad[o:o + p.numel()]
                        nrm = float(gv.norm())
                        if nrm > g["max_grad_norm"]:
                            gv.mul_(g["max_grad_norm"] / (nrm + 1e-6))
            if use_kernel:
                ext.require().fused_bert_adam(b.flat_param.data_ptr() + 4 * s, b.grad.data_ptr() + 4 * s,
                                              m.data_ptr() + 4 * s, v.data_ptr() + 4 * s, e - s, lr, g["b1"], g["b2"],
                                              g["e"], g["weight_decay"], 0 if self._land else 1,
                                              torch.cuda.current_stream().cuda_stream,
                                              self._lr_ptr(gi), self._allreducer.fault_ptr(b.name))
            else:
                gs, ms, vs = b.grad[s:e], m[s:e], v[s:e]
                ms.mul_(g["b1"]).add_(gs, alpha=1 - g["b1"])
                vs.mul_(g["b2"]).addcmul_(gs, gs, value=1 - g["b2"])
                upd = ms / (vs.sqrt() + g["e"])
                if b.flat_param is not None:
                    ps = b.flat_param[s:e]
                    if g["weight_decay"] > 0.0:
                        upd += g["weight_decay"] * ps
                    ps.add_(upd, alpha=-lr)
                else:
                    for p, o in zip(b.params, b.offsets):
                        if s <= o < e:
                            u = upd[o - s:o - s + p.numel()].view_as(p)
                            if g["weight_decay"] > 0.0:
                                u = u + g["weight_decay"] * p.data
                            p.data.add_(u, alpha=-lr)
                gs.zero_()
        for p in b.params:
            self.state[p]["step"] += 1
        b.dirty = False

    def state_dict(self):
        sd = torch.optim.Optimizer.state_dict(self)
        sd["oktopk"] = self._allreducer.state_dict()
        sd["counter"] = self.counter
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        okt = state_dict.pop("oktopk", None)
        self.counter = state_dict.pop("counter", 0)
        torch.optim.Optimizer.load_state_dict(self, state_dict)
        for b in self._buckets:
            fs = self._flat_state.setdefault(b.index, {})
            if any("next_m" in self.state.get(p, {}) for p in b.params):
                if "next_m" not in fs:
                    fs["next_m"] = torch.zeros_like(b.grad)
                    fs["next_v"] = torch.zeros_like(b.grad)
                for p, vm, vv in zip(b.params, b.views(fs["next_m"]), b.views(fs["next_v"])):
                    stp = self.state.get(p, {})
                    if "next_m" in stp:
                        vm.copy_(stp["next_m"])
                        vv.copy_(stp["next_v"])
                    self.state[p]["next_m"], self.state[p]["next_v"] = vm, vv
                    self.state[p].setdefault("step", self.counter)
        if okt is not None:
            self._allreducer.load_state_dict(okt)
