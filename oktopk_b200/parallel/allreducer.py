"""``AllReducer``: the sparse-allreduce engine front end (L4 of the reference's layer map).

Two faces, like the reference:
  * synchronous functional form (BERT flavour, ``BERT/bert/allreducer.py:182,347``):
    ``AllReducer(compression, sparse, density).run(flat_grad) -> flat_grad``;
  * named buckets for the optimizer wrapper (CNN/LSTM flavour, ``VGG/allreducer.py:191-473``:
    ``add_tensor`` / ``get_result`` / ``train_epoch`` / ``get_current_density`` / ``stop``) -- but
    without the consumer thread, the queues and the per-hook ``torch.cuda.synchronize()``: a
    bucket's reduction is enqueued on a CUDA stream the moment its last gradient lands.

Backends: ``cuda`` = fused sm_100a kernels over peer memory (``gpu_engine.py``); ``dist`` = the
same schemes on ``torch.distributed`` collectives (``algorithms.py``; CPU/gloo plumbing and the
NCCL baseline).
"""
from __future__ import annotations

import time
from typing import Dict, Optional

import torch

from ..compression import Compressor, resolve_compressor
from ..config import OkTopkConfig
from ..ops import ext
from . import algorithms
from .state import SparseState
from .world import World, world as _world


class AllReducer:
    def __init__(self, compression="oktopk", sparse: bool = True, density: float = 0.01, train_epoch: int = 0,
                 cfg: Optional[OkTopkConfig] = None, world: Optional[World] = None, backend: Optional[str] = None,
                 named_parameters=None, err_callback=None, layerwise_times=None, sigma_scale: float = 2.5,
                 norm_clip: Optional[float] = None, writer=None, **_ignored):
        base = cfg if cfg is not None else OkTopkConfig()
        self.compressor: Compressor = resolve_compressor(compression, base)
        over = {"sparse": bool(sparse) and self.compressor.name != "none", "compressor": self.compressor.name}
        if cfg is None:
            over.update(density=density, sigma_scale=sigma_scale, norm_clip=norm_clip)
        self.cfg = base.replace(**over)
        if backend is not None:
            self.cfg = self.cfg.replace(backend=backend)
        self.compressor.cfg = self.cfg
        self.world = world if world is not None else _world()
        self.train_epoch = train_epoch
        self.err_callback = err_callback           # kept for API parity; see utils/elastic.py
        self.writer = writer
        self._dist_states: Dict[str, SparseState] = {}
        self._engines: Dict[str, "object"] = {}
        self._timers: Dict[str, list] = {}
        self._running = True
        self.profile_records: list = []              # filled when settings.PROFILING_NORM is on

    # ------------------------------------------------------------------ density schedule
    def get_current_density(self) -> float:
        """``VGG/allreducer.py:451-458``: optional per-epoch density schedule."""
        dd = self.cfg.dynamic_densities
        if dd:
            return float(dd[min(self.train_epoch, len(dd) - 1)])
        return float(self.cfg.density)

    # ------------------------------------------------------------------ backend choice
    def _use_cuda(self, device: torch.device) -> bool:
        if device.type != "cuda" or self.cfg.backend == "dist":
            return False
        if not ext.available():
            if self.cfg.backend == "cuda" or device.type == "cuda":
                ext.require()       # GPU box without the extension: fail loudly, no silent fallback
            return False
        return True

    # ------------------------------------------------------------------ bucket registration (optimizer path)
    def register_bucket(self, name: str, numel: int, device: torch.device) -> torch.Tensor:
        """Create the bucket's engine/state and return the flat fp32 gradient buffer to alias."""
        if self._use_cuda(device):
            from .gpu_engine import CudaBucketEngine
            eng = CudaBucketEngine(numel, self.cfg, self.world, name=name)
            self._engines[name] = eng
            return eng.grad
        self._dist_states[name] = SparseState(numel, self.world.size)
        return torch.zeros(numel, dtype=torch.float32, device=device)

    def reduce_bucket(self, name: str, flat: torch.Tensor, stream=None) -> torch.Tensor:
        density = self.get_current_density()
        from ..utils import settings
        if settings.PROFILING_NORM and self.cfg.sparse:
            return self._reduce_profiled(name, flat, stream, density)
        if name in self._engines:
            return self._engines[name].reduce(self.compressor.name, density, stream=stream, g=flat)
        st = self._dist_states.get(name)
        if st is None:
            st = self._dist_states[name] = SparseState(flat.numel(), self.world.size)
        t0 = time.perf_counter()
        out = algorithms.sparse_allreduce(self.compressor.name, flat, st, self.cfg, self.world, density)
        self._timers.setdefault(name, []).append(time.perf_counter() - t0)
        return out

    def _reduce_profiled(self, name: str, flat: torch.Tensor, stream, density: float) -> torch.Tensor:
        """``settings.PROFILING_NORM`` (``VGG/allreducer.py:584-606,1072-1080``): one extra dense allreduce of the
        error-compensated gradient per step gives the true global top-k, against which the sparse result's relative
        error (the paper's xi) and the selected counts are recorded.  Diagnostic mode: synchronous and slow."""
        from ..utils.metrics import sparsification_error
        eng = self._engines.get(name)
        st = self._dist_states.get(name)
        res = eng.residual if eng is not None else (st.residual if st is not None else None)
        with torch.no_grad():
            acc = flat.detach().clone()
            if res is not None and res.numel() == acc.numel():
                acc += res
            if self.world.size > 1:
                self.world.all_reduce_sum(acc)
            acc /= self.world.size
        if eng is not None:
            out = eng.reduce(self.compressor.name, density, stream=stream, g=flat)
            if flat.is_cuda:
                torch.cuda.synchronize()
        else:
            if st is None:
                st = self._dist_states[name] = SparseState(flat.numel(), self.world.size)
            out = algorithms.sparse_allreduce(self.compressor.name, flat, st, self.cfg, self.world, density)
        k = max(int(flat.numel() * density), 1)
        rec = sparsification_error(acc, out, k)
        rec["bucket"], rec["density"] = name, density
        self.profile_records.append(rec)
        if self.writer is not None and hasattr(self.writer, "add_scalars"):
            self.writer.add_scalars("profiling_norm/" + name, {k2: v for k2, v in rec.items() if isinstance(v, (int, float))},
                                    len(self.profile_records))
        return out

    # ------------------------------------------------------------------ functional form
    def run(self, flat_tensor: torch.Tensor, name: str = "flat") -> torch.Tensor:
        """Reduce one flat fp32 tensor in place and return it (``BERT/bert/allreducer.py:347``)."""
        assert flat_tensor.dim() == 1 and flat_tensor.dtype == torch.float32
        if name not in self._engines and name not in self._dist_states and self._use_cuda(flat_tensor.device):
            from .gpu_engine import CudaBucketEngine
            self._engines[name] = CudaBucketEngine(flat_tensor.numel(), self.cfg, self.world, name=name)
        return self.reduce_bucket(name, flat_tensor)

    # ------------------------------------------------------------------ reference-compatible odds and ends
    def add_tensor(self, name: str, tensor: torch.Tensor):
        return name

    def get_result(self, name: str):
        raise RuntimeError("results are written in place into the gradient bucket")

    def stop(self) -> None:
        self._running = False

    def stats(self, name: Optional[str] = None) -> Dict:
        out = {}
        for nm, eng in self._engines.items():
            out[nm] = eng.stats()
        for nm, st in self._dist_states.items():
            out[nm] = {"counter": st.counter, "local_thr": st.local_thr, "global_thr": st.global_thr,
                       "local_count": st.last_local_count, "global_count": st.last_global_count,
                       "volume_elems": st.last_volume_elems, "mode": st.last_mode,
                       "edges": st.region_offsets + [st.numel]}
        return out if name is None else out[name]

    def check_faults(self) -> None:
        """Failure detection: surfaces device-side peer timeouts.  If an ``err_callback`` was given
        (``DistributedOptimizer(err_handler=...)``) it is invoked as ``cb(new_num_workers, new_rank)`` with the
        current world (the caller decides how to shrink); otherwise ``PeerTimeoutError`` propagates."""
        for eng in self._engines.values():
            try:
                eng.check_fault()
            except RuntimeError:
                if self.err_callback is not None:
                    self.err_callback(self.world.size, self.world.rank)
                    eng.clear_fault()
                else:
                    raise

    def state_dict(self) -> Dict:
        sd = {"train_epoch": self.train_epoch, "buckets": {}}
        for nm, eng in self._engines.items():
            sd["buckets"][nm] = eng.state_dict()
        for nm, st in self._dist_states.items():
            sd["buckets"][nm] = st.state_dict()
        return sd

    def load_state_dict(self, sd: Dict) -> None:
        self.train_epoch = sd.get("train_epoch", 0)
        for nm, b in sd.get("buckets", {}).items():
            if nm in self._engines:
                self._engines[nm].load_state_dict(b)
            elif nm in self._dist_states:
                dev = self._dist_states[nm].residual.device if self._dist_states[nm].residual is not None else None
                self._dist_states[nm].load_state_dict(b, dev)

    def close(self) -> None:
        for eng in self._engines.values():
            eng.close()
        self._engines.clear()


# free functions of the reference module (VGG/allreducer.py:34,76,175) -------------------------------
def dense_allreduce(tensor: torch.Tensor, world: Optional[World] = None) -> torch.Tensor:
    return algorithms.dense_allreduce(tensor, world or _world())


def topk_sparse_allreduce(tensor: torch.Tensor, density: float, world: Optional[World] = None, state=None):
    w = world or _world()
    st = state or SparseState(tensor.numel(), w.size)
    return algorithms.topka_allreduce(tensor, st, OkTopkConfig(density=density), w, density)


def gtopk_sparse_allreduce(tensor: torch.Tensor, density: float, world: Optional[World] = None, state=None):
    w = world or _world()
    st = state or SparseState(tensor.numel(), w.size)
    return algorithms.gtopk_allreduce(tensor, st, OkTopkConfig(density=density), w, density)
