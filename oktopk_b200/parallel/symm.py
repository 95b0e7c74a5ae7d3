"""Symmetric peer-mapped memory: the communication substrate of the fused kernels.

Every rank allocates the same-sized block with the native allocator (``cudaMalloc`` + CUDA IPC
handle, ``csrc/bindings.cpp``), the handles are exchanged once over the bootstrap process group
(``all_gather_object``), and every rank maps every peer's block.  After that the kernels address
peer memory directly over NVLink 5 / NVSwitch; nothing here is on the per-step path.

This replaces the reference's ``MPI.COMM_WORLD`` on host NumPy buffers (``VGG/allreducer.py:219-220``
and every ``.cpu().numpy()`` staging copy listed in SURVEY 2.4 table B).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from ..ops import ext
from .world import World


class SymmBlock:
    """``nbytes`` of zero-initialised device memory on every rank, all of them mapped locally.

    ``ptrs[r]`` is rank r's block in THIS process' address space (``ptrs[rank]`` is the local one).
    """

    def __init__(self, nbytes: int, world: World):
        C = ext.require()
        self.C = C
        self.world = world
        self.nbytes = int(nbytes)
        self.local_ptr, handle = C.symm_alloc(self.nbytes)
        self.ptrs: List[int] = [0] * world.size
        self.ptrs[world.rank] = self.local_ptr
        self._opened: List[int] = []
        if world.size > 1:
            dev = torch.cuda.current_device()
            infos = world.all_gather_object((world.rank, dev, handle))
            for r, rdev, h in infos:
                if r == world.rank:
                    continue
                p = C.symm_open(h)
                self.ptrs[r] = p
                self._opened.append(p)
            world.barrier()

    def tensor(self, offset_bytes: int, numel: int, dtype: str = "float32", rank: Optional[int] = None) -> torch.Tensor:
        r = self.world.rank if rank is None else rank
        return ext.tensor_from_ptr(self.ptrs[r] + offset_bytes, numel, dtype)

    def close(self) -> None:
        if self.local_ptr == 0:
            return
        torch.cuda.synchronize()
        if self.world.size > 1:
            try:
                self.world.barrier()
            except Exception:  # noqa: BLE001
                pass
        for p in self._opened:
            try:
                self.C.symm_close(p)
            except Exception:  # noqa: BLE001
                pass
        self._opened = []
        try:
            self.C.symm_free(self.local_ptr)
        except Exception:  # noqa: BLE001
            pass
        self.local_ptr = 0

    def __del__(self):
        # teardown order at interpreter exit is undefined; leaking device memory at exit is harmless
        pass
