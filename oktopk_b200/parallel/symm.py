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


# ======================================================================================================
# VMM + NVSwitch-multicast flavour (NVLS): same interface plus ``mc_ptr``
# ======================================================================================================
def _exchange_fds(world: World, fd: int, tag: str) -> List[int]:
    """All-to-all exchange of one POSIX file descriptor per rank over abstract-namespace unix datagram sockets
    (``SCM_RIGHTS``).  Returns the list of P descriptors valid in THIS process (``[rank]`` is ``fd`` itself)."""
    import os
    import socket
    import struct
    uid = world.broadcast_object(os.urandom(6).hex() if world.rank == 0 else None, 0)

    def name(r: int) -> str:
        return "\0okt-%s-%s-%d" % (uid, tag, r)

    sock = socket.socket(socket.AF_UNIX, socket.SOCK_DGRAM)
    sock.bind(name(world.rank))
    sock.settimeout(60.0)
    try:
        world.barrier()                                  # every socket is bound
        import array
        for r in range(world.size):
            if r != world.rank:
                # (socket.send_fds() drops its address argument on CPython <= 3.12: build the SCM_RIGHTS message by hand)
                sock.sendmsg([struct.pack("i", world.rank)],
                             [(socket.SOL_SOCKET, socket.SCM_RIGHTS, array.array("i", [fd]))], 0, name(r))
        got = {world.rank: fd}
        while len(got) < world.size:
            msg, fds, _flags, _addr = socket.recv_fds(sock, 16, 1)
            got[struct.unpack("i", msg[:4])[0]] = fds[0]
        world.barrier()
    finally:
        sock.close()
    return [got[r] for r in range(world.size)]


class SymmBlockVMM:
    """Symmetric block allocated with the CUDA virtual-memory API so that, besides the P unicast mappings, the P
    physical copies are bound to ONE NVSwitch multicast object mapped at ``mc_ptr``: a ``multimem.ld_reduce`` on
    ``mc_ptr + off`` returns the switch-side sum of all ranks' words at ``off``, a ``multimem.st`` writes all of
    them (``csrc/dense.cu``).  Descriptor passing and the driver calls are plumbing (``csrc/bindings.cpp``)."""

    def __init__(self, nbytes: int, world: World):
        import os
        C = ext.require()
        self.C, self.world = C, world
        dev = torch.cuda.current_device()
        probe = C.vmm_probe(dev, world.size)
        gran = max(int(probe.get("granularity", 0)), int(probe.get("mc_granularity", 0)), 2 << 20)
        self.gran = gran
        self.nbytes = (int(nbytes) + gran - 1) // gran * gran
        self._handles: List[int] = []
        self._maps: List[int] = []
        h, fd = C.vmm_create(self.nbytes, dev)
        self._handles.append(h)
        self.local_ptr = C.vmm_map(h, self.nbytes, dev, gran)
        self._maps.append(self.local_ptr)
        ext.tensor_from_ptr(self.local_ptr, self.nbytes, "uint8").zero_()
        torch.cuda.synchronize()
        self.ptrs: List[int] = [0] * world.size
        self.ptrs[world.rank] = self.local_ptr
        fds = _exchange_fds(world, fd, "mem")
        for r in range(world.size):
            if r == world.rank:
                continue
            ph = C.vmm_import(fds[r])
            self._handles.append(ph)
            va = C.vmm_map(ph, self.nbytes, dev, gran)
            self._maps.append(va)
            self.ptrs[r] = va
            os.close(fds[r])
        # ---- multicast object: created by rank 0, every device joins, every rank binds its physical memory ----
        if world.rank == 0:
            mc, mfd = C.mc_create(self.nbytes, world.size)
        else:
            mc, mfd = 0, os.dup(fd)                       # placeholder descriptor for the symmetric exchange
        mfds = _exchange_fds(world, mfd, "mc")
        if world.rank != 0:
            mc = C.vmm_import(mfds[0])
        for r in range(world.size):
            if r != world.rank:
                os.close(mfds[r])
        os.close(mfd)
        os.close(fd)
        self._handles.append(mc)
        C.mc_add_device(mc, dev)
        world.barrier()                                   # all devices added before anybody binds
        C.mc_bind(mc, h, self.nbytes)
        world.barrier()
        self.mc_ptr = C.vmm_map(mc, self.nbytes, dev, gran)
        self._maps.append(self.mc_ptr)
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
        for va in self._maps:
            try:
                self.C.vmm_unmap(va, self.nbytes)
            except Exception:  # noqa: BLE001
                pass
        for h in self._handles:
            try:
                self.C.vmm_release(h)
            except Exception:  # noqa: BLE001
                pass
        self._maps, self._handles = [], []
        self.local_ptr = 0


_NVLS_STATE = {"decided": None, "why": ""}
NVLS_MIN_WORLD = 4


def nvls_available(world: World) -> bool:
    """Collective, cached: True iff every rank's device supports VMM + POSIX-fd export + multicast."""
    if _NVLS_STATE["decided"] is not None:
        return _NVLS_STATE["decided"]
    ok, why = False, ""
    if world.size > 1 and torch.cuda.is_available():
        try:
            pr = ext.require().vmm_probe(torch.cuda.current_device(), world.size)
            mine = bool(pr.get("vmm")) and bool(pr.get("posix_fd")) and bool(pr.get("multicast"))
            why = str(dict(pr))
        except Exception as e:  # noqa: BLE001
            mine, why = False, repr(e)
        allr = world.all_gather_object(mine)
        ok = all(allr)
    _NVLS_STATE["decided"], _NVLS_STATE["why"] = ok, why
    return ok


def make_symm_block(nbytes: int, world: World, nvls: str = "auto"):
    """The bucket's symmetric block: VMM + multicast when the box supports NVLS (and ``nvls`` is not 'off'), CUDA IPC
    otherwise.  The decision is collective; a failure while building the multicast flavour falls back to IPC on every
    rank together ('on' raises instead)."""
    import os
    mode = os.environ.get("OKTOPK_NVLS", nvls)
    if mode == "auto" and world.size < NVLS_MIN_WORLD:
        mode = "off"          # measured (profiles/bench): at P=2 plain peer loads beat the switch-side reduction
    if mode != "off" and world.size > 1 and nvls_available(world):
        blk, err = None, None
        try:
            blk = SymmBlockVMM(nbytes, world)
        except Exception as e:  # noqa: BLE001
            err = e
        oks = world.all_gather_object(err is None)
        if all(oks):
            return blk
        if blk is not None:
            try:
                blk.close()
            except Exception:  # noqa: BLE001
                pass
        _NVLS_STATE["decided"], _NVLS_STATE["why"] = False, "multicast setup failed: %r" % (err,)
        if mode == "on":
            raise RuntimeError("nvls='on' but the multicast block could not be built: %r" % (err,))
    elif mode == "on" and world.size > 1:
        raise RuntimeError("nvls='on' but the devices do not support VMM/multicast: %s" % _NVLS_STATE["why"])
    return SymmBlock(nbytes, world)
