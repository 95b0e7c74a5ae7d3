"""Process-group plumbing: one process per GPU, ``torch.distributed`` for bootstrap and for the
library-collective ("dist") backend.

Replaces the reference's ``mpi4py`` ``MPI.COMM_WORLD`` on host NumPy buffers
(``VGG/allreducer.py:10,219-220``; call sites in SURVEY 2.4 table B).  The fused CUDA engine
does not use these collectives on its hot path (it talks through peer-mapped memory,
``symm.py``); they serve the CPU/gloo plumbing configuration, the NCCL baseline, the one-time
parameter broadcast and handle exchange.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class World:
    """A (possibly size-1, possibly un-initialised) communicator over a process group."""

    def __init__(self, group: Optional["dist.ProcessGroup"] = None):
        self.group = group
        self._dist = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self._dist else 0
        self.size = dist.get_world_size(group) if self._dist else 1
        self.backend = dist.get_backend(group) if self._dist else "none"
        # staggered peer schedule (VGG/allreducer.py:246-251)
        self.dsts = [(self.rank + s) % self.size for s in range(self.size)]
        self.srcs = [(self.rank - s) % self.size for s in range(self.size)]
        self.bytes_sent = 0
        self.bytes_recv = 0

    # -- helpers -----------------------------------------------------------
    def _global(self, r: int) -> int:
        return dist.get_global_rank(self.group, r) if (self._dist and self.group is not None) else r

    def _dev(self, like: torch.Tensor) -> torch.device:
        return like.device

    # -- collectives ---------------------------------------------------------
    def barrier(self) -> None:
        if self.size > 1:
            dist.barrier(group=self.group)

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        if self.size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            b = t.numel() * t.element_size()
            self.bytes_sent += b
            self.bytes_recv += b
        return t

    def broadcast(self, t: torch.Tensor, src: int = 0) -> torch.Tensor:
        if self.size > 1:
            dist.broadcast(t, src=self._global(src), group=self.group)
        return t

    def broadcast_object(self, obj, src: int = 0):
        if self.size > 1:
            box = [obj]
            dist.broadcast_object_list(box, src=self._global(src), group=self.group)
            return box[0]
        return obj

    def all_gather_object(self, obj) -> list:
        if self.size == 1:
            return [obj]
        out = [None] * self.size
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def all_gather_fixed(self, t: torch.Tensor) -> torch.Tensor:
        """Equal-size allgather -> [P, *t.shape]."""
        if self.size == 1:
            return t.unsqueeze(0).clone()
        out = torch.empty((self.size,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1), group=self.group)
        b = t.numel() * t.element_size()
        self.bytes_sent += b * (self.size - 1)
        self.bytes_recv += b * (self.size - 1)
        return out

    def all_gather_counts(self, count: int, device) -> List[int]:
        """B7: one int per rank."""
        if self.size == 1:
            return [int(count)]
        t = torch.tensor([int(count)], dtype=torch.int64, device=device)
        return [int(x) for x in self.all_gather_fixed(t).view(-1).tolist()]

    def all_to_all_counts(self, send_counts: Sequence[int], device) -> List[int]:
        """B5: per-destination count handshake (implemented as allgather + column pick so that it
        also works on backends without alltoall)."""
        if self.size == 1:
            return [int(send_counts[0])]
        t = torch.tensor([int(c) for c in send_counts], dtype=torch.int64, device=device)
        m = self.all_gather_fixed(t)          # [src, dst]
        return [int(x) for x in m[:, self.rank].tolist()]

    def all_gatherv(self, tensors: Sequence[torch.Tensor], counts: Optional[List[int]] = None
                    ) -> Tuple[List[torch.Tensor], List[int]]:
        """B8: variable-size allgather of several equally-long 1-D tensors (e.g. int32 idx and fp32
        val).  Returns (concatenated tensors in rank order, per-rank counts)."""
        n = tensors[0].numel()
        dev = tensors[0].device
        if counts is None:
            counts = self.all_gather_counts(n, dev)
        if self.size == 1:
            return [t.clone() for t in tensors], counts
        mx = max(max(counts), 1)
        outs = []
        for t in tensors:
            pad = torch.zeros(mx, dtype=t.dtype, device=dev)
            pad[:n] = t
            g = self.all_gather_fixed(pad)
            outs.append(torch.cat([g[r, :counts[r]] for r in range(self.size)]))
        return outs, counts

    def exchange_pairwise(self, send: Sequence[Sequence[torch.Tensor]], recv_counts: Sequence[int],
                          throttle: int = 4, on_chunk=None) -> List[Optional[List[torch.Tensor]]]:
        """B6: the throttled, staggered pairwise exchange (``VGG/allreducer.py:731-794``).

        ``send[d]`` is the list of tensors for destination d (idx, val); ``recv_counts[s]`` how
        many elements source s will send.  At most ``throttle`` peers are in flight; ``on_chunk``
        (if given) is called with ``[(src, tensors), ...]`` for chunk c-1 while chunk c is in
        flight, which is where the caller overlaps its scatter-add.
        """
        P = self.size
        out: List[Optional[List[torch.Tensor]]] = [None] * P
        out[self.rank] = [t for t in send[self.rank]]
        if P == 1:
            if on_chunk is not None:
                on_chunk([(self.rank, out[self.rank])])
            return out
        throttle = max(1, min(throttle, P))
        steps = list(range(1, P))
        pending = [(self.rank, out[self.rank])]
        for c0 in range(0, len(steps), throttle):
            ops, landed = [], []
            for s in steps[c0:c0 + throttle]:
                dst, src = (self.rank + s) % P, (self.rank - s) % P
                bufs = []
                for t in send[dst]:
                    if t.numel() == 0:      # both sides know the count: skip empty messages
                        continue
                    ops.append(dist.P2POp(dist.isend, t.contiguous(), self._global(dst), group=self.group))
                    self.bytes_sent += t.numel() * t.element_size()
                for t in send[self.rank]:
                    b = torch.empty(int(recv_counts[src]), dtype=t.dtype, device=t.device)
                    if b.numel() == 0:
                        bufs.append(b)
                        continue
                    ops.append(dist.P2POp(dist.irecv, b, self._global(src), group=self.group))
                    self.bytes_recv += b.numel() * b.element_size()
                    bufs.append(b)
                landed.append((src, bufs))
            reqs = dist.batch_isend_irecv(ops) if ops else []
            if on_chunk is not None and pending:
                on_chunk(pending)          # overlap: reduce the previous chunk while this one flies
            for q in reqs:
                q.wait()
            for src, bufs in landed:
                out[src] = bufs
            pending = landed
        if on_chunk is not None and pending:
            on_chunk(pending)
        return out

    def send(self, t: torch.Tensor, dst: int) -> None:
        dist.send(t.contiguous(), dst=self._global(dst), group=self.group)
        self.bytes_sent += t.numel() * t.element_size()

    def recv(self, t: torch.Tensor, src: int) -> torch.Tensor:
        dist.recv(t, src=self._global(src), group=self.group)
        self.bytes_recv += t.numel() * t.element_size()
        return t


_WORLD: Optional[World] = None


def init(backend: Optional[str] = None, device: Optional[torch.device] = None) -> World:
    """Initialise ``torch.distributed`` from the torchrun environment (RANK / WORLD_SIZE /
    MASTER_ADDR / MASTER_PORT / LOCAL_RANK) and return the world.  A single process (no RANK in the
    environment) gets a size-1 world without touching ``torch.distributed``.

    Replaces ``init_distrib_slurm`` (``BERT/bert/main_bert.py:159-203``) and the MPI bootstrap.
    """
    global _WORLD
    if dist.is_available() and not dist.is_initialized() and "RANK" in os.environ \
            and int(os.environ.get("WORLD_SIZE", "1")) >= 1:
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))
            torch.cuda.set_device(local % max(torch.cuda.device_count(), 1))
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        else:
            dist.init_process_group(backend=backend)
    _WORLD = World()
    return _WORLD


def world() -> World:
    global _WORLD
    if _WORLD is None or (dist.is_available() and dist.is_initialized() and not _WORLD._dist):
        _WORLD = World()
    return _WORLD


def rank() -> int:
    return world().rank


def size() -> int:
    return world().size


def shutdown() -> None:
    global _WORLD
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
    _WORLD = None
