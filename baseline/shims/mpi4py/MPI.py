import os
import pickle

import numpy as np
import torch
import torch.distributed as dist


class _Datatype:
    def __init__(self, name, np_dtype=None, size=1):
        self.name, self.np_dtype, self.size = name, np_dtype, size

    def Create_contiguous(self, count):
        return _Datatype("%s[%d]" % (self.name, count), None, self.size * count)

    def Commit(self):
        return self


FLOAT = _Datatype("FLOAT", np.float32, 4)
DOUBLE = _Datatype("DOUBLE", np.float64, 8)
INT = _Datatype("INT", np.int32, 4)
LONG = _Datatype("LONG", np.int64, 8)
BYTE = _Datatype("BYTE", np.uint8, 1)
SUM = "SUM"
ERRORS_RETURN = "ERRORS_RETURN"
ERRORS_ARE_FATAL = "ERRORS_ARE_FATAL"
_typedict = {"f": FLOAT, "d": DOUBLE, "i": INT, "l": LONG, "b": BYTE}


def _ensure_init():
    if dist.is_available() and not dist.is_initialized() and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("gloo")


def _buf(x):
    """MPI buffer spec ([array, TYPE] or array) -> numpy array (no copy)."""
    if isinstance(x, (list, tuple)):
        x = x[0]
    if torch.is_tensor(x):
        x = x.numpy()
    return x


def _t(a):
    """numpy array -> aliasing CPU tensor (uint32 viewed as int32, which torch can carry)."""
    a = np.ascontiguousarray(a) if not a.flags["C_CONTIGUOUS"] else a
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    elif a.dtype == np.uint64:
        a = a.view(np.int64)
    return torch.from_numpy(a.reshape(-1))


class Request:
    def __init__(self, work=None):
        self.work = work

    def Wait(self):
        if self.work is not None:
            self.work.wait()
            self.work = None

    wait = Wait

    @staticmethod
    def Waitall(reqs):
        for r in reqs:
            r.Wait()


class _Comm:
    def __init__(self):
        self._group = None

    # -- identity --------------------------------------------------------------
    @property
    def rank(self):
        _ensure_init()
        return dist.get_rank() if dist.is_initialized() else 0

    @property
    def size(self):
        _ensure_init()
        return dist.get_world_size() if dist.is_initialized() else 1

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def Set_errhandler(self, *_a):
        return None

    def _g(self):
        # host-buffer collectives always go over gloo, also when NCCL is the default group
        if self._group is None and dist.is_initialized():
            self._group = dist.group.WORLD if dist.get_backend() == "gloo" else dist.new_group(backend="gloo")
        return self._group

    # -- collectives ---------------------------------------------------------------
    def Barrier(self):
        if self.size > 1:
            dist.barrier(group=self._g())

    barrier = Barrier

    def Allreduce(self, sendbuf, recvbuf, op=SUM):
        s, r = _buf(sendbuf), _buf(recvbuf)
        if r is not s:
            np.copyto(r.reshape(-1), s.reshape(-1).astype(r.dtype, copy=False))
        if self.size > 1:
            dist.all_reduce(_t(r), op=dist.ReduceOp.SUM, group=self._g())

    def Alltoall(self, sendbuf, recvbuf):
        s, r = _buf(sendbuf), _buf(recvbuf)
        if self.size == 1:
            np.copyto(r.reshape(-1), s.reshape(-1))
            return
        P = self.size
        gathered = [torch.empty(s.size, dtype=_t(s).dtype) for _ in range(P)]
        dist.all_gather(gathered, _t(s).clone(), group=self._g())
        m = s.size // P
        rv = r.reshape(-1)
        for src in range(P):
            rv[src * m:(src + 1) * m] = gathered[src].numpy()[self.rank * m:(self.rank + 1) * m].view(rv.dtype)

    def Allgather(self, sendbuf, recvbuf):
        s, r = _buf(sendbuf), _buf(recvbuf)
        rv = r.reshape(-1)
        if self.size == 1:
            rv[:s.size] = s.reshape(-1)
            return
        P = self.size
        ts = _t(s).clone()
        gathered = [torch.empty_like(ts) for _ in range(P)]
        dist.all_gather(gathered, ts, group=self._g())
        m = s.size
        for src in range(P):
            rv[src * m:(src + 1) * m] = gathered[src].numpy().view(rv.dtype) if rv.dtype.itemsize == ts.element_size() \
                else gathered[src].numpy().astype(rv.dtype)

    def Allgatherv(self, sendbuf, recvspec):
        s = _buf(sendbuf)
        r, counts, offsets = recvspec[0], recvspec[1], recvspec[2]
        r = _buf(r)
        rv = r.reshape(-1)
        counts = [int(c) for c in counts]
        offsets = [int(o) for o in offsets]
        if self.size == 1:
            rv[offsets[0]:offsets[0] + counts[0]] = s.reshape(-1)[:counts[0]]
            return
        P = self.size
        mx = max(max(counts), 1)
        ts = _t(s)
        pad = torch.zeros(mx, dtype=ts.dtype)
        pad[:ts.numel()] = ts
        gathered = [torch.empty_like(pad) for _ in range(P)]
        dist.all_gather(gathered, pad, group=self._g())
        for src in range(P):
            c = counts[src]
            if c:
                rv[offsets[src]:offsets[src] + c] = gathered[src].numpy()[:c].view(rv.dtype) \
                    if rv.dtype.itemsize == pad.element_size() else gathered[src].numpy()[:c].astype(rv.dtype)

    def Bcast(self, buf, root=0):
        if self.size > 1:
            dist.broadcast(_t(_buf(buf)), src=root, group=self._g())

    def bcast(self, obj, root=0):
        if self.size == 1:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=root, group=self._g())
        return box[0]

    # -- point to point -----------------------------------------------------------------
    def Send(self, buf, dest=0, tag=0):
        a = _buf(buf)
        if a.size:
            dist.send(_t(a), dst=dest, group=self._g(), tag=tag)

    def Recv(self, buf, source=0, tag=0):
        a = _buf(buf)
        if a.size:
            dist.recv(_t(a), src=source, group=self._g(), tag=tag)

    def Isend(self, buf, dest=0, tag=0):
        a = _buf(buf)
        if a.size == 0:
            return Request()
        return Request(dist.isend(_t(a), dst=dest, group=self._g(), tag=tag))

    def Irecv(self, buf, source=0, tag=0):
        a = _buf(buf)
        if a.size == 0:
            return Request()
        return Request(dist.irecv(_t(a), src=source, group=self._g(), tag=tag))


COMM_WORLD = _Comm()
