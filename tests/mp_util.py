"""Spawn helper for multi-process tests (gloo on CPU, nccl on GPUs), rendezvous on 127.0.0.1."""
import os
import socket
import traceback

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _plain(x):
    """Tensors cross the result queue as numpy arrays (fd-sharing would race with child exit)."""
    if torch.is_tensor(x):
        return ("__tensor__", x.detach().cpu().numpy())
    if isinstance(x, (list, tuple)):
        return type(x)(_plain(v) for v in x)
    if isinstance(x, dict):
        return {k: _plain(v) for k, v in x.items()}
    return x


def _unplain(x):
    if isinstance(x, tuple) and len(x) == 2 and isinstance(x[0], str) and x[0] == "__tensor__":
        return torch.from_numpy(x[1])
    if isinstance(x, (list, tuple)):
        return type(x)(_unplain(v) for v in x)
    if isinstance(x, dict):
        return {k: _unplain(v) for k, v in x.items()}
    return x


def _entry(rank, world_size, port, backend, fn, args, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size),
                      LOCAL_RANK=str(rank))
    try:
        if backend == "nccl":
            torch.cuda.set_device(rank)
            dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=torch.device("cuda", rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world_size)
        out = fn(rank, world_size, *args)
        q.put((rank, "ok", _plain(out)))
    except Exception:  # noqa: BLE001
        q.put((rank, "err", traceback.format_exc()))
    finally:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass


def run_distributed(fn, world_size: int, args=(), backend: str = "gloo", timeout: float = 300.0):
    """Run ``fn(rank, world_size, *args)`` on ``world_size`` processes; returns results in rank order."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_entry, args=(r, world_size, port, backend, fn, args, q)) for r in range(world_size)]
    for p in procs:
        p.start()
    results = {}
    try:
        for _ in range(world_size):
            rank, status, payload = q.get(timeout=timeout)
            if status != "ok":
                raise RuntimeError("rank %d failed:\n%s" % (rank, payload))
            results[rank] = _unplain(payload)
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.terminate()
    return [results[r] for r in range(world_size)]
