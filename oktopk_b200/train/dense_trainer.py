"""Dense data-parallel baseline driver (capability parity with ``VGG/horovod_trainer.py:22-68``, the reference's
Horovod/NCCL dense-allreduce trainer that none of its launch scripts use).

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 -m oktopk_b200.train.dense_trainer --dnn vgg16 --nccl

Two dense paths: the default is this library's own two-shot peer-memory allreduce kernel (``csrc/dense.cu``);
``--nccl`` routes the gradient buckets through ``torch.distributed.all_reduce`` (``backend='dist'``) -- the honest
strong baseline the sparse schemes are compared against in ``bench/sweep.py``.
"""
from __future__ import annotations

import argparse
import sys


def dense_ssgd(dnn: str = "vgg16", dataset=None, data_dir=None, lr: float = 0.1, batch_size: int = 16, max_epochs: int = 1,
               max_iters=None, nccl: bool = False, **kw):
    import oktopk_b200 as okt
    from .trainer import robust_ssgd
    okt.init()
    return robust_ssgd(dnn, dataset, data_dir, okt.size(), lr, batch_size, 1, max_epochs, compression=False,
                       compressor="none", max_iters=max_iters, backend="dist" if nccl else None, **kw)


def main(argv=None) -> int:
    import oktopk_b200 as okt
    p = argparse.ArgumentParser()
    p.add_argument("--dnn", default="vgg16")
    p.add_argument("--dataset", default=None)
    p.add_argument("--data-dir", default=None)
    p.add_argument("--lr", type=float, default=0.1)
    p.add_argument("--batch-size", type=int, default=16)
    p.add_argument("--max-epochs", type=int, default=1)
    p.add_argument("--max-iters", type=int, default=None)
    p.add_argument("--nccl", action="store_true")
    a = p.parse_args(argv)
    tr = dense_ssgd(a.dnn, a.dataset, a.data_dir, a.lr, a.batch_size, a.max_epochs, a.max_iters, a.nccl)
    if okt.rank() == 0:
        print("final loss %.5f after %d iterations" % (tr.last_loss(), tr.train_iter))
    tr.close()
    okt.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
