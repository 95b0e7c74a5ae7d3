"""Checkpoint-sweep evaluation (capability parity with ``VGG/evaluate.py:1-74``: load every
``<dnn>-rank0-epoch<E>.pth`` under a directory and report the validation metric per epoch).  The
reference's version is dead code because its trainer never writes checkpoints; ours does
(``robust_ssgd(checkpoint_dir=...)``).

    python -m oktopk_b200.train.evaluate --dnn vgg16 --dataset cifar10 --checkpoint-dir ./weights/run
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import re
import sys
from typing import Dict, List

import torch


def evaluate_checkpoints(dnn: str, dataset: str, checkpoint_dir: str, batch_size: int = 64, max_batches: int = 20,
                         data_dir=None, device=None) -> List[Dict]:
    from .trainer import Trainer
    tr = Trainer(dnn=dnn, dataset=dataset, data_dir=data_dir, batch_size=batch_size, compressor="none", compression=False,
                 device=device)
    out = []
    paths = sorted(glob.glob(os.path.join(checkpoint_dir, "*.pth")),
                   key=lambda p: int((re.findall(r"epoch(\d+)", p) or ["0"])[-1]))
    for path in paths:
        tr.load_checkpoint(path, model_only=True)
        epoch = int((re.findall(r"epoch(\d+)", path) or ["0"])[-1])
        res = tr.test(epoch, max_batches=max_batches)
        res["checkpoint"] = os.path.basename(path)
        out.append(res)
    tr.close()
    return out


def main(argv=None) -> int:
    p = argparse.ArgumentParser()
    p.add_argument("--dnn", default="vgg16")
    p.add_argument("--dataset", default=None)
    p.add_argument("--checkpoint-dir", required=True)
    p.add_argument("--data-dir", default=None)
    p.add_argument("--batch-size", type=int, default=64)
    p.add_argument("--max-batches", type=int, default=20)
    a = p.parse_args(argv)
    for r in evaluate_checkpoints(a.dnn, a.dataset, a.checkpoint_dir, a.batch_size, a.max_batches, a.data_dir):
        print(json.dumps(r))
    return 0


if __name__ == "__main__":
    sys.exit(main())
