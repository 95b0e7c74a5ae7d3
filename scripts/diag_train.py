#!/usr/bin/env python
"""Per-step training diagnostics: loss, selected counts, thresholds, overflow -- any model / scheme, eager or graphed.
    python scripts/diag_train.py --model vgg16 --steps 80 [--graph] [--compressor oktopk] [--every 5]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import oktopk_b200 as okt  # noqa: E402
from oktopk_b200.train.trainer import Trainer, preset_for  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--model", default="vgg16")
p.add_argument("--steps", type=int, default=60)
p.add_argument("--every", type=int, default=5)
p.add_argument("--graph", action="store_true")
p.add_argument("--compressor", default="oktopk")
p.add_argument("--density", type=float, default=0.001)
p.add_argument("--lr", type=float, default=None)
p.add_argument("--batch-size", type=int, default=None)
p.add_argument("--warmup-iters", type=int, default=0)
a = p.parse_args()
w = okt.init()
M = {"vgg16": ("vgg16", "cifar10", 16, 0.1), "lstman4": ("lstman4", "an4", 2, 0.001), "bert": ("bert_base", "wikipedia", 8, 2e-4)}
dnn, ds, bs, lr = M[a.model]
cfg = okt.preset(preset_for(dnn), density=a.density, warmup_iters=a.warmup_iters)
tr = Trainer(dnn=dnn, dataset=ds, batch_size=a.batch_size or bs, lr=a.lr or lr, compressor=a.compressor, density=a.density,
             compression=a.compressor != "none", cfg=cfg, world=w, cuda_graph=a.graph, t_total=100000, warmup=0.1)
import math  # noqa: E402
nan_at = None
for i in range(a.steps):
    tr.train_step()
    if nan_at is None and a.model == "lstman4":
        lv = tr.last_loss()
        if not math.isfinite(lv):
            nan_at = i
            print("!! non-finite loss %r first at step %d" % (lv, i), flush=True)
    if i % a.every == 0 or i == a.steps - 1:
        loss = tr.last_loss()
        st = tr.optimizer.comm_stats()
        pn = float(torch.sqrt(sum((p.detach().float() ** 2).sum() for p in tr.net.parameters())))
        s = " | ".join("%s lc=%s gc=%s thr=%.3g gthr=%.3g ovf=%s" % (v.get("mode"), v.get("local_count"), v.get("global_count"),
                                                                   v.get("local_thr", 0), v.get("global_thr", 0),
                                                                   v.get("overflow_send")) for v in st.values())
        if w.rank == 0:
            print("step %3d loss %.4f |param| %.3f  %s" % (i, loss, pn, s), flush=True)
tr.close()
okt.shutdown()
