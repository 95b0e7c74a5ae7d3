#!/usr/bin/env python
"""Smallest end-to-end use of the library (works on CPU/gloo and on B200s):

    python examples/minimal_train.py                                        # 1 process
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/minimal_train.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import oktopk_b200 as okt  # noqa: E402

w = okt.init()
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))) if torch.cuda.is_available() else torch.device("cpu")
torch.manual_seed(0)
model = torch.nn.Sequential(torch.nn.Linear(64, 256), torch.nn.ReLU(), torch.nn.Linear(256, 10)).to(dev)
okt.broadcast_parameters(model)
opt = okt.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9),
                               named_parameters=model.named_parameters(),
                               compression=okt.compressors["oktopk"], is_sparse=True, density=0.01)
g = torch.Generator().manual_seed(w.rank)
teacher = torch.randn(64, 10, generator=torch.Generator().manual_seed(123))
for step in range(200):
    x = torch.randn(32, 64, generator=g)
    y = (x @ teacher).argmax(1)
    opt.zero_grad()
    loss = torch.nn.functional.cross_entropy(model(x.to(dev)), y.to(dev))
    loss.backward()
    opt.step()                     # sparse allreduce of the gradients + fused SGD update
    if step % 50 == 0 and w.rank == 0:
        st = next(iter(opt.comm_stats().values()))
        print("step %3d loss %.3f  selected %s / kept %s" % (step, float(loss.detach()), st["local_count"], st["global_count"]))
opt.close()
okt.shutdown()
