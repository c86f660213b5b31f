#!/usr/bin/env python
"""Narrow-aggregation probe: the kernel timed back to back on its own, and behind a kernel that leaves the L2 full of
dirty lines (what it sees inside the pass).  python profiles/narrow_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import literalkg_b200 as L
from literalkg_b200 import ops

dev = torch.device("cuda:0")
n, e, c = 1_000_000, 20_000_000, 32
kg = L.synthetic.make_kg(n, e, 64)
plan = L.GraphPlan(*(torch.from_numpy(x).to(dev) for x in (kg.h, kg.t, kg.r)), n, 64)
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g, device=dev)
vals = torch.rand(plan.nnz, generator=g, device=dev)
ego, pb, p2 = rnd(n, c), rnd(c, c) * 0.2, rnd(c, c) * 0.2
r = rnd(n, 2 * c)
ln_w, ln_b = torch.ones(c, device=dev), torch.zeros(c, device=dev)
x, xn = torch.empty(n, c, device=dev), torch.empty(n, c, device=dev)
big = torch.empty(256 << 20, dtype=torch.float32, device=dev)          # 1 GB


def call():
    ops.aggregate(plan, vals, ego, c, pb, pb, p2, r[:, :c], r[:, c:], ln_w, ln_b, None, x, xn)


def timed(pre=None, reps=20):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        if pre is not None:
            pre()
        a.record(); call(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2], t[0], t[-1]


for _ in range(3):
    call()
torch.cuda.synchronize()
print("back to back          median / min / max ms: %.3f %.3f %.3f" % timed())
print("after a 1 GB fill     median / min / max ms: %.3f %.3f %.3f" % timed(lambda: big.fill_(1.0)))
print("after a 1 GB read     median / min / max ms: %.3f %.3f %.3f" % timed(lambda: big.sum()))
import time
def idle():
    torch.cuda.synchronize(); time.sleep(0.05)
print("after 50 ms idle      median / min / max ms: %.3f %.3f %.3f" % timed(idle))
