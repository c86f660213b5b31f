"""Attention-update-only probe on the cfg-3 graph (for ncu)."""
import sys, torch
sys.path.insert(0, ".")
import literalkg_b200 as L
from literalkg_b200 import ops
n, e, r, d = 1_000_000, 20_000_000, 64, 300
kg = L.synthetic.make_kg(n, e, r)
h, t, rr = (torch.from_numpy(x).cuda() for x in (kg.h, kg.t, kg.r))
plan = L.GraphPlan(h, t, rr, n, r)
g = torch.Generator(device="cuda").manual_seed(0)
ent = torch.randn(n, d, generator=g, device="cuda") * 0.05
rel = torch.randn(r, d, generator=g, device="cuda") * 0.05
for _ in range(3):
    v = ops.attn_update(plan, ent, rel)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    ops.attn_update(plan, ent, rel)
b.record(); torch.cuda.synchronize()
print("attn_update ms", a.elapsed_time(b) / 5)
