"""Layer-1 aggregate probe on the cfg-3 graph: bi-interaction with the pre-projected sum term (the bench path)."""
import sys, torch
sys.path.insert(0, ".")
import literalkg_b200 as L
from literalkg_b200 import ops
n, e, r, d, c = 1_000_000, 20_000_000, 64, 300, 32
kg = L.synthetic.make_kg(n, e, r)
h, t, rr = (torch.from_numpy(x).cuda() for x in (kg.h, kg.t, kg.r))
plan = L.GraphPlan(h, t, rr, n, r)
g = torch.Generator(device="cuda").manual_seed(0)
ego = torch.randn(n, d, generator=g, device="cuda") * 0.05
vals = torch.rand(plan.nnz, generator=g, device="cuda")
p2 = torch.randn(d, c, generator=g, device="cuda") * 0.05
r12z = torch.randn(n, 3 * c, generator=g, device="cuda")
lw, lb = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
x_out = torch.empty(n, c, device="cuda"); xn = torch.empty(n, c, device="cuda")
def run():
    ops.aggregate(plan, vals, ego, c, None, None, p2, r12z[:, :c], r12z[:, c:2 * c], lw, lb, None, x_out, xn, z=r12z[:, 2 * c:])
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): run()
b.record(); torch.cuda.synchronize()
print("aggregate_d300 (kBiZ) ms", a.elapsed_time(b) / 5)
