"""cfg-3 training step (fine_tuning loss, forward + backward) with per-entry-point timing."""
import argparse, sys, json
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import torch
import literalkg_b200 as L
from literalkg_b200 import ops
import literalkg_oracle as O
n, e, r = 1_000_000, 20_000_000, 64
agg = sys.argv[1] if len(sys.argv) > 1 else "bi-interaction"
cfg = O.OracleConfig(n_conv_layers=3, aggregation_type=agg, mess_dropout=0.1)
kg = L.synthetic.make_kg(n, e, r)
num, txt = L.synthetic.make_literals(n, device="cuda")
args = argparse.Namespace(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
torch.manual_seed(2022)
m = L.LiteralKG(args, n, r, None, num, txt).cuda().train()
h, t, rr = (torch.from_numpy(x).cuda() for x in (kg.h, kg.t, kg.r))
with torch.no_grad():
    m(h, t, rr, list(range(r)), device="cuda", mode="update_att")
bh, bp, bn = (torch.randint(0, n, (681,), device="cuda") for _ in range(3))
opt = torch.optim.Adam(m.parameters(), lr=1e-4)
def step():
    opt.zero_grad(set_to_none=True)
    loss = m(bh, bp, bn, device="cuda", mode="fine_tuning")
    loss.backward()
    opt.step()
    return loss
for _ in range(2): step()
torch.cuda.synchronize()
ops.PROFILE = ops.Profile()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
K = 3
for _ in range(K): loss = step()
b.record(); torch.cuda.synchronize()
print("train step ms", a.elapsed_time(b) / K, "loss", loss.item(), "mem GB", torch.cuda.max_memory_allocated() / 1e9)
s = ops.PROFILE.summary(); ops.PROFILE = None
tot = 0
for k, v in sorted(s.items(), key=lambda kv: -kv[1]["ms_total"]):
    print(f"{k:28s} calls/step {v['calls']/K:5.1f}  ms/step {v['ms_total']/K:8.3f}")
    tot += v["ms_total"] / K
print("sum of entry points ms/step", tot)
