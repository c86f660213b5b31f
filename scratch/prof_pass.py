"""One inference pass + one training step of cfg 3 between cudaProfilerStart/Stop (ncu --profile-from-start off)."""
import argparse, sys
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import torch
import literalkg_b200 as L
import literalkg_oracle as O
n, e, r = 1_000_000, 20_000_000, 64
cfg = O.OracleConfig(n_conv_layers=3, aggregation_type="bi-interaction", mess_dropout=0.1)
kg = L.synthetic.make_kg(n, e, r)
num, txt = L.synthetic.make_literals(n, device="cuda")
args = argparse.Namespace(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
torch.manual_seed(2022)
m = L.LiteralKG(args, n, r, None, num, txt).cuda()
h, t, rr = (torch.from_numpy(x).cuda() for x in (kg.h, kg.t, kg.r))
bh, bp, bn = (torch.randint(0, n, (681,), device="cuda") for _ in range(3))
def infer():
    m.eval()
    with torch.no_grad():
        m(h, t, rr, list(range(r)), device="cuda", mode="update_att")
        return m.gat_embeddings()
def train():
    m.train()
    for p in m.parameters():
        p.grad = None
    loss = m(bh, bp, bn, device="cuda", mode="fine_tuning")
    loss.backward()
    return loss
for _ in range(2):
    infer(); train()
torch.cuda.synchronize()
torch.cuda.profiler.start()
infer(); train()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
