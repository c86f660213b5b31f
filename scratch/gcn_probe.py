import argparse, sys
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import torch
import literalkg_b200 as L
import literalkg_oracle as O
agg, res, layers = sys.argv[1], sys.argv[2] == "1", int(sys.argv[3])
n, n_rel, e = 4000, 6, 40000
cfg = O.OracleConfig(n_conv_layers=layers, aggregation_type=agg, use_residual=res, mess_dropout=0.0)
kg = L.synthetic.make_kg(n, e, n_rel, seed=11, max_out_degree=300)
num, txt = L.synthetic.make_literals(n, seed=11)
p = O.init_params(cfg, n, n_rel, seed=11)
h, t, r = (torch.from_numpy(x) for x in (kg.h, kg.t, kg.r))
idx, val = O.update_attention(p["entity_embed.weight"], p["relation_embed.weight"], h, t, r, range(n_rel), n)
pd = {k: v.double().requires_grad_(v.is_floating_point()) for k, v in p.items()}
emb = O.gat_embeddings(pd, cfg, idx, val.double(), num.double(), txt.double())
gen = torch.Generator().manual_seed(5)
w = torch.randn(emb.shape, generator=gen, dtype=torch.float64)
(emb * w).sum().backward()
args = argparse.Namespace(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
a = torch.sparse_coo_tensor(idx, val, (n, n))
m = L.LiteralKG(args, n, n_rel, a, num, txt)
m.load_state_dict(p, strict=False)
m = m.cuda().train()
with torch.no_grad():
    e0 = m.gat_embeddings()
e1 = m.gat_embeddings()
rel = lambda x, y: ((x.double().cpu() - y).abs().max() / y.abs().max()).item()
print("fwd inference", rel(e0, emb.detach()), "fwd train", rel(e1, emb.detach()), "equal", torch.equal(e0, e1.detach()))
(e1 * w.float().cuda()).sum().backward()
for k, prm in m.named_parameters():
    if k in pd and pd[k].grad is not None and prm.grad is not None:
        print(f"{k:48s} {rel(prm.grad, pd[k].grad):.2e}")
