"""Per-tensor gradient error of the CUDA path vs the fp64 oracle, next to the fp32 oracle's own error."""
import argparse, sys
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import torch
import literalkg_b200 as L
import literalkg_oracle as O

def run(agg, res, layers, scale):
    n, n_rel, e = 4000, 6, 40000
    cfg = O.OracleConfig(n_conv_layers=layers, aggregation_type=agg, use_residual=res, mess_dropout=0.0)
    kg = L.synthetic.make_kg(n, e, n_rel, seed=11, max_out_degree=300)
    num, txt = L.synthetic.make_literals(n, seed=11)
    p = O.init_params(cfg, n, n_rel, seed=11)
    p["entity_embed.weight"] *= scale
    h, t, r = (torch.from_numpy(x) for x in (kg.h, kg.t, kg.r))
    idx, val = O.update_attention(p["entity_embed.weight"], p["relation_embed.weight"], h, t, r, range(n_rel), n)
    gen = torch.Generator().manual_seed(5)
    bh, bp, bn = (torch.randint(0, n, (256,), generator=gen) for _ in range(3))
    br = torch.randint(0, n_rel, (256,), generator=gen)
    out = {}
    for dt in (torch.float64, torch.float32):
        pd = {k: v.to(dt).requires_grad_(v.is_floating_point()) for k, v in p.items()}
        emb = O.gat_embeddings(pd, cfg, idx, val.to(dt), num.to(dt), txt.to(dt))
        O.triplet_loss(pd, emb, cfg, bh, br, bp, bn).backward()
        out[dt] = {k: v.grad for k, v in pd.items() if v.grad is not None}
    args = argparse.Namespace(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    a = torch.sparse_coo_tensor(idx, val, (n, n))
    m = L.LiteralKG(args, n, n_rel, a, num, txt)
    m.load_state_dict(p, strict=False)
    m = m.cuda().train()
    m(bh.cuda(), br.cuda(), bp.cuda(), bn.cuda(), device="cuda", mode="pre_training").backward()
    ref = out[torch.float64]
    print(f"--- {agg} res={res} L={layers} scale={scale}")
    for k, prm in m.named_parameters():
        if k in ref:
            den = ref[k].abs().max().clamp_min(1e-30)
            e_ours = ((prm.grad.double().cpu() - ref[k]).abs().max() / den).item()
            e_f32 = ((out[torch.float32][k].double() - ref[k]).abs().max() / den).item()
            print(f"{k:50s} ours {e_ours:.2e}  fp32-oracle {e_f32:.2e}")

run("bi-interaction", True, 3, 20)
run("bi-interaction", True, 3, 1)
