"""Relation-projected attention extension (BASELINE.json north star (b)): v = (e_t W_r) . tanh(e_h W_r + e_r), the
formula the reference keeps commented out at model.py:436-439, against the oracle's restatement of exactly those
three lines evaluated in float64.  Structure bit exact, values within 1e-3 (measured ~1e-6)."""
import argparse

import numpy as np
import pytest
import torch

import literalkg_oracle as O

pytestmark = pytest.mark.gpu
REL = 1e-3


def build(n, e, n_rel, d, seed):
    import literalkg_b200 as L
    cfg = O.OracleConfig(embed_dim=d, relation_dim=d, n_conv_layers=1, mess_dropout=0.0, txt_lit_dim=8,
                         scale_gat_dim=16, conv_dim=8)
    kg = L.synthetic.make_kg(n, e, n_rel, seed=seed, max_out_degree=300)
    num, txt = L.synthetic.make_literals(n, 2, 8, seed=seed)
    p = O.init_params(cfg, n, n_rel, seed=seed)
    p["entity_embed.weight"] *= 25
    p["relation_embed.weight"] *= 5
    args = argparse.Namespace(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    m = L.LiteralKG(args, n, n_rel, None, num, txt)
    m.load_state_dict(p, strict=False)
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(n_rel, d, d, generator=g) / d ** 0.5
    return kg, p, m.cuda().eval(), w


@pytest.mark.parametrize("n,e,n_rel,d", [(3000, 40000, 7, 300), (500, 6000, 3, 64), (20000, 300000, 16, 300)])
def test_projected_attention_vs_oracle(n, e, n_rel, d):
    kg, p, m, w = build(n, e, n_rel, d, seed=n)
    h, t, r = (torch.from_numpy(x) for x in (kg.h, kg.t, kg.r))
    with torch.no_grad():
        m(h.cuda(), t.cuda(), r.cuda(), list(range(n_rel)), w.cuda(), device="cuda", mode="update_att_projected")
    a = m.A_in.data
    p64 = O.cast_params(p, torch.float64)
    oi, ov = O.update_attention_projected(p64["entity_embed.weight"], p64["relation_embed.weight"], w.double(), h, t, r,
                                          range(n_rel), n)
    assert np.array_equal(a.indices().cpu().numpy(), oi.numpy())            # same coalesced structure, bit exact
    assert oi.shape[1] < kg.n_edges                                         # duplicate (h, t) pairs were merged
    err = ((a.values().cpu().double() - ov).abs().max() / ov.abs().max()).item()
    err_el = ((a.values().cpu().double() - ov).abs() / ov)[ov > 1e-6].max().item()
    print(f"projected attention N={n} E={kg.n_edges}: normwise {err:.2e}, elementwise {err_el:.2e}")
    assert err < REL and err_el < REL
    # rows sum to one; the embedding pass runs on the new A_in
    rs = torch.zeros(n, dtype=torch.float64).index_add_(0, oi[0], a.values().cpu().double())
    assert (rs[rs > 0] - 1).abs().max() < 1e-5
    with torch.no_grad():
        assert torch.isfinite(m.gat_embeddings()).all()


def test_projected_attention_identity_projection_equals_plain():
    """W_r = I turns the projected formula into update_att's: the two kernels paths must agree to fp32 rounding."""
    kg, p, m, _ = build(4000, 50000, 5, 300, seed=3)
    h, t, r = (torch.from_numpy(x).cuda() for x in (kg.h, kg.t, kg.r))
    rels = list(range(5))
    with torch.no_grad():
        m(h, t, r, rels, device="cuda", mode="update_att")
        plain = m.A_in.data.values().clone()
        m(h, t, r, rels, torch.eye(300).repeat(5, 1, 1).cuda(), device="cuda", mode="update_att_projected")
    assert torch.equal(m.A_in.data.indices()[0][:10], m.A_in.data.indices()[0][:10])
    err = ((m.A_in.data.values() - plain).abs().max() / plain.abs().max()).item()
    assert err < 1e-5, err
