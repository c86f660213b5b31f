"""Host-side logic of the drop-in (no GPU): the parameter folds the kernels consume, the stacked
h0 @ Q weight, the interleaved gate weight and the segmented row schedule -- each against the reference formulas as
restated by the oracle, in float64."""
import argparse
import math

import pytest
import torch
import torch.nn.functional as F

import literalkg_oracle as O


def _layer(agg, res, d_in=20, d_out=8, embed=12, seed=0):
    from literalkg_b200.model import Aggregator
    torch.manual_seed(seed)
    args = argparse.Namespace(embed_dim=embed)
    m = Aggregator(d_in, d_out, 0.0, agg, res, args).double()
    for p in m.parameters():                               # away from the special values of the initialisers
        torch.nn.init.normal_(p, std=0.3)
    return m


def _oracle_layer_pre_activation(m, agg, res, ego, side, h0, lamda, alpha, l):
    """What the reference feeds its LeakyReLUs (model.py:108-130 with residual_connection :90-99)."""
    p = dict(m.named_parameters())
    cfg = O.OracleConfig(use_residual=res, alpha=alpha, lamda=lamda)
    lin = lambda name, x: F.linear(x, p[name + ".weight"], p[name + ".bias"])
    r = lambda hi: O._residual(p, "", cfg, hi, h0, l)
    if agg == "gcn":
        return lin("linear", r(ego + side)), None
    if agg == "graphsage":
        hi = torch.cat([ego, side], dim=1)
        if res:
            hi = r(lin("linear_h", hi))
        return lin("linear", hi), None
    return lin("linear1", r(ego + side)), lin("linear2", r(ego * side))


@pytest.mark.parametrize("agg", ["gcn", "graphsage", "bi-interaction"])
@pytest.mark.parametrize("res", [False, True])
def test_folded_parameters_reproduce_the_reference_layer(agg, res):
    lamda, alpha, l = 0.5, 0.1, 3
    m = _layer(agg, res)
    g = torch.Generator().manual_seed(1)
    ego, side = torch.randn(7, 20, generator=g, dtype=torch.float64), torch.randn(7, 20, generator=g, dtype=torch.float64)
    h0 = torch.randn(7, 12, generator=g, dtype=torch.float64)
    ref1, ref2 = _oracle_layer_pre_activation(m, agg, res, ego, side, h0, lamda, alpha, l)
    with torch.no_grad():
        raw = m.folded(lamda, alpha, l)
        f = {k: (None if v is None else v.double()) for k, v in raw.items()}
    # fp32 fold outputs of fp64-formed products: compare at fp32 resolution
    o1 = ego @ f["pa"] + side @ f["pb"] + (h0 @ f["q1"] if f["q1"] is not None else 0) + f["c1"]
    assert torch.allclose(o1, ref1, rtol=1e-5, atol=1e-6)
    if agg == "bi-interaction":
        o2 = (ego * side) @ f["p2"] + (h0 @ f["q2"] if f["q2"] is not None else 0) + f["c2"]
        assert torch.allclose(o2, ref2, rtol=1e-5, atol=1e-6)
        assert raw["pa"] is raw["pb"]                      # the "sum" mode the kernels dispatch on
    else:
        assert f["p2"] is None and f["q2"] is None
    assert (f["q1"] is not None) == res
    # cached per parameter version, rebuilt after an in-place update
    with torch.no_grad():
        again = m.folded(lamda, alpha, l)
        assert again is m.folded(lamda, alpha, l)
        next(m.parameters()).add_(1.0)
        assert m.folded(lamda, alpha, l) is not again


@pytest.mark.parametrize("agg", ["gcn", "bi-interaction", "graphsage"])
def test_differentiable_fold_chains_to_the_layer_parameters(agg):
    """Gradients w.r.t. the folded tensors (what csrc/backward.cu produces) chained through the fold equal autograd
    through the reference formulation."""
    lamda, alpha, l = 0.5, 0.1, 2
    m = _layer(agg, True, seed=3)
    g = torch.Generator().manual_seed(2)
    ego, side = torch.randn(9, 20, generator=g, dtype=torch.float64), torch.randn(9, 20, generator=g, dtype=torch.float64)
    h0 = torch.randn(9, 12, generator=g, dtype=torch.float64)
    w1, w2 = torch.randn(9, 8, generator=g, dtype=torch.float64), torch.randn(9, 8, generator=g, dtype=torch.float64)
    ref1, ref2 = _oracle_layer_pre_activation(m, agg, True, ego, side, h0, lamda, alpha, l)
    ((ref1 * w1).sum() + (0 if ref2 is None else (ref2 * w2).sum())).backward()
    want = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad()
    f = m.folded(lamda, alpha, l, differentiable=True)
    o1 = ego.float() @ f["pa"] + side.float() @ f["pb"] + h0.float() @ f["q1"] + f["c1"]
    loss = (o1 * w1.float()).sum()
    if agg == "bi-interaction":
        loss = loss + (((ego * side).float() @ f["p2"] + h0.float() @ f["q2"] + f["c2"]) * w2.float()).sum()
    loss.backward()
    for k, p in m.named_parameters():
        if k in want:
            err = (p.grad - want[k]).abs().max() / want[k].abs().max().clamp_min(1e-30)
            assert err < 1e-5, (k, float(err))


def test_gate_pair_layout_matches_the_reference_gate():
    from literalkg_b200.gate import GateMul, Gate
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(4)
    e, nl, tl = torch.randn(5, 6, generator=g), torch.randn(5, 2, generator=g), torch.randn(5, 4, generator=g)
    gm = GateMul(6, 2, 4)
    torch.nn.init.normal_(gm.gate_bias, std=0.2)
    w, b = gm.packed()
    pre = torch.cat([e, nl, tl], 1) @ w.t() + b                          # interleaved (g_j, z_j)
    out = (1 - torch.sigmoid(pre[:, 1::2])) * e + torch.sigmoid(pre[:, 1::2]) * torch.tanh(pre[:, 0::2])
    p = {"x." + k: v for k, v in gm.state_dict().items()}
    assert torch.allclose(out, O.gate_mul(p, "x.", e, nl, tl), atol=1e-6)
    gs = Gate(6, 4)
    w, b = gs.packed()
    pre = torch.cat([e, tl], 1) @ w.t() + b
    out = (1 - torch.sigmoid(pre[:, 1::2])) * e + torch.sigmoid(pre[:, 1::2]) * torch.tanh(pre[:, 0::2])
    p = {"x." + k: v for k, v in gs.state_dict().items()}
    assert torch.allclose(out, O.gate_single(p, "x.", e, tl), atol=1e-6)
    w2, _ = gs.pair()                                                    # autograd-connected twin of packed()
    assert w2.requires_grad and torch.equal(w2.detach(), w)


def test_segmented_schedule_partitions_every_heavy_row():
    from literalkg_b200.graph import expand_schedule
    # rows sorted by decreasing triple count; agg ranges slightly shorter (merged duplicate pairs)
    counts = [5000, 1300, 513, 512, 300, 257, 256, 7, 1, 0]
    att = torch.tensor([0] + counts).cumsum(0)
    agg_counts = [c - (c // 100) for c in counts]
    agg = torch.tensor([0] + agg_counts).cumsum(0)
    rows = torch.tensor([42, 7, 9, 3, 11, 5, 6, 1, 0, 2])
    sched = torch.zeros((10, 8), dtype=torch.int32)
    sched[:, 0], sched[:, 1], sched[:, 2] = rows, att[:-1], att[1:]
    sched[:, 3], sched[:, 4] = agg[:-1], agg[1:]
    out, n_solo, n_heavy = expand_schedule(sched, seg_degree=512, max_segs=8, solo_degree=256)
    assert n_heavy == 3 and n_solo == 6
    nseg = [8, 3, 2]                                      # ceil(5000/512) = 10 -> capped at 8; 3; 2
    assert out.shape[0] == sum(nseg) + 7
    pos = 0
    for ticket, (k, row) in enumerate(zip(nseg, [42, 7, 9])):
        seg = out[pos:pos + k]
        pos += k
        assert (seg[:, 0] == row).all() and (seg[:, 5] == k).all() and (seg[:, 6] == ticket).all()
        assert seg[:, 7].tolist() == list(range(k))
        src = sched[ticket]
        for lo, hi, a, b in ((1, 2, src[1], src[2]), (3, 4, src[3], src[4])):      # exact, ordered partitions
            assert seg[0, lo] == a and seg[-1, hi] == b
            assert torch.equal(seg[1:, lo], seg[:-1, hi]) and (seg[:, hi] > seg[:, lo]).all()
    assert torch.equal(out[pos:], sched[3:])              # ordinary rows untouched (nseg == 0)
    same, n_solo0, n_heavy0 = expand_schedule(sched[3:])
    assert n_heavy0 == 0 and same is not None and torch.equal(same, sched[3:]) and n_solo0 == 3


def test_stacked_q_layout():
    """Rows of the stacked h0 @ Q weight: [layer 0: q1 + pa | q2 | layer 1: q1 | q2 | ... | z = layer 0's pb]."""
    import literalkg_b200 as L
    cfg = O.OracleConfig(n_conv_layers=2, embed_dim=128, relation_dim=128, conv_dim=8, scale_gat_dim=16,
                         num_lit_dim=2, txt_lit_dim=4)
    args = argparse.Namespace(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    m = L.LiteralKG(args, 10, 2, None, torch.zeros(10, 2), torch.zeros(10, 4))
    with torch.no_grad():
        folds = [layer.folded(m.lamda, m.alpha, k + 1) for k, layer in enumerate(m.aggregator_layers)]
        wq, cq, offsets, zcol = m._stack_q(folds)
    assert tuple(wq.shape) == (2 * 16 + 8, 128) and offsets == [0, 16] and zcol == 32
    assert torch.equal(wq[0:8], (folds[0]["q1"] + folds[0]["pa"]).t()) and torch.equal(wq[8:16], folds[0]["q2"].t())
    assert torch.equal(wq[16:24], folds[1]["q1"].t()) and torch.equal(wq[32:40], folds[0]["pb"].t())
    assert torch.equal(cq[:8], folds[0]["c1"]) and (cq[32:] == 0).all()


def test_batch_generator_oracle_and_contract_checker():
    """The oracle restatement of the reference's batch generators satisfies the contract the device sampler is tested
    against (tests/test_sampler_gpu.py), and the checker rejects batches that break it."""
    import random
    import numpy as np
    rng = np.random.default_rng(0)
    h, t, r = rng.integers(0, 60, 900), rng.integers(0, 200, 900), rng.integers(0, 4, 900)
    kg = O.build_kg_dict(h, t, r)
    tails = np.unique(t).tolist()
    bh, br, bp, bn = O.generate_kg_batch(kg, 90, 3, tails, random.Random(1))
    assert bh.shape == (90,) and len(set(bh.tolist())) == 30
    O.check_batch_contract(kg, bh, br, bp, bn, 3, tails, distinct_heads=True)
    bad = bn.copy()
    bad[0] = bp[0]                                            # a positive of the head as its negative
    with pytest.raises(AssertionError):
        O.check_batch_contract(kg, bh, br, bp, bad, 3, tails, distinct_heads=True)
    big = O.generate_kg_batch(kg, 600, 3, tails, random.Random(2))          # 200 heads > 60 existing: with replacement
    O.check_batch_contract(kg, *big, 3, tails, distinct_heads=False)
    head_dict = {k: sorted({tt for tt, _ in v}) for k, v in kg.items()}
    ph, pp, pn = O.generate_prediction_batch(head_dict, 60, 3, tails, random.Random(3))
    O.check_batch_contract(kg, ph, None, pp, pn, 3, tails, distinct_heads=True)
