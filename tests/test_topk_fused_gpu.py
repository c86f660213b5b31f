"""Fused all-entity scoring + top-k (lkg_score_index / lkg_score_topk) against (a) the generic path of this
library (3-product score matrix + lkg_topk_rows) and (b) a float64 torch reference with the declared rule
"larger score first, ties -> lower position".  The fused path re-scores its candidates exactly (fp32 products summed
in fp64, one rounding), so positions must match the float64 ranking wherever the float64 scores differ by more than
one fp32 ulp, and exactly-equal scores must come out in position order."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _inference_mode():
    """The parity tests exercise the inference path; with grad enabled gat_embeddings() takes the training path
    (same kernels + saved activations), which tests/test_backward_gpu.py covers."""
    with torch.no_grad():
        yield


def ref_topk(emb, heads, tails, k):
    s = emb[heads].double() @ emb[tails].double().t()
    sf = s.float()                                             # correctly rounded fp32 scores
    order = np.lexsort((np.broadcast_to(np.arange(sf.shape[1]), sf.shape), -sf.cpu().numpy()), axis=1)[:, :k]
    order = torch.from_numpy(order.copy()).to(emb.device)
    return torch.gather(sf, 1, order), order


@pytest.mark.parametrize("n,dim,nh,k", [(40_000, 256, 300, 10), (70_001, 256, 129, 100), (30_000, 64, 64, 5),
                                        (25_000, 200, 513, 10)])
def test_fused_topk_matches_float64_ranking(n, dim, nh, k):
    from literalkg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(n + dim)
    emb = torch.nn.functional.leaky_relu(torch.randn(n, dim, generator=g, device="cuda"), 0.01) * 0.37
    heads = torch.randint(0, n, (nh,), generator=g, device="cuda")
    tails = torch.randperm(n, generator=g, device="cuda")[: n - 7]          # a gathered, permuted tail list
    vals, pos = ops.score_topk(emb, heads, tails, k)
    rv, rp = ref_topk(emb, heads, tails, k)
    assert torch.equal(pos, rp)
    assert torch.equal(vals, rv)
    # identity tail list (tails=None) and a reused index give the same answer
    ti = ops.ScoreIndex(emb, None)
    v2, p2 = ops.score_topk(emb, heads, None, k, tail_index=ti)
    rv2, rp2 = ref_topk(emb, heads, torch.arange(n, device="cuda"), k)
    assert torch.equal(p2, rp2) and torch.equal(v2, rv2)


def test_fused_topk_external_bound_and_sampling_knobs():
    """A caller-supplied lower bound of the k-th best score (theta) and any sampling density give the same result:
    the bound only decides how many candidates are re-scored."""
    from literalkg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    n, dim, k = 60_000, 128, 7
    emb = torch.randn(n, dim, generator=g, device="cuda") * 2.0
    heads = torch.randint(0, n, (70,), generator=g, device="cuda")
    rv, rp = ref_topk(emb, heads, torch.arange(n, device="cuda"), k)
    for kw in (dict(sample_tiles=16), dict(sample_tiles=469), dict(theta=rv[:, -1] - 0.5), dict(theta=rv[:, -1].clone())):
        v, p_ = ops.score_topk(emb, heads, None, k, **kw)
        assert torch.equal(p_, rp) and torch.equal(v, rv), kw


def test_fused_topk_plateau_overflow_and_ties():
    """20 000 identical best tails: the candidate band overflows every head's list; the exact fallback must return
    the lowest positions of the plateau, in order."""
    from literalkg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    n, dim, k = 50_000, 128, 10
    emb = torch.randn(n, dim, generator=g, device="cuda") * 0.1
    plateau = torch.randperm(n, generator=g, device="cuda")[:20_000]
    emb[plateau] = emb[plateau[0]].clone() * 0 + 1.0                        # identical rows with the largest scores
    heads = torch.arange(40, device="cuda") * 13 + 1
    emb[heads] = emb[heads].abs() + 0.5                                     # positive heads: plateau rows win
    vals, pos = ops.score_topk(emb, heads, None, k, cap=1024)
    rv, rp = ref_topk(emb, heads, torch.arange(n, device="cuda"), k)
    assert torch.equal(pos, rp)
    assert torch.equal(vals, rv)
    assert (pos[:, 1:] > pos[:, :-1]).all()                                  # ties in position order


def test_fused_topk_vs_generic_path_and_model_api():
    import argparse
    import literalkg_b200 as L
    import literalkg_oracle as O
    from literalkg_b200 import ops
    cfg = O.OracleConfig(n_conv_layers=2, mess_dropout=0.0)
    n, n_rel = 20_000, 8
    kg = L.synthetic.make_kg(n, 120_000, n_rel, seed=5, max_out_degree=200)
    num, txt = L.synthetic.make_literals(n, seed=5)
    args = argparse.Namespace(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    kt = L.KGTensors(kg.h, kg.t, kg.r, n_entities=n)
    torch.manual_seed(0)
    m = L.LiteralKG(args, n, n_rel, kt.A_in, num, txt).cuda().eval()
    with torch.no_grad():
        m.entity_embed.weight.mul_(30)
    m(kt.h_list, kt.t_list, kt.r_list, kt.relations, device="cuda", mode="update_att")
    emb = m.gat_embeddings()
    heads = torch.arange(0, 200, device="cuda") * 37 % n
    tails = torch.arange(n, device="cuda")
    v, p, r = m.topk(heads, tails, 10, all_embed=emb)                        # n >= 16384: fused path
    assert r is None
    s = ops.score(emb, heads, tails)                                        # generic path: 3-product score matrix
    gv, gp, _ = ops.topk_rows(s, 10)
    # the two paths round the scores differently (exact vs 2^-22): compare where the generic top-11 is separated
    tv, _ = torch.topk(s, 11, dim=1)
    clear = ((tv[:, :-1] - tv[:, 1:]).min(dim=1).values > 1e-5 * s.abs().max())
    assert clear.sum() > 20
    assert torch.equal(p[clear], gp[clear])
    assert ((v - gv).abs().max() / gv.abs().max()).item() < 1e-5
    rv, rp = ref_topk(emb, heads, tails, 10)
    assert torch.equal(p, rp) and torch.equal(v, rv)


# ---- rank epilogue: lkg_rank_prepare / lkg_score_rank / lkg_rank_finalize ---------------------------------------
def ref_ranks(emb, heads, tails, target_pos, chunk=128):
    """O.rank_of's rule on the exact scores (fp64 dot products rounded once to fp32), chunked over the heads."""
    e64 = emb.double()
    t64 = e64 if tails is None else e64[tails]
    pos = torch.arange(t64.shape[0], device=emb.device).unsqueeze(0)
    out = []
    for i in range(0, heads.numel(), chunk):
        s = (e64[heads[i:i + chunk]] @ t64.t()).float()
        tp = target_pos[i:i + chunk].unsqueeze(1)
        tgt = torch.gather(s, 1, tp)
        out.append(((s > tgt) | ((s == tgt) & (pos < tp))).sum(1))
    return torch.cat(out)


@pytest.mark.parametrize("n,dim,nh", [(40_000, 256, 300), (70_001, 256, 129), (25_000, 200, 513), (9_000, 64, 77)])
def test_rank_epilogue_bit_exact(n, dim, nh):
    from literalkg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(n + nh)
    emb = torch.nn.functional.leaky_relu(torch.randn(n, dim, generator=g, device="cuda"), 0.01) * 0.37
    emb[500:520] = emb[7]                                           # exact ties: resolved by position
    heads = torch.randint(0, n, (nh,), generator=g, device="cuda")
    heads[:4] = torch.tensor([7, 505, 7, 519], device="cuda")
    tails = torch.randperm(n, generator=g, device="cuda")[: n - 5]
    target = torch.randint(0, n - 5, (nh,), generator=g, device="cuda")
    where = {int(t): i for i, t in enumerate(tails.tolist())} if n < 10_000 else None
    if where is not None:                                           # make some targets members of the tie group
        target[0], target[1] = where[505], where[7]
    ti = ops.ScoreIndex(emb, tails)
    got = ops.score_rank(emb, heads, target, ti)
    assert torch.equal(got, ref_ranks(emb, heads, tails, target))
    # identity tail list; the best and the worst tail of every head get ranks 0 and n - 1
    ti2 = ops.ScoreIndex(emb, None)
    s = (emb[heads[:16]].double() @ emb.double().t()).float()
    best = torch.sort(s, dim=1, descending=True, stable=True).indices
    assert torch.equal(ops.score_rank(emb, heads[:16], best[:, 0], ti2), torch.zeros(16, dtype=torch.int64, device="cuda"))
    # a tiny band capacity forces the overflow path (exact scan): same answer
    got3 = ops.score_rank(emb, heads[:32], target[:32], ti, band_cap=1)
    assert torch.equal(got3, got[:32])


def test_model_topk_returns_ranks_without_score_matrix():
    """LiteralKG.topk(target_tails=...) on a large candidate set goes through the fused kernels for values, positions
    AND ranks; identical to the dense path (score matrix + lkg_topk_rows) on the same embeddings."""
    import argparse
    import literalkg_b200 as L
    import literalkg_oracle as O
    from literalkg_b200 import ops
    cfg = O.OracleConfig(n_conv_layers=1, mess_dropout=0.0)
    n = 30_000
    kg = L.synthetic.make_kg(n, 120_000, 4, seed=1, max_out_degree=100)
    num, txt = L.synthetic.make_literals(n, seed=1)
    args = argparse.Namespace(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    kt = L.KGTensors(kg.h, kg.t, kg.r, n_entities=n)
    torch.manual_seed(0)
    m = L.LiteralKG(args, n, 4, kt.A_in, num, txt).cuda().eval()
    heads = torch.arange(0, 200, device="cuda") * 37 % n
    tails = torch.arange(n, device="cuda")
    target = (heads * 11 + 3) % n
    vals, pos, ranks = m.topk(heads, tails, 10, target_tails=target)
    emb = m.gat_embeddings()
    assert torch.equal(ranks, ref_ranks(emb, heads, None, target))
    rv, rp = ref_topk(emb, heads, tails, 10)
    assert torch.equal(pos, rp) and torch.equal(vals, rv)
