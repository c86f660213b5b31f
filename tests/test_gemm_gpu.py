"""The tcgen05 GEMM engine (scaled hi/lo fp16 planes, three products per k-step, fp32 TMEM accumulation) against
a float64 torch reference of the same op.  Claimed bound: 5e-6 relative to max|C| (fp32-class; the split keeps 22
significand bits); the path-level bound of BASELINE.json is 1e-3."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 5e-6


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("m,n,ks", [(128, 64, [64]), (300, 192, [300]), (1000, 256, [396]), (257, 600, [300, 302]),
                                    (77, 16, [12]), (513, 40, [12, 10]), (4096, 272, [300, 2, 300])])
@pytest.mark.parametrize("act", [0, 1])
def test_linear(m, n, ks, act):
    from literalkg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n)
    segs = [torch.randn(m, k, generator=g, device="cuda") for k in ks]
    w = torch.randn(n, sum(ks), generator=g, device="cuda") / sum(ks) ** 0.5
    b = torch.randn(n, generator=g, device="cuda")
    out = ops.linear([ops.split_planes(s) for s in segs], w, b, act)
    ref = torch.cat(segs, 1).double() @ w.double().t() + b.double()
    if act:
        ref = torch.nn.functional.leaky_relu(ref, 0.01)
    assert out.shape == ref.shape
    assert rel(out, ref) < TOL


def test_linear_planes_output_and_strided_out():
    from literalkg_b200 import _lib, ops
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(700, 300, generator=g, device="cuda")
    w = torch.randn(96, 300, generator=g, device="cuda") * 0.05
    big = torch.zeros(700, 200, device="cuda")
    ref = x.double() @ w.double().t()
    planes = _lib.Planes(700, 96, "cuda", rec=ops.scale_from_bound(float(ref.abs().max()) * 1.01, "cuda"))
    out = ops.linear([ops.split_planes(x)], w, None, 0, out=big[:, 40:136], out_planes=planes)
    assert rel(out, ref) < TOL
    assert (big[:, :40] == 0).all() and (big[:, 136:] == 0).all()
    assert rel(planes.dequant(), ref) < TOL


def test_scale_records_and_mixed_magnitudes():
    """Segments of very different magnitude (entity embeddings ~1e-3 next to literals ~1) keep fp32-class accuracy:
    every segment has its own power-of-two scale and the weight absorbs the ratio."""
    from literalkg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    a0 = torch.randn(513, 300, generator=g, device="cuda") * 2e-3
    a1 = torch.rand(513, 2, generator=g, device="cuda")
    a2 = torch.randn(513, 300, generator=g, device="cuda") * 40.0
    w = torch.randn(64, 602, generator=g, device="cuda")
    w[:, :300] *= 100.0
    p0 = ops.split_planes(a0)
    rec = p0.rec.cpu()
    assert rec[1] * rec[2] == 1.0 and 2048 <= rec[0] * rec[1] < 4096        # power of two, absmax parked in [2^11, 2^12)
    assert float(rec[0]) == float(a0.abs().max())
    assert rel(p0.dequant(), a0) < 1e-6
    out = ops.linear([p0, ops.split_planes(a1), ops.split_planes(a2)], w, None, 0)
    ref = torch.cat([a0, a1, a2], 1).double() @ w.double().t()
    assert rel(out, ref) < TOL
    z = ops.split_planes(torch.zeros(8, 16, device="cuda"))                     # all-zero operand: scale 1, no NaN
    assert float(z.rec[1]) == 1.0 and (z.dequant() == 0).all()


@pytest.mark.parametrize("m,dim,lits", [(500, 300, [2, 300]), (130, 12, [2, 8]), (64, 12, [8])])
def test_gate(m, dim, lits):
    import literalkg_b200 as L
    g = torch.Generator().manual_seed(dim + m)
    mod = (L.GateMul(dim, *lits) if len(lits) == 2 else L.Gate(dim, lits[0]))
    with torch.no_grad():
        mod.gate_bias.add_(torch.randn(dim, generator=g) * 0.2)
    mod = mod.cuda()
    xs = [torch.randn(m, dim, generator=g).cuda()] + [torch.randn(m, k, generator=g).cuda() for k in lits]
    with pytest.raises(RuntimeError):          # standalone modules are inference modules: no silent detach under grad
        mod(*xs)
    with torch.no_grad():
        out = mod(*xs)
    d = lambda t: t.detach().double()
    x = torch.cat([d(t) for t in xs], 1)
    gg = torch.tanh(x @ d(mod.g.weight).t() + d(mod.g.bias))
    if len(lits) == 2:
        z = d(xs[0]) @ d(mod.gate_ent.weight).t() + d(xs[1]) @ d(mod.gate_num_lit.weight).t() + d(xs[2]) @ d(mod.gate_txt_lit.weight).t()
    else:
        z = d(xs[0]) @ d(mod.gate_ent.weight).t() + d(xs[1]) @ d(mod.gate_lit.weight).t()
    z = torch.sigmoid(z + d(mod.gate_bias))
    ref = (1 - z) * d(xs[0]) + z * gg
    assert rel(out, ref) < TOL


@pytest.mark.parametrize("nh,nt,dim", [(9, 23, 16), (300, 5000, 256), (2048, 3000, 256)])
def test_score_minmax(nh, nt, dim):
    from literalkg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(nh)
    emb = torch.randn(6000, dim, generator=g, device="cuda")
    heads = torch.randint(0, 6000, (nh,), generator=g, device="cuda")
    tails = torch.randint(0, 6000, (nt,), generator=g, device="cuda")
    mm = torch.empty(2, dtype=torch.int32, device="cuda")
    s = ops.score(emb, heads, tails, mm)
    ref = emb[heads].double() @ emb[tails].double().t()
    assert rel(s, ref) < TOL
    pred = ops.predict(emb, heads, tails, 0.5)
    norm = (s - s.min()) / (s.max() - s.min())
    assert torch.equal(pred, (norm > 0.5).int())            # min / max / threshold are exact on our own scores


def test_exact_integers():
    """Small integers are exact in fp16 and in fp32 accumulation: the engine must be bit exact."""
    from literalkg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randint(-8, 9, (640, 128), generator=g, device="cuda").float()     # power-of-two scaling keeps integers exact
    w = torch.randint(-8, 9, (256, 128), generator=g, device="cuda").float()
    out = ops.linear([ops.split_planes(x)], w, None, 0)
    assert torch.equal(out, x @ w.t())


# ---- CTA pairs (tcgen05.mma.cta_group::2 over 256-row tiles) -------------------------------------------------------
@pytest.fixture
def cta_group():
    """Forces the CTA-group size of the GEMM engine for one test and restores the automatic choice afterwards."""
    from literalkg_b200 import _lib

    def set_(cg):
        _lib.check(_lib.load().lkg_gemm_set_cta_group(cg))
    yield set_
    set_(0)


@pytest.mark.parametrize("m,n,ks", [(256, 64, [64]), (129, 16, [12]), (1000, 600, [300, 302]), (38000, 600, [300, 302]),
                                    (20001, 224, [300]), (19999, 256, [300, 96]), (513, 208, [64, 64, 64, 64])])
def test_cta_pair_matches_single_cta(m, n, ks, cta_group):
    """Same products in the same order: a CTA pair must reproduce the single-CTA result bit for bit, ragged last
    tiles (rows of the second CTA entirely out of range) and every K-segment layout included."""
    from literalkg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(m + n)
    segs = [torch.randn(m, k, generator=g, device="cuda") for k in ks]
    w = torch.randn(n, sum(ks), generator=g, device="cuda") / sum(ks) ** 0.5
    b = torch.randn(n, generator=g, device="cuda")
    planes = [ops.split_planes(s) for s in segs]
    outs = []
    for cg in (1, 2):
        cta_group(cg)
        outs.append(ops.linear(planes, w, b, 1))
    ref = torch.nn.functional.leaky_relu(torch.cat(segs, 1).double() @ w.double().t() + b.double(), 0.01)
    assert rel(outs[1], ref) < TOL
    assert torch.equal(outs[0], outs[1])


def test_cta_pair_gate_and_score(cta_group):
    import literalkg_b200 as L
    from literalkg_b200 import ops
    g = torch.Generator().manual_seed(11)
    m, dim, lits = 30001, 300, [2, 300]
    mod = L.GateMul(dim, *lits).cuda()
    xs = [torch.randn(m, dim, generator=g).cuda()] + [torch.randn(m, k, generator=g).cuda() for k in lits]
    res = []
    for cg in (1, 2):
        cta_group(cg)
        with torch.no_grad():
            res.append(mod(*xs))
    d = lambda t: t.detach().double()
    x = torch.cat([d(t) for t in xs], 1)
    gg = torch.tanh(x @ d(mod.g.weight).t() + d(mod.g.bias))
    z = torch.sigmoid(d(xs[0]) @ d(mod.gate_ent.weight).t() + d(xs[1]) @ d(mod.gate_num_lit.weight).t()
                      + d(xs[2]) @ d(mod.gate_txt_lit.weight).t() + d(mod.gate_bias))
    assert rel(res[1], (1 - z) * d(xs[0]) + z * gg) < TOL
    assert torch.equal(res[0], res[1])
    emb = torch.randn(50000, 256, generator=g).cuda()
    heads = torch.arange(0, 700, device="cuda") * 3
    tails = torch.arange(0, 45003, device="cuda")
    sc = []
    for cg in (1, 2):
        cta_group(cg)
        mm = torch.empty(2, dtype=torch.int32, device="cuda")
        sc.append((ops.score(emb, heads, tails, mm), mm.clone()))
    assert rel(sc[1][0], emb[heads].double() @ emb[tails].double().t()) < TOL
    assert torch.equal(sc[0][0], sc[1][0]) and torch.equal(sc[0][1], sc[1][1])


@pytest.mark.parametrize("m,k", [(1000, 32), (513, 30), (77, 12), (2049, 300), (300, 302), (5, 4), (1, 1), (4097, 64)])
@pytest.mark.parametrize("layout", ["contiguous", "view", "gather", "odd_view"])
def test_split_planes_layouts(m, k, layout):
    """The flat-indexed absmax / split kernels on narrow, ragged, strided and gathered operands: record == the data's
    absmax and its power-of-two scale, hi + lo reproduces the value to 2^-22 of the absmax, pad columns are zero."""
    from literalkg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(m * 31 + k)
    rows = None
    if layout == "contiguous":
        src = torch.randn(m, k, generator=g, device="cuda")
    elif layout == "view":
        src = torch.randn(m, k + 24, generator=g, device="cuda")[:, 8:8 + k]          # 16-byte aligned window
    elif layout == "odd_view":
        src = torch.randn(m, k + 7, generator=g, device="cuda")[:, 3:3 + k]           # unaligned: scalar path
    else:
        src = torch.randn(3 * m + 1, k, generator=g, device="cuda")
        rows = torch.randint(0, 3 * m + 1, (m,), generator=g, device="cuda")
    src[m // 2, k // 2] = -7.5                                                        # the absmax, negative
    ref = src if rows is None else src[rows]
    pl = ops.split_planes(src, rows)
    amax = ref.abs().max()
    assert pl.rec[0] == amax
    scaled = float(amax * pl.rec[1])
    assert 2048.0 <= scaled < 4096.0 and float(pl.rec[1] * pl.rec[2]) == 1.0
    assert (pl.dequant() - ref).abs().max() <= float(amax) * 2.0 ** -22
    assert (pl.t[:, :m, k:] == 0).all()
    # record built in two halves by the accumulate entry point == the one-pass record
    rec = ops.raw_record("cuda")
    if rows is None and m > 1:
        ops.absmax_accumulate(src[: m // 2], rec)
        ops.absmax_accumulate(src[m // 2:], rec)
        ops.scale_finish(rec)
        assert torch.equal(rec[:3], pl.rec[:3])
        assert torch.equal(ops.split_planes(src, rec=rec).t, pl.t)


@pytest.mark.parametrize("m,n,k,split", [(1000, 224, 300, 192), (257, 96, 64, 4), (4096, 272, 300, 256), (300, 40, 12, 36)])
def test_linear_split_output(m, n, k, split):
    """lkg_linear_fwd_split: the result columns from ``split`` on go to a second buffer with its own row stride (the
    [h0 | z] table), bit-identical to the one-destination GEMM."""
    from literalkg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(m + n)
    x = torch.randn(m, k, generator=g, device="cuda")
    w = torch.randn(n, k, generator=g, device="cuda") / k ** 0.5
    b = torch.randn(n, generator=g, device="cuda")
    xp = ops.split_planes(x)
    ref = ops.linear([xp], w, b, 0)
    big = torch.full((m, n - split + 24), -7.0, device="cuda")
    out = torch.full((m, n), -7.0, device="cuda")
    ops.linear([xp], w, b, 0, out=out, out2=big[:, 8:8 + n - split], split_col=split)
    assert torch.equal(out[:, :split], ref[:, :split]) and (out[:, split:] == -7.0).all()
    assert torch.equal(big[:, 8:8 + n - split], ref[:, split:])
    assert (big[:, :8] == -7.0).all() and (big[:, 8 + n - split:] == -7.0).all()
    assert rel(ref, x.double() @ w.double().t() + b.double()) < TOL
