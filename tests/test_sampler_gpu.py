"""Device-side minibatch assembly (csrc/sample.cu) against the contract of the reference's generators
(dataloader.py:192-333).  The random streams differ by construction (Python's `random` vs a counter-based device
generator), so parity is the set of properties the reference's loops guarantee + distribution checks."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _kg(n=3000, e=40000, n_rel=7, seed=3):
    import literalkg_b200 as L
    kg = L.synthetic.make_kg(n, e, n_rel, seed=seed, max_out_degree=200)
    h, t, r = (torch.from_numpy(x).cuda() for x in (kg.h, kg.t, kg.r))
    return kg, L.GraphPlan(h, t, r, n, n_rel)


def test_kg_batch_contract():
    import literalkg_b200 as L
    kg, plan = _kg()
    triples = set(zip(kg.h.tolist(), kg.r.tolist(), kg.t.tolist()))
    cand = np.unique(kg.t)                                    # training_tails
    s = L.BatchSampler(plan, cand, neg_rate=3, use_relation=True, seed=11)
    h, r, p, ng = (x.cpu().numpy() for x in s.sample(2048))
    n = 2048 // 3
    assert h.shape == r.shape == p.shape == ng.shape == (n * 3,)
    heads = h.reshape(n, 3)
    assert (heads == heads[:, :1]).all() and (r.reshape(n, 3) == r.reshape(n, 3)[:, :1]).all()   # repeated by neg rate
    assert (p.reshape(n, 3) == p.reshape(n, 3)[:, :1]).all()
    assert len(set(heads[:, 0].tolist())) == n                # random.sample: heads without replacement
    candset = set(cand.tolist())
    for i in range(n):
        assert (int(heads[i, 0]), int(r[3 * i]), int(p[3 * i])) in triples          # a real positive triple
        negs = ng[3 * i:3 * i + 3].tolist()
        assert len(set(negs)) == 3 and all(x in candset for x in negs)              # distinct, from the candidates
        assert all((int(heads[i, 0]), int(r[3 * i]), int(x)) not in triples for x in negs)
    assert int(s.n_failed.item()) == 0
    import literalkg_oracle as O                              # the same checker the oracle's generators pass on CPU
    O.check_batch_contract(O.build_kg_dict(kg.h, kg.t, kg.r), h, r, p, ng, 3, cand, distinct_heads=True)
    # reproducible for a seed, different across calls
    s2 = L.BatchSampler(plan, cand, neg_rate=3, use_relation=True, seed=11)
    again = s2.sample(2048)
    assert all(torch.equal(a.cpu(), torch.from_numpy(b)) for a, b in zip(again, (h, r, p, ng)))
    assert not torch.equal(s2.sample(2048)[3].cpu(), torch.from_numpy(ng))


def test_prediction_batch_and_oversampling():
    """Fine-tuning pairs: relation-free rejection; a batch larger than the head list draws heads with replacement."""
    import literalkg_b200 as L
    rng = np.random.default_rng(0)
    n = 500
    heads = rng.integers(0, 40, 300)                          # 40 heads only
    tails = rng.integers(400, 420, 300)                       # 20 candidate tails: most are positives of someone
    pairs = np.unique(np.stack([heads, tails], 1), axis=0)
    plan = L.GraphPlan(torch.from_numpy(pairs[:, 0]).cuda(), torch.from_numpy(pairs[:, 1]).cuda(),
                       torch.zeros(len(pairs), dtype=torch.int64).cuda(), n, 1)
    pos = {}
    for a, b in pairs.tolist():
        pos.setdefault(a, set()).add(b)
    s = L.BatchSampler(plan, np.arange(400, 420), neg_rate=3, use_relation=False, seed=5)
    h, r, p, ng = s.sample(600)                               # 200 heads > 40 existing heads
    assert r is None
    h, p, ng = h.cpu().numpy(), p.cpu().numpy(), ng.cpu().numpy()
    assert len(h) == 600 and set(h.tolist()) <= set(pos)
    failed = 0
    assert ng.min() >= 400 and ng.max() < 420                 # emitted ids are always valid candidates (no -1 sentinel)
    for i in range(0, 600, 3):
        assert int(p[i]) in pos[int(h[i])]
        negs = [int(x) for x in ng[i:i + 3]]
        ok = list(dict.fromkeys(x for x in negs if x not in pos[int(h[i])]))          # admissible and distinct
        # a head whose positives leave fewer than 3 admissible tails cannot be served (the reference loops forever):
        # such draws are flagged in n_failed and repeat / reuse a candidate
        assert len(ok) == min(3, 20 - len(pos[int(h[i])])) or len(ok) == 3
        failed += 3 - len(ok)
    assert int(s.n_failed.item()) >= failed
    if failed:
        with pytest.raises(RuntimeError):
            s.check()
    with pytest.raises(ValueError):                           # rows sorted by (relation, tail): no tail-only search
        kg3 = L.GraphPlan(torch.from_numpy(pairs[:, 0]).cuda(), torch.from_numpy(pairs[:, 1]).cuda(),
                          (torch.arange(len(pairs)) % 3).cuda(), n, 3)
        L.BatchSampler(kg3, np.arange(400, 420), neg_rate=3, use_relation=False)


def test_positive_draw_is_uniform():
    import literalkg_b200 as L
    n = 50
    t = torch.arange(10, 30)
    h = torch.zeros(20, dtype=torch.int64)
    plan = L.GraphPlan(h.cuda(), t.cuda(), torch.zeros(20, dtype=torch.int64).cuda(), n, 1)
    s = L.BatchSampler(plan, np.arange(30, 50), neg_rate=1, use_relation=True, seed=1)
    counts = np.zeros(20)
    negc = np.zeros(20)
    for _ in range(300):
        _, _, p, ng = s.sample(64)                            # 64 > 1 head: drawn with replacement
        counts += np.bincount(p.cpu().numpy() - 10, minlength=20)
        negc += np.bincount(ng.cpu().numpy() - 30, minlength=20)
    for c in (counts, negc):                                  # 19 200 draws over 20 bins: chi-square, 19 dof
        chi2 = ((c - c.mean()) ** 2 / c.mean()).sum()
        assert chi2 < 60, (chi2, c)
