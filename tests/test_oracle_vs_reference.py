"""Live cross-check of the oracle against the UNMODIFIED reference classes, on inputs that are not among the
committed golden vectors.  Runs only where the reference checkout is mounted (the build container); the GPU box has
no /root/reference and the committed golden vectors (tests/test_oracle_golden.py) play this role there."""
import os
import sys
import warnings

import numpy as np
import pytest
import torch

import literalkg_oracle as O

REF = os.environ.get("LKG_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "model.py")),
                                reason="reference checkout not mounted")

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))


def close(a, b, rtol, atol=1e-7):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    assert a.shape == b.shape
    assert (a - b).abs().max().item() <= atol + rtol * b.abs().max().item(), (a - b).abs().max().item()


@pytest.mark.parametrize("agg,res,layers,seed", [("bi-interaction", True, 3, 101), ("gcn", False, 2, 102),
                                                 ("graphsage", True, 2, 103), ("bi-interaction", False, 1, 104)])
def test_oracle_matches_live_reference(agg, res, layers, seed):
    import make_golden as G
    ref_model, _ = G.import_reference()
    cfg = O.OracleConfig(aggregation_type=agg, use_residual=res, n_conv_layers=layers, mess_dropout=0.0,
                         embed_dim=20, relation_dim=20, scale_gat_dim=12, conv_dim=8, num_lit_dim=2, txt_lit_dim=6)
    n, n_rel = 83, 5
    torch.manual_seed(seed)
    h, t, r = G.make_kg(n, n_rel, 400, seed)
    rng = np.random.default_rng(seed)
    num = torch.from_numpy(rng.uniform(0, 1, (n, 2)).astype(np.float32)) * (torch.rand(n, 1) < 0.3)
    txt = torch.from_numpy(rng.normal(0, 0.3, (n, 6)).astype(np.float32)) * (torch.rand(n, 1) < 0.3)
    lap_idx, lap_val, relations = G.ref_laplacian(h, t, r, n, "random-walk")
    oi, ov = O.laplacian_A_in(h, t, r, n)
    assert np.array_equal(oi, lap_idx) and np.array_equal(ov, lap_val)            # bit exact
    a0 = torch.sparse_coo_tensor(torch.from_numpy(lap_idx), torch.from_numpy(lap_val), (n, n))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = ref_model.LiteralKG(G.namespace(cfg), n, n_rel, a0, num, txt).eval()
        with torch.no_grad():
            model.entity_embed.weight.mul_(8)
            model.relation_embed.weight.mul_(4)
        p = {k: v.detach().clone() for k, v in model.state_dict().items() if k != "A_in"}
        hl, tl, rl = (torch.from_numpy(x) for x in (h, t, r))
        model(hl, tl, rl, relations, device="cpu", mode="update_att")
        a = model.A_in.data.coalesce()
        idx, val = O.update_attention(p["entity_embed.weight"], p["relation_embed.weight"], hl, tl, rl, relations, n)
        assert torch.equal(idx, a.indices())
        close(val, a.values(), 2e-6)
        with torch.no_grad():
            ref_emb = model.gat_embeddings()
        emb = O.gat_embeddings(p, cfg, idx, val, num, txt)
        close(emb, ref_emb, 2e-5)
        heads, tails = torch.arange(0, 9), torch.arange(5, 40)
        with torch.no_grad():
            close(O.calc_score(emb, heads, tails), model.calc_score(heads, tails), 5e-5)
        bh, bp, bn = (torch.from_numpy(rng.integers(0, n, 31)) for _ in range(3))
        br = torch.from_numpy(rng.integers(0, n_rel, 31))
        pg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in p.items()}
        for mode, ref_in, loss_fn in (
                ("pre_training", (bh, br, bp, bn), lambda e: O.triplet_loss(pg, e, cfg, bh, br, bp, bn)),
                ("fine_tuning", (bh, bp, bn), lambda e: O.prediction_loss(e, cfg, bh, bp, bn))):
            model.zero_grad()
            ref_loss = model(*ref_in, device="cpu", mode=mode)
            ref_loss.backward()
            for v in pg.values():
                v.grad = None
            loss = loss_fn(O.gat_embeddings(pg, cfg, idx, val, num, txt))
            loss.backward()
            assert abs(loss.item() - ref_loss.item()) <= 2e-5 * abs(ref_loss.item())
            for k, v in model.named_parameters():
                if k != "A_in" and v.grad is not None:
                    close(pg[k].grad, v.grad, 5e-4)


def test_reference_checkpoint_loads_into_the_drop_in(tmp_path):
    """a checkpoint written by the reference (``torch.save(model.state_dict())``,
    utils/model_utils.py:19-38, sparse ``A_in`` inside) loads into the drop-in class key for key, and what the drop-in
    saves loads back into the reference.  Host-side only: no kernel runs."""
    import literalkg_b200 as L
    import make_golden as G
    ref_model, _ = G.import_reference()
    cfg = O.OracleConfig(n_conv_layers=2, embed_dim=20, relation_dim=20, scale_gat_dim=12, conv_dim=8, num_lit_dim=2,
                         txt_lit_dim=6)
    n, n_rel = 30, 3
    h, t, r = G.make_kg(n, n_rel, 90, 7)
    idx, val, _ = G.ref_laplacian(h, t, r, n, "random-walk")
    a0 = torch.sparse_coo_tensor(torch.from_numpy(idx), torch.from_numpy(val), (n, n))
    num, txt = torch.zeros(n, 2), torch.zeros(n, 6)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ref_model.LiteralKG(G.namespace(cfg), n, n_rel, a0, num, txt)
        path = tmp_path / "ckpt.pth"
        torch.save(ref.state_dict(), path)
        ours = L.LiteralKG(G.namespace(cfg), n, n_rel, None, num, txt)
        missing, unexpected = ours.load_state_dict(torch.load(path), strict=True)
        assert not missing and not unexpected
        for (k1, v1), (k2, v2) in zip(ref.state_dict().items(), ours.state_dict().items()):
            assert k1 == k2 and v1.shape == v2.shape and v1.dtype == v2.dtype
            if v1.is_sparse:
                assert torch.equal(v1.coalesce().indices(), v2.coalesce().indices())
                assert torch.equal(v1.coalesce().values(), v2.coalesce().values())
            else:
                assert torch.equal(v1, v2)
        path2 = tmp_path / "ours.pth"
        torch.save(ours.state_dict(), path2)
        ref2 = ref_model.LiteralKG(G.namespace(cfg), n, n_rel, None, num, txt)
        ref2.load_state_dict(torch.load(path2))
        assert torch.equal(ref2.entity_embed.weight, ref.entity_embed.weight)
