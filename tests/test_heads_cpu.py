"""The oracle's restatement of the variant heads (`mlp` mode, model.py:499-519 /
model_bce.py:423-436; TransE loss, model_bce.py:329-368) against vectors produced by the unmodified reference
(oracle/make_golden.py --heads -> tests/golden_heads/*.npz)."""
import pytest
import torch

import literalkg_oracle as O
from _golden import HEAD_CASES, HEADS_DIR, Golden


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module", params=HEAD_CASES)
def g(request):
    return Golden(request.param, HEADS_DIR)


def test_cases_present():
    assert len(HEAD_CASES) >= 3 and any(c.startswith("bce") for c in HEAD_CASES) and any(c.startswith("mlp") for c in HEAD_CASES)


def test_trunk_and_mlp_head(g):
    p = {k: v.clone() for k, v in g.sd.items()}
    emb = O.gat_embeddings(p, g.cfg, g.t("att/idx"), g.t("att/val"), g.num_lit, g.txt_lit)
    assert rel(emb, g.z["final"]) < 2e-4
    h, t = g.t("mlp/h"), g.t("mlp/t")
    assert rel(O.mlp_head(p, emb, h, t, training=False), g.z["mlp/eval_out"]) < 1e-5
    y = O.mlp_head(p, emb, h, t, training=True)
    assert rel(y, g.z["mlp/train_out"]) < 1e-5
    for nm in ("norm1", "norm2"):                                   # running buffers after one training batch
        assert rel(p[f"{nm}.running_mean"], g.z[f"mlp/after/{nm}.running_mean"]) < 1e-5
        assert rel(p[f"{nm}.running_var"], g.z[f"mlp/after/{nm}.running_var"]) < 1e-5
    loss = torch.nn.functional.binary_cross_entropy(y.reshape(-1), g.t("mlp/labels"))
    assert abs(loss.item() - float(g.z["mlp/bce_loss"])) < 1e-5 * abs(float(g.z["mlp/bce_loss"]))


def test_mlp_head_gradients(g):
    """autograd through the oracle (fp64) reproduces the reference's BCE gradients of the head parameters."""
    p = {k: (v.double().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in g.sd.items()}
    emb = O.gat_embeddings(p, g.cfg, g.t("att/idx"), g.t("att/val"), g.num_lit.double(), g.txt_lit.double())
    y = O.mlp_head(p, emb, g.t("mlp/h"), g.t("mlp/t"), training=True, update=False)
    torch.nn.functional.binary_cross_entropy(y.reshape(-1), g.t("mlp/labels").double()).backward()
    for k in ("fc1.weight", "fc2.weight", "fc3.weight", "norm1.weight", "norm2.bias", "linear_gat.weight"):
        assert rel(p[k].grad, g.z["grad_mlp/" + k]) < 2e-3, k


def test_transe_loss(g):
    if "loss/pre_training" not in g.z:
        pytest.skip("model.py case: TransR, covered by the main golden vectors")
    p = g.sd
    emb = torch.from_numpy(g.z["final"])
    loss = O.transe_loss(p, emb, g.cfg, g.t("loss/h"), g.t("loss/r"), g.t("loss/pos"), g.t("loss/neg"))
    # the golden loss was taken in train() mode with dropout 0: same trunk output
    assert abs(loss.item() - float(g.z["loss/pre_training"])) < 2e-4 * abs(float(g.z["loss/pre_training"]))
    assert "gat_trans_M" not in p                                   # the BCE variant has no TransR projection
