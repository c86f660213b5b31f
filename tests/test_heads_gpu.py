"""CUDA path of the variant heads (csrc/mlp_head.cu through the drop-in classes) against the vectors of the unmodified
reference (tests/golden_heads/*.npz): mode 'mlp' in eval / train mode, BatchNorm running buffers, nn.BCELoss
gradients of every parameter (head and trunk), the TransE loss of model_bce.py and its gradients."""
import argparse

import pytest
import torch

from _golden import HEAD_CASES, HEADS_DIR, Golden

pytestmark = pytest.mark.gpu
REL = 1e-3


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module", params=HEAD_CASES)
def g(request):
    return Golden(request.param, HEADS_DIR)


def make(g):
    import literalkg_b200 as L
    from literalkg_b200 import model_bce
    args = argparse.Namespace(**{k: getattr(g.cfg, k) for k in g.cfg.__dataclass_fields__})
    a = torch.sparse_coo_tensor(g.t("att/idx"), g.t("att/val"), (g.n, g.n))
    bce = g.name.startswith("bce")
    m = (model_bce.LiteralKG if bce else L.LiteralKG)(args, g.n, g.n_rel, a, g.num_lit, g.txt_lit)
    if not bce:
        m.initialize_MLP()
    missing, unexpected = m.load_state_dict(g.sd, strict=False)
    assert missing == ["A_in"] and not unexpected           # identical state-dict keys (BatchNorm buffers included)
    return m.cuda()


def test_mlp_mode_eval_and_train(g):
    m = make(g).eval()
    h, t = g.t("mlp/h").cuda(), g.t("mlp/t").cuda()
    with torch.no_grad():
        assert rel(m.gat_embeddings(), g.z["final"]) < REL
        out = m(h, t, device="cuda", mode="mlp")
    assert tuple(out.shape) == (h.numel(), 1)
    assert rel(out, g.z["mlp/eval_out"]) < REL
    m.train()
    y = m(h, t, device="cuda", mode="mlp")
    assert rel(y, g.z["mlp/train_out"]) < REL
    for nm in ("norm1", "norm2"):
        assert rel(getattr(m, nm).running_mean, g.z[f"mlp/after/{nm}.running_mean"]) < 1e-4
        assert rel(getattr(m, nm).running_var, g.z[f"mlp/after/{nm}.running_var"]) < 1e-4
        assert int(getattr(m, nm).num_batches_tracked) == int(g.sd[f"{nm}.num_batches_tracked"]) + 1
    loss = torch.nn.BCELoss()(y.reshape(-1), g.t("mlp/labels").cuda())          # main_finetuning_BCE.py:88,117-124
    assert abs(loss.item() - float(g.z["mlp/bce_loss"])) < REL * abs(float(g.z["mlp/bce_loss"]))
    loss.backward()
    params = dict(m.named_parameters())
    ref = {k[len("grad_mlp/"):]: v for k, v in g.z.items() if k.startswith("grad_mlp/")}
    assert len(ref) > 20
    bad = {}
    for k, gr in ref.items():
        assert params[k].grad is not None, k
        e = rel(params[k].grad, gr)
        if not e < 2e-3:
            bad[k] = e
    assert not bad, bad


def test_transe_loss_and_gradients(g):
    if not g.name.startswith("bce"):
        pytest.skip("TransR variant: covered by test_backward_gpu.py")
    m = make(g).train()
    batch = tuple(g.t(k).cuda() for k in ("loss/h", "loss/r", "loss/pos", "loss/neg"))
    for mode, inp in (("pre_training", batch), ("fine_tuning", (batch[0], batch[2], batch[3]))):
        m.zero_grad(set_to_none=True)
        loss = m(*inp, device="cuda", mode=mode)
        assert abs(loss.item() - float(g.z[f"loss/{mode}"])) < REL * abs(float(g.z[f"loss/{mode}"])), mode
        loss.backward()
        params = dict(m.named_parameters())
        ref = {k[len(f"grad_{mode}/"):]: v for k, v in g.z.items() if k.startswith(f"grad_{mode}/")}
        bad = {k: rel(params[k].grad, gr) for k, gr in ref.items() if not rel(params[k].grad, gr) < 2e-3}
        assert not bad, (mode, bad)


def test_mlp_head_kernels_vs_torch_float64():
    """The head alone at the reference's widths (2 x 256 -> 128 -> 64 -> 1) and a realistic batch, against torch in
    float64: forward (train + eval statistics) and every gradient."""
    import literalkg_b200 as L
    from literalkg_b200.model import _MlpHeadFn
    torch.manual_seed(3)
    n, gdim, b = 5000, 256, 2048
    emb = (torch.randn(n, gdim) * 0.3).cuda().requires_grad_(True)
    h, t = torch.randint(0, n, (b,)).cuda(), torch.randint(0, n, (b,)).cuda()

    class Owner(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1, self.norm1 = torch.nn.Linear(2 * gdim, 128), torch.nn.BatchNorm1d(128)
            self.fc2, self.norm2 = torch.nn.Linear(128, 64), torch.nn.BatchNorm1d(64)
            self.fc3 = torch.nn.Linear(64, 1)

    own = Owner().cuda()
    ref = Owner().double().cuda()
    ref.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in own.state_dict().items()})
    for training in (True, False):
        own.train(training); ref.train(training)
        own.zero_grad(); ref.zero_grad(); emb.grad = None
        y = _MlpHeadFn.apply(own, emb, h, t, own.fc1.weight, own.fc1.bias, own.norm1.weight, own.norm1.bias,
                             own.fc2.weight, own.fc2.bias, own.norm2.weight, own.norm2.bias, own.fc3.weight, own.fc3.bias)
        e64 = emb.detach().double().requires_grad_(True)
        x = torch.cat([e64[h], e64[t]], 1)
        x = ref.norm1(torch.relu(ref.fc1(x)))
        x = ref.norm2(torch.relu(ref.fc2(x)))
        y64 = torch.sigmoid(ref.fc3(x))
        assert rel(y, y64) < 1e-5
        w = torch.randn(b, 1, device="cuda")
        (y * w).sum().backward()
        (y64 * w.double()).sum().backward()
        assert rel(emb.grad, e64.grad) < 1e-4
        for (k, p1), (_, p2) in zip(own.named_parameters(), ref.named_parameters()):
            assert rel(p1.grad, p2.grad) < 1e-4, (training, k)
        for nm in ("norm1", "norm2"):
            assert rel(getattr(own, nm).running_var, getattr(ref, nm).running_var) < 1e-5
