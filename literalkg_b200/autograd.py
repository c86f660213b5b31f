"""Glue between nn.Module parameters and the forward kernels.

The forward of every stage runs in the hand-written kernels.  Gradient support is attached by
``literalkg_b200.model`` at the level of the whole ``gat_embeddings`` pass (``_GatEmbeddingsFn``: one autograd
node whose backward is csrc/backward.cu + csrc/xty_tc.cu) and of the loss heads (``_BprLossFn``,
``_TransRLossFn``: csrc/loss.cu); the standalone ``Gate`` / ``GateMul`` / ``Aggregator`` modules are forward
(inference) modules.
"""
from __future__ import annotations

import torch

from . import ops


def require_no_grad(module, tensors, what: str) -> None:
    """The standalone modules run the forward kernels only.  The reference's modules are differentiable, so a caller
    that trains through one of them must not get a silently detached result: fail loudly instead."""
    if torch.is_grad_enabled() and (any(p.requires_grad for p in module.parameters())
                                    or any(t is not None and t.requires_grad for t in tensors)):
        raise RuntimeError(f"{what} is a forward (inference) module of literalkg_b200: call it under torch.no_grad(); "
                           "gradients of the path flow through LiteralKG.gat_embeddings() (one fused autograd node)")


def gate_apply(module, inputs, out=None, out_planes=None, ent_planes=None, lit_planes=None, packed=None, gz_out=None):
    """Fused literal gate forward for ``Gate`` / ``GateMul`` (gate.py:22-28, 45-51).  The A operand is the K
    concatenation (entity | literals); ``ent_planes`` / ``lit_planes`` let the caller reuse cached fp16 planes."""
    x_ent = inputs[0]
    require_no_grad(module, inputs, type(module).__name__)
    with torch.no_grad():
        w_pair, b_pair = module.packed() if packed is None else packed
        if ent_planes is None:
            ent_planes = ops.split_planes(x_ent.detach())
        if lit_planes is None:
            lits = [x.detach().float() for x in inputs[1:]]
            lit_planes = ops.split_planes(lits[0] if len(lits) == 1 else torch.cat(lits, dim=1))
        return ops.gate([ent_planes, lit_planes], w_pair, b_pair, x_ent.detach(), out, out_planes, gz_out=gz_out)
