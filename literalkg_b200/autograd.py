"""Glue between nn.Module parameters and the forward kernels.

The forward of every stage runs in the hand-written kernels.  Gradient support is attached by
``literalkg_b200.model`` at the level of the whole ``gat_embeddings`` pass (see ``GatEmbeddingsFn``);
the standalone ``Gate`` / ``GateMul`` / ``Aggregator`` modules are forward (inference) modules.
"""
from __future__ import annotations

import torch

from . import ops


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def gate_apply(module, inputs, out=None):
    """Fused literal gate forward for ``Gate`` / ``GateMul`` (gate.py:22-28, 45-51)."""
    x_ent = inputs[0]
    with torch.no_grad():
        w_pair, b_pair = module.packed()
        return ops.gate([x.detach() for x in inputs], w_pair, b_pair, x_ent.detach(), out)
