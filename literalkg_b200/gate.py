"""Literal gate modules -- drop-in for the reference's ``gate.py`` (GateMul :5-28, Gate :30-51).

Same constructor signatures, parameter names and shapes (state-dict keys ``g.weight``, ``g.bias``,
``gate_ent.weight``, ``gate_num_lit.weight`` / ``gate_txt_lit.weight`` / ``gate_lit.weight``,
``gate_bias``).  The forward is ONE fused kernel: the virtual concat ``[x_ent | literals]`` times the
interleaved weight ``[2*emb, K]`` with tanh / sigmoid / mix in the epilogue, instead of the
reference's cat + 4 GEMMs + ~6 elementwise passes.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


def _pair(g_weight, gate_weights, g_bias, gate_bias):
    """Row 2j = g.weight[j], row 2j+1 = [gate_ent | gate_lit...][j]; biases likewise."""
    z_weight = torch.cat(list(gate_weights), dim=1)
    w = torch.stack([g_weight, z_weight], dim=1).reshape(2 * g_weight.shape[0], g_weight.shape[1])
    b = torch.stack([g_bias, gate_bias], dim=1).reshape(-1)
    return w.contiguous(), b.contiguous()


class GateMul(nn.Module):
    """gate.py:5-28: out = (1 - z) * x_ent + z * tanh(g([x_ent, x_num, x_txt]))."""

    def __init__(self, emb_size, num_lit_size, txt_lit_size, gate_activation=torch.sigmoid):
        super().__init__()
        if gate_activation is not torch.sigmoid:
            raise NotImplementedError("the fused gate kernel implements the sigmoid gate of the reference")
        self.emb_size, self.num_lit_size, self.txt_lit_size = emb_size, num_lit_size, txt_lit_size
        self.gate_activation = gate_activation
        self.g = nn.Linear(emb_size + num_lit_size + txt_lit_size, emb_size)
        self.gate_ent = nn.Linear(emb_size, emb_size, bias=False)
        self.gate_num_lit = nn.Linear(num_lit_size, emb_size, bias=False)
        self.gate_txt_lit = nn.Linear(txt_lit_size, emb_size, bias=False)
        self.gate_bias = nn.Parameter(torch.zeros(emb_size))

    def packed(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if getattr(self, "_packed", None) is None or self._packed[0] != key:      # parameter derived: cached per version
            self._packed = (key, _pair(self.g.weight, (self.gate_ent.weight, self.gate_num_lit.weight,
                                                       self.gate_txt_lit.weight), self.g.bias, self.gate_bias))
        return self._packed[1]

    def pair(self):
        """The interleaved (g, z) weight and bias built under autograd (training path)."""
        return _pair(self.g.weight, (self.gate_ent.weight, self.gate_num_lit.weight, self.gate_txt_lit.weight),
                     self.g.bias, self.gate_bias)

    def forward(self, x_ent, x_lit_num, x_lit_txt, out=None, **planes):
        from .autograd import gate_apply
        return gate_apply(self, (x_ent, x_lit_num, x_lit_txt), out, **planes)


class Gate(nn.Module):
    """gate.py:30-51: single-literal variant."""

    def __init__(self, emb_size, lit_size, gate_activation=torch.sigmoid):
        super().__init__()
        if gate_activation is not torch.sigmoid:
            raise NotImplementedError("the fused gate kernel implements the sigmoid gate of the reference")
        self.emb_size, self.lit_size = emb_size, lit_size
        self.gate_activation = gate_activation
        self.g = nn.Linear(emb_size + lit_size, emb_size)
        self.gate_ent = nn.Linear(emb_size, emb_size, bias=False)
        self.gate_lit = nn.Linear(lit_size, emb_size, bias=False)
        self.gate_bias = nn.Parameter(torch.zeros(emb_size))

    def packed(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if getattr(self, "_packed", None) is None or self._packed[0] != key:
            self._packed = (key, _pair(self.g.weight, (self.gate_ent.weight, self.gate_lit.weight), self.g.bias,
                                       self.gate_bias))
        return self._packed[1]

    def pair(self):
        return _pair(self.g.weight, (self.gate_ent.weight, self.gate_lit.weight), self.g.bias, self.gate_bias)

    def forward(self, x_ent, x_lit, out=None, **planes):
        from .autograd import gate_apply
        return gate_apply(self, (x_ent, x_lit), out, **planes)
