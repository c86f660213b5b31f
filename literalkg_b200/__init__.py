"""literalkg_b200 -- B200-native (sm_100a) message-passing + scoring path of LiteralKG.

Drop-in surface (same names / signatures / state-dict keys as the reference's model.py, gate.py):
    LiteralKG, Aggregator, Gate, GateMul
plus the device-side graph tensors (KGTensors, GraphPlan) and the functional kernel wrappers (ops).
The compute lives in liblkg.so (include/lkg.h); importing this package never falls back to PyTorch math.
"""
from .gate import Gate, GateMul
from .graph import GraphPlan
from .model import Aggregator, LiteralKG
from .dataloader import BatchSampler, KGTensors
from . import ops, synthetic

__all__ = ["LiteralKG", "Aggregator", "Gate", "GateMul", "GraphPlan", "KGTensors", "BatchSampler", "ops", "synthetic"]
