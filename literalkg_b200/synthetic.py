"""Synthetic knowledge graphs and literal tables of the shapes named in BASELINE.json.

The Drive-hosted literal pickles and most KG blobs of the reference are not available offline, so
benchmarks and large-scale tests run on generated data: power-law out-degrees (Zipf exponent 1 over a
random permutation of entity ids, clipped), skewed tails, uniform relations with every id present,
de-duplicated on (h, r, t), and a small fraction of (h, t) pairs repeated under a second relation to
exercise the duplicate-merge path of ``update_att``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

SEED = 2022  # the reference's default seed (argument.py:7)


@dataclass
class SyntheticKG:
    h: np.ndarray          # int64 [E] file order
    t: np.ndarray
    r: np.ndarray
    n_entities: int
    n_relations: int

    @property
    def n_edges(self) -> int:
        return int(self.h.shape[0])


def make_kg(n_entities: int, n_edges: int, n_relations: int, seed: int = SEED, max_out_degree: int = 4096,
            dup_fraction: float = 2e-4, tail_skew: float = 2.0) -> SyntheticKG:
    rng = np.random.default_rng(seed)
    n = int(n_entities)
    # out-degrees ~ 1/rank, clipped, rescaled to the requested edge count
    w = 1.0 / np.arange(1, n + 1, dtype=np.float64)
    deg = w / w.sum() * n_edges
    for _ in range(8):                                   # redistribute the clipped mass
        over = deg > max_out_degree
        excess = (deg[over] - max_out_degree).sum()
        if excess < 1:
            break
        deg[over] = max_out_degree
        deg[~over] += excess * deg[~over] / deg[~over].sum()
    deg = np.floor(deg + rng.random(n)).astype(np.int64)
    deg = np.minimum(deg, max_out_degree)
    perm = rng.permutation(n)
    h = np.repeat(perm, deg)
    e = h.shape[0]
    # skewed tails: rank = N * u^tail_skew over an independent permutation
    perm_t = rng.permutation(n)
    t = perm_t[np.minimum((n * rng.random(e) ** tail_skew).astype(np.int64), n - 1)]
    r = rng.integers(0, n_relations, size=e, dtype=np.int64)
    r[:n_relations] = np.arange(n_relations)             # every relation id present
    # (h, t) pairs repeated under a second relation
    n_dup = int(e * dup_fraction)
    if n_dup:
        pick = rng.choice(e, size=n_dup, replace=False)
        h = np.concatenate([h, h[pick]])
        t = np.concatenate([t, t[pick]])
        r = np.concatenate([r, (r[pick] + 1) % n_relations])
    # de-duplicate on (h, r, t), then shuffle into an arbitrary file order
    bits_n = max(1, int(np.ceil(np.log2(max(n, 2)))))
    bits_r = max(1, int(np.ceil(np.log2(max(n_relations, 2)))))
    assert 2 * bits_n + bits_r <= 63
    key = (h << (bits_n + bits_r)) | (r << bits_n) | t
    key = np.unique(key)
    rng.shuffle(key)
    t = key & ((1 << bits_n) - 1)
    r = (key >> bits_n) & ((1 << bits_r) - 1)
    h = key >> (bits_n + bits_r)
    return SyntheticKG(h.astype(np.int64), t.astype(np.int64), r.astype(np.int64), n, int(n_relations))


def make_literals(n_entities: int, num_dim: int = 2, txt_dim: int = 300, seed: int = SEED,
                  device="cpu", num_fraction: float = 0.14, txt_fraction: float = 0.15
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Numeric table: one non-zero column per populated row with a value in (0, 1] (the reference's
    (v+1)/max normalisation); text table: ~N(0, 0.1) on ``txt_fraction`` of the rows, zero elsewhere."""
    g = torch.Generator(device=device).manual_seed(seed)
    n = int(n_entities)
    num = torch.zeros((n, num_dim), dtype=torch.float32, device=device)
    rows = torch.rand(n, generator=g, device=device) < num_fraction
    cols = torch.randint(0, num_dim, (n,), generator=g, device=device)
    vals = torch.rand(n, generator=g, device=device) * 0.95 + 0.05
    num[torch.arange(n, device=device)[rows], cols[rows]] = vals[rows]
    txt = torch.randn((n, txt_dim), generator=g, device=device) * 0.1
    txt *= (torch.rand(n, 1, generator=g, device=device) < txt_fraction).float()
    return num, txt
