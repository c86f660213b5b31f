"""Graph / literal tensors of the reference's ``DataLoader`` (dataloader.py:345-512), built on device.

Only the tensor-producing part of the reference loader is on the accelerated path (SURVEY.md 8(a) rows
a1-a3): ``h_list / t_list / r_list``, the ``relations`` order, ``n_entities / n_relations``, the initial
``A_in`` (sum of per-relation normalised adjacencies) and the dense literal tables -- plus the minibatch
samplers of the two training loops (section 8(f) rank 1), which run as one kernel per batch (``BatchSampler``).
Label files stay with the reference.
"""
from __future__ import annotations

import os
import pickle
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .graph import GraphPlan


def read_triples(path: str) -> np.ndarray:
    """``h r t`` per line, space separated; exact duplicate rows dropped keeping the first occurrence and
    the file order (``load_graph`` dataloader.py:186-190).  Returns int64 [E, 3] in (h, r, t) columns.

    The reference walks the frame with ``iterrows`` (~9 s per 200 k rows); here the file goes
    through pandas' C parser and the de-duplication is one sort of packed 64-bit keys."""
    try:
        import pandas as pd
        arr = pd.read_csv(path, sep=" ", header=None, names=["h", "r", "t"], dtype=np.int64,
                          engine="c").to_numpy(dtype=np.int64)
    except Exception:                       # ragged whitespace etc.: the tolerant reader
        arr = np.loadtxt(path, dtype=np.int64, ndmin=2)
    if arr.ndim != 2 or arr.shape[1] != 3:
        raise ValueError(f"{path}: expected 3 columns 'h r t'")
    if arr.shape[0] == 0:
        return arr
    if arr.min() >= 0:
        bits = [max(1, int(arr[:, c].max()).bit_length()) for c in range(3)]
        if sum(bits) <= 63:                 # one sortable key per row
            key = (arr[:, 0] << (bits[1] + bits[2])) | (arr[:, 1] << bits[2]) | arr[:, 2]
            _, first = np.unique(key, return_index=True)
            return arr[np.sort(first)]
    _, first = np.unique(arr, axis=0, return_index=True)
    return arr[np.sort(first)]


def relation_order(r: np.ndarray) -> List[int]:
    """Order of first appearance == ``laplacian_dict.keys()`` (dataloader.py:403,491), which main.py:150
    passes to ``update_att`` as ``relations``."""
    _, first = np.unique(r, return_index=True)
    return [int(x) for x in r[np.sort(first)]]


def read_numeric_literals(paths: Sequence[str], n_entities: Optional[int], numeric_dim: int):
    """``load_attributes`` + ``embed_num_literal`` (dataloader.py:111-137, 426-431): file i fills column i
    with (v + 1) / max(v); lines without a tab (the leading count) are skipped; a later file resets the
    whole row of an entity it mentions.  Returns (table float32 [n, numeric_dim], max entity id)."""
    rows: Dict[int, np.ndarray] = {}
    for col, path in enumerate(paths):
        ids, vals = [], []
        with open(path) as fh:
            for line in fh:
                parts = line.split("\t")
                if len(parts) > 1:
                    ids.append(int(parts[0]))
                    vals.append(float(parts[1]))
        vmax = max([0.0] + vals)
        latest = dict(zip(ids, vals))                      # a repeated id keeps its last value (dict semantics)
        for ent, v in latest.items():
            row = np.zeros(numeric_dim)
            if vmax != 0:
                row[col] = (v + 1) / vmax
            rows[ent] = row
    max_id = max(rows) if rows else -1
    n = max(n_entities or 0, max_id + 1)
    table = np.zeros((n, numeric_dim), dtype=np.float32)
    if rows:
        idx = np.fromiter(rows.keys(), dtype=np.int64)
        table[idx] = np.stack(list(rows.values())).astype(np.float32)
    return table, max_id


NUMERIC_FILES = ("age_dict.txt", "weight_dict.txt")                                     # dataloader.py:29-30
TEXT_FILES = ("cc_dict.pickle", "disease_dict.pickle", "memo_dict.pickle", "prescription_dict.pickle",
              "treatment_dict.pickle")                                                   # dataloader.py:31-32


def read_text_literals(paths: Sequence[str]) -> Dict[int, np.ndarray]:
    """The text-literal pickles (dataloader.py:139-152): each holds ``{entity id: vector}``; a later file overrides an
    earlier one for the same entity (dict assignment in file order)."""
    text: Dict[int, np.ndarray] = {}
    for p in paths:
        with open(p, "rb") as fh:
            for k, v in pickle.load(fh).items():
                text[int(k)] = np.asarray(v, dtype=np.float32)
    return text


def read_data_dir(data_dir: str, kg_file: str = "pre_training_train.txt", numeric_dim: int = 2, text_dim: int = 300,
                  numeric_files=NUMERIC_FILES, text_files=TEXT_FILES, use_num_lit: bool = True,
                  use_txt_lit: bool = True):
    """Host side of ``DataLoader.__init__`` for the tensors on the path (dataloader.py:111-152, 186-190, 405-438):
    -> (triples int64 [E, 3] as (h, r, t), numeric table or None, text table or None), both tables [n, dim] float32
    with n = max(KG ids, literal ids) + 1.  An entity that appears in a text pickle has its numeric row zeroed and an
    entity of a numeric file gets a zero text row unless a pickle sets it (load_attributes, :133-150)."""
    trip = read_triples(os.path.join(data_dir, kg_file))
    n_kg = int(max(trip[:, 0].max(), trip[:, 2].max()) + 1)
    num_paths = [os.path.join(data_dir, f) for f in numeric_files if os.path.exists(os.path.join(data_dir, f))]
    txt_paths = [os.path.join(data_dir, f) for f in text_files if os.path.exists(os.path.join(data_dir, f))]
    text = read_text_literals(txt_paths) if use_txt_lit else {}
    num_table = text_table = None
    num_max = -1
    if num_paths and (use_num_lit or use_txt_lit):
        table, num_max = read_numeric_literals(num_paths, n_kg, numeric_dim)
        if use_num_lit:
            if text:                                                  # dataloader.py:147-150
                n_all = max(table.shape[0], max(text) + 1)
                if n_all > table.shape[0]:
                    table = np.concatenate([table, np.zeros((n_all - table.shape[0], numeric_dim), np.float32)])
                table[np.fromiter(text.keys(), dtype=np.int64)] = 0
            num_table = table
    if use_txt_lit and (text or num_max >= 0):
        n_txt = max([n_kg, num_max + 1] + ([max(text) + 1] if text else []))
        if num_table is not None:
            n_txt = max(n_txt, num_table.shape[0])
        tt = np.zeros((n_txt, text_dim), dtype=np.float32)
        for k, v in text.items():
            tt[k] = v
        text_table = tt
        if num_table is not None and num_table.shape[0] < n_txt:
            num_table = np.concatenate([num_table, np.zeros((n_txt - num_table.shape[0], numeric_dim), np.float32)])
    return trip, num_table, text_table


class KGTensors:
    """Device tensors with the attribute names the reference's training loop reads from its DataLoader
    (main.py:42-43,147-151): ``A_in``, ``h_list``, ``t_list``, ``r_list``, ``laplacian_dict`` (keys only),
    ``num_embedding_table``, ``text_embedding_table``, ``n_entities``, ``n_relations``."""

    def __init__(self, h, t, r, n_entities: Optional[int] = None, laplacian_type: str = "random-walk",
                 device="cuda", num_table=None, text_table=None, min_entities: int = 0):
        h = torch.as_tensor(h, dtype=torch.int64)
        t = torch.as_tensor(t, dtype=torch.int64)
        r = torch.as_tensor(r, dtype=torch.int64)
        self.device = torch.device(device)
        self.n_relations = len(torch.unique(r))            # dataloader.py:374 (ids must be 0..R-1)
        n = int(max(h.max().item() + 1, t.max().item() + 1, min_entities))   # dataloader.py:405-418
        for tab in (num_table, text_table):
            if tab is not None:
                n = max(n, tab.shape[0])
        self.n_entities = n if n_entities is None else int(n_entities)
        self.h_list, self.t_list, self.r_list = h.to(self.device), t.to(self.device), r.to(self.device)
        self.relations = relation_order(r.cpu().numpy())
        self.laplacian_dict = {rel: None for rel in self.relations}   # the training loop only uses .keys()
        self.laplacian_type = laplacian_type
        self.plan = GraphPlan(self.h_list, self.t_list, self.r_list, self.n_entities, self.n_relations)
        vals = self.plan.laplacian(laplacian_type)                          # dataloader.py:449-495
        if laplacian_type == "symmetric":
            # scipy's diags() drops zero diagonal entries structurally (dataloader.py:466-470), so a
            # triple whose tail has no outgoing edge under that relation leaves no entry at all
            live = vals != 0
            self.A_in = torch.sparse_coo_tensor(self.plan.indices[:, live], vals[live],
                                                (self.n_entities, self.n_entities), is_coalesced=True)
        else:
            self.A_in = self.plan.sparse(vals)
        self.num_embedding_table = None if num_table is None else self._table(num_table)
        self.text_embedding_table = None if text_table is None else self._table(text_table)

    def _table(self, tab) -> torch.Tensor:
        tab = torch.as_tensor(tab, dtype=torch.float32)
        if tab.shape[0] < self.n_entities:
            tab = torch.cat([tab, tab.new_zeros(self.n_entities - tab.shape[0], tab.shape[1])])
        return tab.to(self.device).contiguous()

    @classmethod
    def from_dir(cls, data_dir: str, kg_file: str = "pre_training_train.txt", numeric_dim: int = 2,
                 text_dim: int = 300, numeric_files=NUMERIC_FILES, text_files=TEXT_FILES,
                 use_num_lit: bool = True, use_txt_lit: bool = True, **kw) -> "KGTensors":
        """Reads a reference data directory (dataloader.py:24-32).  Missing literal files are skipped
        (the Drive-hosted pickles are not bundled with the reference)."""
        trip, num_table, text_table = read_data_dir(data_dir, kg_file, numeric_dim, text_dim, numeric_files, text_files,
                                                    use_num_lit, use_txt_lit)
        return cls(trip[:, 0], trip[:, 2], trip[:, 1], num_table=num_table, text_table=text_table, **kw)


class BatchSampler:
    """Device-side ``generate_kg_batch`` / ``generate_prediction_batch`` (dataloader.py:221-318).

    ``plan``: GraphPlan of the triples the positives come from (its att arrays are the per-head (tail, relation)
    lists of ``train_kg_dict`` sorted by (relation, tail)); for fine-tuning, a plan of the (head, tail) pairs with
    relation 0 (``head_dict``).  ``candidates``: the tails negatives are drawn from (``training_tails`` /
    ``prediction_tail_ids``; repeats weigh a tail like they do in the reference's list).  Heads are drawn on the
    host side of the binding exactly as upstream: without replacement when the batch fits the head list
    (``random.sample``), with replacement otherwise; everything per head is one kernel (csrc/sample.cu)."""

    def __init__(self, plan: GraphPlan, candidates, neg_rate: int, use_relation: bool, seed: int = 2022,
                 max_tries: int = 10_000):
        self.plan, self.neg_rate, self.use_relation = plan, int(neg_rate), bool(use_relation)
        if not self.use_relation and plan.n_relations > 1:
            # the rejection test is a binary search by tail in a row sorted by (relation, tail)
            raise ValueError("use_relation=False needs a single-relation plan (build the (head, tail) plan with "
                             "relation 0, like head_dict of the reference)")
        self.device = plan.device
        self.candidates = torch.as_tensor(candidates, dtype=torch.int64).to(self.device).contiguous()
        deg = plan.att_rowptr[1:] - plan.att_rowptr[:-1]
        self.exist_heads = torch.nonzero(deg > 0).reshape(-1)            # list(kg_dict.keys())
        self.gen = torch.Generator(device=self.device).manual_seed(int(seed))
        self.seed, self.calls, self.max_tries = int(seed), 0, int(max_tries)
        self.n_failed = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.check_every = 256          # calls between two looks at n_failed (one host sync each)

    def check(self) -> None:
        """Raises when a draw found no admissible negative within ``max_tries`` or a head had no triple (the reference
        loops forever / raises KeyError there); ``sample`` calls it every ``check_every`` batches, callers at epoch end."""
        n = int(self.n_failed.item())
        if n:
            raise RuntimeError(f"BatchSampler: {n} draws failed (no admissible negative tail within {self.max_tries} tries, "
                               "or a head without triples)")

    def sample(self, batch_size: int):
        """-> (head, relation or None, pos_tail, neg_tail), int64 [n * neg_rate] with n = batch_size // neg_rate
        (the division of dataloader.py:224 / :287)."""
        n = int(batch_size / self.neg_rate)
        ne = self.exist_heads.numel()
        if n <= ne:
            pick = torch.randperm(ne, device=self.device, generator=self.gen)[:n]
        else:
            pick = torch.randint(0, ne, (n,), device=self.device, generator=self.gen)
        heads = self.exist_heads[pick].contiguous()
        m = n * self.neg_rate
        i64 = dict(dtype=torch.int64, device=self.device)
        out_h, out_pos, out_neg = torch.empty(m, **i64), torch.empty(m, **i64), torch.empty(m, **i64)
        out_r = torch.empty(m, **i64) if self.use_relation else None
        self.calls += 1
        if self.check_every and self.calls % self.check_every == 0:
            self.check()
        p = self.plan
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().lkg_sample_batch(
                p.att_rowptr.data_ptr(), p.att_tail.data_ptr(), p.att_rel.data_ptr(), heads.data_ptr(), n,
                self.candidates.data_ptr(), self.candidates.numel(), self.neg_rate, int(self.use_relation),
                (self.seed * 0x9E3779B1 + self.calls) & 0xFFFFFFFFFFFFFFFF, self.max_tries, out_h.data_ptr(),
                _lib.ptr(out_r), out_pos.data_ptr(), out_neg.data_ptr(), self.n_failed.data_ptr(), _lib.stream()))
        return out_h, out_r, out_pos, out_neg
