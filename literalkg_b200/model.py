"""B200-native drop-in for the reference's ``model.py`` (``Aggregator`` :12-164, ``LiteralKG`` :167-532).

Same class names, constructor signatures, ``forward(*input, device=, mode=)`` modes, public helpers,
attribute names and state-dict keys, so the reference's training / evaluation
scripts run unchanged.  What differs is the inside: every stage of the message-passing + scoring path
runs in the sm_100a kernels of liblkg.so (include/lkg.h); there is no PyTorch-math or CPU fallback.

Stage map (reference -> here)
    gate_embeddings   model.py:265-279 -> ops.gate        (one fused GEMM + tanh/sigmoid/mix epilogue)
    Aggregator        model.py:101-164 -> ops.aggregate   (CSR SpMM + folded combine + LayerNorm + L2 norm)
    cat + linear_gat  model.py:308-311 -> ops.linear      (the concat buffer is written in place by the layers)
    update_attention  model.py:444-471 -> ops.attn_update (logits + duplicate merge + row softmax, on device)
    calc_score        model.py:473-486 -> ops.score
    predict_links     model.py:488-491 -> ops.predict
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from .gate import Gate, GateMul
from .graph import GraphPlan


def _param_key(module: nn.Module):
    """Identity + in-place version of every parameter: cache key for tensors derived from the parameters only
    (folded matrices, stacked gate weights).  Optimizer steps and load_state_dict bump the versions."""
    return tuple((p.data_ptr(), p._version) for p in module.parameters())


_EYE = {}


def _identity(n: int, device) -> torch.Tensor:
    key = (n, str(device))
    if key not in _EYE:
        _EYE[key] = torch.eye(n, dtype=torch.float32, device=device)
    return _EYE[key]


def _L2_loss_mean(x):
    return torch.mean(torch.sum(torch.pow(x, 2), dim=1, keepdim=False) / 2.)


class Aggregator(nn.Module):
    """One propagation layer (model.py:12-164): gcn / graphsage / bi-interaction, optional GCNII-style
    residual connection, LeakyReLU, LayerNorm, message dropout."""

    def __init__(self, in_dim, out_dim, dropout, aggregator_type, use_residual=False, args=None):
        super().__init__()
        self.in_dim, self.out_dim, self.dropout = in_dim, out_dim, dropout
        self.aggregator_type = aggregator_type
        self.use_residual = use_residual
        self.weight = nn.Parameter(torch.empty(in_dim, in_dim))
        if use_residual:
            self.linear_h0 = nn.Linear(args.embed_dim, in_dim)
            nn.init.xavier_uniform_(self.linear_h0.weight)
        self.reset_parameters()
        self.message_dropout = nn.Dropout(dropout)
        self.activation = nn.LeakyReLU()
        self.layer_normalize = nn.LayerNorm(out_dim)
        if aggregator_type == 'gcn':
            self.linear = nn.Linear(in_dim, out_dim)
            nn.init.xavier_uniform_(self.linear.weight)
        elif aggregator_type == 'graphsage':
            if use_residual:
                self.linear_h = nn.Linear(in_dim * 2, in_dim)
                nn.init.xavier_uniform_(self.linear_h.weight)
                self.linear = nn.Linear(in_dim, out_dim)
            else:
                self.linear = nn.Linear(in_dim * 2, out_dim)
            nn.init.xavier_uniform_(self.linear.weight)
        elif aggregator_type == 'bi-interaction':
            self.linear1 = nn.Linear(in_dim, out_dim)
            self.linear2 = nn.Linear(in_dim, out_dim)
            nn.init.xavier_uniform_(self.linear1.weight)
            nn.init.xavier_uniform_(self.linear2.weight)
        else:
            # 'gin' exists in the reference but is outside the accelerated path
            raise NotImplementedError(aggregator_type)
        self._plan_cache: Optional[Tuple[int, int, GraphPlan, torch.Tensor]] = None
        self._fold_cache = None

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.out_dim)
        self.weight.data.uniform_(-stdv, stdv)

    # ---- parameter folding -------------------------------------------------
    def folded(self, lamda: float, alpha: float, l: int) -> Dict[str, Optional[torch.Tensor]]:
        """linear(residual(hi)) == hi @ P + h0 @ Q + c with  M = (1-b) + b*W,  b = ln(lamda/l + 1)
        (model.py:90-99): P = (1-a) M W_lin^T, Q = a W_h0^T M W_lin^T, c = a b_h0 M W_lin^T + b_lin.
        Formed in float64 from the live parameters, returned in fp32.  Keys: pa, pb, p2 ([d_in, d_out]),
        q1, q2 ([embed_dim, d_out] or None), c1, c2 ([d_out])."""
        key = (float(lamda), float(alpha), int(l), _param_key(self))
        if self._fold_cache is not None and self._fold_cache[0] == key:
            return self._fold_cache[1]
        dd = torch.float64
        t = self.aggregator_type
        d = self.in_dim

        def lin(m):
            return m.weight.to(dd), m.bias.to(dd)

        out: Dict[str, Optional[torch.Tensor]] = dict(pa=None, pb=None, p2=None, q1=None, q2=None, c1=None, c2=None)
        if self.use_residual:
            beta = math.log(lamda / l + 1)
            m_id = (1 - beta) + beta * self.weight.to(dd)
            wh0, bh0 = lin(self.linear_h0)

            def fold(w_lin, b_lin):
                mw = m_id @ w_lin.t()                                  # [d, C]
                return (1 - alpha) * mw, alpha * (wh0.t() @ mw), alpha * (bh0 @ mw) + b_lin
        else:
            def fold(w_lin, b_lin):
                return w_lin.t(), None, b_lin

        if t == 'gcn':
            p, q, c = fold(*lin(self.linear))
            out.update(pa=p, pb=p, q1=q, c1=c)
        elif t == 'bi-interaction':
            p, q, c = fold(*lin(self.linear1))
            p2, q2, c2 = fold(*lin(self.linear2))
            out.update(pa=p, pb=p, p2=p2, q1=q, q2=q2, c1=c, c2=c2)
        else:  # graphsage
            if self.use_residual:
                wh, bh = lin(self.linear_h)
                p, q, c = fold(*lin(self.linear))
                out.update(pa=wh[:, :d].t() @ p, pb=wh[:, d:].t() @ p, q1=q, c1=bh @ p + c)
            else:
                w, b = lin(self.linear)
                out.update(pa=w[:, :d].t(), pb=w[:, d:].t(), c1=b)
        res = {}
        for k, v in out.items():
            res[k] = None if v is None else v.float().contiguous()
        if res["pa"] is not None and out["pa"] is out["pb"]:
            res["pa"] = res["pb"]                                      # keep identity: "sum" mode of the kernel
        self._fold_cache = (key, res)       # derived from the parameters only: reused until one of them changes
        return res

    def _drop_mask(self, n: int, device) -> Optional[torch.Tensor]:
        if self.training and self.dropout > 0:
            keep = 1.0 - self.dropout
            return (torch.rand((n, self.out_dim), device=device) < keep).float() / keep
        return None

    def run(self, plan: GraphPlan, a_values: torch.Tensor, ego: torch.Tensor, f: Dict[str, Optional[torch.Tensor]],
            r1: Optional[torch.Tensor], r2: Optional[torch.Tensor], x_out: torch.Tensor,
            xn_out: Optional[torch.Tensor], fold_ego: bool = False, xn_planes=None, rows=None, z=None
            ) -> torch.Tensor:
        """``rows`` = (begin, end): the head rows this rank owns; r1 / r2 / xn_out / xn_planes then hold those rows
        only, ``ego`` and ``x_out`` stay indexed by the global row."""
        pa = None if fold_ego else f["pa"]
        rb, re = (0, ego.shape[0]) if rows is None else rows
        pb, p2 = f["pb"], f["p2"]
        if z is not None:
            # the neighbour sum term arrives pre-projected (z = ego @ Pb):
            #   bi-interaction: the wide kernel gathers z next to the ego rows and only combines the product term;
            #   gcn / graphsage: nothing but z is gathered -- the layer runs on the d_out-wide z table
            assert pa is None
            if p2 is None:
                ego, z, pb = z, None, _identity(self.out_dim, ego.device)
            else:
                pb = None
        return ops.aggregate(plan, a_values, ego, self.out_dim, pa, pb, p2, r1, r2,
                             self.layer_normalize.weight, self.layer_normalize.bias,
                             self._drop_mask(re - rb, ego.device), x_out, xn_out, xn_planes, local_row_base=rb, z=z)

    def forward(self, ego_embeddings, A_in, all_layers, lamda, alpha, l):
        """Reference signature (model.py:101): ``A_in`` is a sparse COO tensor, ``all_layers[0]`` the gate
        output used by the residual connection, ``l`` the 1-based layer index."""
        with torch.no_grad():
            ego = _lib.f32c(ego_embeddings)
            _lib.require_cuda(ego, "ego_embeddings")
            plan, vals = self._plan_for(A_in)
            f = self.folded(lamda, alpha, l)
            r1, r2 = f["c1"], f["c2"]
            if self.use_residual:
                h0 = _lib.f32c(all_layers[0])
                qs = [f["q1"]] + ([f["q2"]] if f["q2"] is not None else [])
                cs = [f["c1"]] + ([f["c2"]] if f["c2"] is not None else [])
                r = ops.linear([ops.split_planes(h0)], torch.cat(qs, dim=1).t().contiguous(), torch.cat(cs))
                r1 = r[:, :self.out_dim]
                r2 = r[:, self.out_dim:] if f["q2"] is not None else None
            x = torch.empty((ego.shape[0], self.out_dim), dtype=torch.float32, device=ego.device)
            return self.run(plan, vals, ego, f, r1, r2, x, None)

    def _plan_for(self, A_in: torch.Tensor) -> Tuple[GraphPlan, torch.Tensor]:
        key = (A_in._values().data_ptr(), A_in._nnz())
        if self._plan_cache is None or self._plan_cache[:2] != key:
            plan = GraphPlan.from_coo(A_in._indices(), A_in.shape[0])
            self._plan_cache = (*key, plan, plan.import_values(A_in._values()))
        return self._plan_cache[2], self._plan_cache[3]


class LiteralKG(nn.Module):
    """model.py:167-532."""

    def __init__(self, args, n_entities, n_relations, A_in=None, numerical_literals=None, text_literals=None):
        super().__init__()
        self.use_pretrain = args.use_pretrain
        self.args = args
        self.device = args.device
        self.n_entities, self.n_relations = n_entities, n_relations
        self.embed_dim, self.relation_dim = args.embed_dim, args.relation_dim
        self.scale_gat_dim = args.scale_gat_dim
        self.use_residual, self.alpha, self.lamda = args.use_residual, args.alpha, args.lamda
        self.aggregation_type = args.aggregation_type
        self.n_layers = args.n_conv_layers
        self.conv_dim_list = [args.embed_dim] + [args.conv_dim] * self.n_layers
        self.total_conv_dim = sum(self.conv_dim_list)
        self.mess_dropout = [args.mess_dropout] * self.n_layers
        self.kg_l2loss_lambda = args.kg_l2loss_lambda
        self.prediction_l2loss_lambda = args.fine_tuning_l2loss_lambda
        self.pre_training_neg_rate = args.pre_training_neg_rate
        self.fine_tuning_neg_rate = args.fine_tuning_neg_rate
        self.n_num_lit, self.n_txt_lit = args.num_lit_dim, args.txt_lit_dim

        self.entity_embed = nn.Embedding(n_entities, self.embed_dim)
        self.relation_embed = nn.Embedding(n_relations, self.relation_dim)
        if self.scale_gat_dim is not None:
            self.linear_gat = nn.Linear(self.total_conv_dim, self.scale_gat_dim)
            self.gat_activation = nn.LeakyReLU()
            nn.init.xavier_uniform_(self.linear_gat.weight)
            self.gat_trans_M = nn.Parameter(torch.empty(n_relations, self.scale_gat_dim, self.relation_dim))
        else:
            self.gat_trans_M = nn.Parameter(torch.empty(n_relations, self.total_conv_dim, self.relation_dim))
        nn.init.xavier_uniform_(self.entity_embed.weight)
        nn.init.xavier_uniform_(self.relation_embed.weight)
        nn.init.xavier_uniform_(self.gat_trans_M)

        self.aggregator_layers = nn.ModuleList()
        # plain attributes, not buffers, exactly like the reference (model.py:241-242)
        self.numerical_literals_embed = numerical_literals
        self.text_literals_embed = text_literals
        if args.use_num_lit and args.use_txt_lit:
            self.emb_mul_lit = GateMul(self.embed_dim, self.n_num_lit, self.n_txt_lit)
        elif args.use_num_lit:
            self.emb_num_lit = Gate(self.embed_dim, self.n_num_lit)
        elif args.use_txt_lit:
            self.emb_txt_lit = Gate(self.embed_dim, self.n_txt_lit)
        for k in range(self.n_layers):
            self.aggregator_layers.append(
                Aggregator(self.conv_dim_list[k], self.conv_dim_list[k + 1], self.mess_dropout[k],
                           self.aggregation_type, self.use_residual, args))

        self.A_in = nn.Parameter(torch.sparse_coo_tensor(size=(n_entities, n_entities), dtype=torch.float32))
        if A_in is not None:
            self.A_in.data = A_in
        self.A_in.requires_grad = False
        self.milestone_score = args.milestone_score

        # device-side plan state (not part of the state dict)
        self._agg_plan: Optional[GraphPlan] = None        # CSR of the current A_in
        self._agg_values: Optional[torch.Tensor] = None   # its values, plan order (shared with A_in.data)
        self._att_plan: Optional[GraphPlan] = None        # plan of the (h, t, r) lists given to update_att
        self._att_key = None
        self._lit_planes = None                           # fp16 hi/lo planes of the (constant) literal tables
        self._unit_rec = None                             # scale record of planes bounded by 1 (normalised rows)
        self._h0q_cache = None                            # stacked h0 @ Q weight of all layers (parameter derived)
        self._part = None                                 # parallel.RowPartition when the path is row partitioned
        self.sync_attention = True                        # partitioned update_att: all-reduce the A_in values
        self._lit_key = None

    # ---- helpers -------------------------------------------------------------------------------
    def _param_device(self) -> torch.device:
        dev = self.entity_embed.weight.device
        if dev.type != "cuda":
            raise RuntimeError("literalkg_b200.LiteralKG runs on CUDA sm_100 only: move the model with "
                               ".to('cuda') first (there is no CPU fallback)")
        return dev

    def _literal(self, name: str) -> torch.Tensor:
        t = getattr(self, name)
        if t is None:
            raise RuntimeError(f"{name} was not given to the constructor")
        dev = self._param_device()
        if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(device=dev, dtype=torch.float32).contiguous()
            setattr(self, name, t)           # the reference also caches the moved table (model.py:269-276)
        return t

    def _current_plan(self) -> Tuple[GraphPlan, torch.Tensor]:
        """CSR plan + values of whatever ``self.A_in`` currently holds (constructor argument, checkpoint,
        external assignment or our own update_att result)."""
        dev = self._param_device()
        a = self.A_in.data
        vals = a._values()
        if (self._agg_plan is not None and self._agg_values is not None and a.device == dev
                and vals.data_ptr() == self._agg_values.data_ptr() and a._nnz() == self._agg_plan.nnz):
            return self._agg_plan, self._agg_values
        a = a.to(dev)
        plan = GraphPlan.from_coo(a._indices(), self.n_entities)
        values = plan.import_values(a._values())
        self._agg_plan, self._agg_values = plan, values
        self.A_in.data = plan.sparse(values)              # same matrix, coalesced, values shared with the kernels
        return plan, values

    # ---- gate ------------------------------------------------------------------------------------
    def _literal_planes(self, tables, rows=None):
        """The literal tables are constants (plain attributes in the reference): their fp16 hi/lo planes for
        the gate GEMM are built once and reused, like the reference caches the ``.to(device)`` copy."""
        key = (tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tables), rows)
        if self._lit_key != key:
            if rows is not None:
                tables = [t[rows[0]:rows[1]] for t in tables]
            src = tables[0] if len(tables) == 1 else torch.cat(list(tables), dim=1)
            self._lit_planes = ops.split_planes(src)
            self._lit_key = key
        return self._lit_planes

    def set_partition(self, part) -> None:
        """Row-partition the path over the ranks of ``part`` (``parallel.RowPartition``); None = single GPU."""
        self._part = part
        self._lit_key = None

    def _unit_record(self, dev) -> torch.Tensor:
        if self._unit_rec is None or self._unit_rec.device != dev:
            self._unit_rec = ops.scale_from_bound(1.0, dev)
        return self._unit_rec

    def gate_embeddings(self, out: Optional[torch.Tensor] = None, planes_window=None, rows=None):
        """model.py:265-279.  ``planes_window``: optional (Planes, col, k) column window that receives the scaled
        fp16 hi/lo copy of the result (A operand of the GEMMs that follow).  ``rows`` = (begin, end): only these
        entity rows (the gate is row local; used by the row partition)."""
        ent = self.entity_embed.weight
        dev = self._param_device()
        sl = slice(None) if rows is None else slice(rows[0], rows[1])
        with torch.no_grad():
            ent_rows = ent.detach()[sl]
            gate_mod, tables = None, ()
            if self.args.use_num_lit and self.args.use_txt_lit:
                gate_mod = self.emb_mul_lit
                tables = (self._literal("numerical_literals_embed"), self._literal("text_literals_embed"))
            elif self.args.use_num_lit:
                gate_mod, tables = self.emb_num_lit, (self._literal("numerical_literals_embed"),)
            elif self.args.use_txt_lit:
                gate_mod, tables = self.emb_txt_lit, (self._literal("text_literals_embed"),)
            if gate_mod is not None:
                ent_planes = ops.split_planes(ent_rows)
                out_planes = None
                if planes_window is not None:
                    # |gate output| <= max(1, max|entity|): convex mix of the entity row and a tanh
                    base, col, k = planes_window
                    out_planes = base.view(col, k, rec=ops.scale_from_bound(1.0, dev, other=ent_planes.rec))
                res = gate_mod(ent_rows, *[t[sl] for t in tables], out=out, out_planes=out_planes,
                               ent_planes=ent_planes, lit_planes=self._literal_planes(tables, rows))
                return (res, out_planes) if planes_window is not None else res
            if out is not None:
                out.copy_(ent_rows)
                if planes_window is not None:
                    base, col, k = planes_window
                    view = base.view(col, k, rec=torch.empty(_lib.LKG_SCALE_FLOATS, dtype=torch.float32, device=dev))
                    ops.split_planes(ent_rows, out=view)
                    return out, view
                return out
        return ent if rows is None else ent[sl]

    # ---- full-graph embedding pass -------------------------------------------------------------
    def gat_embeddings(self, gather: bool = True):
        """model.py:298-314.  Row partitioned: ``gather=False`` returns this rank's rows only (what the sharded
        scoring consumes); the default all-gathers the [N, G] result like the reference returns it."""
        with torch.no_grad():
            out = self._gat_embeddings_native()
            part = self._part
            if part is None or part.world == 1 or not gather:
                return out
            full = torch.empty((part.padded, out.shape[1]), dtype=torch.float32, device=out.device)
            full[part.begin:part.end] = out
            return part.all_gather_rows(full)[:self.n_entities]

    def _gat_embeddings_native(self, keep: Optional[dict] = None) -> torch.Tensor:
        dev = self._param_device()
        plan, a_values = self._current_plan()
        n, d, total = self.n_entities, self.embed_dim, self.total_conv_dim
        part = self._part if (self._part is not None and self._part.world > 1) else None
        rb, re = (0, n) if part is None else (part.begin, part.end)
        rows = None if part is None else (rb, re)
        n_own, n_tab = re - rb, (n if part is None else part.padded)
        plan.set_row_range(rb, re)
        cat = torch.empty((n_own, total), dtype=torch.float32, device=dev)   # concat buffer, this rank's rows
        if part is None:
            h0_tab = h0 = cat[:, :d]                      # gate output lives in the concat buffer
        else:
            h0_tab = torch.empty((n_tab, d), dtype=torch.float32, device=dev)   # row index == entity id
            h0 = h0_tab[rb:re]
        # scaled fp16 hi/lo planes of the concat buffer: A operand of the h0 @ Q and linear_gat tensor-core GEMMs.
        # Two K segments with their own scale records: the gate output and the L2-normalised layer outputs
        # (|x| <= 1); the second window starts on a 16-byte boundary.
        xcol = (d + 7) // 8 * 8
        cat_planes = _lib.Planes(n_own, xcol + (total - d), dev)
        _, h0_planes = self.gate_embeddings(out=h0, planes_window=(cat_planes, 0, d), rows=rows)
        xn_all = cat_planes.view(xcol, total - d, rec=self._unit_record(dev))
        folds = [layer.folded(self.lamda, self.alpha, k + 1) for k, layer in enumerate(self.aggregator_layers)]
        h0q = None
        offsets: List[int] = []
        zoff = -1                                         # column of the pre-projected layer-1 sum term in h0q
        if self.use_residual and self.n_layers > 0:
            qkey = tuple(id(f) for f in folds)                     # the fold dicts are cached per parameter version
            if self._h0q_cache is None or self._h0q_cache[0] != qkey:
                qs, cs, off = [], [], 0
                for k, f in enumerate(folds):
                    q1 = f["q1"] + f["pa"] if k == 0 else f["q1"]  # layer 0: ego == h0, fold ego @ Pa into h0 @ Q
                    offsets.append(off)
                    qs.append(q1); cs.append(f["c1"]); off += q1.shape[1]
                    if f["q2"] is not None:
                        qs.append(f["q2"]); cs.append(f["c2"]); off += f["q2"].shape[1]
                # layer 0 again: z = h0 @ Pb, the neighbour sum term projected BEFORE the aggregation
                # ((A h0) Pb == A (h0 Pb)): 32 more GEMM columns replace a 300 x 32 combine per row, and for
                # gcn / graphsage the 1 200-byte neighbour gather of layer 1 altogether
                zcol = off if d >= 128 and folds[0]["pb"].shape[1] % 4 == 0 else -1
                if zcol >= 0:
                    qs.append(folds[0]["pb"]); cs.append(torch.zeros_like(folds[0]["c1"])); off += folds[0]["pb"].shape[1]
                self._h0q_cache = (qkey, torch.cat(qs, dim=1).t().contiguous(), torch.cat(cs), offsets, folds, zcol)
            _, wq, cq, offsets, _, zoff = self._h0q_cache
            h0q = ops.linear([h0_planes], wq, cq)                  # this rank's rows

        c0 = self.aggregator_layers[0].out_dim if self.n_layers > 0 else 0
        z_tab = None
        if zoff >= 0:
            if part is None:
                z_tab = h0q[:, zoff:zoff + c0]                     # [N, C] view, row index == entity id
            else:
                z_tab = torch.empty((n_tab, c0), dtype=torch.float32, device=dev)
                z_tab[rb:re] = h0q[:, zoff:zoff + c0]
                part.all_gather_rows(z_tab)
        needs_h0_rows = self.n_layers > 0 and (z_tab is None or self.aggregation_type == 'bi-interaction')
        if part is not None:
            if needs_h0_rows:
                part.all_gather_rows(h0_tab)              # layer 1 gathers arbitrary neighbour rows of h0
            if self.scale_gat_dim is None:
                cat[:, :d] = h0

        x = h0_tab
        col = d
        for k, (layer, f) in enumerate(zip(self.aggregator_layers, folds)):
            c = layer.out_dim
            if h0q is not None:
                r1 = h0q[:, offsets[k]:offsets[k] + c]
                r2 = h0q[:, offsets[k] + c:offsets[k] + 2 * c] if f["q2"] is not None else None
            else:
                r1, r2 = f["c1"], f["c2"]
            x_out = torch.empty((n_tab, c), dtype=torch.float32, device=dev)
            layer.run(plan, a_values, x, f, r1, r2, x_out, cat[:, col:col + c], fold_ego=(h0q is not None and k == 0),
                      xn_planes=_lib.PlanesView(cat_planes, xcol + col - d, c, rec=xn_all.rec), rows=rows,
                      z=z_tab if k == 0 else None)
            if part is not None and k + 1 < self.n_layers:
                part.all_gather_rows(x_out)               # the next layer reads every row of this one
            x = x_out
            col += c
        plan.set_row_range(0, n)
        if keep is not None:
            keep["cat"] = cat
        if self.scale_gat_dim is not None:
            segs = [h0_planes] + ([xn_all] if total > d else [])
            return ops.linear(segs, self.linear_gat.weight, self.linear_gat.bias, _lib.ACT_LEAKY_RELU)
        return cat

    # ---- losses ----------------------------------------------------------------------------------
    def calculate_prediction_loss(self, head_ids, tail_pos_ids, tail_neg_ids):
        """model.py:316-348 (BPR)."""
        self.gat_embed = self.gat_embeddings()
        head_embed = self.gat_embed[head_ids]
        tail_pos_embed = self.gat_embed[tail_pos_ids]
        tail_neg_embed = self.gat_embed[tail_neg_ids]
        pos_score = torch.sum(head_embed * tail_pos_embed, dim=1)
        neg_score = torch.sum(head_embed * tail_neg_embed, dim=1)
        prediction_loss = torch.mean((-1.0) * F.logsigmoid(pos_score - neg_score))
        l2_loss = _L2_loss_mean(head_embed) + _L2_loss_mean(tail_pos_embed) + _L2_loss_mean(tail_neg_embed)
        return prediction_loss + self.prediction_l2loss_lambda * l2_loss

    def calc_triplet_loss(self, h, r, pos_t, neg_t):
        """model.py:364-428 (TransR on the GAT embeddings)."""
        r_embed = self.relation_embed(r)
        W_r = self.gat_trans_M[r]
        self.gat_embed = self.gat_embeddings()
        head_embed, tail_pos_embed, tail_neg_embed = self.gat_embed[h], self.gat_embed[pos_t], self.gat_embed[neg_t]
        r_mul_h = torch.bmm(head_embed.unsqueeze(1), W_r).squeeze(1)
        r_mul_pos_t = torch.bmm(tail_pos_embed.unsqueeze(1), W_r).squeeze(1)
        r_mul_neg_t = torch.bmm(tail_neg_embed.unsqueeze(1), W_r).squeeze(1)
        pos_score = torch.sum(torch.pow(r_mul_h + r_embed - r_mul_pos_t, 2), dim=1)
        neg_score = torch.sum(torch.pow(r_mul_h + r_embed - r_mul_neg_t, 2), dim=1)
        triplet_loss = torch.mean((-1.0) * F.logsigmoid(neg_score - pos_score))
        l2_loss = (_L2_loss_mean(r_mul_h) + _L2_loss_mean(r_embed) + _L2_loss_mean(r_mul_pos_t)
                   + _L2_loss_mean(r_mul_neg_t))
        return triplet_loss + self.kg_l2loss_lambda * l2_loss

    # ---- attention update ------------------------------------------------------------------------
    def update_attention(self, h_list, t_list, r_list, relations):
        """model.py:444-471.  Entirely on device: no host round trip, no per-relation Python loop."""
        dev = self._param_device()
        i64 = dict(device=dev, dtype=torch.int64, non_blocking=True)
        h, t, r = h_list.to(**i64).contiguous(), t_list.to(**i64).contiguous(), r_list.to(**i64).contiguous()
        # the CSR plan only depends on the edge list: recognise an unchanged list by content, not by address
        key = (GraphPlan.fingerprint(h, t, r), tuple(int(x) for x in relations))
        if self._att_plan is None or self._att_key != key:
            self._att_plan = GraphPlan(h, t, r, self.n_entities, self.n_relations, relations)
            self._att_key = key
        plan = self._att_plan
        part = self._part if (self._part is not None and self._part.world > 1) else None
        with torch.no_grad():
            if part is None:
                values = ops.attn_update(plan, self.entity_embed.weight.detach(), self.relation_embed.weight.detach())
            else:   # rows are independent: every rank fills the values of its own head rows
                plan.set_row_range(part.begin, part.end)
                values = torch.zeros(max(plan.nnz, 1), dtype=torch.float32, device=dev)[:plan.nnz]
                ops.attn_update(plan, self.entity_embed.weight.detach(), self.relation_embed.weight.detach(), out=values)
                plan.set_row_range(0, self.n_entities)
                if self.sync_attention:       # complete A_in on every rank (state dict); the layers only need own rows
                    part.all_reduce(values)
        self._agg_plan, self._agg_values = plan, values
        self.A_in.data = plan.sparse(values)

    # ---- scoring -----------------------------------------------------------------------------------
    def calc_score(self, head_ids, tail_ids):
        """model.py:473-486."""
        all_embed = self.gat_embeddings()
        return ops.score(all_embed, head_ids, tail_ids)

    def predict_links(self, head_ids, tail_ids):
        """model.py:488-491."""
        all_embed = self.gat_embeddings()
        return ops.predict(all_embed, head_ids, tail_ids, self.milestone_score)

    def topk(self, head_ids, tail_ids, k, target_tails=None, all_embed=None, tail_index=None):
        """Extension (BASELINE.json north star; no reference counterpart): per head the k best tails among
        ``tail_ids`` (larger score first, ties -> lower position), as (values, positions, ranks-of-targets).
        Large candidate sets go through the fused scoring + top-k kernels (no B x Nt score matrix);
        ``tail_index`` (``ops.ScoreIndex`` of the tails) can be reused across head batches."""
        if all_embed is None:
            all_embed = self.gat_embeddings()
        n_tails = tail_index.m if tail_index is not None else len(tail_ids)
        fused = target_tails is None and ops.fused_topk_applicable(n_tails, all_embed.shape[1], k)
        if fused:
            vals, pos = ops.score_topk(all_embed, head_ids, tail_ids, k, tail_index=tail_index)
            return vals, pos, None
        scores = ops.score(all_embed, head_ids, tail_ids)
        return ops.topk_rows(scores, k, target_tails)

    def topk_sharded(self, head_ids, k, local_embed, tail_index=None):
        """Row-partitioned all-entity top-k: every rank scores the heads against the entities it owns
        (``local_embed`` = ``gat_embeddings(gather=False)``) and the per-rank survivors are merged.  Returns
        (values, entity ids) [B, k], identical on every rank."""
        part = self._part
        assert part is not None, "topk_sharded needs set_partition()"
        dev = local_embed.device
        head_ids = head_ids.to(device=dev, dtype=torch.int64)
        # head rows from their owners: a zero-padded [B, G] block summed over the ranks (x + 0 is exact)
        mine = (head_ids >= part.begin) & (head_ids < part.end)
        hmat = torch.zeros((head_ids.numel(), local_embed.shape[1]), dtype=torch.float32, device=dev)
        hmat[mine] = local_embed[head_ids[mine] - part.begin]
        part.all_reduce(hmat)
        if tail_index is None:
            tail_index = self.sharded_index(local_embed)
        if ops.fused_topk_applicable(tail_index.m, local_embed.shape[1], k):
            vals, pos = ops.score_topk(local_embed, None, None, k, tail_index=tail_index, head_emb=hmat)
        else:
            both = torch.cat([local_embed, hmat])
            scores = ops.score(both, torch.arange(hmat.shape[0], device=dev) + local_embed.shape[0],
                               torch.arange(local_embed.shape[0], device=dev), rec=tail_index.rec)
            vals, pos, _ = ops.topk_rows(scores, k)
        ids = torch.where(pos >= 0, pos + part.begin, pos)
        from .parallel import merge_topk
        return merge_topk(part.all_gather_stack(vals), part.all_gather_stack(ids), k,
                          lambda sc, kk: ops.topk_rows(sc, kk)[:2])

    def sharded_index(self, local_embed):
        """Tail index of this rank's rows with a scale record that bounds EVERY rank's embeddings (the head rows
        come from other ranks): the absmax is all-reduced before the power-of-two scale is derived."""
        part = self._part
        rec = ops.scale_from_data(local_embed)
        part.all_reduce(rec[0:1], op=torch.distributed.ReduceOp.MAX)
        rec = ops.scale_from_bound(0.0, local_embed.device, other=rec)
        return ops.ScoreIndex(local_embed, None, rec=rec)

    def get_final_embeddings(self, entity_ids):
        """model.py:493-497."""
        return self.gat_embeddings()[entity_ids]

    def forward(self, *input, device, mode):
        """model.py:521-532."""
        self.device = device
        if mode == 'fine_tuning':
            return self.calculate_prediction_loss(*input)
        if mode == 'pre_training':
            return self.calc_triplet_loss(*input)
        if mode == 'update_att':
            return self.update_attention(*input)
        if mode == 'predict':
            return self.predict_links(*input)
        raise NotImplementedError(f"mode {mode!r} is outside the accelerated path (SURVEY.md section 8(f))")
