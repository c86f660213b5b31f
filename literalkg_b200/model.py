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

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.out_dim)
        self.weight.data.uniform_(-stdv, stdv)

    # ---- parameter folding -------------------------------------------------
    def folded(self, lamda: float, alpha: float, l: int) -> Dict[str, Optional[torch.Tensor]]:
        """linear(residual(hi)) == hi @ P + h0 @ Q + c with  M = (1-b) + b*W,  b = ln(lamda/l + 1)
        (model.py:90-99): P = (1-a) M W_lin^T, Q = a W_h0^T M W_lin^T, c = a b_h0 M W_lin^T + b_lin.
        Formed in float64 from the live parameters, returned in fp32.  Keys: pa, pb, p2 ([d_in, d_out]),
        q1, q2 ([embed_dim, d_out] or None), c1, c2 ([d_out])."""
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
        return res

    def _drop_mask(self, n: int, device) -> Optional[torch.Tensor]:
        if self.training and self.dropout > 0:
            keep = 1.0 - self.dropout
            return (torch.rand((n, self.out_dim), device=device) < keep).float() / keep
        return None

    def run(self, plan: GraphPlan, a_values: torch.Tensor, ego: torch.Tensor, f: Dict[str, Optional[torch.Tensor]],
            r1: Optional[torch.Tensor], r2: Optional[torch.Tensor], x_out: torch.Tensor,
            xn_out: Optional[torch.Tensor], fold_ego: bool = False, xn_planes=None) -> torch.Tensor:
        pa = None if fold_ego else f["pa"]
        return ops.aggregate(plan, a_values, ego, self.out_dim, pa, f["pb"], f["p2"], r1, r2,
                             self.layer_normalize.weight, self.layer_normalize.bias,
                             self._drop_mask(ego.shape[0], ego.device), x_out, xn_out, xn_planes)

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
    def _literal_planes(self, tables):
        """The literal tables are constants (plain attributes in the reference): their bf16 hi/lo planes for
        the gate GEMM are built once and reused, like the reference caches the ``.to(device)`` copy."""
        key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tables)
        if self._lit_key != key:
            src = tables[0] if len(tables) == 1 else torch.cat(tables, dim=1)
            self._lit_planes = ops.split_planes(src)
            self._lit_key = key
        return self._lit_planes

    def _unit_record(self, dev) -> torch.Tensor:
        if self._unit_rec is None or self._unit_rec.device != dev:
            self._unit_rec = ops.scale_from_bound(1.0, dev)
        return self._unit_rec

    def gate_embeddings(self, out: Optional[torch.Tensor] = None, planes_window=None):
        """model.py:265-279.  ``planes_window``: optional (Planes, col, k) column window that receives the scaled
        fp16 hi/lo copy of the result (A operand of the GEMMs that follow)."""
        ent = self.entity_embed.weight
        dev = self._param_device()
        with torch.no_grad():
            gate_mod, tables = None, ()
            if self.args.use_num_lit and self.args.use_txt_lit:
                gate_mod = self.emb_mul_lit
                tables = (self._literal("numerical_literals_embed"), self._literal("text_literals_embed"))
            elif self.args.use_num_lit:
                gate_mod, tables = self.emb_num_lit, (self._literal("numerical_literals_embed"),)
            elif self.args.use_txt_lit:
                gate_mod, tables = self.emb_txt_lit, (self._literal("text_literals_embed"),)
            if gate_mod is not None:
                ent_planes = ops.split_planes(ent.detach())
                out_planes = None
                if planes_window is not None:
                    # |gate output| <= max(1, max|entity|): convex mix of the entity row and a tanh
                    base, col, k = planes_window
                    out_planes = base.view(col, k, rec=ops.scale_from_bound(1.0, dev, other=ent_planes.rec))
                res = gate_mod(ent, *tables, out=out, out_planes=out_planes, ent_planes=ent_planes,
                               lit_planes=self._literal_planes(tables))
                return (res, out_planes) if planes_window is not None else res
            if out is not None:
                out.copy_(ent.detach())
                if planes_window is not None:
                    base, col, k = planes_window
                    view = base.view(col, k, rec=torch.empty(_lib.LKG_SCALE_FLOATS, dtype=torch.float32, device=dev))
                    ops.split_planes(ent.detach(), out=view)
                    return out, view
                return out
        return ent

    # ---- full-graph embedding pass -------------------------------------------------------------
    def gat_embeddings(self):
        """model.py:298-314."""
        with torch.no_grad():
            return self._gat_embeddings_native()

    def _gat_embeddings_native(self, keep: Optional[dict] = None) -> torch.Tensor:
        dev = self._param_device()
        plan, a_values = self._current_plan()
        n, d, total = self.n_entities, self.embed_dim, self.total_conv_dim
        cat = torch.empty((n, total), dtype=torch.float32, device=dev)
        h0 = cat[:, :d]                                   # gate output lives in the concat buffer
        # scaled fp16 hi/lo planes of the concat buffer: A operand of the h0 @ Q and linear_gat tensor-core GEMMs.
        # Two K segments with their own scale records: the gate output and the L2-normalised layer outputs
        # (|x| <= 1); the second window starts on a 16-byte boundary.
        xcol = (d + 7) // 8 * 8
        cat_planes = _lib.Planes(n, xcol + (total - d), dev)
        _, h0_planes = self.gate_embeddings(out=h0, planes_window=(cat_planes, 0, d))
        xn_all = cat_planes.view(xcol, total - d, rec=self._unit_record(dev))

        folds = [layer.folded(self.lamda, self.alpha, k + 1) for k, layer in enumerate(self.aggregator_layers)]
        h0q = None
        offsets: List[int] = []
        if self.use_residual and self.n_layers > 0:
            qs, cs, off = [], [], 0
            for k, f in enumerate(folds):
                q1 = f["q1"] + f["pa"] if k == 0 else f["q1"]      # layer 0: ego == h0, fold ego @ Pa into h0 @ Q
                offsets.append(off)
                qs.append(q1); cs.append(f["c1"]); off += q1.shape[1]
                if f["q2"] is not None:
                    qs.append(f["q2"]); cs.append(f["c2"]); off += f["q2"].shape[1]
            h0q = ops.linear([h0_planes], torch.cat(qs, dim=1).t().contiguous(), torch.cat(cs))

        x = h0
        col = d
        for k, (layer, f) in enumerate(zip(self.aggregator_layers, folds)):
            c = layer.out_dim
            if h0q is not None:
                r1 = h0q[:, offsets[k]:offsets[k] + c]
                r2 = h0q[:, offsets[k] + c:offsets[k] + 2 * c] if f["q2"] is not None else None
            else:
                r1, r2 = f["c1"], f["c2"]
            x_out = torch.empty((n, c), dtype=torch.float32, device=dev)
            layer.run(plan, a_values, x, f, r1, r2, x_out, cat[:, col:col + c], fold_ego=(h0q is not None and k == 0),
                      xn_planes=_lib.PlanesView(cat_planes, xcol + col - d, c, rec=xn_all.rec))
            x = x_out
            col += c
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
        with torch.no_grad():
            values = ops.attn_update(plan, self.entity_embed.weight.detach(), self.relation_embed.weight.detach())
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
        fused = (target_tails is None and all_embed.shape[1] <= ops.FUSED_TOPK_MAX_DIM
                 and (tail_index is not None or len(tail_ids) >= ops.FUSED_TOPK_MIN_TAILS))
        if fused:
            vals, pos = ops.score_topk(all_embed, head_ids, tail_ids, k, tail_index=tail_index)
            return vals, pos, None
        scores = ops.score(all_embed, head_ids, tail_ids)
        return ops.topk_rows(scores, k, target_tails)

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
