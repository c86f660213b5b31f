"""Row partition of the path over the GPUs of one box: one process per GPU, NCCL / NVLink.

The reference has no working multi-GPU code (a dead ``nn.DataParallel``, main.py:82-84); this is the B200
analogue of the scaling axis it lacks.  Head rows are split into contiguous ranges -- equal row counts by default,
nnz-balanced with ``RowPartition.balanced`` (real KGs group ids by entity type) -- and every rank owns its rows of
every activation:

    update_att       rows independent, no collective (``complete_attention`` gathers ``A_in`` for checkpoints)
    gate, h0 @ Q     row local
    layer k          reads the whole ego table -> one exchange of the row blocks per layer, in place in a
                     [rows, d_k] table whose row index is the global entity id
    linear_gat       row local; the final embeddings stay sharded -- each rank scores its own candidate tails
    scoring / top-k  head rows from their owners, local fused top-k over the local tails, all-gather of the
                     P x [B, k] survivors, k-way merge on every rank

Two transports for the per-layer exchange:
  * ``PeerTable`` (default on NCCL boxes): the table lives in symmetric memory (CUDA VMM mappings of every rank's
    copy, ``torch.distributed._symmetric_memory`` for allocation / rendezvous / barrier only); a rank PUSHES its row
    block into every peer's copy with device-to-device copies on a side stream -- copy engines over NVLink, no SM is
    involved, so the 1.2 GB ``h0`` exchange really runs next to the HBM-bound attention kernel (NCCL's all-gather
    kernel competes with it for the SMs: ~5 % overlap measured in round 1);
  * NCCL all-gather / reduce-scatter (training pass, and the fallback when symmetric memory is unavailable).
Collectives go straight to ``torch.distributed``; host staging for CPU-only test backends lives in the tests.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


class _comm_profile:
    """Brackets a blocking collective with CUDA events when ``ops.PROFILE`` is on (bench.py's per-call timing)."""

    def __init__(self, t: torch.Tensor, kind: str, nbytes: int, async_op: bool):
        from . import ops
        self.on = ops.PROFILE is not None and ops.PROFILE.only is None and t.is_cuda and not async_op
        self.name = f"{kind}_{nbytes / 1e6:.0f}MB"

    def __enter__(self):
        if self.on:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record()

    def __exit__(self, *exc):
        if self.on:
            from . import ops
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            ops.PROFILE.events.append((self.name, self.start, end))
        return False


class RowPartition:
    """Rank p owns head rows [bounds[p], bounds[p + 1]).  Default: equal chunks of ceil(N / world) rows (tables are
    then padded to ``chunk * world`` rows so that one flat all-gather moves them); ``bounds``: any non-decreasing
    list of world + 1 row indices from 0 to N (tables have exactly N rows, exchanges are per-rank views)."""

    def __init__(self, n_entities: int, rank: Optional[int] = None, world: Optional[int] = None, group=None,
                 bounds: Optional[Sequence[int]] = None):
        self.group = group
        self.world = dist.get_world_size(group) if world is None else int(world)
        self.rank = dist.get_rank(group) if rank is None else int(rank)
        self.n = int(n_entities)
        self.chunk = (self.n + self.world - 1) // self.world
        if bounds is None:
            self.uniform = True
            self.bounds = [min(self.n, r * self.chunk) for r in range(self.world)] + [self.n]
            self.padded = self.chunk * self.world          # rows of an exchange table (row index == entity id)
        else:
            b = [int(x) for x in bounds]
            if len(b) != self.world + 1 or b[0] != 0 or b[-1] != self.n or any(y < x for x, y in zip(b, b[1:])):
                raise ValueError("bounds must be world + 1 non-decreasing row indices from 0 to n_entities")
            self.uniform = False
            self.bounds = b
            self.padded = self.n
        self.begin, self.end = self.bounds[self.rank], self.bounds[self.rank + 1]

    @classmethod
    def balanced(cls, n_entities: int, heads, rank: Optional[int] = None, world: Optional[int] = None, group=None,
                 row_cost: float = 4.0) -> "RowPartition":
        """Contiguous ranges of (nearly) equal cost, cost(row) = its triples + ``row_cost`` (the row-local work: gate,
        GEMM rows, table rows; ~2.7 KB per row against ~1.2 KB per triple at the reference dims).  ``heads``: the
        head column of the edge list (numpy / tensor, any order).  Deterministic: every rank computes the same cut."""
        w = dist.get_world_size(group) if world is None else int(world)
        h = heads.detach().cpu().numpy() if isinstance(heads, torch.Tensor) else np.asarray(heads)
        cost = np.bincount(h.astype(np.int64), minlength=int(n_entities)).astype(np.float64) + float(row_cost)
        cum = np.cumsum(cost)
        cuts = np.searchsorted(cum, cum[-1] * np.arange(1, w) / w, side="left") + 1 if n_entities else np.zeros(w - 1)
        bounds = [0] + [int(min(max(c, 0), n_entities)) for c in cuts] + [int(n_entities)]
        for i in range(1, len(bounds)):
            bounds[i] = max(bounds[i], bounds[i - 1])
        return cls(n_entities, rank, w, group, bounds)

    @property
    def n_own(self) -> int:
        return self.end - self.begin

    def rows_of(self, r: int) -> Tuple[int, int]:
        return self.bounds[r], self.bounds[r + 1]

    def owner_of(self, rows: torch.Tensor) -> torch.Tensor:
        b = torch.as_tensor(self.bounds[1:], device=rows.device, dtype=rows.dtype)
        return torch.searchsorted(b, rows, right=True)

    def _backend(self) -> str:
        return dist.get_backend(self.group) if dist.is_initialized() else "none"

    def _views(self, buf: torch.Tensor) -> List[torch.Tensor]:
        return [buf[b:e] for b, e in zip(self.bounds[:-1], self.bounds[1:])]

    # ---- collectives (overridden by the CPU-staging subclass of the tests) ---------------------------------------
    def _all_gather_flat(self, out: torch.Tensor, own: torch.Tensor, async_op: bool):
        return dist.all_gather_into_tensor(out, own, group=self.group, async_op=async_op)

    def _all_gather_list(self, outs: List[torch.Tensor], own: torch.Tensor, async_op: bool):
        """Uneven row blocks: one broadcast per owner into its (in place) view.  The works complete in issue order on
        the backend's stream, so waiting for the last one waits for all of them."""
        last = None
        for r, o in enumerate(outs):
            if o.numel():
                src = r if self.group is None else dist.get_global_rank(self.group, r)
                last = dist.broadcast(o, src=src, group=self.group, async_op=async_op)
        return last if async_op else None

    def _reduce_scatter_flat(self, out: torch.Tensor, full: torch.Tensor):
        dist.reduce_scatter_tensor(out, full, group=self.group)

    def _all_reduce(self, t: torch.Tensor, op):
        dist.all_reduce(t, op=op, group=self.group)

    def all_gather_rows(self, buf: torch.Tensor, async_op: bool = False):
        """``buf`` [padded, d] contiguous with this rank's rows already written: fetches every other rank's
        row block in place.  ``async_op``: returns a handle whose ``wait()`` orders the current stream after the
        transfer (NCCL: the collective runs on NCCL's stream next to the kernels launched meanwhile), or None when
        the transfer already happened."""
        assert buf.is_contiguous() and buf.shape[0] == self.padded
        if self.world == 1:
            return None if async_op else buf
        with _comm_profile(buf, "all_gather", buf.numel() * buf.element_size(), async_op):
            if self.uniform:
                own = buf[self.rank * self.chunk:(self.rank + 1) * self.chunk]
                work = self._all_gather_flat(buf.view(-1), own.reshape(-1), async_op)
            else:
                work = self._all_gather_list(self._views(buf), buf[self.begin:self.end], async_op)
        return work if async_op else buf

    def reduce_scatter_rows(self, full: torch.Tensor) -> torch.Tensor:
        """``full`` [padded, d]: every rank's partial sums for ALL rows -> the sum over the ranks of this rank's row
        block ([chunk, d] for equal chunks, [n_own, d] otherwise); the dual of ``all_gather_rows``, used by the
        backward of ``A_in @ x``."""
        assert full.is_contiguous() and full.shape[0] == self.padded
        if self.world == 1:
            return full
        with _comm_profile(full, "reduce_scatter", full.numel() * full.element_size(), False):
            if self.uniform:
                out = torch.empty((self.chunk, full.shape[1]), dtype=full.dtype, device=full.device)
                self._reduce_scatter_flat(out.view(-1), full.view(-1))
                return out
            self._all_reduce(full, dist.ReduceOp.SUM)          # uneven blocks: sum everywhere, keep the own rows
            return full[self.begin:self.end]

    def all_reduce(self, t: torch.Tensor, op=dist.ReduceOp.SUM) -> torch.Tensor:
        if self.world == 1:
            return t
        with _comm_profile(t, "all_reduce", t.numel() * t.element_size(), False):
            self._all_reduce(t, op)
        return t

    def all_gather_stack(self, t: torch.Tensor) -> torch.Tensor:
        """[...] on every rank -> [world, ...]."""
        if self.world == 1:
            return t.unsqueeze(0)
        t = t.contiguous()
        out = torch.empty(self.world * t.numel(), dtype=t.dtype, device=t.device)
        self._all_gather_flat(out, t.reshape(-1), False)
        return out.view(self.world, *t.shape)

    def peer_tables_enabled(self) -> bool:
        return (os.environ.get("LKG_P2P_GATHER", "1") != "0" and self.world > 1 and self._backend() == "nccl"
                and torch.cuda.is_available())


class PeerTable:
    """Exchange table [padded, d] fp32 in symmetric memory: every rank maps every other rank's copy, a rank writes
    its own row block and PUSHES it into all peer copies (see the module docstring).  ``slots`` copies alternate per
    call: slot s is rewritten two calls later, after the barrier of the call in between, which every rank only
    joins once its compute stream is past the kernels that read slot s (the callers' ``wait()``, kernel launches and
    the next ``begin()`` are issued on that stream in program order)."""

    def __init__(self, part: RowPartition, d: int, device, slots: int = 2):
        import torch.distributed._symmetric_memory as symm_mem
        self.part, self.d, self.slots = part, int(d), int(slots)
        self.rows = part.padded
        group = part.group if part.group is not None else dist.group.WORLD
        self.t = symm_mem.empty((self.slots, self.rows, self.d), dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.t, group)
        self.stream = torch.cuda.Stream(device=device)
        # the row blocks go out as device-to-device copies (copy engines), one peer at a time on one stream
        # (measured at 8 GPUs, r02j: 4 concurrent streams made the 1.2 GB exchange SLOWER -- pass 5.0 ms against 4.1 ms
        # with one stream; the staggered destinations only keep the ingress links disjoint when every rank sends to one
        # peer at a time.  LKG_PUSH_STREAMS overrides for experiments.)
        n_copy = max(1, min(int(os.environ.get("LKG_PUSH_STREAMS", "1")), max(1, part.world - 1)))
        self.copy_streams = [self.stream] + [torch.cuda.Stream(device=device) for _ in range(n_copy - 1)]
        # Optional (LKG_PUSH_SM_FRACTION > 0): a fraction of the row block goes out through lkg_peer_push -- a kernel on
        # LKG_PUSH_SM_CTAS CTAs that stores straight into the peers' copies -- at the same time as the copy engines move
        # the rest.  Measured at 8 GPUs (r02t, 1.33 GB table): pass 4.03 ms with the copy engines alone, 4.28 ms with
        # 40 % on 24 CTAs, 4.51 ms with 70 % -- the SM stores do not add NVLink throughput here, so the default is off.
        self.sm_fraction = float(os.environ.get("LKG_PUSH_SM_FRACTION", "0"))
        self.sm_ctas = int(os.environ.get("LKG_PUSH_SM_CTAS", "24"))
        self.sm_stream = torch.cuda.Stream(device=device) if self.sm_fraction > 0 else None
        self.peer_ptrs = list(self.hdl.buffer_ptrs)
        self.calls = 0

    def begin(self):
        """-> (table, handle): ``table`` [padded, d] is this call's copy -- write this rank's rows ``[begin, end)``
        into it on the current stream, then call ``handle.start()``; ``handle.wait()`` orders the current stream
        after the arrival of every other rank's rows."""
        slot = self.calls % self.slots
        self.calls += 1
        return self.t[slot], _PeerHandle(self, slot)


class _PeerHandle:
    def __init__(self, owner: PeerTable, slot: int):
        self.owner, self.slot, self.done = owner, slot, None

    def start(self):
        o, part = self.owner, self.owner.part
        main = torch.cuda.current_stream()
        for cs in o.copy_streams:
            cs.wait_stream(main)                        # the rows are written, earlier readers of the slot are queued
        from . import ops
        timed = ops.PROFILE is not None and ops.PROFILE.only is None
        if timed:                                       # bench.py's per-call breakdown: transfer + barrier time
            t0 = torch.cuda.Event(enable_timing=True)
            t0.record(o.stream)
        b, e = part.begin, part.end
        if e > b and o.sm_stream is not None:            # the tail of the block: pushed by SMs, concurrently
            import ctypes as C
            from . import _lib
            m = b + int((e - b) * (1.0 - o.sm_fraction))
            o.sm_stream.wait_stream(main)
            nbytes = (e - m) * o.d * 4
            if nbytes > 0 and nbytes % 16 == 0:
                offb = (self.slot * o.rows + m) * o.d * 4
                peers = [r for r in range(part.world) if r != part.rank]
                arr = (C.c_void_p * len(peers))(*[o.peer_ptrs[r] + offb for r in peers])
                with torch.cuda.stream(o.sm_stream):
                    _lib.check(_lib.load().lkg_peer_push(o.t[self.slot, m:e].data_ptr(), nbytes, arr, len(peers), o.sm_ctas,
                                                         o.sm_stream.cuda_stream))
                e = m
        if e > b:
            rows = o.t[self.slot, b:e]
            off = (self.slot * o.rows + b) * o.d
            for step in range(1, part.world):           # staggered destinations: no two ranks start on the same peer
                r = (part.rank + step) % part.world
                with torch.cuda.stream(o.copy_streams[step % len(o.copy_streams)]):
                    o.hdl.get_buffer(r, (e - b, o.d), torch.float32, off).copy_(rows, non_blocking=True)
        for cs in o.copy_streams[1:]:
            o.stream.wait_stream(cs)
        if o.sm_stream is not None:
            o.stream.wait_stream(o.sm_stream)
        with torch.cuda.stream(o.stream):
            o.hdl.barrier()                             # stream ordered: every rank's pushes have landed
            self.done = torch.cuda.Event(enable_timing=timed)
            self.done.record(o.stream)
        if timed:
            ops.PROFILE.events.append((f"peer_push_{o.rows * o.d * 4 / 1e6:.0f}MB", t0, self.done))
        return self

    def wait(self):
        torch.cuda.current_stream().wait_event(self.done)


class PeerExchange:
    """The exchange tables of one model: created on first use per (name, width), NCCL fallback when symmetric memory
    cannot be set up on this box (reported once)."""

    def __init__(self, part: RowPartition, device):
        self.part, self.device = part, device
        self.tables = {}
        self.failed: Optional[str] = None

    def table(self, name: str, d: int) -> Optional[PeerTable]:
        if self.failed is not None:
            return None
        key = (name, int(d))
        if key not in self.tables:
            try:
                self.tables[key] = PeerTable(self.part, d, self.device)
            except Exception as ex:                     # noqa: BLE001 -- any setup failure means "use NCCL"
                self.failed = f"{type(ex).__name__}: {ex}"
                if self.part.rank == 0:
                    import sys
                    print(f"[literalkg_b200] symmetric-memory exchange unavailable ({self.failed}); using NCCL "
                          "all-gather", file=sys.stderr, flush=True)
                return None
        return self.tables[key]


def merge_topk(vals: torch.Tensor, ids: torch.Tensor, k: int, topk_fn) -> Tuple[torch.Tensor, torch.Tensor]:
    """k-way merge of per-rank results.  ``vals`` / ``ids`` [world, B, k] (ids = global tail positions, -1 pads with
    value -inf).  Ranks own ascending, disjoint position ranges and each list is ordered (score desc, position asc),
    so "ties -> lower column" on the rank-major concatenation equals "ties -> lower global position".
    ``topk_fn(scores [B, world * k], k) -> (values, columns)`` is the row-wise top-k (``ops.topk_rows`` on the GPU)."""
    world, b, kk = vals.shape
    flat_v = vals.permute(1, 0, 2).reshape(b, world * kk).contiguous()
    flat_i = ids.permute(1, 0, 2).reshape(b, world * kk).contiguous()
    top_v, cols = topk_fn(flat_v, k)
    top_i = torch.gather(flat_i, 1, cols.clamp_min(0))
    top_i = torch.where(cols < 0, torch.full_like(top_i, -1), top_i)
    return top_v, top_i
