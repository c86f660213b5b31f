"""Row partition of the path over the GPUs of one box: one process per GPU, NCCL over NVLink.

The reference has no working multi-GPU code (a dead ``nn.DataParallel``, main.py:82-84); this is the B200
analogue of the scaling axis it lacks.  Head rows are split into equal contiguous ranges; every rank keeps the
full CSR plan (0.4 GB at 20 M triples) and the raw parameter tables, and owns its rows of every activation:

    update_att       rows independent, no collective (optional all-reduce to complete the ``A_in`` values)
    gate, h0 @ Q     row local
    layer k          reads the whole ego table -> one all-gather of the (N/P x d_k) row blocks per layer, written in
                     place into a [P * chunk, d_k] buffer whose row index is the global entity id
    linear_gat       row local; the final embeddings stay sharded -- each rank scores its own candidate tails
    scoring / top-k  head rows are summed from their owners (all-reduce of a zero-padded [B, G] block), local fused
                     top-k over the local tails, all-gather of the P x [B, k] survivors, k-way merge on every rank

The collectives go through ``torch.distributed`` (NCCL on the GPU box).  With the ``gloo`` backend (CPU tests, or two
processes sharing one GPU in the test-suite) CUDA tensors are staged through host memory.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

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
    """Rank p owns head rows [p * chunk, min(N, (p + 1) * chunk)), chunk = ceil(N / world)."""

    def __init__(self, n_entities: int, rank: Optional[int] = None, world: Optional[int] = None, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if world is None else int(world)
        self.rank = dist.get_rank(group) if rank is None else int(rank)
        self.n = int(n_entities)
        self.chunk = (self.n + self.world - 1) // self.world
        self.begin = min(self.n, self.rank * self.chunk)
        self.end = min(self.n, self.begin + self.chunk)
        self.padded = self.chunk * self.world          # rows of an all-gather buffer (row index == entity id)

    @property
    def n_own(self) -> int:
        return self.end - self.begin

    def owner_of(self, rows: torch.Tensor) -> torch.Tensor:
        return torch.div(rows, self.chunk, rounding_mode="floor")

    def _backend(self) -> str:
        return dist.get_backend(self.group) if dist.is_initialized() else "none"

    # ---- collectives -------------------------------------------------------------------------------------
    def all_gather_rows(self, buf: torch.Tensor, async_op: bool = False):
        """``buf`` [padded, d] contiguous with this rank's rows already written: fetches every other rank's
        row block in place.  ``async_op``: returns a handle whose ``wait()`` orders the current stream after the
        transfer (NCCL: the collective runs on NCCL's stream next to the kernels launched meanwhile), or None when
        the transfer already happened."""
        assert buf.is_contiguous() and buf.shape[0] == self.padded
        if self.world == 1:
            return None if async_op else buf
        own = buf[self.rank * self.chunk:(self.rank + 1) * self.chunk]
        with _comm_profile(buf, "all_gather", buf.numel() * buf.element_size(), async_op):
            if buf.is_cuda and self._backend() != "nccl":      # gloo has no CUDA all-gather: stage through the host
                full = torch.empty(buf.numel(), dtype=buf.dtype)
                dist.all_gather_into_tensor(full, own.cpu().reshape(-1), group=self.group)
                buf.copy_(full.view(buf.shape))
                return None if async_op else buf
            work = dist.all_gather_into_tensor(buf.view(-1), own.reshape(-1), group=self.group, async_op=async_op)
        return work if async_op else buf

    def reduce_scatter_rows(self, full: torch.Tensor) -> torch.Tensor:
        """``full`` [padded, d]: every rank's partial sums for ALL rows -> [chunk, d], the sum over the ranks of this
        rank's row block (the dual of ``all_gather_rows``; used by the backward of ``A_in @ x``)."""
        assert full.is_contiguous() and full.shape[0] == self.padded
        if self.world == 1:
            return full
        with _comm_profile(full, "reduce_scatter", full.numel() * full.element_size(), False):
            if self._backend() == "nccl" and full.is_cuda:
                out = torch.empty((self.chunk, full.shape[1]), dtype=full.dtype, device=full.device)
                dist.reduce_scatter_tensor(out.view(-1), full.view(-1), group=self.group)
                return out
            self.all_reduce(full)                              # gloo: no reduce-scatter for these tensors
            return full[self.rank * self.chunk:(self.rank + 1) * self.chunk]

    def all_reduce(self, t: torch.Tensor, op=dist.ReduceOp.SUM) -> torch.Tensor:
        if self.world == 1:
            return t
        with _comm_profile(t, "all_reduce", t.numel() * t.element_size(), False):
            if t.is_cuda and self._backend() != "nccl":
                c = t.cpu()
                dist.all_reduce(c, op=op, group=self.group)
                t.copy_(c)
            else:
                dist.all_reduce(t, op=op, group=self.group)
        return t

    def all_gather_stack(self, t: torch.Tensor) -> torch.Tensor:
        """[...] on every rank -> [world, ...]."""
        if self.world == 1:
            return t.unsqueeze(0)
        t = t.contiguous()
        staged = t.is_cuda and self._backend() != "nccl"
        src = t.cpu() if staged else t
        out = torch.empty(self.world * src.numel(), dtype=src.dtype, device=src.device)
        dist.all_gather_into_tensor(out, src.reshape(-1), group=self.group)     # flat: accepted by every backend
        return out.view(self.world, *t.shape).to(t.device)


class PeerGather:
    """All-gather of row blocks by peer-to-peer copies instead of NCCL (opt-in: ``LKG_P2P_GATHER=1``, NCCL backend,
    one GPU per rank on one NVLink box).

    Why: the layer-1 all-gather of ``h0`` (1.2 GB at N = 1 M) is the largest transfer of the pass and NCCL's kernel
    competes for the SMs with the HBM-bound attention kernel it is supposed to hide under (measured overlap ~5 %,
    ``scratch/nccl_probe.py``).  Device-to-device ``copy_`` between peer-mapped buffers goes through the copy engines:
    no SM is involved, so the transfer really runs next to the kernels.

    Every rank owns two ``[padded, d]`` buffers whose CUDA IPC handles are exchanged once; a gather PUSHES the rank's
    own rows into the same slot of every peer on a side stream and closes with a one-element NCCL all-reduce issued
    from that stream (a stream-ordered barrier: it completes when every rank's pushes are done).  Slots alternate per
    call; slot s is rewritten two calls later, after the barrier of the call in between, which every rank only joins
    once its main stream is past the kernels that read slot s -- the callers' ``wait()`` / kernel launches and the next
    ``begin()`` are issued on that main stream in program order.

    Status (round 1): results identical to the NCCL path (tests/test_parallel.py with LKG_P2P_GATHER=1 on two GPUs),
    but torch's cross-device ``copy_`` into the IPC-mapped peer buffers moves only ~30 GB/s on the test box (2-GPU
    pass 29.9 ms vs 10.1 ms with NCCL) -- it does not take the NVLink peer path.  Off by default; the transfer needs
    its own copy kernel over the peer mapping (or explicit cudaMemcpyPeerAsync) before it can replace NCCL."""

    def __init__(self, part: "RowPartition", d: int, device, dtype=torch.float32):
        from torch.multiprocessing.reductions import reduce_tensor
        self.part, self.d = part, int(d)
        self.local = [torch.zeros((part.padded, d), dtype=dtype, device=device) for _ in range(2)]
        mine = [reduce_tensor(b) for b in self.local]
        everyone = [None] * part.world
        dist.all_gather_object(everyone, mine, group=part.group)
        self.peers = []                                 # peers[r][slot]: rank r's buffer mapped into this process
        for r in range(part.world):
            self.peers.append(self.local if r == part.rank else [fn(*args) for fn, args in everyone[r]])
        self.stream = torch.cuda.Stream(device=device)
        self.flag = torch.zeros(1, dtype=torch.float32, device=device)
        self.calls = 0

    def begin(self):
        """-> (table, handle): ``table`` [padded, d] is this call's buffer -- write this rank's rows
        ``[begin, end)`` into it on the current stream, then call ``handle.start()``; ``handle.wait()`` orders the
        current stream after the arrival of every other rank's rows."""
        slot = self.calls & 1
        self.calls += 1
        return self.local[slot], _PeerHandle(self, slot)


class _PeerHandle:
    def __init__(self, owner: PeerGather, slot: int):
        self.owner, self.slot, self.done = owner, slot, None

    def start(self):
        o, part = self.owner, self.owner.part
        main = torch.cuda.current_stream()
        o.stream.wait_stream(main)                      # the rows are written, earlier readers of the slot are queued
        with torch.cuda.stream(o.stream):
            rows = o.local[self.slot][part.begin:part.end]
            for r in range(part.world):
                if r != part.rank:
                    o.peers[r][self.slot][part.begin:part.end].copy_(rows, non_blocking=True)
            dist.all_reduce(o.flag, group=part.group)   # stream-ordered barrier: every rank's pushes have landed
            self.done = torch.cuda.Event()
            self.done.record(o.stream)
        return self

    def wait(self):
        torch.cuda.current_stream().wait_event(self.done)


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
