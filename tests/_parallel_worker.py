"""Workers of tests/test_parallel*.py (spawned with torch.multiprocessing; one process per rank)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist


def _init(rank, world, port, backend="gloo"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)


def staged_partition_cls():
    """RowPartition whose collectives stage CUDA tensors through the host: lets two test ranks share one GPU over
    gloo.  Test-only (the product talks to torch.distributed / NCCL directly)."""
    from literalkg_b200.parallel import RowPartition

    class Staged(RowPartition):
        def _all_gather_flat(self, out, own, async_op):
            c = torch.empty(out.numel(), dtype=out.dtype)
            dist.all_gather_into_tensor(c, own.detach().cpu().reshape(-1), group=self.group)
            out.copy_(c)
            return None

        def _all_gather_list(self, outs, own, async_op):
            for r, o in enumerate(outs):
                if o.numel():
                    c = o.detach().cpu().contiguous()
                    dist.broadcast(c, src=r, group=self.group)
                    o.copy_(c)
            return None

        def _reduce_scatter_flat(self, out, full):
            c = full.detach().cpu()
            dist.all_reduce(c, group=self.group)
            n = out.numel()
            out.copy_(c[self.rank * n:(self.rank + 1) * n])

        def _all_reduce(self, t, op):
            c = t.detach().cpu()
            dist.all_reduce(c, op=op, group=self.group)
            t.copy_(c)

    return Staged


def cpu_collectives(rank, world, port, out_dir):
    """Host-side logic of the row partition on CPU tensors over gloo."""
    _init(rank, world, port)
    from literalkg_b200.parallel import merge_topk
    RowPartition = staged_partition_cls()               # gloo has no reduce-scatter: host-staged in the test subclass
    n, d = 1003, 8                                      # not divisible by the world size
    # nnz-balanced ranges (uneven row counts): every rank computes the same cut; exchanges go through per-rank views
    gh = torch.Generator().manual_seed(3)
    heads = (torch.rand(5000, generator=gh) ** 3 * n).long()          # skewed: low ids are heavy
    bal = RowPartition.balanced(n, heads)
    assert bal.bounds[0] == 0 and bal.bounds[-1] == n and not bal.uniform and bal.padded == n
    cost = torch.bincount(heads, minlength=n).double() + 4.0
    per = [cost[b:e].sum().item() for b, e in zip(bal.bounds[:-1], bal.bounds[1:])]
    assert max(per) <= 1.25 * (sum(per) / world) + cost.max().item()
    assert bal.n_own != RowPartition(n).n_own or world == 1
    rows = torch.arange(n)
    own = bal.owner_of(rows)
    assert all(bool(((rows[own == r] >= bal.bounds[r]) & (rows[own == r] < bal.bounds[r + 1])).all()) for r in range(world))
    refb = torch.arange(n * d, dtype=torch.float32).reshape(n, d)
    bufb = torch.full((n, d), -1.0)
    bufb[bal.begin:bal.end] = refb[bal.begin:bal.end]
    bal.all_gather_rows(bufb)
    assert torch.equal(bufb, refb)
    rsb = bal.reduce_scatter_rows(torch.full((n, d), float(rank + 1)))
    assert rsb.shape == (bal.n_own, d) and bool((rsb == float(sum(range(1, world + 1)))).all())
    part = RowPartition(n)
    assert part.world == world and part.rank == rank
    assert part.chunk * world >= n and part.begin == min(n, rank * part.chunk)
    ref = torch.arange(part.padded * d, dtype=torch.float32).reshape(part.padded, d)
    buf = torch.full((part.padded, d), -1.0)
    buf[part.begin:part.end] = ref[part.begin:part.end]
    part.all_gather_rows(buf)
    assert torch.equal(buf[:n], ref[:n])
    buf2 = torch.full((part.padded, d), -1.0)
    buf2[part.begin:part.end] = ref[part.begin:part.end]
    work = part.all_gather_rows(buf2, async_op=True)     # handle form used by the overlapped gate stage
    if work is not None:
        work.wait()
    assert torch.equal(buf2[:n], ref[:n])
    partial = torch.full((part.padded, d), float(rank + 1))
    rs = part.reduce_scatter_rows(partial)
    assert rs.shape == (part.chunk, d) and torch.equal(rs, torch.full((part.chunk, d), float(sum(range(1, world + 1)))))
    t = torch.full((5,), float(rank + 1))
    part.all_reduce(t)
    assert torch.equal(t, torch.full((5,), float(sum(range(1, world + 1)))))
    st = part.all_gather_stack(torch.tensor([rank, rank * 10]))
    assert st.shape == (world, 2) and st[:, 0].tolist() == list(range(world))
    # sharded top-k merge against a global top-k, ties included
    g = torch.Generator().manual_seed(5)
    scores = torch.randint(0, 50, (6, n), generator=g).float()            # many exact ties
    k = 9
    local = scores[:, part.begin:part.end]
    order = torch.argsort(local, dim=1, descending=True, stable=True)[:, :k]
    lv, li = torch.gather(local, 1, order), order + part.begin
    def topk_fn(sc, kk):
        o = torch.argsort(sc, dim=1, descending=True, stable=True)[:, :kk]
        return torch.gather(sc, 1, o), o
    mv, mi = merge_topk(part.all_gather_stack(lv), part.all_gather_stack(li), k, topk_fn)
    go = torch.argsort(scores, dim=1, descending=True, stable=True)[:, :k]
    assert torch.equal(mi, go) and torch.equal(mv, torch.gather(scores, 1, go))
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def gpu_partitioned_model(rank, world, port, out_dir):
    """Both ranks share cuda:0 (gloo stages the collectives through the host): the row-partitioned path must
    reproduce the single-GPU results."""
    # one GPU per rank over NCCL when the box has them, else both ranks share cuda:0 and gloo stages through the host
    nccl = torch.cuda.device_count() >= world
    _init(rank, world, port, "nccl" if nccl else "gloo")
    import literalkg_b200 as L
    import literalkg_oracle as O
    from literalkg_b200.parallel import RowPartition
    if not nccl:
        RowPartition = staged_partition_cls()
    torch.cuda.set_device(rank if nccl else 0)
    cfg = O.OracleConfig(n_conv_layers=3, mess_dropout=0.0)
    n, n_rel = 20_001, 8
    kg = L.synthetic.make_kg(n, 150_000, n_rel, seed=4, max_out_degree=300)
    num, txt = L.synthetic.make_literals(n, seed=4)
    args = argparse.Namespace(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    kt = L.KGTensors(kg.h, kg.t, kg.r, n_entities=n)

    def make():
        torch.manual_seed(1)
        m = L.LiteralKG(args, n, n_rel, kt.A_in, num, txt).cuda().eval()
        with torch.no_grad():
            m.entity_embed.weight.mul_(30)
        return m

    torch.set_grad_enabled(False)                        # inference path (the backward pass is single-GPU)
    single, multi = make(), make()
    part = RowPartition(n)
    multi.set_partition(part)
    for m in (single, multi):
        m(kt.h_list, kt.t_list, kt.r_list, kt.relations, device="cuda", mode="update_att")
    try:                                                 # a checkpoint before the (collective) completion must not
        multi.state_dict()                               # silently hold one rank's rows, nor hide a collective
        raise AssertionError("state_dict() must refuse while A_in is partial")
    except RuntimeError as ex:
        assert "complete_attention" in str(ex)
    multi.complete_attention()                           # lazy by default: the pass only reads a rank's own rows
    assert "A_in" in multi.state_dict()
    assert torch.equal(multi.A_in.data.indices(), single.A_in.data.indices())
    assert torch.equal(multi.A_in.data.values(), single.A_in.data.values())        # rows are independent: bit exact
    ref = single.gat_embeddings()
    local = multi.gat_embeddings(gather=False)
    full = multi.gat_embeddings()
    assert local.shape[0] == part.n_own and full.shape == ref.shape
    err = ((full - ref).abs().max() / ref.abs().max()).item()
    assert err < 2e-5, err                               # per-rank operand scales differ: not bit exact
    assert torch.equal(full[part.begin:part.end], local)
    # pre-partitioned edge list + nnz-balanced (uneven) row ranges: a rank is given only the triples of its own heads
    bal = RowPartition.balanced(n, kg.h)
    assert bal.bounds != part.bounds
    loc = make()
    loc.set_partition(bal, local_edges=True)
    mine = (kt.h_list >= bal.begin) & (kt.h_list < bal.end)
    loc(kt.h_list[mine].int(), kt.t_list[mine].int(), kt.r_list[mine].int(), kt.relations, device="cuda", mode="update_att")
    full_l = loc.gat_embeddings()
    assert full_l.shape == ref.shape and ((full_l - ref).abs().max() / ref.abs().max()).item() < 2e-5
    loc.complete_attention()                             # all-gather of the per-rank COO blocks = the full A_in
    assert torch.equal(loc.A_in.data.indices(), single.A_in.data.indices())
    assert torch.equal(loc.A_in.data.values(), single.A_in.data.values())
    full_l2 = loc.gat_embeddings()                       # plan rebuilt from the completed matrix
    assert ((full_l2 - ref).abs().max() / ref.abs().max()).item() < 2e-5
    del loc
    heads = torch.arange(0, 150, device="cuda") * 131 % n
    k = 10
    sv, si = multi.topk_sharded(heads, k, local)
    # same embeddings on one GPU.  The shards (10 001 tails) are below the fused path's minimum, so the two sides
    # round the scores differently (3-product GEMM vs exact re-score): compare positions where the top-(k+1)
    # scores are separated, values to 1e-5.
    rv, rp, _ = single.topk(heads, torch.arange(n, device="cuda"), k, all_embed=full)
    s = L.ops.score(full, heads, torch.arange(n, device="cuda"))
    tv, _ = torch.topk(s, k + 1, dim=1)
    clear = (tv[:, :-1] - tv[:, 1:]).min(dim=1).values > 1e-5 * s.abs().max()
    assert clear.sum() > 50
    assert torch.equal(si[clear], rp[clear])
    assert ((sv - rv).abs().max() / rv.abs().max()).item() < 1e-5
    # and with the fused kernels on both sides (every tail a candidate set of its own rank): bit exact
    L.ops.FUSED_TOPK_MIN_TAILS = 4096
    sv2, si2 = multi.topk_sharded(heads, k, local)
    rv2, rp2, _ = single.topk(heads, torch.arange(n, device="cuda"), k, all_embed=full)
    assert torch.equal(si2, rp2) and torch.equal(sv2, rv2)
    # a list of head batches shares one round of collectives: same results per batch
    many = multi.topk_sharded([heads[:70], heads[70:]], k, local)
    assert torch.equal(torch.cat([m_[1] for m_ in many]), si2) and torch.equal(torch.cat([m_[0] for m_ in many]), sv2)
    # training modes: the row-partitioned backward (reduce-scatter of the A^T partials, all-reduce of the parameter
    # gradients, all-gather of the entity-gradient rows) must reproduce the single-GPU gradients on every rank
    torch.set_grad_enabled(True)
    gen = torch.Generator().manual_seed(9)
    bh, bp, bn = (torch.randint(0, n, (512,), generator=gen).cuda() for _ in range(3))
    br = torch.randint(0, n_rel, (512,), generator=gen).cuda()
    for mode, batch in (("fine_tuning", (bh, bp, bn)), ("pre_training", (bh, br, bp, bn))):
        losses = []
        for m in (single, multi):
            m.train()
            m.zero_grad(set_to_none=True)
            loss = m(*batch, device="cuda", mode=mode)
            loss.backward()
            losses.append(loss.item())
        assert abs(losses[0] - losses[1]) <= 1e-5 * abs(losses[0]), (mode, losses)
        worst = 0.0
        for (k1, p1), (k2, p2) in zip(single.named_parameters(), multi.named_parameters()):
            if k1 == "A_in":
                continue
            assert (p1.grad is None) == (p2.grad is None), k1
            if p1.grad is not None:
                e_ = ((p1.grad - p2.grad).abs().max() / p1.grad.abs().max().clamp_min(1e-30)).item()
                assert e_ < 2e-4, (mode, k1, e_)          # per-rank operand scales and summation orders differ
                worst = max(worst, e_)
    open(os.path.join(out_dir, f"ok{rank}"), "w").write(f"{err:.3e} grad {worst:.3e}")
    dist.destroy_process_group()
