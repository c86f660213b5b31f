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


def cpu_collectives(rank, world, port, out_dir):
    """Host-side logic of the row partition on CPU tensors over gloo."""
    _init(rank, world, port)
    from literalkg_b200.parallel import RowPartition, merge_topk
    n, d = 1003, 8                                      # not divisible by the world size
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
    multi.complete_attention()                           # lazy by default: the pass only reads a rank's own rows
    assert torch.equal(multi.A_in.data.indices(), single.A_in.data.indices())
    assert torch.equal(multi.A_in.data.values(), single.A_in.data.values())        # rows are independent: bit exact
    ref = single.gat_embeddings()
    local = multi.gat_embeddings(gather=False)
    full = multi.gat_embeddings()
    assert local.shape[0] == part.n_own and full.shape == ref.shape
    err = ((full - ref).abs().max() / ref.abs().max()).item()
    assert err < 2e-5, err                               # per-rank operand scales differ: not bit exact
    assert torch.equal(full[part.begin:part.end], local)
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
