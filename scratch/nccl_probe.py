"""Raw NCCL all-gather / reduce-scatter bandwidth on this box, alone and next to a streaming kernel."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, ".")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
torch.cuda.set_device(local)
def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
res = {}
for mb in (128, 1200):
    n = mb * 1000 * 1000 // 4 // world * world
    buf = torch.empty(n, dtype=torch.float32, device="cuda")
    own = buf[rank * (n // world):(rank + 1) * (n // world)]
    ms = timed(lambda: dist.all_gather_into_tensor(buf, own))
    res[f"ag_{mb}MB"] = (round(ms, 3), round(n * 4 * (world - 1) / world / ms / 1e6, 1))
    out = torch.empty(n // world, dtype=torch.float32, device="cuda")
    ms = timed(lambda: dist.reduce_scatter_tensor(out, buf))
    res[f"rs_{mb}MB"] = (round(ms, 3), round(n * 4 * (world - 1) / world / ms / 1e6, 1))
# next to a bandwidth-bound kernel on the main stream (a big copy), all-gather launched first, async
n = 1200 * 1000 * 1000 // 4 // world * world
buf = torch.empty(n, dtype=torch.float32, device="cuda")
own = buf[rank * (n // world):(rank + 1) * (n // world)]
src = torch.empty(1 << 30, dtype=torch.float32, device="cuda"); dst = torch.empty_like(src)
def overlapped():
    w = dist.all_gather_into_tensor(buf, own, async_op=True)
    dst.copy_(src)          # 8 GB of HBM traffic ~ 1.3 ms
    w.wait()
res["ag_1200MB_with_copy"] = round(timed(overlapped), 3)
res["copy_alone"] = round(timed(lambda: dst.copy_(src)), 3)
if rank == 0:
    print(world, {k: v for k, v in os.environ.items() if k.startswith("NCCL")}, res, flush=True)
dist.destroy_process_group()
