"""Scoring-only probe: candidate statistics and per-call timing of the fused top-k on the bench embeddings."""
import sys, time, torch
sys.path.insert(0, ".")
from literalkg_b200 import ops
torch.manual_seed(0)
n, g, b, k = 1_000_000, 256, 2048, int(sys.argv[1]) if len(sys.argv) > 1 else 10
kind = sys.argv[2] if len(sys.argv) > 2 else "leaky"
x = torch.randn(n, g, device="cuda")
emb = torch.nn.functional.leaky_relu(x, 0.01) * 0.3 if kind == "leaky" else x
heads = (torch.arange(b, device="cuda") * 487) % n
ti = ops.ScoreIndex(emb, None)
for st in (None, 64, 156, 400, 1000):
    stats = {}
    ops.score_topk(emb, heads, None, k, tail_index=ti, sample_tiles=st, stats=stats)
    torch.cuda.synchronize()
    c = stats["candidates"].float()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        ops.score_topk(emb, heads, None, k, tail_index=ti, sample_tiles=st)
    t1.record(); torch.cuda.synchronize()
    print(f"sample_tiles={stats['sample_tiles']:5d} cand mean {c.mean():8.1f} max {c.max():8.0f} overflow {(c > stats['cap']).sum().item():4d}  {t0.elapsed_time(t1)/5:.3f} ms/batch")
