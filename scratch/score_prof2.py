import sys, torch
sys.path.insert(0, ".")
from literalkg_b200 import ops
torch.manual_seed(0)
n, g, b, k = 1_000_000, 256, 2048, 10
emb = torch.nn.functional.leaky_relu(torch.randn(n, g, device="cuda"), 0.01) * 0.3
heads = (torch.arange(b, device="cuda") * 487) % n
ti = ops.ScoreIndex(emb, None)
for st in (None, 64, 1000):
    for _ in range(2):
        ops.score_topk(emb, heads, None, k, tail_index=ti, sample_tiles=st)
torch.cuda.synchronize()
