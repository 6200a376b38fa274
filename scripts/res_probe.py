"""Timing probe for the resident kernel (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import make_patches_torch, H, W, B
from vpint_b200.engine import Solver
from vpint_b200.vpint2 import _band_fill_values

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else -1
dev = torch.device("cuda", 0)
target, feats = make_patches_torch(n, 1003, dev)[:2]
fill = _band_fill_values(target.cpu().numpy()).reshape(-1)
s = Solver(2, n * B, H, W, nprob_inner=B, planes=True, k_block=0, device=dev)
def ev(): return torch.cuda.Event(enable_timing=True)
for rep in range(3):
    e = [ev() for _ in range(5)]
    e[0].record(); s.load(target.permute(0, 3, 1, 2), fill=fill)
    e[1].record(); s.weights(feats.permute(0, 3, 1, 2).unsqueeze(-1), method="exact")
    e[2].record()
    if iters >= 0: s.run(iters, auto_terminate=False)
    else: s.run(5000, auto_terminate=True)
    e[3].record(); out, nit = s.result()
    e[4].record(); torch.cuda.synchronize()
    tot = int(nit.sum().item())
    print("rep %d load %.3f ms weights %.3f ms run %.3f ms result %.3f ms | sweeps total %d max %d -> %.3f us per problem-sweep/cluster33" % (
        rep, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), e[3].elapsed_time(e[4]), tot, int(nit.max().item()),
        e[2].elapsed_time(e[3]) * 1e3 / (tot / 33.0)))
