import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from bench import make_patches_torch, H, W, B
import vpint_b200.vpint2 as V
from vpint_b200.engine import Solver
from vpint_b200.wp import wp_run_options
dev = torch.device("cuda", 0)
n = 128
target, feats = make_patches_torch(n, 1003, dev)
t_host = torch.empty(target.shape, dtype=torch.float32).pin_memory(); t_host.copy_(target)
f_host = torch.empty(feats.shape, dtype=torch.float32).pin_memory(); f_host.copy_(feats)
torch.cuda.synchronize()
opts = wp_run_options()
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T()
    s = Solver(2, 64 * B, H, W, nprob_inner=B, planes=True, k_block=0, device=dev)
    t1 = T()
    td = torch.empty((64, H, W, B), dtype=torch.float32, device=dev); fd = torch.empty_like(td)
    td.copy_(t_host[:64], non_blocking=True); fd.copy_(f_host[:64], non_blocking=True)
    t2 = T()
    s.load(td.permute(0, 3, 1, 2), fill=None)
    t3 = T()
    s.weights(fd.permute(0, 3, 1, 2).unsqueeze(-1), method="exact")
    t4 = T()
    s.run(5000)
    t5 = T()
    out, nit = s.result()
    t6 = T()
    s.close()
    t7 = T()
    print("rep", rep, "create %.1f h2d %.1f load %.1f weights %.1f run %.1f result %.1f close %.1f ms" % tuple((b - a) * 1e3 for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5), (t5, t6), (t6, t7))))
