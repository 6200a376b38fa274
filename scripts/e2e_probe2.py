import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from bench import make_patches_torch, H, W, B
import vpint_b200.vpint2 as V
from vpint_b200.wp import wp_run_options
dev = torch.device("cuda", 0)
n = 512
target, feats = make_patches_torch(n, 1003, dev)
t_host = torch.empty(target.shape, dtype=torch.float32).pin_memory(); t_host.copy_(target)
f_host = torch.empty(feats.shape, dtype=torch.float32).pin_memory(); f_host.copy_(feats)
out_host = torch.empty((n, B, H, W), dtype=torch.float64).pin_memory()
torch.cuda.synchronize()
opts = wp_run_options()
# instrument _solve_chunk
orig = V._solve_chunk
log = []
def timed_chunk(solver, t_dev, f_dev, opts_, res_dev, out_layout, fill):
    t0 = time.perf_counter()
    r = orig(solver, t_dev, f_dev, opts_, res_dev, out_layout, fill)
    log.append((t0, time.perf_counter()))
    return r
V._solve_chunk = timed_chunk
for rep in range(3):
    log.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res, sw = V.solve_bands(t_host, f_host, opts, out_layout="bhw", return_sweeps=True, device=dev, out=out_host)
    torch.cuda.synchronize(); t1 = time.perf_counter()
print("total %.1f ms" % ((t1 - t0) * 1e3))
print("chunk start/end (ms):", [(round((a - t0) * 1e3, 1), round((b - t0) * 1e3, 1)) for a, b in log])
# all-resident device step for comparison
t_dev, f_dev = t_host.cuda(), f_host.cuda()
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res, sw = V.solve_bands(t_dev, f_dev, opts, out_layout="bhw", return_sweeps=True, device=dev)
    torch.cuda.synchronize(); t1 = time.perf_counter()
print("device-resident single solve incl. D2H of result %.1f ms" % ((t1 - t0) * 1e3))
