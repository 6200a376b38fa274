import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from bench import make_patches_torch, H, W, B, bind_to_gpu_node
print("affinity:", bind_to_gpu_node(0))
import vpint_b200.vpint2 as V
from vpint_b200.wp import wp_run_options
dev = torch.device("cuda", 0)
n = 512
target, feats = make_patches_torch(n, 1003, dev)
t_host = torch.empty(target.shape, dtype=torch.float32).pin_memory(); t_host.copy_(target)
f_host = torch.empty(feats.shape, dtype=torch.float32).pin_memory(); f_host.copy_(feats)
out_host = torch.empty((n, B, H, W), dtype=torch.float64).pin_memory()
torch.cuda.synchronize()
# raw PCIe
d = torch.empty_like(target)
for name, fn in (("h2d", lambda: d.copy_(t_host, non_blocking=True)), ("d2h", lambda: t_host.copy_(d, non_blocking=True))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(name, "%.1f GB/s" % (t_host.numel() * 4 / dt / 1e9))
opts = wp_run_options()
import itertools
for workers, chunk in itertools.product((4, 6, 8), (16, 32, 64)):
    V.PIPELINE_CHUNK_PATCHES = chunk
    V.PIPELINE_WORKERS = workers
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res, sw = V.solve_bands(t_host, f_host, opts, out_layout="bhw", return_sweeps=True, device=dev, out=out_host)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("workers", workers, "chunk", chunk, "e2e %.1f ms" % (dt * 1e3), "sweeps", int(sw.sum()))
