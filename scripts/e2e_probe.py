import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from bench import make_patches_torch, H, W, B, bind_to_gpu_node
print("affinity:", bind_to_gpu_node(0))
import vpint_b200.vpint2 as V
from vpint_b200.wp import wp_run_options
dev = torch.device("cuda", 0)
n = 512
target, feats = make_patches_torch(n, 1003, dev)[:2]
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
# both directions at once (what the pipeline asks of the link): 3.49 GB in + 3.49 GB out
d_in = torch.empty((n, H, W, B), dtype=torch.float32, device=dev); d_in2 = torch.empty_like(d_in)
d_out = torch.empty((n, B, H, W), dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s1):
        d_in.copy_(t_host, non_blocking=True); d_in2.copy_(f_host, non_blocking=True)
    with torch.cuda.stream(s2):
        out_host.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("bidirectional: 3.49 GB in + 3.49 GB out in %.1f ms (%.1f GB/s per direction)" % (dt * 1e3, t_host.numel() * 8 / dt / 1e9))
del d_in, d_in2, d_out
opts = wp_run_options()
import itertools
for workers, chunk in itertools.product((6,), (16,)):
    V.PIPELINE_CHUNK_PATCHES = chunk
    V.PIPELINE_WORKERS = workers
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res, sw = V.solve_bands(t_host, f_host, opts, out_layout="bhw", return_sweeps=True, device=dev, out=out_host)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("workers", workers, "chunk", chunk, "e2e %.1f ms" % (dt * 1e3), "sweeps", int(sw.sum()))
