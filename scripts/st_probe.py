import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from vpint_b200.engine import Solver
dev = torch.device("cuda", 0)
Hc, Wc, Tc = 2048, 2048, 64
g = torch.Generator(device=dev); g.manual_seed(1005)
yy = (torch.arange(Hc, device=dev, dtype=torch.float64) / Hc).view(-1, 1, 1)
xx = (torch.arange(Wc, device=dev, dtype=torch.float64) / Wc).view(1, -1, 1)
cube = 20.0 + 5.0 * torch.sin(6.0 * yy) * torch.cos(8.0 * xx) + 0.3 * torch.randn((Hc, Wc, Tc), generator=g, device=dev, dtype=torch.float64)
miss = torch.zeros((Hc, Wc, Tc), dtype=torch.bool, device=dev)
discs = torch.rand((Tc, 8, 3), generator=g, device=dev, dtype=torch.float64)
for t in range(Tc):
    m2 = torch.zeros((Hc, Wc), dtype=torch.bool, device=dev)
    for k in range(8):
        cy, cx, r = discs[t, k]
        rad = 0.05 + 0.12 * r
        m2 |= (yy[:, :, 0] - cy) ** 2 + (xx[:, :, 0] - cx) ** 2 < rad * rad
    miss[:, :, t] = m2
cube[miss] = float("nan")
s = Solver(3, 1, Hc, Wc, Tc, planes=False, device=dev)
s.load(cube.unsqueeze(0), fill=[20.0])
s.discounts(0.9, 0.9)
s.run(int(sys.argv[1]) if len(sys.argv) > 1 else 6, auto_terminate=False)
torch.cuda.synchronize()
print("done", s.launches)
