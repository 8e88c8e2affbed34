"""Quick device-side timing of one training step (not the contract bench; see bench.py)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plenoctree_b200.nerf.models import NerfModel, Rays  # noqa: E402
from plenoctree_b200.nerf import train as T  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
model = NerfModel(sh_deg=3, max_rays=R, sparsity_npoints=10000)
model.init_params(1)
state = T.TrainState(model)
g = torch.Generator(device="cuda").manual_seed(0)
o = torch.randn((R, 3), device="cuda", generator=g) * 0.1 + torch.tensor([0.0, 0.0, 4.0], device="cuda")
d = torch.randn((R, 3), device="cuda", generator=g) * 0.2 + torch.tensor([0.0, 0.0, -1.0], device="cuda")
v = d / d.norm(dim=-1, keepdim=True)
px = torch.rand((R, 3), device="cuda", generator=g)
batch = {"rays": Rays(o, d, v), "pixels": px}
for _ in range(3):
    T.train_step(model, state, batch, 5e-4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
t0 = time.time()
e0.record()
for _ in range(n):
    T.train_step(model, state, batch, 5e-4)
e1.record()
torch.cuda.synchronize()
wall = (time.time() - t0) / n * 1e3
ms = e0.elapsed_time(e1) / n
flops = R * 756.9e6 + 29.57e9
res = dict(rays=R, ms_per_step=ms, wall_ms=wall, rays_per_s=R / ms * 1e3, tflops_alg=flops / ms / 1e9)
print(res)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/bench_train_quick_{R}.json", "w"))
