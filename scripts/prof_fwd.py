import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nerf_sh_oracle as O
from plenoctree_b200 import ops
flat = O.init_flat_params(3, 1, bias_scale=0.05)
blob = ops.pack_weights(torch.from_numpy(flat).cuda(), 3)
m = 1 << 20
pts = (torch.rand((m, 3), device="cuda") * 3 - 1.5).contiguous()
for _ in range(3):
    ops.eval_points_raw(blob, 3, pts, want_rgb=False)
torch.cuda.synchronize()
