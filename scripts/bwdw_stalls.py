"""Cycle accounting of the opt-in fused backward kernel (POB_FUSED_BWD=1)."""
import os, sys
os.environ.setdefault("POB_FUSED_BWD", "1")
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plenoctree_b200.nerf.models import NerfModel, Rays, ctypes_ref
from plenoctree_b200.nerf import train as T
from plenoctree_b200._lib import check, lib, ptr
R = 4096
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
out = np.zeros((3, 256, 4), dtype=np.uint64)
check(lib.pob_debug_bwdw_stalls(ctypes_ref(model.cfg), ptr(model.workspace(True)), out.ctypes.data))
NP = int(os.environ.get("POB_BWDW_NP", "0")) or (148 * 59 + 50) // 100
for j, name in enumerate(["coarse", "fine", "sparsity"]):
    prod = out[j, :NP].astype(np.float64)
    cons = out[j, NP:148].astype(np.float64)
    print(name, "producers: total cycles mean %.0f, spin-on-consumers mean %.0f (%.0f%%), max %.0f" % (
        prod[:, 2].mean(), prod[:, 0].mean(), 100 * prod[:, 0].mean() / max(1, prod[:, 2].mean()), prod[:, 0].max()))
    for role in range(10):
        m = cons[cons[:, 3] == role]
        if len(m):
            print("   role %d x%d: total %.0f  wait-producer %.0f (%.0f%%)  wait-stage %.0f (%.0f%%)" % (
                role, len(m), m[:, 2].mean(), m[:, 0].mean(), 100 * m[:, 0].mean() / m[:, 2].mean(),
                m[:, 1].mean(), 100 * m[:, 1].mean() / m[:, 2].mean()))
