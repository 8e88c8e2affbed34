"""Cycle timeline of one CTA of mlp_fwd (inference, FP16): MMA issuer vs epilogue warps."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nerf_sh_oracle as O
from plenoctree_b200 import ops
from plenoctree_b200._lib import check, lib, ptr
flat = O.init_flat_params(3, 1, bias_scale=0.05)
blob = ops.pack_weights(torch.from_numpy(flat).cuda(), 3)
m = 148 * 256 * int(os.environ.get('TRACE_ITERS', '4'))
pts = (torch.rand((m, 3), device="cuda") * 3 - 1.5).contiguous()
sig = torch.empty(m, device="cuda")
tr = torch.zeros((5, 256), dtype=torch.int64, device="cuda")
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
saving = len(sys.argv) > 2 and sys.argv[2] == "save"
tiles = (m + 255) // 256 * 2
sh = torch.empty(tiles * 8 * 65536, dtype=torch.uint8, device="cuda") if saving else None
se = torch.empty(tiles * 16384, dtype=torch.uint8, device="cuda") if saving else None
sm = torch.empty(8 * tiles * 128 * 8, dtype=torch.int32, device="cuda") if saving else None
for _ in range(2):
    tr.zero_()
    check(lib.pob_debug_trace_fwd(ptr(blob), 3, ptr(pts), m, ptr(sig), ptr(tr), flags, ptr(sh), ptr(se), ptr(sm), None))
torch.cuda.synchronize()
print("debug_flags", flags)
t = tr.cpu().numpy()
t0 = t[t > 0].min()
mma = t[0][t[0] > 0] - t0
e0 = t[1][t[1] > 0] - t0
e1 = t[2][t[2] > 0] - t0
print("MMA stamps per layer: [a_ready X, a_ready Y, issued]")
mm = mma[: (len(mma) // 3) * 3].reshape(-1, 3)
ee0 = e0[: (len(e0) // 4) * 4].reshape(-1, 4)
mm2 = mma[: (len(mma) // 2) * 2].reshape(-1, 2)     # [a_ready X observed, layer issued]
print("saving", saving)
for k in range(min(18, len(ee0))):
    a = ee0[k]
    li = k + k // 8          # epilogue index k -> layer-step index (heads step has no epilogue stamps)
    print("epi", k, "d_ready", a[0], "first_chunk", a[1]-a[0], "rest_7", a[2]-a[1], "signal", a[3]-a[2],
          "| MMA a_rdyX", mm2[li][0], "issued", mm2[li][1], "phase", mm2[li][1]-mm2[li][0])
iss = t[3][t[3] > 0] - t0
obs = t[4][t[4] > 0] - t0
n = min(len(iss), len(obs), 60)
print("slot: issued, observed_full, latency, gap_between_observed")
for i in range(n):
    print(i, iss[i], obs[i], obs[i] - iss[i], (obs[i] - obs[i-1]) if i else 0)
print("iteration period (cycles):", mm2[9][0] - mm2[0][0] if len(mm2) > 9 else None)
print("all iteration periods:", [int(mm2[9*(k+1)][0] - mm2[9*k][0]) for k in range((len(mm2)-1)//9)])
sys.exit(0)
print("epi stamps per trunk layer: [d_ready, drained, signalled]  (heads epilogue has no stamps)")
print("layer-step | MMA a_rdyX a_rdyY issued | epiX d_ready drained signalled | epiY d_ready drained signalled")
n = min(len(mm), 6)
k = 0
for i in range(n):
    line = f"{i:3d} | {mm[i][0]:8d} {mm[i][1]:8d} {mm[i][2]:8d} |"
    li = i % 9
    if li < 8 and k < len(ee0) and k < len(ee1):
        line += f" {ee0[k][0]:8d} {ee0[k][1]:8d} {ee0[k][2]:8d} | {ee1[k][0]:8d} {ee1[k][1]:8d} {ee1[k][2]:8d}"
        line += f"   epi={ee0[k][1]-ee0[k][0]} sig={ee0[k][2]-ee0[k][1]} mma_wait->issue={mm[i][2]-mm[i][0]} d_ready-issued={ee0[k][0]-mm[i][2]}"
        k += 1
    print(line)
