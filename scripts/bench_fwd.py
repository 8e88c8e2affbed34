"""Quick device-side timing of the fused point evaluator (not the contract bench; see bench.py)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nerf_sh_oracle as O  # noqa: E402
from plenoctree_b200 import ops  # noqa: E402

sh_deg = 3
flat = O.init_flat_params(sh_deg, 1, bias_scale=0.05)
blob = ops.pack_weights(torch.from_numpy(flat).cuda(), sh_deg)
res = {}
for m in (1 << 20, 1 << 22):
    pts = (torch.rand((m, 3), device="cuda") * 3 - 1.5).contiguous()
    for prec, name in ((ops.PREC_FP16, "fp16"), (ops.PREC_FP16X3, "fp16x3")):
        for want_rgb in (False, True):
            for _ in range(2):
                ops.eval_points_raw(blob, sh_deg, pts, want_rgb=want_rgb, precision=prec)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 5
            for _ in range(n):
                ops.eval_points_raw(blob, sh_deg, pts, want_rgb=want_rgb, precision=prec)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            tf = m * 1007104 / (ms * 1e-3) / 1e12
            key = f"m{m}_{name}_{'raw' if want_rgb else 'sigma'}"
            res[key] = dict(ms=ms, mpts_per_s=m / ms / 1e3, tflops_alg=tf)
            print(key, res[key], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_fwd.json", "w"), indent=1)
