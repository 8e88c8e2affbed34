"""BASELINE configs[4] at N GPUs: octree.optimization, 256^3-equivalent SH16 tree, ray-parallel.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      scripts/bench_octree_dist.py [--depth 7] [--images 20] [--hw 800]

Every rank holds the tree, renders its pixel-row slab of each training image (fused render + MSE gradient + scatter),
the dense gradient is all-reduced (NCCL, SUM) and every rank applies the same SGD step: the reference's sequential
per-image SGD, exactly.  Timing: CUDA events, barrier + synchronize on both sides, max over ranks; rank 0 prints one
JSON line (images/s and rays/s of the whole job, plus the split march / all-reduce / SGD times of rank 0).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

from bench_octree import build_tree  # noqa: E402
from plenoctree_b200.nerf.utils import pose_spherical  # noqa: E402
from plenoctree_b200.octree import VolumeRenderer  # noqa: E402
from plenoctree_b200.octree.optimization import row_slab  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=7)
    ap.add_argument("--images", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--hw", type=int, default=800)
    ap.add_argument("--step", type=float, default=1e-4)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tree, n_occ, _ = build_tree(args.depth, dev)
    H = W = args.hw
    focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    rs = np.random.RandomState(20200823)
    poses = [pose_spherical(rs.uniform(-180, 180), rs.uniform(-90, 0), 4.0) for _ in range(8)]
    r = VolumeRenderer(tree, step_size=args.step)
    r0, nr = row_slab(H, rank, world)
    with torch.no_grad():
        gts = [(r.render_persp(p, W, H, focal, rows=(r0, nr)) + 0.05 * torch.randn((nr, W, 3), device=dev)).clamp_(0, 1)
               for p in poses]
    sq = torch.zeros(1, dtype=torch.float64, device=dev)
    g = tree.grad_buffer()[:tree.n_internal]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def image(i, timed=False):
        if timed:
            ev[0].record()
        r.train_persp(poses[i % 8], gts[i % 8], W, H, focal, rows=(r0, nr), sq_err=sq)
        if timed:
            ev[1].record()
        if world > 1:
            dist.all_reduce(g)
        if timed:
            ev[2].record()
        tree.sgd_step(1e-3)
        if timed:
            ev[3].record()

    for i in range(args.warmup):
        image(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(args.images):
        image(i)
    b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    image(0, timed=True)
    torch.cuda.synchronize()
    parts = [ev[k].elapsed_time(ev[k + 1]) for k in range(3)]
    if rank == 0:
        per = float(ms.item()) / args.images
        out = {"metric": "octree.optimization images/s (SH16 tree, ray-parallel, exact per-image SGD)",
               "value": 1e3 / per, "unit": "images/s", "rays_per_s": H * W / per * 1e3, "n_gpus": world,
               "ms_per_image": per, "scaling": "strong",
               "config": {"workload": f"configs[4]: {2 ** (args.depth + 1)}^3-equivalent SH16 tree, {H}x{W} images",
                          "nodes": tree.n_internal, "occupied_voxels": n_occ,
                          "grad_allreduce_mb": tree.n_internal * 8 * tree.data_dim * 4 / 1e6},
               "rank0_ms": {"march_grad_scatter": parts[0], "allreduce": parts[1], "sgd": parts[2]}}
        print(json.dumps(out), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"bench_octree_dist_n{world}.json"), "w") as f:
            json.dump(out, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
