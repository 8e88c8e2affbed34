"""`python -m plenoctree_b200.octree.evaluation` and `eval_octree` (octree/evaluation.py:75-123,
octree/nerf/utils.py:448-498): render every test view of a PlenOctree, PSNR / SSIM against the ground truth
LPIPS is added when its downloaded weights can be found (nerf/lpips.py), else reported as nan."""
import os

import numpy as np
import torch

from ..nerf.utils import compute_psnr, compute_ssim, save_img
from .n3tree import N3Tree
from .renderer import VolumeRenderer


def eval_octree(t, dataset, args, want_frames=False, lpips_fn=None, metrics=None):
    """utils.eval_octree (octree/nerf/utils.py:448-498): -> (avg_psnr, avg_ssim[, frames]).  With `lpips_fn`
    (nerf/lpips.py::load_lpips) the mean LPIPS(gt, render) is left in `metrics["lpips"]`."""
    w, h, focal = dataset.w, dataset.h, dataset.focal
    r = VolumeRenderer(t, step_size=args.renderer_step_size)
    avg_psnr = avg_ssim = avg_lpips = 0.0
    frames = []
    with torch.no_grad():
        for idx in range(dataset.size):
            c2w = dataset.camtoworlds[idx]
            im_gt = torch.from_numpy(dataset.images[idx]).float().to(t.device)
            im = r.render_persp(c2w, width=w, height=h, fx=focal, fast=not args.no_early_stop).clamp_(0.0, 1.0)
            mse = float(((im - im_gt) ** 2).mean())
            avg_psnr += float(compute_psnr(mse))
            avg_ssim += float(compute_ssim(im, im_gt, max_val=1.0, padding="same"))   # octree/nerf/utils.py twin
            if lpips_fn is not None:
                avg_lpips += lpips_fn(im_gt.permute(2, 0, 1).contiguous(), im.permute(2, 0, 1).contiguous())
            if want_frames:
                frames.append((im.cpu().numpy() * 255).astype(np.uint8))
    n = max(dataset.size, 1)
    if metrics is not None:
        metrics["lpips"] = avg_lpips / n if lpips_fn is not None else float("nan")
    return (avg_psnr / n, avg_ssim / n, frames) if want_frames else (avg_psnr / n, avg_ssim / n)


def main(unused_argv):
    from ..nerf import datasets, flags as F
    F.define_flags(octree=True)
    F.define({"input": ("string", "./tree.npz", "Input octree npz"),
              "write_images": ("string", None, "If specified, writes rendered images to this directory")})
    FLAGS = F.FLAGS
    F.update_flags(FLAGS)
    dev = torch.device("cuda")
    dataset = datasets.get_dataset("test", FLAGS, device=dev)
    t = N3Tree.load(FLAGS.input, map_location=dev)
    from ..nerf.lpips import load_lpips
    extra = {}
    psnr, ssim, frames = eval_octree(t, dataset, FLAGS, want_frames=True, lpips_fn=load_lpips(dev), metrics=extra)
    print("Average PSNR", psnr, "SSIM", ssim, "LPIPS", extra["lpips"])
    if FLAGS.write_images:
        os.makedirs(FLAGS.write_images, exist_ok=True)
        for i, fr in enumerate(frames):
            save_img(fr.astype(np.float32) / 255.0, os.path.join(FLAGS.write_images, f"{i:04d}.png"))
    return psnr, ssim


if __name__ == "__main__":
    from absl import app
    app.run(main)
