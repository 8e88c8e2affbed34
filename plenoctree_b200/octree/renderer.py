"""Host-side mirror of `svox.VolumeRenderer` for the calls the reference makes.

    reference call                                                   here
    svox.VolumeRenderer(t, step_size=..., ndc=None)                  VolumeRenderer(tree, step_size, ...)
      octree/optimization.py:174, octree/nerf/utils.py:456
    r.render_persp(c2w, height=H, width=W, fx=focal, fast=False)     render_persp (autograd-aware: the image
      octree/optimization.py:178,202, octree/nerf/utils.py:471        carries a grad_fn that fills tree.data.grad)
    r.forward(rays)  (svox.Rays(origins, dirs, viewdirs))            forward / __call__
    mse.backward(); optimizer.step()  optimization.py:205-208        train_persp + N3Tree.sgd_step: one launch for
                                                                     render + clamp-MSE gradient + scatter

All arithmetic is in the CUDA library (csrc/octree.cu) behind include/plenoctree_b200.h; torch carries device
memory, streams and the autograd edge only.  NDC rays (LLFF) are outside the scope of this path.
"""
import collections
import ctypes

import numpy as np
import torch

from .. import _lib
from .._lib import check, lib, ptr, stream_ptr

Rays = collections.namedtuple("Rays", ("origins", "dirs", "viewdirs"))


def make_camera(c2w, width, height, fx, fy=None):
    c2w = c2w.detach().cpu().numpy() if isinstance(c2w, torch.Tensor) else np.asarray(c2w)
    c2w = np.asarray(c2w, dtype=np.float32)
    if c2w.shape not in ((4, 4), (3, 4)):
        raise ValueError("c2w must be [4,4] or [3,4]")
    cam = _lib.Camera()
    for i in range(3):
        for j in range(4):
            cam.c2w[4 * i + j] = float(c2w[i, j])
    cam.fx = float(fx)
    cam.fy = float(fx if fy is None else fy)
    cam.width = float(int(width))
    cam.height = float(int(height))
    return cam


def camera_array(c2ws, width, height, fx, fy=None, device="cuda"):
    """[n,16] float32 device array of pob_camera records (pob_grid_weight_render)."""
    c2ws = np.asarray(c2ws, dtype=np.float32)
    n = c2ws.shape[0]
    out = np.zeros((n, 16), dtype=np.float32)
    out[:, :12] = c2ws[:, :3, :4].reshape(n, 12)
    out[:, 12] = fx
    out[:, 13] = fx if fy is None else fy
    out[:, 14] = int(width)
    out[:, 15] = int(height)
    return torch.from_numpy(out).to(device)


class _RenderFn(torch.autograd.Function):
    """autograd edge: d loss / d tree.data through pob_octree_render_backward."""

    @staticmethod
    def forward(ctx, data, renderer, rays, cam, row0, nrows, opts):
        ctx.renderer, ctx.rays, ctx.cam, ctx.row0, ctx.nrows = renderer, rays, cam, row0, nrows
        return renderer._render_raw(rays, cam, row0, nrows, opts)

    @staticmethod
    def backward(ctx, grad_out):
        r = ctx.renderer
        tree = r.tree
        g = torch.zeros_like(tree.data)
        t = tree.c_struct()
        o = r._opts(False)
        go = grad_out.reshape(-1, 3).contiguous().float()
        if ctx.cam is None:
            ro, rd, rv = ctx.rays
            check(lib.pob_octree_render_backward(ctypes.byref(t), ctypes.byref(o), ptr(ro), ptr(rd), ptr(rv),
                                                 ro.shape[0], None, 0, 0, ptr(go), ptr(g), stream_ptr()))
        else:
            check(lib.pob_octree_render_backward(ctypes.byref(t), ctypes.byref(o), None, None, None, 0,
                                                 ctypes.byref(ctx.cam), ctx.row0, ctx.nrows, ptr(go), ptr(g),
                                                 stream_ptr()))
        return g, None, None, None, None, None, None


class VolumeRenderer:
    def __init__(self, tree, step_size=1e-3, background_brightness=1.0, ndc=None):
        if ndc is not None:
            raise NotImplementedError("NDC rays (LLFF) are outside the scope of this path")
        self.tree = tree
        self.step_size = float(step_size)
        self.background_brightness = float(background_brightness)

    def _opts(self, fast):
        o = _lib.OctreeOpts()
        o.step_size = self.step_size
        o.background_brightness = self.background_brightness
        o.sigma_thresh = 1e-2 if fast else 0.0   # svox VolumeRenderer._get_options(fast)
        o.stop_thresh = 1e-2 if fast else 0.0
        return o

    def _render_raw(self, rays, cam, row0, nrows, opts, counters=None):
        tree = self.tree
        t = tree.c_struct()
        if cam is None:
            ro, rd, rv = rays
            n = ro.shape[0]
            out = torch.empty((n, 3), dtype=torch.float32, device=tree.device)
            check(lib.pob_octree_render(ctypes.byref(t), ctypes.byref(opts), ptr(ro), ptr(rd), ptr(rv), n, None, 0, 0,
                                        ptr(out), ptr(counters), stream_ptr()))
            return out
        W = int(cam.width)
        out = torch.empty((nrows, W, 3), dtype=torch.float32, device=tree.device)
        check(lib.pob_octree_render(ctypes.byref(t), ctypes.byref(opts), None, None, None, 0, ctypes.byref(cam), row0,
                                    nrows, ptr(out), ptr(counters), stream_ptr()))
        return out

    @staticmethod
    def _f32(t, dev):
        if isinstance(t, np.ndarray):
            t = torch.from_numpy(t)
        return t.to(device=dev, dtype=torch.float32).reshape(-1, 3).contiguous()

    def forward(self, rays, fast=False, counters=None):
        """VolumeRenderer.forward(rays: Rays(origins, dirs, viewdirs)) -> rgb [n,3]."""
        dev = self.tree.device
        r3 = (self._f32(rays.origins, dev), self._f32(rays.dirs, dev), self._f32(rays.viewdirs, dev))
        opts = self._opts(fast)
        data = self.tree.data
        if torch.is_grad_enabled() and data.requires_grad:
            return _RenderFn.apply(data, self, r3, None, 0, 0, opts)
        return self._render_raw(r3, None, 0, 0, opts, counters)

    __call__ = forward

    def render_persp(self, c2w, width=256, height=256, fx=1111.111, fy=None, fast=False, cuda=True, rows=None,
                     counters=None):
        """render_persp(c2w, width, height, fx) -> [H,W,3] (octree/optimization.py:178,202).  rows=(row0,nrows)
        renders one pixel-row slab (rank sharding of evaluation renders)."""
        cam = make_camera(c2w, width, height, fx, fy)
        row0, nrows = (0, int(height)) if rows is None else rows
        opts = self._opts(fast)
        data = self.tree.data
        if torch.is_grad_enabled() and data.requires_grad:
            return _RenderFn.apply(data, self, None, cam, row0, nrows, opts)
        return self._render_raw(None, cam, row0, nrows, opts, counters)

    def train_persp(self, c2w, gt, width, height, fx, fy=None, rows=None, want_image=False, sq_err=None):
        """One training image of octree.optimization (octree/optimization.py:201-207) in ONE kernel: render the slab,
        mse = mean((clamp(im,0,1) - gt)^2), scatter d mse / d data into tree.grad_buffer().  gt: [H,W,3] (or the
        slab's rows).  Returns (sum of squared errors as a 1-element float64 device tensor, image or None)."""
        tree = self.tree
        cam = make_camera(c2w, width, height, fx, fy)
        H, W = int(height), int(width)
        row0, nrows = (0, H) if rows is None else rows
        gt = gt.to(device=tree.device, dtype=torch.float32)
        if gt.shape[0] == H and nrows != H:
            gt = gt[row0:row0 + nrows]
        gt = gt.reshape(-1, 3).contiguous()
        if gt.shape[0] != nrows * W:
            raise ValueError("gt does not match the rendered slab")
        g = tree.grad_buffer()
        if sq_err is None:
            sq_err = torch.zeros(1, dtype=torch.float64, device=tree.device)
        out = torch.empty((nrows, W, 3), dtype=torch.float32, device=tree.device) if want_image else None
        t = tree.c_struct()
        o = self._opts(False)
        check(lib.pob_octree_train_persp(ctypes.byref(t), ctypes.byref(o), ctypes.byref(cam), row0, nrows, ptr(gt),
                                         1.0 / float(H * W * 3), ptr(g), ptr(sq_err), ptr(out), stream_ptr()))
        return sq_err, out
