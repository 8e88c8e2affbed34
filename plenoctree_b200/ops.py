"""Thin torch-tensor front end over the C ABI (device memory and streams only; no math here)."""
import ctypes

import numpy as np
import torch

from ._lib import PREC_FP16, PREC_FP16X3, check, lib, ptr, stream_ptr  # noqa: F401  (precision modes re-exported)
from .layouts import K_of


def _f32c(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise ValueError(f"{name} must be a contiguous float32 CUDA tensor")
    return t


def param_count(sh_deg):
    n = lib.pob_param_count(sh_deg)
    if n < 0:
        raise ValueError("sh_deg must be in [-1, 4]")
    return int(n)


def pack_weights(flat, sh_deg, out=None):
    """flat fp32 parameters of one MLP (reference order) -> packed operand blob (uint8 tensor)."""
    _f32c(flat, "flat")
    if flat.numel() != param_count(sh_deg):
        raise ValueError(f"expected {param_count(sh_deg)} parameters, got {flat.numel()}")
    nbytes = int(lib.pob_packed_bytes(sh_deg))
    if out is None:
        out = torch.zeros(nbytes, dtype=torch.uint8, device=flat.device)
    check(lib.pob_pack_weights(ptr(flat), sh_deg, ptr(out), stream_ptr()))
    return out


def eval_points_raw(blob, sh_deg, points, want_rgb=True, precision=PREC_FP16):
    """NerfModel.eval_points_raw (nerf_sh/nerf/models.py:143-181): -> (raw_rgb [M,3K] | None, raw_sigma [M,1])."""
    _f32c(points, "points")
    m = points.shape[0]
    K = K_of(sh_deg)
    rgb = torch.empty((m, 3 * K), dtype=torch.float32, device=points.device) if want_rgb else None
    sig = torch.empty((m, 1), dtype=torch.float32, device=points.device)
    check(lib.pob_eval_points_raw(ptr(blob), sh_deg, ptr(points), m, ptr(rgb), ptr(sig), precision,
                                  stream_ptr()))
    return rgb, sig


def eval_points(blob, sh_deg, points, viewdirs, precision=PREC_FP16):
    """NerfModel.eval_points (models.py:183-214): -> (rgb [M,3], sigma [M,1]) after sigmoid / relu."""
    _f32c(points, "points")
    if viewdirs is not None:
        _f32c(viewdirs, "viewdirs")
    m = points.shape[0]
    out = torch.empty((m, 4), dtype=torch.float32, device=points.device)
    check(lib.pob_eval_points(ptr(blob), sh_deg, ptr(points), ptr(viewdirs), m, ptr(out), precision,
                              stream_ptr()))
    return out[:, :3], out[:, 3:4]


def eval_cells_mean(blob, sh_deg, points, samples_per_cell, precision=PREC_FP16):
    """extraction step 2 (octree/extraction.py:367-394): points [n_cells, S, 3] -> [n_cells, 3K+1] means."""
    _f32c(points, "points")
    pts = points.reshape(-1, 3)
    if pts.shape[0] % samples_per_cell:
        raise ValueError("points must hold samples_per_cell points per cell")
    n_cells = pts.shape[0] // samples_per_cell
    out = torch.empty((n_cells, 3 * K_of(sh_deg) + 1), dtype=torch.float32, device=points.device)
    check(lib.pob_eval_cells_mean(ptr(blob), sh_deg, ptr(pts), n_cells, samples_per_cell, ptr(out), precision,
                                  stream_ptr()))
    return out


def eval_grid(blob, sh_deg, reso, offset, scale, x0=0, nx=None, ny=None, nz=None, want_rgb=False,
              precision=PREC_FP16, device="cuda"):
    """Dense-grid sweep of octree.extraction (octree/extraction.py:244-320) for one x-slab."""
    nx = reso - x0 if nx is None else nx
    ny = reso if ny is None else ny
    nz = reso if nz is None else nz
    m = nx * ny * nz
    K = K_of(sh_deg)
    rgb = torch.empty((m, 3 * K), dtype=torch.float32, device=device) if want_rgb else None
    sig = torch.empty((m,), dtype=torch.float32, device=device)
    off = (ctypes.c_float * 3)(*[float(v) for v in offset])
    sc = (ctypes.c_float * 3)(*[float(v) for v in scale])
    check(lib.pob_eval_grid(ptr(blob), sh_deg, reso, x0, nx, ny, nz, off, sc, ptr(rgb), ptr(sig),
                            precision, stream_ptr()))
    return rgb, sig


def umma_probe(a_img, b_img, b_off, adesc, bdesc, dcol, accum, idesc, out_cols, pair=False):
    """Run tcgen05.mma ops on raw smem images; returns the [128, out_cols] fp32 accumulator
    (pair=True: cta_group::2 on a CTA pair, images are [2, bytes], returns [256, out_cols])."""
    dev = "cuda"
    a = torch.from_numpy(np.ascontiguousarray(a_img)).to(dev)
    b = torch.from_numpy(np.ascontiguousarray(b_img)).to(dev)
    ad = torch.from_numpy(np.asarray(adesc, dtype=np.uint64).view(np.int64)).to(dev)
    bd = torch.from_numpy(np.asarray(bdesc, dtype=np.uint64).view(np.int64)).to(dev)
    dc = torch.from_numpy(np.asarray(dcol, dtype=np.uint32).view(np.int32)).to(dev)
    ac = torch.from_numpy(np.asarray(accum, dtype=np.uint32).view(np.int32)).to(dev)
    out = torch.zeros((256 if pair else 128, out_cols), dtype=torch.float32, device=dev)
    if pair:
        check(lib.pob_umma_probe_pair(ptr(a), a.numel() // 2, ptr(b), b.numel() // 2, b_off, ptr(ad), ptr(bd),
                                      ptr(dc), ptr(ac), len(adesc), idesc, out_cols, ptr(out), stream_ptr()))
    else:
        check(lib.pob_umma_probe(ptr(a), a.numel(), ptr(b), b.numel(), b_off, ptr(ad), ptr(bd), ptr(dc),
                                 ptr(ac), len(adesc), idesc, out_cols, ptr(out), stream_ptr()))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def grid_slab(reso, rank, world):
    """x-slab [x0, x0+nx) of the extraction grid owned by `rank` (voxel slabs, no collective)."""
    base, rem = divmod(reso, world)
    x0 = rank * base + min(rank, rem)
    return x0, base + (1 if rank < rem else 0)
