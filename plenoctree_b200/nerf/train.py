"""train_step of the reference (nerf_sh/train.py:51-121) on the CUDA library.

    loss_fn + value_and_grad   -> lib.pob_loss_and_grad   (fused tcgen05 forward / dgrad / wgrad)
    lax.pmean(grad, "batch")   -> one torch.distributed all-reduce on the flat gradient (NCCL)
    optimizer.apply_gradient   -> lib.pob_adam_update      (flax Adam + operand re-pack)

Data parallel layout = one process per GPU; `batch` holds this rank's shard of the global batch
(reference: batch_size is global and split over devices, nerf_sh/nerf/utils.py:518-522, F8).
"""
import collections
import math

import numpy as np
import torch
import torch.distributed as dist

from .._lib import TrainHParams, check, lib, ptr, stream_ptr
from .models import _cuda_f32, ctypes_ref

# nerf_sh/nerf/utils.py:43-50
Stats = collections.namedtuple("Stats", ("loss", "psnr", "loss_c", "loss_sp", "psnr_c", "weight_l2"))


def learning_rate_decay(step, lr_init, lr_final, max_steps, lr_delay_steps=0, lr_delay_mult=1):
    """nerf_sh/nerf/utils.py:483-515."""
    if lr_delay_steps > 0:
        delay_rate = lr_delay_mult + (1 - lr_delay_mult) * np.sin(0.5 * np.pi * np.clip(step / lr_delay_steps, 0, 1))
    else:
        delay_rate = 1.0
    t = np.clip(step / max_steps, 0, 1)
    log_lerp = np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t)
    return delay_rate * log_lerp


class TrainState:
    """utils.TrainState(optimizer) (nerf_sh/nerf/utils.py:38-41): parameters live in model.params; the Adam
    moments and the step counter (flax optimizer.state.step) live here."""

    def __init__(self, model):
        self.model = model
        self.m = torch.zeros_like(model.params)
        self.v = torch.zeros_like(model.params)
        self.step = 0
        # gradient buffer: [params | 8 stats] so that one all-reduce carries both pmean calls
        self.gbuf = torch.zeros(model.params.numel() + 8, dtype=torch.float32, device=model.device)

    @property
    def grads(self):
        return self.gbuf[:self.model.params.numel()]

    @property
    def stats_raw(self):
        return self.gbuf[self.model.params.numel():]


def default_loss_scale(n_rays):
    """power-of-two scale that keeps the fp16 gradient chain in range: dL/dC = scale*2(C-px)/(3R)."""
    return float(2 ** int(round(math.log2(128.0 * max(1, n_rays)))))


def loss_and_grad(model, state, batch, sparsity_weight=1e-3, sparsity_length=0.05, sparsity_radius=1.5,
                  randomized=True, t_rand=None, u=None, sp_points=None, loss_scale=None, z_fine=None,
                  sigma_noise=None):
    """value_and_grad(loss_fn) for this rank's shard; fills state.grads / state.stats_raw (device)."""
    rays = batch["rays"]
    o = _cuda_f32(rays.origins, "rays.origins", 3)
    d = _cuda_f32(rays.directions, "rays.directions", 3)
    v = _cuda_f32(rays.viewdirs, "rays.viewdirs", 3)
    px = _cuda_f32(batch["pixels"], "pixels")[..., :3].contiguous()
    n = o.shape[0]
    t_rand, u, upr = model._uniforms(n, randomized, t_rand, u)
    use_sp = sparsity_weight > 0.0 and model.sparsity_npoints > 0
    if use_sp:
        if sp_points is None:
            # random.uniform(key, (npoints,3), minval=-radius, maxval=radius)  (train.py:79)
            sp_points = (torch.rand((model.sparsity_npoints, 3), device=model.device) * 2 - 1) * sparsity_radius
        sp_points = _cuda_f32(sp_points, "sp_points", 3)
        if sp_points.shape[0] != model.sparsity_npoints:
            raise ValueError("sp_points must have sparsity_npoints rows")
    z_fine = None if z_fine is None else _cuda_f32(z_fine, "z_fine")   # keep alive until the launch
    noise = model._set_sigma_noise(n, randomized, sigma_noise)          # noqa: F841  (same)
    ws = model.workspace(True)
    hp = TrainHParams(float(sparsity_weight if use_sp else 0.0), float(sparsity_length),
                      float(loss_scale or default_loss_scale(n)))
    check(lib.pob_loss_and_grad(ctypes_ref(model.cfg), ctypes_ref(hp), ptr(model.blobs[0]),
                                ptr(model.blobs[1]) if model.num_mlps == 2 else None, ptr(o), ptr(d), ptr(v),
                                ptr(px), n, ptr(model.z_base), ptr(t_rand), ptr(u), upr,
                                ptr(z_fine), ptr(sp_points) if use_sp else None, ptr(state.grads),
                                ptr(state.stats_raw), ptr(ws), stream_ptr()))
    return n


def stats_from_raw(raw, n_rays, sparsity_weight, sparsity_npoints, two_level, world=1):
    """device sums -> reference Stats (train.py:86-112); `raw` already averaged over ranks."""
    raw = [float(x) for x in raw.tolist()]
    loss = raw[0] / (3.0 * n_rays)
    loss_c = raw[1] / (3.0 * n_rays) if two_level else 0.0
    loss_sp = sparsity_weight * (1.0 - raw[2] / sparsity_npoints) if sparsity_npoints > 0 and sparsity_weight > 0 else 0.0
    psnr = -10.0 * math.log10(loss) if loss > 0 else float("inf")
    psnr_c = (-10.0 * math.log10(loss_c) if loss_c > 0 else float("inf")) if two_level else 0.0
    return Stats(loss, psnr, loss_c, loss_sp, psnr_c, float("nan"))


def allreduce_gradients(gbuf):
    """lax.pmean(grad) + lax.pmean(stats) (nerf_sh/train.py:117-118) as ONE all-reduce(SUM) on the flat
    [grads | stats] buffer; returns the world size whose reciprocal the caller folds into Adam / stats."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(gbuf, op=dist.ReduceOp.SUM)
        return dist.get_world_size()
    return 1


def shard_batch(batch_size, rank, world):
    """reference semantics: batch_size is global and split evenly over devices (utils.py:518-522,252)."""
    if batch_size % world != 0:
        raise ValueError("Batch size must be divisible by the number of devices.")
    per = batch_size // world
    return rank * per, (rank + 1) * per


def train_step(model, state, batch, lr, sparsity_weight=1e-3, sparsity_length=0.05, sparsity_radius=1.5,
               weight_decay_mult=0.0, randomized=True, t_rand=None, u=None, sp_points=None, loss_scale=None,
               sync_stats=False):
    """One optimisation step (nerf_sh/train.py:51-121).  Returns Stats when sync_stats (forces a
    device->host read of the six scalars, like the reference's periodic logging), else None."""
    n = loss_and_grad(model, state, batch, sparsity_weight, sparsity_length, sparsity_radius, randomized, t_rand, u,
                      sp_points, loss_scale)
    world = allreduce_gradients(state.gbuf)   # pmean(grad) and pmean(stats) in one bucket
    # weight_l2 = sum(theta^2)/numel  ->  d/dtheta = 2*theta/numel  (train.py:101-108,114)
    wd = 2.0 * weight_decay_mult / model.params.numel() if weight_decay_mult else 0.0
    check(lib.pob_adam_update(model.sh_deg, model.num_mlps, ptr(model.params), ptr(state.grads), ptr(state.m),
                              ptr(state.v), float(lr), float(state.step), 1.0 / world, wd, ptr(model.blobs[0]),
                              ptr(model.blobs[1]) if model.num_mlps == 2 else None, stream_ptr()))
    state.step += 1
    if sync_stats:
        raw = state.stats_raw / world
        st = stats_from_raw(raw, n, sparsity_weight, model.sparsity_npoints if sparsity_weight > 0 else 0,
                            model.num_mlps == 2)
        wl2 = float((model.params.double() ** 2).sum() / model.params.numel())
        return st._replace(weight_l2=wl2)
    return None
