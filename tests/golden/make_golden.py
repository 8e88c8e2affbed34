"""Generate the golden vectors under tests/golden/ by IMPORTING THE REFERENCE (read-only tree at
/root/reference).  Runs only in the build container; the GPU box has no reference tree, it uses the
committed .npz files.

    python tests/golden/make_golden.py

What is generated (all fp32, seeds fixed):
  eval_points_sh16.npz / eval_points_sh25.npz
      reference torch twin octree.nerf.models.NerfModel.eval_points_raw (octree/nerf/models.py:211)
      on weights produced by oracle.init_flat_params(seed) (loaded into the twin's nn.Linear
      modules, kernels transposed like octree/nerf/models.py:79-102 does for flax checkpoints).
  eval_sh.npz      reference nerf_sh/nerf/sh.py::eval_sh for deg 0..4 on random coefficients/dirs.
  posenc.npz       reference octree/nerf/model_utils.py::posenc.
  rays.npz         reference octree/nerf/utils.py::generate_rays on two spherical poses.
  ref_render.npz   /root/reference/nerf_sh/nerf/model_utils.py and models.py EXECUTED UNMODIFIED over numpy-backed
                   stand-ins for jax / flax (tests/golden/jax_stub.py): sample_along_rays (plain and with an injected
                   t_rand), volumetric_rendering, piecewise_constant_pdf (peaky / flat / zero / single-bin weights,
                   deterministic and injected u), sample_pdf, add_gaussian_noise, the reference MLP class, and the
                   whole NerfModel.__call__ (64 + 128 samples, SH16, white background, randomized with injected
                   draws) on oracle-initialised weights.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import nerf_sh_oracle as O  # noqa: E402


def load_ref_module(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _use_reference_octree():
    """`octree` must resolve to the reference's package, not to this repo's drop-in shim of the same name (a regular
    package beats the reference's namespace package whatever the sys.path order): register a package object whose
    search path is the reference's directory, and forget any shim modules imported so far."""
    import types
    cur = sys.modules.get("octree")
    if cur is not None and list(getattr(cur, "__path__", [])) == [os.path.join(REF, "octree")]:
        return
    for name in [m for m in sys.modules if m == "octree" or m.startswith("octree.")]:
        del sys.modules[name]
    pkg = types.ModuleType("octree")
    pkg.__path__ = [os.path.join(REF, "octree")]
    sys.modules["octree"] = pkg


def ref_model(sh_deg, flat_c, flat_f):
    _use_reference_octree()
    from octree.nerf import models as ref_models
    K = (sh_deg + 1) ** 2
    model = ref_models.NerfModel(use_viewdirs=False, sh_deg=sh_deg, num_rgb_channels=3 * K,
                                 num_coarse_samples=64, num_fine_samples=128)
    for name, flat in (("MLP_0", flat_c), ("MLP_1", flat_f)):
        mlp = getattr(model, name)
        params = O.unflatten(flat, sh_deg)
        with torch.no_grad():
            for i in range(8):
                mlp.input_layers[i].weight.copy_(params[i][0].T)
                mlp.input_layers[i].bias.copy_(params[i][1])
            mlp.sigma_layer.weight.copy_(params[8][0].T)
            mlp.sigma_layer.bias.copy_(params[8][1])
            mlp.rgb_layer.weight.copy_(params[9][0].T)
            mlp.rgb_layer.bias.copy_(params[9][1])
    return model.eval()


def gen_eval_points(sh_deg, n, seed, fname):
    flat_c = O.init_flat_params(sh_deg, seed, bias_scale=0.05)
    flat_f = O.init_flat_params(sh_deg, seed + 1, bias_scale=0.05)
    rs = np.random.RandomState(seed + 2)
    pts = rs.uniform(-1.5, 1.5, size=(n, 3)).astype(np.float32)
    pts[: n // 8] *= 3.0  # far points (|x| up to 4.5) stress the 2^9 posenc octave
    pts[0] = 0.0
    model = ref_model(sh_deg, flat_c, flat_f)
    with torch.no_grad():
        rgb_f, sig_f = model.eval_points_raw(torch.from_numpy(pts))
        rgb_c, sig_c = model.eval_points_raw(torch.from_numpy(pts), coarse=True)
    np.savez_compressed(os.path.join(HERE, fname), sh_deg=sh_deg, seed=seed, points=pts,
                        raw_rgb_fine=rgb_f.numpy(), raw_sigma_fine=sig_f.numpy(),
                        raw_rgb_coarse=rgb_c.numpy(), raw_sigma_coarse=sig_c.numpy(),
                        flat_c_checksum=np.float64(flat_c.astype(np.float64).sum()),
                        flat_f_checksum=np.float64(flat_f.astype(np.float64).sum()))
    print(fname, rgb_f.shape, float(sig_f.abs().mean()))


def gen_eval_sh():
    ref_sh = load_ref_module("ref_sh", "nerf_sh/nerf/sh.py")
    rs = np.random.RandomState(7)
    out = {}
    for deg in range(5):
        K = (deg + 1) ** 2
        sh = rs.normal(size=(64, 3, K)).astype(np.float32)
        d = rs.normal(size=(64, 3)).astype(np.float32)
        d /= np.linalg.norm(d, axis=-1, keepdims=True)
        out[f"sh{deg}"] = sh
        out[f"dirs{deg}"] = d
        out[f"res{deg}"] = ref_sh.eval_sh(deg, sh, d).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "eval_sh.npz"), **out)
    print("eval_sh.npz")


def gen_posenc():
    from octree.nerf import model_utils as ref_mu
    rs = np.random.RandomState(11)
    x = rs.uniform(-4, 4, size=(256, 3)).astype(np.float32)
    enc = ref_mu.posenc(torch.from_numpy(x), 0, 10).numpy()
    np.savez_compressed(os.path.join(HERE, "posenc.npz"), x=x, enc=enc)
    print("posenc.npz", enc.shape)


def gen_rays():
    ref_utils = load_ref_module("ref_octree_utils", "octree/nerf/utils.py")
    poses = np.stack([O.pose_spherical(30.0, -30.0, 4.0), O.pose_spherical(-120.0, -75.0, 4.0)])
    w, h, focal = 40, 30, 55.5
    rays = ref_utils.generate_rays(w, h, focal, poses)
    np.savez_compressed(os.path.join(HERE, "rays.npz"), poses=poses, w=w, h=h, focal=focal,
                        origins=np.asarray(rays.origins), directions=np.asarray(rays.directions),
                        viewdirs=np.asarray(rays.viewdirs))
    print("rays.npz")


def gen_ckpt_bridge():
    """Let the REFERENCE's own loader (octree/nerf/models.py:66-113 restore_model_state_from_jaxnerf) consume a
    flax-format checkpoint written by plenoctree_b200.nerf.checkpoints.  flax is not installed, so the one call the
    loader makes into it (flax.training.checkpoints.restore_checkpoint(train_dir, target=None)) is served by this
    package's msgpack reader; everything after that — key names, Dense index mapping, transposes, load_state_dict
    into the reference torch model, eval_points_raw — is reference code."""
    import hashlib
    import tempfile
    import types
    from octree.nerf import models as ref_models
    from plenoctree_b200.nerf import checkpoints as C
    sh_deg = 3
    flat_c = O.init_flat_params(sh_deg, 4101, bias_scale=0.05)
    flat_f = O.init_flat_params(sh_deg, 4102, bias_scale=0.05)
    flat = np.concatenate([flat_c, flat_f])
    rs = np.random.RandomState(4103)
    m = rs.normal(size=flat.shape).astype(np.float32) * 1e-3
    v = (rs.uniform(size=flat.shape).astype(np.float32) * 1e-6)
    step = 12345
    blob = C.msgpack_serialize(C.train_state_dict(flat, m, v, step, sh_deg))
    tmp = tempfile.mkdtemp()
    with open(os.path.join(tmp, f"checkpoint_{step}"), "wb") as f:
        f.write(blob)
    fake_flax = types.ModuleType("flax")
    fake_training = types.ModuleType("flax.training")
    fake_ckpt = types.ModuleType("flax.training.checkpoints")
    fake_ckpt.restore_checkpoint = lambda train_dir, target=None: C.restore_flax_state_dict(train_dir)
    fake_training.checkpoints = fake_ckpt
    fake_flax.training = fake_training
    sys.modules.update({"flax": fake_flax, "flax.training": fake_training, "flax.training.checkpoints": fake_ckpt})
    try:
        model = ref_models.NerfModel(use_viewdirs=False, sh_deg=sh_deg, num_rgb_channels=3 * 16,
                                     num_coarse_samples=64, num_fine_samples=128)
        args = types.SimpleNamespace(train_dir=tmp)
        model = ref_models.restore_model_state_from_jaxnerf(args, model).eval()
    finally:
        for k in ("flax", "flax.training", "flax.training.checkpoints"):
            sys.modules.pop(k, None)
    pts = rs.uniform(-1.5, 1.5, size=(64, 3)).astype(np.float32)
    with torch.no_grad():
        rgb_f, sig_f = model.eval_points_raw(torch.from_numpy(pts))
        rgb_c, sig_c = model.eval_points_raw(torch.from_numpy(pts), coarse=True)
    sd = model.state_dict()
    keys = sorted(sd.keys())
    np.savez_compressed(os.path.join(HERE, "ckpt_bridge.npz"), sh_deg=sh_deg, seeds=np.array([4101, 4102, 4103]),
                        step=step, sha256=hashlib.sha256(blob).hexdigest(), nbytes=len(blob), points=pts,
                        raw_rgb_fine=rgb_f.numpy(), raw_sigma_fine=sig_f.numpy(), raw_rgb_coarse=rgb_c.numpy(),
                        raw_sigma_coarse=sig_c.numpy(), keys=np.array(keys),
                        shapes=np.array([";".join(map(str, sd[k].shape)) for k in keys]),
                        sums=np.array([float(sd[k].double().sum()) for k in keys]),
                        abs_sums=np.array([float(sd[k].double().abs().sum()) for k in keys]))
    print("ckpt_bridge.npz", len(keys), "tensors,", len(blob), "bytes")


def gen_ref_render():
    """Execute the reference's JAX forward path (model_utils.py, models.py::NerfModel.__call__) over numpy."""
    import types
    import jax_stub
    names = jax_stub.install()
    # nerf_sh.nerf.utils pulls in flax.optim, jax.dlpack, datasets, ...: models.py only needs these two names of it
    fake_utils = types.ModuleType("nerf_sh.nerf.utils")
    import collections
    fake_utils.Rays = collections.namedtuple("Rays", ("origins", "directions", "viewdirs"))
    fake_utils.TrainState = object
    sys.modules["nerf_sh.nerf.utils"] = fake_utils
    try:
        from nerf_sh.nerf import model_utils as MU
        from nerf_sh.nerf import models as RM
        import flax.linen as nn
        rs = np.random.RandomState(20200823)
        out = {}
        B, N, NF = 24, 64, 128
        poses = np.stack([O.pose_spherical(rs.uniform(-180, 180), rs.uniform(-90, 0), 4.0) for _ in range(4)])
        rays_all = O.generate_rays(40, 30, 55.5, poses)
        pick = rs.choice(4 * 30 * 40, B, replace=False)
        o, d, v = [np.ascontiguousarray(np.asarray(r).reshape(-1, 3)[pick]).astype(np.float32) for r in rays_all]
        out.update(origins=o, directions=d, viewdirs=v)
        # ---- sample_along_rays ----
        t_rand = rs.uniform(size=(B, N)).astype(np.float32)
        for tag, rnd, lin in (("plain", False, False), ("rand", True, False), ("lindisp_rand", True, True)):
            z, pts = MU.sample_along_rays(jax_stub.Key(uniform=t_rand), o, d, N, 2.0, 6.0, rnd, lin)
            out[f"sar_{tag}_z"], out[f"sar_{tag}_pts"] = np.asarray(z), np.asarray(pts)
        out["t_rand"] = t_rand
        # ---- volumetric_rendering ----
        z_sorted = np.sort(rs.uniform(2.0, 6.0, size=(B, N)).astype(np.float32), axis=-1)
        rgb = rs.uniform(size=(B, N, 3)).astype(np.float32)
        sigma = (rs.exponential(3.0, size=(B, N, 1)) * (rs.uniform(size=(B, N, 1)) < 0.4)).astype(np.float32)
        sigma[0] = 0.0                    # an empty ray: the disp guard and the white background take over
        sigma[1, :, 0] = 80.0             # an opaque one
        for wb in (True, False):
            c, di, ac, w = MU.volumetric_rendering(rgb, sigma, z_sorted, d, wb)
            out[f"vr_rgb_{int(wb)}"], out[f"vr_disp_{int(wb)}"] = np.asarray(c), np.asarray(di)
            out[f"vr_acc_{int(wb)}"], out[f"vr_weights_{int(wb)}"] = np.asarray(ac), np.asarray(w)
        out.update(vr_in_rgb=rgb, vr_in_sigma=sigma, vr_in_z=z_sorted)
        # ---- piecewise_constant_pdf / sample_pdf ----
        bins = np.sort(rs.uniform(2.0, 6.0, size=(5, N - 1)).astype(np.float32), axis=-1)
        wts = rs.uniform(size=(5, N - 2)).astype(np.float32)
        wts[0] = wts[0] ** 8                           # peaky
        wts[1] = 0.37                                  # flat
        wts[2] = 0.0                                   # all zero: the eps padding decides
        wts[3] = 0.0
        wts[3, 17] = 1.0                               # a single bin
        u = rs.uniform(size=(5, NF)).astype(np.float32)
        u[4, :3] = (0.0, np.float32(1.0) - np.finfo(np.float32).eps, 0.5)
        out.update(pdf_bins=bins, pdf_weights=wts, pdf_u=u)
        out["pdf_det"] = np.asarray(MU.piecewise_constant_pdf(jax_stub.Key(), bins, wts.copy(), NF, False))
        out["pdf_rand"] = np.asarray(MU.piecewise_constant_pdf(jax_stub.Key(uniform=u), bins, wts.copy(), NF, True))
        zc = np.sort(rs.uniform(2.0, 6.0, size=(5, N)).astype(np.float32), axis=-1)
        zs, ps = MU.sample_pdf(jax_stub.Key(uniform=u), bins, wts.copy(), o[:5], d[:5], zc, NF, True)
        out.update(spdf_z_coarse=zc, spdf_z=np.asarray(zs), spdf_pts=np.asarray(ps))
        # ---- add_gaussian_noise ----
        raw = rs.normal(size=(B, N, 1)).astype(np.float32)
        nz = rs.normal(size=(B, N, 1)).astype(np.float32)
        out.update(noise_raw=raw, noise_draw=nz,
                   noise_on=np.asarray(MU.add_gaussian_noise(jax_stub.Key(normal=nz), raw, 0.7, True)),
                   noise_off_std=np.asarray(MU.add_gaussian_noise(jax_stub.Key(normal=nz), raw, None, True)),
                   noise_off_rand=np.asarray(MU.add_gaussian_noise(jax_stub.Key(normal=nz), raw, 0.7, False)))
        # ---- the reference MLP class and NerfModel.__call__ ----
        sh_deg = 3
        flat_c = O.init_flat_params(sh_deg, 5101, bias_scale=0.05)
        flat_f = O.init_flat_params(sh_deg, 5102, bias_scale=0.05)

        def plist(flat):
            return [(w.numpy(), b.numpy()) for w, b in O.unflatten(flat, sh_deg)]
        model = RM.NerfModel(num_coarse_samples=N, num_fine_samples=NF, use_viewdirs=False, sh_deg=sh_deg, sg_dim=-1,
                             near=2.0, far=6.0, noise_std=None, net_depth=8, net_width=256, net_depth_condition=1,
                             net_width_condition=128, net_activation=nn.relu, skip_layer=4, num_rgb_channels=48,
                             num_sigma_channels=1, white_bkgd=True, min_deg_point=0, max_deg_point=10, deg_view=4,
                             lindisp=False, rgb_activation=nn.sigmoid, sigma_activation=nn.relu,
                             legacy_posenc_order=False)
        model.MLP_0._params = plist(flat_c)
        model.MLP_1._params = plist(flat_f)
        enc = np.asarray(MU.posenc(rs.uniform(-1.5, 1.5, size=(3, 7, 3)).astype(np.float32), 0, 10))
        raw_rgb, raw_sigma = model.MLP_1(enc)
        out.update(mlp_enc=enc, mlp_raw_rgb=np.asarray(raw_rgb), mlp_raw_sigma=np.asarray(raw_sigma))
        u_f = rs.uniform(size=(B, NF)).astype(np.float32)
        for tag, rnd in (("det", False), ("rand", True)):
            ret = model(jax_stub.Key(uniform=t_rand), jax_stub.Key(uniform=u_f), fake_utils.Rays(o, d, v), rnd)
            for lvl, (c, di, ac) in zip(("coarse", "fine"), ret):
                out[f"call_{tag}_{lvl}_rgb"], out[f"call_{tag}_{lvl}_disp"] = np.asarray(c), np.asarray(di)
                out[f"call_{tag}_{lvl}_acc"] = np.asarray(ac)
        out.update(call_u=u_f, seeds=np.array([5101, 5102]), sh_deg=sh_deg)
    finally:
        jax_stub.uninstall(names)
        for k in [k for k in sys.modules if k.startswith("nerf_sh")]:
            sys.modules.pop(k, None)
    assert all(np.asarray(a).dtype != np.float64 for a in out.values()), "float64 leaked through the jnp stand-in"
    np.savez_compressed(os.path.join(HERE, "ref_render.npz"), **out)
    print("ref_render.npz", len(out), "arrays")


def gen_ref_loss():
    """Execute the reference's train_step (nerf_sh/train.py:51-121) up to its loss value: loss_fn — the two MSE
    terms, the sparsity term on injected points, weight_l2, the PSNRs — runs unmodified over the numpy stand-ins
    (jax.value_and_grad returns loss_fn's value and zero gradients; pmean and apply_gradient are identities)."""
    import collections
    import dataclasses
    import types
    import jax_stub
    names = jax_stub.install()
    fake_ds = types.ModuleType("nerf_sh.nerf.datasets")       # loaders: not on this path (define_flags lists the names)
    fake_ds.dataset_dict = {"blender": None, "llff": None, "nsvf": None}
    sys.modules["nerf_sh.nerf.datasets"] = fake_ds
    try:
        from absl import flags
        import nerf_sh.train as RT            # defines the reference's flags (utils.define_flags) at import
        from nerf_sh.nerf import models as RM, utils as RU
        import flax.linen as nn
        FLAGS = flags.FLAGS
        FLAGS(["make_golden"])
        sh_deg, B, N, NF, NSP = 3, 16, 64, 128, 40
        FLAGS.randomized = True
        FLAGS.sparsity_weight = 1e-3
        FLAGS.sparsity_npoints = NSP
        FLAGS.sparsity_radius = 1.5
        FLAGS.sparsity_length = 0.05
        FLAGS.weight_decay_mult = 0.25
        rs = np.random.RandomState(777)
        poses = np.stack([O.pose_spherical(rs.uniform(-180, 180), rs.uniform(-90, 0), 4.0) for _ in range(3)])
        rays_all = O.generate_rays(40, 30, 55.5, poses)
        pick = rs.choice(3 * 30 * 40, B, replace=False)
        o, d, v = [np.ascontiguousarray(np.asarray(r).reshape(-1, 3)[pick]).astype(np.float32) for r in rays_all]
        px = rs.uniform(size=(B, 3)).astype(np.float32)
        t_rand = rs.uniform(size=(B, N)).astype(np.float32)
        u_f = rs.uniform(size=(B, NF)).astype(np.float32)
        sp01 = rs.uniform(size=(NSP, 3)).astype(np.float32)
        flat_c = O.init_flat_params(sh_deg, 6101, bias_scale=0.05)
        flat_f = O.init_flat_params(sh_deg, 6102, bias_scale=0.05)

        def ptree(flat):
            return {f"Dense_{j}": {"kernel": w.numpy(), "bias": b.numpy()} for j, (w, b) in enumerate(O.unflatten(flat, sh_deg))}
        variables = {"params": {"MLP_0": ptree(flat_c), "MLP_1": ptree(flat_f)}}
        model = RM.NerfModel(num_coarse_samples=N, num_fine_samples=NF, use_viewdirs=False, sh_deg=sh_deg, sg_dim=-1,
                             near=2.0, far=6.0, noise_std=None, net_depth=8, net_width=256, net_depth_condition=1,
                             net_width_condition=128, net_activation=nn.relu, skip_layer=4, num_rgb_channels=48,
                             num_sigma_channels=1, white_bkgd=True, min_deg_point=0, max_deg_point=10, deg_view=4,
                             lindisp=False, rgb_activation=nn.sigmoid, sigma_activation=nn.relu,
                             legacy_posenc_order=False)

        @dataclasses.dataclass
        class Opt:
            target: dict
            def apply_gradient(self, grad, learning_rate=None):
                return self
        state = RU.TrainState(optimizer=Opt(variables))
        # train_step splits rng into (rng, key_0, key_1, key_2): every key carries the injected draws; they are told
        # apart by shape ([B,N] jitter, [B,NF] inverse-CDF uniforms, [NSP,3] sparsity points)
        class MultiKey(jax_stub.Key):
            pass
        import jax.random as jr
        table = {tuple(t_rand.shape): t_rand, tuple(u_f.shape): u_f, tuple(sp01.shape): sp01}
        orig_uniform = jr.uniform

        def uniform(key, shape, dtype=np.float32, minval=0.0, maxval=1.0):
            base = table[tuple(shape)]
            return (base * np.float32(maxval - minval) + np.float32(minval)).astype(np.float32)
        jr.uniform = uniform
        RT.random.uniform = uniform
        import nerf_sh.nerf.model_utils as MU
        MU.random.uniform = uniform
        batch = {"rays": RU.Rays(o, d, v), "pixels": px}
        _, stats, _ = RT.train_step(model, jax_stub.Key(seed=1), state, batch, 5e-4)
        jr.uniform = orig_uniform
        out = dict(origins=o, directions=d, viewdirs=v, pixels=px, t_rand=t_rand, u=u_f, sp01=sp01, sh_deg=sh_deg,
                   seeds=np.array([6101, 6102]), sparsity_weight=1e-3, sparsity_radius=1.5, sparsity_length=0.05,
                   weight_decay_mult=0.25, loss=np.float32(stats.loss), psnr=np.float32(stats.psnr),
                   loss_c=np.float32(stats.loss_c), psnr_c=np.float32(stats.psnr_c),
                   loss_sp=np.float32(stats.loss_sp), weight_l2=np.float32(stats.weight_l2))
    finally:
        jax_stub.uninstall(names)
        for k in [k for k in sys.modules if k.startswith("nerf_sh")]:
            sys.modules.pop(k, None)
    np.savez_compressed(os.path.join(HERE, "ref_loss.npz"), **out)
    print("ref_loss.npz", {k: float(out[k]) for k in ("loss", "loss_c", "loss_sp", "weight_l2", "psnr")})


def write_llff_scene(data_dir, images_u8, poses_bounds, factor=0):
    """images[_<factor>]/NNN.png + poses_bounds.npy, the layout nerf_sh/nerf/datasets.py:238-266 reads."""
    from PIL import Image
    sub = os.path.join(data_dir, "images" + (f"_{factor}" if factor > 0 else ""))
    os.makedirs(sub, exist_ok=True)
    for i, im in enumerate(images_u8):
        Image.fromarray(im, mode="RGB").save(os.path.join(sub, f"{i:03d}.png"))
    np.save(os.path.join(data_dir, "poses_bounds.npy"), poses_bounds)


def synthetic_llff(seed, n, h, w, focal, ring=False):
    """n cameras in the LLFF on-disk convention ([down, right, back | t | h,w,f] + near/far bounds): forward-facing
    around the origin looking down -z, or (ring) on a circle looking inwards."""
    rs = np.random.RandomState(seed)
    rows = []
    for i in range(n):
        if ring:
            a = 2 * np.pi * i / n + rs.uniform(-0.1, 0.1)
            pos = np.array([2.5 * np.cos(a), 2.5 * np.sin(a), 0.4 + rs.uniform(-0.2, 0.2)])
            back = pos / np.linalg.norm(pos)
            world_up = np.array([0.0, 0.0, 1.0])
        else:
            pos = rs.uniform(-0.6, 0.6, size=3) * np.array([1.0, 1.0, 0.2])
            back = np.array([0.0, 0.0, 1.0]) + rs.uniform(-0.15, 0.15, size=3)
            back /= np.linalg.norm(back)
            world_up = np.array([0.0, 1.0, 0.0])
        right = np.cross(world_up, back); right /= np.linalg.norm(right)
        up = np.cross(back, right)
        m = np.stack([-up, right, back, pos, np.array([h, w, focal], dtype=np.float64)], axis=1)      # [3,5]
        near = rs.uniform(1.2, 2.0)
        rows.append(np.concatenate([m.reshape(-1), [near, near * rs.uniform(4.0, 8.0)]]))
    images = rs.randint(0, 256, size=(n, h, w, 3)).astype(np.uint8)
    return images, np.stack(rows).astype(np.float64)


def gen_ref_llff():
    """Execute the reference's LLFF loader (nerf_sh/nerf/datasets.py:235-487: pose re-ordering, bound rescale,
    recentring, spiral / spherical render paths, llffhold split, NDC rays via convert_to_ndc :40-60) unmodified on
    synthetic scenes written to a temporary directory; jax only enters through `import jax` at module level."""
    import tempfile
    import types
    import jax_stub
    names = jax_stub.install()
    try:
        from nerf_sh.nerf import utils as RU            # pulls in the real nerf_sh.nerf.datasets
        RD = sys.modules["nerf_sh.nerf.datasets"]
        out = {}
        cases = {"fwd": dict(seed=11, n=9, h=12, w=16, focal=20.0, ring=False, factor=0, spherify=False),
                 "ring": dict(seed=12, n=10, h=10, w=14, focal=36.0, ring=True, factor=2, spherify=True)}
        for name, c in cases.items():
            images, pb = synthetic_llff(c["seed"], c["n"], c["h"], c["w"], c["focal"] * max(c["factor"], 1), c["ring"])
            out[f"{name}_images"], out[f"{name}_poses_bounds"] = images, pb
            with tempfile.TemporaryDirectory() as d:
                write_llff_scene(d, images, pb, c["factor"])
                for split in ("train", "test"):
                    args = types.SimpleNamespace(data_dir=d, factor=c["factor"], spherify=c["spherify"], llffhold=4,
                                                 render_path=(split == "test"))
                    ds = RD.LLFF.__new__(RD.LLFF)          # no thread / queue: the two loader stages only
                    ds.split = split
                    ds._load_renderings(args)
                    ds._generate_rays()
                    k = f"{name}_{split}_"
                    out[k + "images"] = np.asarray(ds.images, dtype=np.float32)
                    out[k + "camtoworlds"] = np.asarray(ds.camtoworlds)
                    out[k + "focal"] = np.asarray(ds.focal)
                    out[k + "hw_n"] = np.array([ds.h, ds.w, ds.n_examples])
                    for f, r in zip(("o", "d", "v"), ds.rays):
                        out[k + "rays_" + f] = np.asarray(r)
                    if split == "test":
                        out[k + "render_poses"] = np.asarray(ds.render_poses)
                        for f, r in zip(("o", "d", "v"), ds.render_rays):
                            out[k + "render_rays_" + f] = np.asarray(r)[::15]        # every 15th path pose
        # convert_to_ndc on free rays (near != 1 too)
        rs = np.random.RandomState(5)
        o = rs.uniform(-1, 1, size=(64, 3)).astype(np.float32)
        dd = rs.uniform(-1, 1, size=(64, 3)).astype(np.float32); dd[:, 2] = -np.abs(dd[:, 2]) - 0.2
        out["ndc_in_o"], out["ndc_in_d"] = o, dd
        for near in (1.0, 0.5):
            no, nd = RD.convert_to_ndc(o, dd, np.float32(21.5), 16, 12, near=near)
            out[f"ndc_o_{near}"], out[f"ndc_d_{near}"] = no, nd
        # pose_spherical incl. the up-axis re-orientation gen_video uses (nerf_sh/nerf/utils.py:656-685)
        sph = [(th, ph, rad, ua) for ua in range(6) for th, ph, rad in ((-180.0, -30.0, 4.0), (37.5, -75.0, 2.5))]
        out["pose_sph_in"] = np.array(sph, dtype=np.float64)
        out["pose_sph_out"] = np.stack([RU.pose_spherical(th, ph, rad, ua) for th, ph, rad, ua in sph])
        # the JAX-side compute_ssim (nerf_sh/nerf/utils.py:396-466, "valid" borders; what nerf_sh.train / eval report) on
        # the image pair of ssim.npz; convolve2d = scipy's, vmap = a slice loop
        zs = np.load(os.path.join(HERE, "ssim.npz"))
        out["ssim_jax_valid"] = np.asarray(RU.compute_ssim(zs["a"], zs["b"], 1.0), dtype=np.float64)
        out["ssim_jax_valid_map"] = np.asarray(RU.compute_ssim(zs["a"], zs["b"], 1.0, return_map=True))
        # learning-rate schedule (nerf_sh/nerf/utils.py:483-515), with and without the warm-up
        lr_in = [(s_, 5e-4, 5e-6, 2000000, d_, m_) for s_ in (0, 1, 999, 250000, 1999999, 2000000, 3000000)
                 for d_, m_ in ((0, 1.0), (2500, 0.01))]
        out["lr_in"] = np.array(lr_in, dtype=np.float64)
        out["lr_out"] = np.array([float(RU.learning_rate_decay(*a)) for a in lr_in], dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, "ref_llff.npz"), **out)
        print("ref_llff.npz:", {k: v.shape for k, v in out.items() if "test_render_poses" in k or "train_rays_o" in k})
    finally:
        jax_stub.uninstall(names)


def gen_configs():
    """The reference's shipped configuration files as parsed values (nerf_sh/config/{blender,tt}.yaml,
    octree/config/{syn_sh16,tt_sh25}.json): plenoctree_b200/presets.py must reproduce them."""
    import json
    import yaml
    out = {"nerf_sh": {}, "octree": {}}
    for n in ("blender", "tt"):
        out["nerf_sh"][n] = yaml.safe_load(open(os.path.join(REF, "nerf_sh", "config", n + ".yaml")))
    for n in ("syn_sh16", "tt_sh25"):
        out["octree"][n] = json.load(open(os.path.join(REF, "octree", "config", n + ".json")))
    json.dump(out, open(os.path.join(HERE, "ref_configs.json"), "w"), indent=1, sort_keys=True)
    print("ref_configs.json")


def gen_ref_extraction():
    """Execute the reference's own extraction control flow — octree/extraction.py `auto_scale` (:244-286), `step1`
    (:288-353, sigma mask) and `step2` (:355-394) plus the tail of `main` (:503-504: relu on sigma, shrink) —
    unmodified, on the CPU, with (a) the reference's torch NerfModel twin holding seeded weights and (b) a stand-in
    for the absent third-party `svox` package whose N3Tree is backed by oracle/octree_oracle.py (so the tree
    internals are the oracle's; what is pinned is the reference's SEQUENCE: grid construction, thresholds, the
    refinement schedule, leaf selection / chunking, sample -> eval -> mean -> assign, relu).  `.cuda()` is the
    identity here; the uniforms `sample` draws are recorded."""
    import types
    from oracle import octree_oracle as OO
    draws = []
    rs = np.random.RandomState(4242)

    class Format:
        RGBA = 0
        SH = 1

        def __init__(self, name):
            self.format = Format.RGBA if name in (None, "RGBA") else Format.SH
            self.name = name

    class View:
        def __init__(self, tree, key):
            self.tree, self.key = tree, key

        def refine(self):
            return self.tree.o.refine_at(self.key.numpy().astype(np.float32))

        def sample(self, n):
            lv = self.tree.o.leaves()[self.key.numpy()]
            u = rs.uniform(size=(lv.shape[0], n, 3)).astype(np.float32)
            draws.append(u)
            return torch.from_numpy(self.tree.o.sample(lv, n, u))

        def relu_(self):           # tree[:, -1:].relu_(): last channel of every leaf
            o = self.tree.o
            lv = o.leaves()
            vals = o.data[lv[:, 0], lv[:, 1], lv[:, 2], lv[:, 3], -1]
            o.data[lv[:, 0], lv[:, 1], lv[:, 2], lv[:, 3], -1] = np.maximum(vals, 0.0)

    class N3Tree:
        def __init__(self, N=2, data_dim=4, init_refine=0, init_reserve=1, geom_resize_fact=1.0, depth_limit=10,
                     radius=0.5, center=(0.5, 0.5, 0.5), data_format=None, extra_data=None, map_location="cpu"):
            self.o = OO.N3Tree(N=N, data_dim=data_dim, depth_limit=depth_limit, init_reserve=init_reserve,
                               geom_resize_fact=geom_resize_fact, radius=radius, center=center, data_format=data_format)
            self.data_dim = data_dim
            self.data_format = Format(data_format)

        offset = property(lambda self: torch.from_numpy(self.o.offset))
        invradius = property(lambda self: torch.from_numpy(self.o.invradius))
        max_depth = property(lambda self: self.o.max_depth)
        depths = property(lambda self: torch.from_numpy(self.o.leaf_depths(self.o.leaves()).astype(np.int64)))

        def __getitem__(self, key):
            if isinstance(key, tuple):
                assert key == (slice(None), slice(-1, None))
                return View(self, None)
            return View(self, key)

        def __setitem__(self, key, value):
            lv = self.o.leaves()[key.numpy()]
            self.o.data[lv[:, 0], lv[:, 1], lv[:, 2], lv[:, 3]] = value.numpy()

        def shrink_to_fit(self):
            self.o.shrink_to_fit()

        def __repr__(self):
            return f"svox-stand-in N3Tree(nodes={self.o.n_internal})"

    svox = types.ModuleType("svox")
    svox.N3Tree, svox.NDCConfig, svox.VolumeRenderer = N3Tree, object, object
    helpers = types.ModuleType("svox.helpers")
    helpers._get_c_extension = lambda: types.SimpleNamespace()
    svox.helpers = helpers
    sys.modules["svox"], sys.modules["svox.helpers"] = svox, helpers
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.empty_cache = lambda: None
    _use_reference_octree()
    from octree import extraction as RE
    RE.FLAGS(["make_golden"])
    RE.FLAGS.masking_mode = "sigma"
    sh_deg, L, S = 3, 3, 4
    seed = 31337
    flat_c = O.init_flat_params(sh_deg, seed, bias_scale=0.05)
    flat_f = O.init_flat_params(sh_deg, seed + 1, bias_scale=0.05)
    nerf = ref_model(sh_deg, flat_c, flat_f)
    center, radius = [0.1, -0.05, 0.0], [1.2, 1.0, 1.1]
    args = types.SimpleNamespace(init_grid_depth=L, alpha_thresh=0.0485, scale_alpha_thresh=0.1, chunk=1000, z_min=None,
                                 z_max=None, samples_per_cell=S, use_viewdirs=False, projection_samples=10000,
                                 sh_deg=sh_deg)
    with torch.no_grad():
        as_center, as_radius = RE.auto_scale(args, center, radius, nerf)
        tree = N3Tree(N=2, data_dim=49, init_refine=0, init_reserve=64, geom_resize_fact=1.0, depth_limit=L,
                      radius=radius, center=center, data_format="SH16", extra_data=None, map_location="cpu")
        RE.step1(args, tree, nerf, None)
        n_after_step1 = tree.o.n_internal
        RE.step2(args, tree, nerf)
        tree[:, -1:].relu_()
        tree.shrink_to_fit()
    o = tree.o
    np.savez_compressed(os.path.join(HERE, "ref_extraction.npz"), sh_deg=sh_deg, seed=seed, init_grid_depth=L,
                        samples_per_cell=S, chunk=args.chunk, alpha_thresh=args.alpha_thresh,
                        scale_alpha_thresh=args.scale_alpha_thresh, center=np.array(center), radius=np.array(radius),
                        autoscale_center=np.array(as_center), autoscale_radius=np.array(as_radius),
                        child=o.child, parent_depth=o.parent_depth, n_internal=o.n_internal,
                        n_after_step1=n_after_step1, data=o.data.astype(np.float32),
                        uniforms=np.concatenate(draws, axis=0),
                        flat_c_checksum=np.float64(flat_c.astype(np.float64).sum()))
    print("ref_extraction.npz nodes", o.n_internal, "leaves at max depth", int((tree.depths == L).sum()),
          "sample chunks", len(draws), "autoscale", as_center, as_radius)


def synthetic_octree(seed=77, data_dim=13, data_format="SH4"):
    """a small depth-2 oracle tree (root refined, three of its cells refined again) with random leaf data:
    SH coefficients ~ N(0, 1), densities in [0, 12)."""
    from oracle import octree_oracle as OO
    rs = np.random.RandomState(seed)
    tree = OO.N3Tree(N=2, data_dim=data_dim, depth_limit=4, init_reserve=16, geom_resize_fact=1.5, radius=1.0,
                     center=[0.0, 0.0, 0.0], data_format=data_format)
    tree.refine_at(np.array([[0.5, 0.5, 0.5]], dtype=np.float32))
    tree.refine_at(np.array([[0.5, 0.5, 0.5], [-0.5, 0.5, -0.5], [0.5, -0.5, 0.5]], dtype=np.float32))
    n = tree.n_internal
    tree.data[:n] = rs.normal(size=tree.data[:n].shape).astype(np.float32)
    tree.data[:n, ..., -1] = rs.uniform(0.0, 12.0, size=tree.data[:n, ..., -1].shape).astype(np.float32)
    return tree


def gen_ref_optimization():
    """Execute the reference's octree/optimization.py `main` (:133-248) unmodified on the CPU: Blender loader of the
    octree side, per-image render -> clamp -> MSE -> backward -> torch.optim.SGD step, validation PSNR every epoch,
    best-model bookkeeping and save.  `svox` is a stand-in: N3Tree = an nn.Module holding the oracle tree's arrays
    (`data` is the Parameter SGD updates), VolumeRenderer.render_persp = an autograd Function around the oracle's
    forward / backward march.  The renderer internals are therefore the oracle's; what is pinned is the reference's
    training-step semantics around it (clamp gradient, mean normalisation, update order, PSNR, best-of-validation)."""
    import contextlib
    import copy
    import io
    import json
    import tempfile
    import types
    from PIL import Image
    from oracle import octree_oracle as OO
    H = W = 6
    focal_angle = 0.9
    focal = 0.5 * W / np.tan(0.5 * focal_angle)
    step_size = 1e-3
    teacher = synthetic_octree(77)
    student = synthetic_octree(77)
    rs = np.random.RandomState(5)
    n = student.n_internal
    student.data[:n] = (student.data[:n] + rs.normal(scale=0.3, size=student.data[:n].shape)).astype(np.float32)
    student.data[:n, ..., -1] = np.maximum(student.data[:n, ..., -1], 0.0)
    poses = {"train": [O.pose_spherical(40.0 * i - 60.0, -30.0, 3.0) for i in range(3)],
             "val": [O.pose_spherical(25.0, -20.0, 3.0), O.pose_spherical(-110.0, -45.0, 3.0)]}

    class TreeModule(torch.nn.Module):
        def __init__(self, o):
            super().__init__()
            self.o = o
            self.data = torch.nn.Parameter(torch.from_numpy(o.data[:o.n_internal].copy()))

        @classmethod
        def load(cls, path, map_location="cpu"):
            z = np.load(path)
            o = OO.N3Tree(N=2, data_dim=int(z["data_dim"]), depth_limit=int(z["depth_limit"]), init_reserve=int(z["n_internal"]),
                          geom_resize_fact=float(z["geom_resize_fact"]), data_format=str(z["data_format"]))
            o.invradius, o.offset = z["invradius3"].astype(np.float32), z["offset"].astype(np.float32)
            o.child, o.parent_depth = z["child"].copy(), z["parent_depth"].copy()
            o.data = z["data"].astype(np.float32)
            o.n_internal = int(z["n_internal"])
            return cls(o)

        def clone(self, device="cpu"):
            c = TreeModule(copy.deepcopy(self.o))
            c.data = torch.nn.Parameter(self.data.detach().clone())
            return c

        def save(self, path, compress=False):
            self.o.data = self.data.detach().numpy().copy()
            st = self.o.state()
            st["data"] = self.o.data[:self.o.n_internal].astype(np.float32)          # keep fp32 for the comparison
            np.savez(path, **st)

    class March(torch.autograd.Function):
        @staticmethod
        def forward(ctx, data, tree, rays):
            tree.o.data = data.detach().numpy().copy()
            ctx.tree, ctx.rays = tree, rays
            return torch.from_numpy(OO.volume_render(tree.o, *rays, step_size=step_size))

        @staticmethod
        def backward(ctx, g):
            grad = OO.volume_render_backward(ctx.tree.o, *ctx.rays, g.numpy().astype(np.float32), step_size=step_size)
            return torch.from_numpy(grad), None, None

    class VolumeRenderer:
        def __init__(self, tree, step_size=1e-3, ndc=None):
            assert ndc is None
            self.tree = tree

        def render_persp(self, c2w, height, width, fx, fast=False, cuda=True):
            assert not fast
            rays = OO.persp_rays(c2w.numpy(), width, height, fx)
            return March.apply(self.tree.data, self.tree, rays).reshape(height, width, 3)

    svox = types.ModuleType("svox")
    svox.N3Tree, svox.VolumeRenderer, svox.NDCConfig = TreeModule, VolumeRenderer, object
    sys.modules["svox"] = svox
    sys.modules["imageio"] = types.SimpleNamespace(imwrite=lambda *a, **k: None)
    _use_reference_octree()
    from octree import optimization as RO
    with tempfile.TemporaryDirectory() as d:
        gts = {}
        for split, ps in poses.items():
            os.makedirs(os.path.join(d, split))
            frames = []
            for i, c2w in enumerate(ps):
                im = OO.volume_render(teacher, *OO.persp_rays(c2w, W, H, focal), step_size=step_size).reshape(H, W, 3)
                rgba = np.concatenate([np.clip(im, 0, 1), np.ones((H, W, 1), np.float32)], axis=-1)
                Image.fromarray((rgba * 255.0 + 0.5).astype(np.uint8), mode="RGBA").save(os.path.join(d, split, f"r_{i}.png"))
                frames.append({"file_path": f"./{split}/r_{i}", "transform_matrix": np.asarray(c2w, dtype=np.float64).tolist()})
            json.dump({"camera_angle_x": focal_angle, "frames": frames}, open(os.path.join(d, f"transforms_{split}.json"), "w"))
        st = student.state()
        st["data"] = student.data[:student.n_internal].astype(np.float32)
        np.savez(os.path.join(d, "tree.npz"), **st)
        lr = 40.0
        open(os.path.join(d, "cfg.yaml"), "w").write("dataset: blender\nfactor: 0\nwhite_bkgd: true\n")
        RO.FLAGS(["make_golden", "--config", os.path.join(d, "cfg"), "--input", os.path.join(d, "tree.npz"), "--output", os.path.join(d, "tree_opt.npz"),
                  "--data_dir", d, "--dataset", "blender", "--factor", "0", "--white_bkgd", "--num_epochs", "3",
                  "--val_interval", "1", "--sgd", "--lr", str(lr), "--continue_on_decrease", "--renderer_step_size", str(step_size)])
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            RO.main(None)
        log = buf.getvalue()
        out = np.load(os.path.join(d, "tree_opt.npz"))
        train_psnr = [float(l.split()[-1]) for l in log.splitlines() if l.startswith("** train_psnr")]
        val_psnr = [float(l.split()[3]) for l in log.splitlines() if l.startswith("** val psnr")]
        init_val = [float(l.split()[-1]) for l in log.splitlines() if l.startswith("** initial val psnr")][0]
        ds = {s_: RO.datasets.get_dataset(s_, RO.FLAGS) for s_ in ("train", "val")}
        np.savez_compressed(os.path.join(HERE, "ref_optimization.npz"), H=H, W=W, focal=np.float64(ds["train"].focal),
                            step_size=step_size, lr=lr, epochs=3,
                            train_c2w=ds["train"].camtoworlds, val_c2w=ds["val"].camtoworlds,
                            train_gt=ds["train"].images.astype(np.float32), val_gt=ds["val"].images.astype(np.float32),
                            child=student.child[:n], parent_depth=student.parent_depth[:n],
                            data0=student.data[:n].astype(np.float32), invradius=student.invradius, offset=student.offset,
                            train_psnr=np.array(train_psnr), val_psnr=np.array(val_psnr), initial_val_psnr=init_val,
                            data_best=out["data"].astype(np.float32))
    print("ref_optimization.npz", "initial val", init_val, "train", train_psnr, "val", val_psnr)


def gen_ssim():
    """reference torch twin octree/nerf/utils.py::compute_ssim on two random images."""
    ref_utils = load_ref_module("ref_octree_utils2", "octree/nerf/utils.py")
    rs = np.random.RandomState(21)
    a = rs.uniform(0, 1, size=(40, 36, 3)).astype(np.float32)
    b = np.clip(a + rs.normal(0, 0.1, size=a.shape), 0, 1).astype(np.float32)
    val = float(ref_utils.compute_ssim(torch.from_numpy(a), torch.from_numpy(b), max_val=1.0))
    np.savez_compressed(os.path.join(HERE, "ssim.npz"), a=a, b=b, ssim=val)
    print("ssim.npz", val)


def gen_flags():
    """every flags.DEFINE_* of the reference's four flag tables (name -> [kind, default]), read from the source AST
    (nerf_sh/nerf/utils.py imports jax at module level and cannot be imported here)."""
    import ast
    import json

    def extract(path):
        out = {}
        for node in ast.walk(ast.parse(open(os.path.join(REF, path)).read())):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr.startswith("DEFINE_"):
                try:
                    out[ast.literal_eval(node.args[0])] = [node.func.attr[len("DEFINE_"):], ast.literal_eval(node.args[1])]
                except Exception:
                    continue
        return out
    res = {p: extract(p) for p in ("nerf_sh/nerf/utils.py", "octree/nerf/utils.py", "octree/extraction.py",
                                   "octree/optimization.py")}
    with open(os.path.join(HERE, "flags.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print("flags.json", {k: len(v) for k, v in res.items()})


if __name__ == "__main__":
    torch.manual_seed(20200823)
    torch.set_num_threads(8)
    sys.path.insert(0, HERE)
    if len(sys.argv) > 1 and sys.argv[1] == "ref_render":
        gen_ref_render()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ref_loss":
        gen_ref_loss()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ref_extraction":    # own process: defines the octree-side flags, patches .cuda()
        gen_ref_extraction()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ref_optimization":  # own process: octree-side flags, svox / imageio stand-ins
        gen_ref_optimization()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "configs":
        gen_configs()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ref_llff":         # own process: imports the real loaders module
        gen_ref_llff()
        sys.exit(0)
    gen_ref_render()
    gen_ref_loss()
    gen_eval_points(3, 2048, 20200823, "eval_points_sh16.npz")
    gen_eval_points(4, 512, 20200900, "eval_points_sh25.npz")
    gen_eval_sh()
    gen_posenc()
    gen_ckpt_bridge()
    gen_ssim()
    gen_flags()
    gen_configs()
    try:
        gen_rays()
    except Exception as e:  # octree/nerf/utils.py pulls optional deps
        print("rays.npz skipped:", repr(e))
