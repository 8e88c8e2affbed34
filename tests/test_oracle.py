"""Pins the CPU oracle (oracle/nerf_sh_oracle.py) against golden vectors produced by the reference's
own importable code (tests/golden/make_golden.py) and against closed-form identities (SURVEY §4 iii)."""
import os

import numpy as np
import torch

from oracle import nerf_sh_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_param_counts():
    # SURVEY.md F5 / BASELINE.md §2
    assert O.param_count(3) == 505649
    assert O.param_count(4) == 512588


def test_eval_points_raw_matches_reference_twin(golden_dir):
    for name, sh_deg in (("eval_points_sh16.npz", 3), ("eval_points_sh25.npz", 4)):
        g = _load(golden_dir, name)
        seed = int(g["seed"])
        flat_c = O.init_flat_params(sh_deg, seed, bias_scale=0.05)
        flat_f = O.init_flat_params(sh_deg, seed + 1, bias_scale=0.05)
        assert abs(flat_c.astype(np.float64).sum() - float(g["flat_c_checksum"])) < 1e-9
        assert abs(flat_f.astype(np.float64).sum() - float(g["flat_f_checksum"])) < 1e-9
        pts = torch.from_numpy(g["points"])
        for flat, tag in ((flat_f, "fine"), (flat_c, "coarse")):
            rgb, sig = O.eval_points_raw(O.unflatten(flat, sh_deg), pts)
            np.testing.assert_allclose(rgb.numpy(), g[f"raw_rgb_{tag}"], rtol=2e-5, atol=2e-6)
            np.testing.assert_allclose(sig.numpy(), g[f"raw_sigma_{tag}"], rtol=2e-5, atol=2e-6)


def test_eval_sh_matches_reference(golden_dir):
    g = _load(golden_dir, "eval_sh.npz")
    for deg in range(5):
        res = O.eval_sh(deg, torch.from_numpy(g[f"sh{deg}"]), torch.from_numpy(g[f"dirs{deg}"]))
        np.testing.assert_allclose(res.numpy(), g[f"res{deg}"], rtol=1e-5, atol=1e-6)


def test_posenc_matches_reference(golden_dir):
    g = _load(golden_dir, "posenc.npz")
    enc = O.posenc(torch.from_numpy(g["x"])).numpy()
    np.testing.assert_array_equal(enc, g["enc"])
    # layout: [x, sin(2^j x) j-major, sin(2^j x + pi/2)]
    x = g["x"][:4]
    np.testing.assert_allclose(enc[:4, 3:6], np.sin(x), atol=1e-6)
    np.testing.assert_allclose(enc[:4, 6:9], np.sin(2 * x), atol=1e-6)
    np.testing.assert_allclose(enc[:4, 33:36], np.cos(x), atol=1e-6)


def test_generate_rays_matches_reference(golden_dir):
    g = _load(golden_dir, "rays.npz")
    o, d, v = O.generate_rays(int(g["w"]), int(g["h"]), float(g["focal"]), g["poses"])
    np.testing.assert_allclose(o, g["origins"], atol=1e-6)
    np.testing.assert_allclose(d, g["directions"], atol=1e-6)
    np.testing.assert_allclose(v, g["viewdirs"], atol=1e-6)


def test_compositing_identities():
    torch.manual_seed(0)
    B, N = 5, 64
    z = torch.linspace(2, 6, N).expand(B, N).contiguous()
    dirs = torch.randn(B, 3)
    sigma = torch.rand(B, N, 1) * 3
    color = torch.tensor([0.2, 0.5, 0.9])
    rgb = color.expand(B, N, 3)
    comp, disp, acc, w = O.volumetric_rendering(rgb, sigma, z, dirs, white_bkgd=False)
    # last alpha is 1 (sigma > 0, dist 1e10) => weights sum to ~1 and a constant field renders itself
    np.testing.assert_allclose(acc.numpy(), 1.0, atol=1e-5)
    np.testing.assert_allclose(comp.numpy(), color.expand(B, 3).numpy(), atol=1e-5)
    # empty space + white background = white, disp falls back to 1e10
    comp, disp, acc, w = O.volumetric_rendering(rgb, torch.zeros(B, N, 1), z, dirs, white_bkgd=True)
    np.testing.assert_allclose(comp.numpy(), 1.0, atol=1e-6)
    assert torch.all(disp == 1e10) and torch.all(acc == 0)


def test_piecewise_constant_pdf_uniform():
    B = 3
    bins = torch.linspace(2, 6, 64)[None].expand(B, 64)
    mids = 0.5 * (bins[..., 1:] + bins[..., :-1])
    w = torch.ones(B, 62)
    z = O.piecewise_constant_pdf(mids, w, 128)
    # uniform pdf + deterministic u = evenly spaced samples over [mids[0], mids[-1])
    expect = mids[0, 0] + torch.linspace(0, 1 - 2 ** -23, 128) * (mids[0, -1] - mids[0, 0])
    np.testing.assert_allclose(z[0].numpy(), expect.numpy(), rtol=1e-5)
    # zero weights hit the eps padding branch and still give finite, sorted samples
    z0 = O.piecewise_constant_pdf(mids, torch.zeros(B, 62), 128)
    assert torch.isfinite(z0).all() and torch.all(z0[:, 1:] >= z0[:, :-1])


def test_nerf_forward_shapes_and_determinism():
    sh_deg = 3
    fc = O.init_flat_params(sh_deg, 1)
    ff = O.init_flat_params(sh_deg, 2)
    rs = np.random.RandomState(0)
    o = torch.from_numpy(rs.normal(size=(6, 3)).astype(np.float32))
    d = torch.from_numpy(rs.normal(size=(6, 3)).astype(np.float32))
    v = d / d.norm(dim=-1, keepdim=True)
    with torch.no_grad():
        ret, aux = O.nerf_forward(O.unflatten(fc, sh_deg), O.unflatten(ff, sh_deg), sh_deg, (o, d, v), 64, 128,
                                  2.0, 6.0, return_aux=True)
    assert len(ret) == 2 and ret[1][0].shape == (6, 3) and ret[1][1].shape == (6,)
    assert aux["z_fine"].shape == (6, 192)
    assert torch.all(aux["z_fine"][:, 1:] >= aux["z_fine"][:, :-1])


def test_adam_matches_torch():
    rs = np.random.RandomState(3)
    p = rs.normal(size=100).astype(np.float32)
    tp = torch.nn.Parameter(torch.from_numpy(p.copy()))
    opt = torch.optim.Adam([tp], lr=5e-4, betas=(0.9, 0.999), eps=1e-8)
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    for step in range(3):
        g = rs.normal(size=100).astype(np.float32)
        tp.grad = torch.from_numpy(g.copy())
        opt.step()
        p, m, v = O.adam_step(p, g, m, v, float(step), 5e-4)
    np.testing.assert_allclose(p, tp.detach().numpy(), rtol=1e-5, atol=1e-7)


def test_lr_schedule():
    assert abs(O.learning_rate_decay(0, 5e-4, 5e-6, 1000) - 5e-4) < 1e-12
    assert abs(O.learning_rate_decay(1000, 5e-4, 5e-6, 1000) - 5e-6) < 1e-12
    assert abs(O.learning_rate_decay(500, 5e-4, 5e-6, 1000) - 5e-5) < 1e-10


# ----------------------------------------------------------------------------------------------------------------
# The oracle against the reference's own JAX forward path, EXECUTED unmodified over numpy stand-ins for jax / flax
# (tests/golden/make_golden.py::gen_ref_render, tests/golden/jax_stub.py): /root/reference/nerf_sh/nerf/
# model_utils.py:97-332 and models.py:216-348.  Arithmetic without transcendentals must agree bit for bit (both
# sides are IEEE fp32); exp / sin / reductions may differ in the last bits between numpy and torch.
# ----------------------------------------------------------------------------------------------------------------
def _ref(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_render.npz"))


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def test_sample_along_rays_matches_executed_reference(golden_dir):
    g = _ref(golden_dir)
    o, d = _t(g["origins"]), _t(g["directions"])
    for tag, t_rand, lin in (("plain", None, False), ("rand", _t(g["t_rand"]), False), ("lindisp_rand", _t(g["t_rand"]), True)):
        z, pts = O.sample_along_rays(o, d, 64, 2.0, 6.0, t_rand, lin)
        # one ulp: linspace(0, 1, 64) is fl32(i/63) in jnp / the numpy stand-in (computed in float64, then rounded)
        # and fl32(i * fl32(1/63)) (mirrored from the far end for i > 31) in torch
        np.testing.assert_allclose(z.numpy(), g[f"sar_{tag}_z"], rtol=2.5e-7, atol=0, err_msg=tag)
        np.testing.assert_allclose(pts.numpy(), g[f"sar_{tag}_pts"], rtol=0, atol=2e-6, err_msg=tag)


def test_volumetric_rendering_matches_executed_reference(golden_dir):
    g = _ref(golden_dir)
    rgb, sigma, z, d = _t(g["vr_in_rgb"]), _t(g["vr_in_sigma"]), _t(g["vr_in_z"]), _t(g["directions"])
    for wb in (1, 0):
        c, disp, acc, w = O.volumetric_rendering(rgb, sigma, z, d, bool(wb))
        np.testing.assert_allclose(w.numpy(), g[f"vr_weights_{wb}"], rtol=2e-6, atol=1.5e-7)   # cumprod order, exp ulps
        np.testing.assert_allclose(c.numpy(), g[f"vr_rgb_{wb}"], rtol=2e-6, atol=2e-7)
        np.testing.assert_allclose(acc.numpy(), g[f"vr_acc_{wb}"], rtol=2e-6, atol=2e-7)
        np.testing.assert_allclose(disp.numpy(), g[f"vr_disp_{wb}"], rtol=5e-6)
    # the guarded cases: an empty ray renders the background with disp = 1/eps, an opaque one has acc = 1
    assert g["vr_disp_1"][0] == np.float32(1e10) and np.allclose(g["vr_rgb_1"][0], 1.0)
    assert abs(float(g["vr_acc_1"][1]) - 1.0) < 1e-6


def test_piecewise_constant_pdf_matches_executed_reference(golden_dir):
    g = _ref(golden_dir)
    bins, wts, u = _t(g["pdf_bins"]), _t(g["pdf_weights"]), _t(g["pdf_u"])
    det = O.piecewise_constant_pdf(bins, wts, 128, None).numpy()
    rnd = O.piecewise_constant_pdf(bins, wts, 128, u).numpy()
    # rows: peaky, flat, all-zero (eps padding), single bin, generic with u in {0, 1-eps, 0.5}
    # cumsum order (numpy sequential, torch blocked) moves the cdf by ulps: 1e-5 of a z in [2, 6], 2e-4 of a bin width
    np.testing.assert_allclose(det, g["pdf_det"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(rnd, g["pdf_rand"], rtol=0, atol=1e-5)
    # bracket selection itself (no rounding involved) must be identical: the samples fall in the same bins
    bn = g["pdf_bins"]
    for a, b in ((det, g["pdf_det"]), (rnd, g["pdf_rand"])):
        ia = np.stack([np.searchsorted(bn[r], a[r], side="right") for r in range(bn.shape[0])])
        ib = np.stack([np.searchsorted(bn[r], b[r], side="right") for r in range(bn.shape[0])])
        assert (ia != ib).mean() < 0.01     # a sample sitting on a bin edge may round to either side
    z, pts = O.sample_pdf(bins, wts, _t(g["origins"][:5]), _t(g["directions"][:5]), _t(g["spdf_z_coarse"]), 128, u)
    np.testing.assert_allclose(z.numpy(), g["spdf_z"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(pts.numpy(), g["spdf_pts"], rtol=0, atol=5e-5)
    assert np.all(np.diff(g["spdf_z"], axis=-1) >= 0)


def test_add_gaussian_noise_matches_executed_reference(golden_dir):
    g = _ref(golden_dir)
    raw, nz = _t(g["noise_raw"]), _t(g["noise_draw"])
    assert np.array_equal(O.add_gaussian_noise(raw, nz * np.float32(0.7)).numpy(), g["noise_on"])
    assert np.array_equal(O.add_gaussian_noise(raw, None).numpy(), g["noise_off_std"])
    assert np.array_equal(g["noise_off_rand"], g["noise_raw"])


def test_mlp_matches_executed_reference_class(golden_dir):
    """model_utils.MLP.__call__ (skip-concat after layer 4, sigma head = Dense_8, rgb head = Dense_9)."""
    g = _ref(golden_dir)
    params = O.unflatten(O.init_flat_params(int(g["sh_deg"]), int(g["seeds"][1]), bias_scale=0.05), int(g["sh_deg"]))
    with torch.no_grad():
        rgb, sigma = O.mlp(params, _t(g["mlp_enc"]))
    scale = float(np.abs(g["mlp_raw_rgb"]).max())
    assert float(np.abs(rgb.numpy() - g["mlp_raw_rgb"]).max()) < 2e-5 * scale
    assert float(np.abs(sigma.numpy() - g["mlp_raw_sigma"]).max()) < 2e-5 * max(1.0, float(np.abs(g["mlp_raw_sigma"]).max()))


def test_nerf_forward_matches_executed_reference_call(golden_dir):
    """NerfModel.__call__ (models.py:216-348), both levels, deterministic and with injected draws."""
    g = _ref(golden_dir)
    sh_deg = int(g["sh_deg"])
    pc = O.unflatten(O.init_flat_params(sh_deg, int(g["seeds"][0]), bias_scale=0.05), sh_deg)
    pf = O.unflatten(O.init_flat_params(sh_deg, int(g["seeds"][1]), bias_scale=0.05), sh_deg)
    rays = (_t(g["origins"]), _t(g["directions"]), _t(g["viewdirs"]))
    for tag, t_rand, u in (("det", None, None), ("rand", _t(g["t_rand"]), _t(g["call_u"]))):
        with torch.no_grad():
            ret = O.nerf_forward(pc, pf, sh_deg, rays, 64, 128, 2.0, 6.0, True, t_rand=t_rand, u=u)
        for lvl, (rgb, disp, acc) in zip(("coarse", "fine"), ret):
            # random-init field: the fine level's samples move with 1e-6 changes of the coarse weights
            tol = 2e-5 if lvl == "coarse" else 5e-4
            np.testing.assert_allclose(rgb.numpy(), g[f"call_{tag}_{lvl}_rgb"], rtol=0, atol=tol)
            np.testing.assert_allclose(acc.numpy(), g[f"call_{tag}_{lvl}_acc"], rtol=0, atol=tol)
            np.testing.assert_allclose(disp.numpy(), g[f"call_{tag}_{lvl}_disp"], rtol=20 * tol)


def test_package_generate_rays_matches_reference(golden_dir):
    """the product's host copy (plenoctree_b200.nerf.rays.generate_rays), not only the oracle's, against the
    reference's generate_rays (octree/nerf/utils.py:401-445 = nerf_sh/nerf/utils.py:545-589)."""
    from plenoctree_b200.nerf.rays import generate_rays
    g = _load(golden_dir, "rays.npz")
    rays = generate_rays(int(g["w"]), int(g["h"]), float(g["focal"]), g["poses"])
    np.testing.assert_allclose(rays.origins, g["origins"], rtol=0, atol=0)
    np.testing.assert_allclose(rays.directions, g["directions"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(rays.viewdirs, g["viewdirs"], rtol=1e-6, atol=1e-6)


def test_loss_fn_matches_executed_reference_train_step(golden_dir):
    """train_step.loss_fn (nerf_sh/train.py:68-114) executed unmodified over the numpy stand-ins
    (tests/golden/make_golden.py::gen_ref_loss): both MSE terms, the PSNRs, the sparsity term on injected points
    (random.uniform(key, (n,3), minval=-r, maxval=r), train.py:79) and weight_l2 over the whole parameter tree."""
    g = np.load(os.path.join(golden_dir, "ref_loss.npz"))
    sh_deg = int(g["sh_deg"])
    pc = O.unflatten(O.init_flat_params(sh_deg, int(g["seeds"][0]), bias_scale=0.05), sh_deg)
    pf = O.unflatten(O.init_flat_params(sh_deg, int(g["seeds"][1]), bias_scale=0.05), sh_deg)
    r = np.float32(g["sparsity_radius"])
    sp = (g["sp01"] * np.float32(r - (-r)) + np.float32(-r)).astype(np.float32)      # the stand-in's uniform()
    cfg = dict(num_coarse_samples=64, num_fine_samples=128, near=2.0, far=6.0, white_bkgd=True,
               sparsity_weight=float(g["sparsity_weight"]), sparsity_length=float(g["sparsity_length"]),
               weight_decay_mult=float(g["weight_decay_mult"]))
    with torch.no_grad():
        total, st = O.loss_fn(pc, pf, sh_deg, (_t(g["origins"]), _t(g["directions"]), _t(g["viewdirs"])),
                              _t(g["pixels"]), cfg, _t(g["t_rand"]), _t(g["u"]), _t(sp))
    for k, tol in (("loss", 3e-4), ("loss_c", 2e-5), ("loss_sp", 1e-4), ("weight_l2", 1e-6), ("psnr", 3e-4), ("psnr_c", 2e-5)):
        assert abs(float(st[k]) - float(g[k])) <= tol * abs(float(g[k])), (k, float(st[k]), float(g[k]))
    want_total = float(g["loss"]) + float(g["loss_c"]) + float(g["loss_sp"]) + float(g["weight_decay_mult"]) * float(g["weight_l2"])
    assert abs(float(total) - want_total) < 3e-4 * want_total
