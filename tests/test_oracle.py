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
