import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nerf_sh_oracle as O
from plenoctree_b200._lib import check, lib, ptr
rs = np.random.RandomState(5)
R, Nc, Nf = 301, 64, 128
o = rs.normal(size=(R,3)).astype(np.float32); d = rs.normal(size=(R,3)).astype(np.float32)
z_o, _ = O.sample_along_rays(torch.from_numpy(o), torch.from_numpy(d), Nc, 2.0, 6.0, None)
w = rs.uniform(0, 1, size=(R, Nc)).astype(np.float32) ** 8
w[:10] = 0.0; w[10:20] = 1.0 / Nc; w[20:30] = 0.0; w[20:30, 17] = 0.9
z_c = z_o.contiguous()
mids = 0.5 * (z_c[..., 1:] + z_c[..., :-1])
ut = torch.linspace(0.0, 1.0 - float(np.finfo(np.float32).eps), Nf).cuda()
z_g = torch.empty((R, Nc + Nf), device="cuda")
check(lib.pob_sample_pdf(ptr(z_c.cuda()), ptr(torch.from_numpy(w).cuda()), ptr(ut), 0, R, Nc, Nf, ptr(z_g), None))
torch.cuda.synchronize()
zg = z_g.cpu().numpy()
z_ref, _ = O.sample_pdf(mids, torch.from_numpy(w)[..., 1:-1], torch.from_numpy(o), torch.from_numpy(d), z_c, Nf, None)
zr = z_ref.numpy()
print("sorted:", np.all(zg[:, 1:] >= zg[:, :-1]), "finite:", np.isfinite(zg).all())
err = np.abs(zg - zr)
print("err percentiles", np.percentile(err, [50, 90, 99, 99.9, 100]))
for name, sl in (("zero rows", slice(0, 10)), ("flat rows", slice(10, 20)), ("single-bin rows", slice(20, 30)), ("peaky", slice(30, R))):
    print(name, "max err", err[sl].max(), "median", np.median(err[sl]))
r = int(np.argmax(err.max(axis=1)))
print("worst ray", r, "gpu", zg[r, :12], "ref", zr[r, :12])
print("row 12 gpu", zg[12, 60:70], "ref", zr[12, 60:70])
