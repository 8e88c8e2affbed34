"""Ray containers and synthetic ray generation (numpy only; imports neither torch nor the CUDA library, so
the CPU reference arm of bench.py and the oracle-side tests can use it without mapping libplenoctree_b200.so).

    Rays namedtuple          nerf_sh/nerf/utils.py:53
    pose_spherical           nerf_sh/nerf/utils.py:656-685
    generate_rays            nerf_sh/nerf/utils.py:545-589
    convert_to_ndc           nerf_sh/nerf/datasets.py:40-60
"""
import collections

import numpy as np

# nerf_sh/nerf/utils.py:53
Rays = collections.namedtuple("Rays", ("origins", "directions", "viewdirs"))


def pose_spherical(theta, phi, radius, up_axis=0):
    """camera-to-world matrix looking at the origin (angles in degrees).  up_axis 0 keeps the NeRF-synthetic frame
    (z up); 1..5 re-orient the world so that -z / +y / -y / +x / -x is up (gen_video's --up_axis minus one)."""
    th, ph = np.deg2rad(theta), np.deg2rad(phi)
    trans = np.eye(4, dtype=np.float64)
    trans[2, 3] = radius
    rot_phi = np.array([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0], [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1.0]])
    rot_theta = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1.0]])
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1.0]])
    c2w = flip @ rot_theta @ rot_phi @ trans
    if up_axis != 0:
        up_dim = 2 - up_axis // 2                     # axis that becomes "up" ...
        up = np.zeros(3)
        up[up_dim] = -1.0 if up_axis % 2 else 1.0     # ... and its sign
        e1 = np.zeros(3)
        e1[1 if up_dim == 0 else 0] = 1.0
        frame = np.eye(4)
        frame[:3, 0], frame[:3, 1], frame[:3, 2] = e1, np.cross(up, e1), up
        c2w = frame @ c2w
    return c2w.astype(np.float32)


def generate_rays(w, h, focal, camtoworlds):
    """per-pixel origins, un-normalised directions and unit viewdirs, each [n, h, w, 3]; pixel (x, y) on the
    integer grid (no +0.5 centre shift, README.md:184)."""
    x, y = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32), indexing="xy")
    cam = np.stack([(x - w * 0.5) / focal, -(y - h * 0.5) / focal, -np.ones_like(x)], axis=-1)
    dirs = np.einsum("nij,hwj->nhwi", camtoworlds[:, :3, :3], cam).astype(np.float32)
    origins = np.broadcast_to(camtoworlds[:, None, None, :3, 3], dirs.shape).astype(np.float32).copy()
    viewdirs = dirs / np.linalg.norm(dirs, axis=-1, keepdims=True)
    return Rays(origins, dirs, viewdirs.astype(np.float32))


def convert_to_ndc(origins, directions, focal, w, h, near=1.0):
    """Rays of a forward-facing scene in normalised device coordinates: origins moved onto the plane z = -near,
    then the perspective projection that maps the frustum to [-1,1]^3 (z: near -> -1, infinity -> +1)."""
    shift = -(near + origins[..., 2]) / directions[..., 2]
    o = origins + shift[..., None] * directions
    sx, sy = -((2 * focal) / w), -((2 * focal) / h)
    ox_z, oy_z = o[..., 0] / o[..., 2], o[..., 1] / o[..., 2]
    ndc_o = np.stack([sx * ox_z, sy * oy_z, 1 + 2 * near / o[..., 2]], axis=-1)
    ndc_d = np.stack([sx * (directions[..., 0] / directions[..., 2] - ox_z),
                      sy * (directions[..., 1] / directions[..., 2] - oy_z),
                      -2 * near / o[..., 2]], axis=-1)
    return ndc_o, ndc_d


def random_rays_np(n, seed, w=800, h=800, camera_angle_x=0.6911112070083618, radius=4.0, n_poses=100, focal=None):
    """Synthetic benchmark rays: n_poses random spherical poses (theta ~ U(-180,180),
    phi ~ U(-90,0)), each ray = one uniformly random pixel of one random pose; targets ~ U[0,1).
    Returns origins, directions, viewdirs, pixels as float32 [n,3] arrays."""
    rs = np.random.RandomState(seed)
    if focal is None:
        focal = 0.5 * w / np.tan(0.5 * camera_angle_x)
    poses = np.stack([pose_spherical(rs.uniform(-180, 180), rs.uniform(-90, 0), radius) for _ in range(n_poses)])
    pi = rs.randint(0, n_poses, size=n)
    x = rs.randint(0, w, size=n).astype(np.float32)
    y = rs.randint(0, h, size=n).astype(np.float32)
    cam = np.stack([(x - w * 0.5) / focal, -(y - h * 0.5) / focal, -np.ones_like(x)], axis=-1).astype(np.float32)
    d = np.einsum("nij,nj->ni", poses[pi, :3, :3], cam).astype(np.float32)
    o = poses[pi, :3, 3].astype(np.float32)
    v = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32)
    px = rs.uniform(0, 1, size=(n, 3)).astype(np.float32)
    return o, d, v, px
