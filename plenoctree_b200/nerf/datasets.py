"""Blender-, LLFF- and NSVF-format datasets (nerf_sh/nerf/datasets.py:58-552; octree/nerf/datasets.py same classes,
plus bbox.txt :72-78) with the ray pool
resident on the GPU: images and per-pixel rays are built once on the host with the reference expressions
(`generate_rays`, white-background compositing, INTER_AREA half-resolution for factor 2), moved to HBM, and training
batches are drawn on the device (one random image + `batch_size` random pixels with replacement, datasets.py:159-166,
or pixels from all images with `image_batching`, :152-158).  The reference feeds batches through a host thread and a
3-deep queue (datasets.py:63-118); with the pool in HBM no host->device traffic is left on the training path.
"""
import json
import os

import numpy as np
import torch

from .models import Rays
from .rays import convert_to_ndc
from .utils import generate_rays


class Dataset:
    def __init__(self, split, args, device="cuda", rank=0, world=1):
        self.split = split
        self.render_path = bool(getattr(args, "render_path", False))
        bbox_path = os.path.join(os.path.expanduser(args.data_dir), "bbox.txt")      # octree/nerf/datasets.py:72-78
        self.bbox = np.loadtxt(bbox_path)[:-1] if os.path.isfile(bbox_path) else None
        self.device = torch.device(device)
        self.batch_size = int(args.batch_size) // world
        self.image_batching = bool(args.image_batching)
        self._render_rays_np = None                          # LLFF test split only (datasets.py:349-355)
        self._load_renderings(args)
        if not hasattr(self, "n_examples"):
            self.n_examples = self.images.shape[0]
        # Rays are built on first use: the octree CLIs only read camtoworlds / images, and the per-pixel ray set of a
        # Tanks&Temples scene (1920x1080, hundreds of views) is tens of GB on the host and in HBM.
        self._rays_np = None
        self._pool = None
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(20201473 + rank)                # np.random.seed(20201473 + host_id), train.py:128
        self.it = 0

    @property
    def rays_np(self):
        """per-pixel Rays of every view, [n,h,w,3] x3 numpy (utils.py:545-589)."""
        if self._rays_np is None:
            self._rays_np = self._build_rays()
        return self._rays_np

    def _build_rays(self):
        return generate_rays(self.w, self.h, self.focal, self.camtoworlds)

    @property
    def render_rays_np(self):
        """Rays of the generated camera path (`render_path`), [n_path,h,w,3] x3; built together with rays_np."""
        self.rays_np
        if self._render_rays_np is None:
            raise ValueError("this dataset / split has no render path")
        return self._render_rays_np

    def _device_pool(self):
        """pixels [n,hw,3] and Rays [n,hw,3] x3 resident in HBM (training batches are gathered on the device)."""
        if self._pool is None:
            n, hw = self.images.shape[0], self.h * self.w
            pixels = torch.from_numpy(self.images.reshape(n, hw, 3)).to(self.device)
            rays = Rays(*[torch.from_numpy(np.ascontiguousarray(r.reshape(n, hw, 3))).to(self.device)
                          for r in self.rays_np])
            self._pool = (pixels, rays)
        return self._pool

    @property
    def pixels(self):
        return self._device_pool()[0]

    @property
    def rays(self):
        return self._device_pool()[1]

    def _load_renderings(self, args):
        raise NotImplementedError

    @property
    def size(self):
        return self.n_examples

    def __len__(self):
        return self.n_examples

    def next_train(self):
        """{"pixels": [B,3], "rays": Rays([B,3] x3)} on the device (datasets.py:148-168)."""
        B, hw, n = self.batch_size, self.h * self.w, self.images.shape[0]
        if self.image_batching:
            idx = torch.randint(0, n * hw, (B,), device=self.device, generator=self.gen)
            img, pix = idx // hw, idx % hw
        else:
            img = torch.randint(0, n, (1,), device=self.device, generator=self.gen).expand(B)
            pix = torch.randint(0, hw, (B,), device=self.device, generator=self.gen)
        return {"pixels": self.pixels[img, pix], "rays": Rays(*[r[img, pix] for r in self.rays])}

    def next_test(self):
        """{"pixels": [h,w,3], "rays": Rays([h,w,3] x3)} as numpy (datasets.py:170-182)."""
        idx = self.it
        self.it = (self.it + 1) % self.n_examples
        if self.render_path:
            return {"rays": Rays(*[r[idx] for r in self.render_rays_np])}
        return {"pixels": self.images[idx], "rays": Rays(*[r[idx] for r in self.rays_np])}

    def __iter__(self):
        return self

    def __next__(self):
        return self.next_train() if self.split == "train" else self.next_test()


class Blender(Dataset):
    # datasets.py:189-232
    def _load_renderings(self, args):
        from PIL import Image
        if self.render_path:
            raise ValueError("render_path cannot be used for the blender dataset.")
        with open(os.path.join(args.data_dir, f"transforms_{self.split}.json"), "r") as fp:
            meta = json.load(fp)
        images, cams = [], []
        for frame in meta["frames"]:
            fname = os.path.join(args.data_dir, frame["file_path"] + ".png")
            image = np.array(Image.open(fname), dtype=np.float32) / 255.0
            if args.factor == 2:
                import cv2
                image = cv2.resize(image, (image.shape[1] // 2, image.shape[0] // 2), interpolation=cv2.INTER_AREA)
            elif args.factor > 0:
                raise ValueError(f"Blender dataset only supports factor=0 or 2, {args.factor} set.")
            cams.append(frame["transform_matrix"])
            if image.shape[-1] == 4:
                if args.white_bkgd:
                    mask = image[..., -1:]
                    image = image[..., :3] * mask + (1.0 - mask)
                else:
                    image = image[..., :3]
            images.append(image[..., :3])
        self.images = np.stack(images, axis=0).astype(np.float32)
        self.h, self.w = self.images.shape[1:3]
        self.resolution = self.h * self.w
        self.camtoworlds = np.stack(cams, axis=0).astype(np.float32)
        self.focal = 0.5 * self.w / np.tan(0.5 * float(meta["camera_angle_x"]))


def _unit(x):
    return x / np.linalg.norm(x)


def _look_at(z, up, pos):
    """[3,4] camera frame with back axis along z, x = up x z (re-orthogonalised up), position pos."""
    back = _unit(z)
    right = _unit(np.cross(up, back))
    return np.stack([right, _unit(np.cross(back, right)), back, pos], axis=1)


def _mean_pose(poses):
    """[3,5] average camera: mean position, summed back / up axes, the first camera's (h, w, focal) column."""
    return np.concatenate([_look_at(_unit(poses[:, :3, 2].sum(0)), poses[:, :3, 1].sum(0), poses[:, :3, 3].mean(0)),
                           poses[0, :3, -1:]], axis=1)


def _homogeneous(m34):
    last = np.zeros(m34.shape[:-2] + (1, 4)); last[..., 0, 3] = 1.0
    return np.concatenate([m34, last], axis=-2)


class LLFF(Dataset):
    """Forward-facing / 360 real scenes in the LLFF layout (nerf_sh/nerf/datasets.py:235-487): images[_<factor>]/ +
    poses_bounds.npy ([n,17]: a 3x5 [down, right, back | t | h,w,f] block and near/far depth bounds per view).
    Poses are re-ordered to [right, up, back], the scene is rescaled so that the nearest bound sits at 1/0.75,
    recentred on the average camera (or, with `spherify`, normalised onto the unit sphere around the point closest
    to all optical axes); every `llffhold`-th view is the test split.  Rays are handed out in NDC (near plane 1)
    unless `spherify`; the test split also carries a 120-pose spiral (or circle) path for `render_path`."""

    N_PATH = 120

    def _load_renderings(self, args):
        from PIL import Image
        d = os.path.expanduser(args.data_dir)
        factor = args.factor if args.factor > 0 else 1
        imgdir = os.path.join(d, "images" + (f"_{args.factor}" if args.factor > 0 else ""))
        if not os.path.exists(imgdir):
            raise ValueError(f"Image folder {imgdir} doesn't exist.")
        files = [f for f in sorted(os.listdir(imgdir)) if f.endswith(("JPG", "jpg", "png"))]
        images = np.stack([np.array(Image.open(os.path.join(imgdir, f)), dtype=np.float32) / 255.0 for f in files])
        raw = np.load(os.path.join(d, "poses_bounds.npy"))
        if raw.shape[0] != images.shape[0]:
            raise RuntimeError(f"Mismatch between imgs {images.shape[0]} and poses {raw.shape[0]}")
        m = raw[:, :15].reshape(-1, 3, 5).copy()
        m[:, 0, 4], m[:, 1, 4] = images.shape[1], images.shape[2]           # (h, w) of the images actually loaded
        m[:, 2, 4] /= factor
        poses = np.concatenate([m[:, :, 1:2], -m[:, :, 0:1], m[:, :, 2:]], axis=2).astype(np.float32)
        bds = raw[:, 15:17].astype(np.float32)
        scale = 1.0 / (bds.min() * 0.75)
        poses[:, :3, 3] *= scale
        bds *= scale
        # recentre: express every camera in the frame of the average camera
        world_from_avg = _homogeneous(_mean_pose(poses)[:, :4])
        poses[:, :3, :4] = (np.linalg.inv(world_from_avg) @ _homogeneous(poses[:, :, :4]))[:, :3, :4]
        self.spherify = bool(args.spherify)
        if self.spherify:
            poses = self._spherify(poses, bds)
        elif self.split == "test":
            self.render_poses = self._spiral_path(poses, bds)
        i_test = np.arange(images.shape[0])[:: args.llffhold]
        keep = i_test if self.split != "train" else np.setdiff1d(np.arange(images.shape[0]), i_test)
        images, poses = images[keep], poses[keep]
        self.images = images[..., :3]
        self.camtoworlds = poses[:, :3, :4]
        self.focal = poses[0, -1, -1]
        self.h, self.w = images.shape[1:3]
        self.resolution = self.h * self.w
        self.n_examples = self.render_poses.shape[0] if self.render_path else images.shape[0]

    def _build_rays(self):
        """per-pixel rays of the split's views — and of the render path, test split — in NDC unless spherify."""
        n_path = self.render_poses.shape[0] if self.split == "test" else 0
        cams = np.concatenate([self.render_poses, self.camtoworlds], axis=0) if n_path else self.camtoworlds
        rays = generate_rays(self.w, self.h, self.focal, cams)
        if not self.spherify:
            o, dd = convert_to_ndc(rays.origins, rays.directions, self.focal, self.w, self.h)
            rays = Rays(o, dd, rays.viewdirs)
        if n_path:
            self._render_rays_np = Rays(*[r[:n_path] for r in rays])
            rays = Rays(*[r[n_path:] for r in rays])
        return rays

    def _spiral_path(self, poses, bds):
        """N_PATH poses on a two-turn spiral around the average camera, looking at a focus depth between the
        nearest and farthest bounds (weighted 1:3 in disparity); radii = 90th percentile of |camera positions|."""
        avg = _mean_pose(poses)
        up = _unit(poses[:, :3, 1].sum(0))
        close, far = bds.min() * 0.9, bds.max() * 5.0
        focus = 1.0 / (0.25 / close + 0.75 / far)
        radii = np.append(np.percentile(np.abs(poses[:, :3, 3]), 90, 0).astype(np.float64), 1.0)
        frame = avg[:3, :4]
        target = frame @ np.array([0.0, 0.0, -focus, 1.0])
        path = []
        for theta in np.linspace(0.0, 4.0 * np.pi, self.N_PATH + 1)[:-1]:
            eye = frame @ (np.array([np.cos(theta), -np.sin(theta), -np.sin(theta * 0.5), 1.0]) * radii)
            path.append(_look_at(eye - target, up, eye))
        return np.stack(path).astype(np.float32)

    def _spherify(self, poses, bds):
        """360 scenes: move the origin to the least-squares intersection of the optical axes, z along the mean
        camera offset, scale so that the cameras' RMS distance is 1; the render path is a circle at the cameras'
        mean height.  `bds` is rescaled in place."""
        axis, pos = poses[:, :3, 2:3], poses[:, :3, 3:4]
        proj = np.eye(3) - axis * np.transpose(axis, [0, 2, 1])            # projector off each optical axis
        centre = np.squeeze(np.linalg.inv((np.transpose(proj, [0, 2, 1]) @ proj).mean(0)) @ (proj @ pos).mean(0))
        z = _unit((poses[:, :3, 3] - centre).mean(0))
        x = _unit(np.cross([0.1, 0.2, 0.3], z))
        y = _unit(np.cross(z, x))
        world = _homogeneous(np.stack([x, y, z, centre], axis=1)[None])
        local = np.linalg.inv(world) @ _homogeneous(poses[:, :3, :4])
        radius = np.sqrt(np.mean(np.sum(np.square(local[:, :3, 3]), -1)))
        local[:, :3, 3] *= 1.0 / radius
        bds *= 1.0 / radius
        height = np.mean(local[:, :3, 3], 0)[2]
        r_circle = np.sqrt(1.0 - height ** 2)
        hwf = poses[0, :3, -1:]
        if self.split == "test":
            ring = []
            for th in np.linspace(0.0, 2.0 * np.pi, self.N_PATH):
                eye = np.array([r_circle * np.cos(th), r_circle * np.sin(th), height])
                back = _unit(eye)
                right = _unit(np.cross(back, [0.0, 0.0, -1.0]))
                ring.append(np.stack([right, _unit(np.cross(back, right)), back, eye], axis=1))
            self.render_poses = np.stack(ring)
        return np.concatenate([local[:, :3, :4], np.broadcast_to(hwf, local[:, :3, -1:].shape)], axis=-1)


class NSVF(Dataset):
    """NSVF generic dataset (nerf_sh/nerf/datasets.py:491-552): intrinsics.txt, pose/<split>_*.txt, rgb/<split>_*.png
    with split prefixes 0_ train / 1_ val / 2_ test (1_ when there is no 2_), camera flip diag(1,-1,-1,1)."""

    def _load_renderings(self, args):
        from PIL import Image
        if self.render_path:
            raise ValueError("render_path cannot be used for the NSVF dataset.")
        d = os.path.expanduser(args.data_dir)
        K = np.loadtxt(os.path.join(d, "intrinsics.txt"))
        pose_files = sorted(os.listdir(os.path.join(d, "pose")))
        img_files = sorted(os.listdir(os.path.join(d, "rgb")))
        pick = lambda files, pre: [x for x in files if x.startswith(pre)]
        if self.split == "train":
            pose_files, img_files = pick(pose_files, "0_"), pick(img_files, "0_")
        elif self.split == "val":
            pose_files, img_files = pick(pose_files, "1_"), pick(img_files, "1_")
        elif self.split == "test":
            tp, ti = pick(pose_files, "2_"), pick(img_files, "2_")
            if len(tp) == 0:
                tp, ti = pick(pose_files, "1_"), pick(img_files, "1_")
            pose_files, img_files = tp, ti
        assert len(img_files) == len(pose_files)
        cam_trans = np.diag(np.array([1, -1, -1, 1], dtype=np.float32))
        images, cams = [], []
        for img_fname, pose_fname in zip(img_files, pose_files):
            image = np.array(Image.open(os.path.join(d, "rgb", img_fname)), dtype=np.float32) / 255.0
            cams.append(np.loadtxt(os.path.join(d, "pose", pose_fname)) @ cam_trans)
            if image.shape[-1] == 4:
                if args.white_bkgd:
                    mask = image[..., -1:]
                    image = image[..., :3] * mask + (1.0 - mask)
                else:
                    image = image[..., :3]
            if args.factor > 1:
                import cv2
                image = cv2.resize(image, (image.shape[1] // args.factor, image.shape[0] // args.factor),
                                   interpolation=cv2.INTER_AREA)
            images.append(image[..., :3])
        self.images = np.stack(images, axis=0).astype(np.float32)
        self.h, self.w = self.images.shape[1:3]
        self.resolution = self.h * self.w
        self.camtoworlds = np.stack(cams, axis=0).astype(np.float32)
        self.focal = (K[0, 0] + K[1, 1]) * 0.5          # fx and fy assumed equal
        if args.factor > 1:
            self.focal /= args.factor


def get_dataset(split, args, **kw):
    """datasets.get_dataset (nerf_sh/nerf/datasets.py:39-40)."""
    classes = {"blender": Blender, "llff": LLFF, "nsvf": NSVF}
    if args.dataset not in classes:
        raise NotImplementedError(f"dataset {args.dataset!r}: one of {sorted(classes)} expected")
    return classes[args.dataset](split, args, **kw)


def write_blender_scene(data_dir, images_by_split, poses_by_split, camera_angle_x):
    """Write a scene in the Blender (NeRF-synthetic) layout: transforms_<split>.json + <split>/r_<i>.png (RGBA with
    alpha 1).  Used to build synthetic scenes from a teacher model where no dataset can be downloaded."""
    from PIL import Image
    os.makedirs(data_dir, exist_ok=True)
    for split, images in images_by_split.items():
        os.makedirs(os.path.join(data_dir, split), exist_ok=True)
        frames = []
        for i, (im, c2w) in enumerate(zip(images, poses_by_split[split])):
            rgba = np.concatenate([np.clip(im, 0, 1), np.ones_like(im[..., :1])], axis=-1)
            Image.fromarray((rgba * 255.0 + 0.5).astype(np.uint8), mode="RGBA").save(
                os.path.join(data_dir, split, f"r_{i}.png"))
            frames.append({"file_path": f"./{split}/r_{i}", "transform_matrix": np.asarray(c2w, dtype=np.float64).tolist()})
        with open(os.path.join(data_dir, f"transforms_{split}.json"), "w") as fp:
            json.dump({"camera_angle_x": float(camera_angle_x), "frames": frames}, fp)


def write_nsvf_scene(data_dir, images_by_split, poses_by_split, focal, bbox=None):
    """Write a scene in the NSVF layout (intrinsics.txt, pose/, rgb/, optional bbox.txt); poses are the c2w matrices
    the loader should return (the on-disk files carry the inverse camera flip)."""
    from PIL import Image
    os.makedirs(os.path.join(data_dir, "pose"), exist_ok=True)
    os.makedirs(os.path.join(data_dir, "rgb"), exist_ok=True)
    h, w = next(iter(images_by_split.values()))[0].shape[:2]
    K = np.array([[focal, 0, w * 0.5, 0], [0, focal, h * 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)
    np.savetxt(os.path.join(data_dir, "intrinsics.txt"), K)
    flip = np.diag([1.0, -1.0, -1.0, 1.0])
    for split, pre in (("train", "0_"), ("val", "1_"), ("test", "2_")):
        for i, (im, c2w) in enumerate(zip(images_by_split.get(split, []), poses_by_split.get(split, []))):
            Image.fromarray((np.clip(im, 0, 1) * 255.0 + 0.5).astype(np.uint8), mode="RGB").save(
                os.path.join(data_dir, "rgb", f"{pre}{i:04d}.png"))
            np.savetxt(os.path.join(data_dir, "pose", f"{pre}{i:04d}.txt"), np.asarray(c2w, dtype=np.float64) @ flip)
    if bbox is not None:
        np.savetxt(os.path.join(data_dir, "bbox.txt"), np.asarray(list(bbox) + [0.4], dtype=np.float64)[None])
