"""LPIPS (VGG) — the third metric of the reference's octree evaluation (`lpips.LPIPS(net="vgg")(gt, im,
normalize=True)`, octree/nerf/utils.py:461-486; `lpips` is requirements.txt:5, third-party, not installed here).

The metric needs two sets of downloaded weights, which cannot ship with this repository and cannot be fetched in
an offline build: torchvision's ImageNet VGG-16 (`vgg16-397923af.pth`) and the LPIPS v0.1 linear heads
(`vgg.pth`, keys `lin<l>.model.1.weight`).  `load_lpips()` therefore looks for them (the `lpips` package itself if
it is importable; else `$POB_LPIPS_DIR`, else the torch hub cache) and returns None when they are absent — callers
then report `nan`, as the task manager's results.txt does.

The network is restated from the published algorithm (Zhang et al. 2018, PerceptualSimilarity v0.1):
inputs in [0,1] -> [-1,1] (`normalize=True`) -> per-channel shift / scale -> VGG-16 features at relu1_2, relu2_2,
relu3_3, relu4_3, relu5_3 -> unit-normalise every feature vector over channels -> squared difference -> learned
non-negative 1x1 weights -> spatial mean -> sum over the five taps.  The feature stack is checked against
torchvision's `vgg16().features` layer for layer (tests/test_pipeline.py); the linear heads have no reference here:
parity unpinned.
"""
import os

import torch
import torch.nn.functional as F

# (convolutions per block, output channels): VGG-16 configuration "D"; a 2x2 max-pool precedes blocks 2..5
_BLOCKS = ((2, 64), (2, 128), (3, 256), (3, 512), (3, 512))
# index of every convolution inside torchvision's `features` Sequential (for state-dict key mapping)
_TV_CONV_INDEX = (0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28)
_SHIFT = (-0.030, -0.088, -0.188)
_SCALE = (0.458, 0.448, 0.450)


class LPIPSVGG(torch.nn.Module):
    def __init__(self):
        super().__init__()
        convs, c_in = [], 3
        for n, c_out in _BLOCKS:
            for _ in range(n):
                convs.append(torch.nn.Conv2d(c_in, c_out, kernel_size=3, padding=1))
                c_in = c_out
        self.convs = torch.nn.ModuleList(convs)
        self.lins = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(1, c, 1, 1)) for _, c in _BLOCKS])
        self.register_buffer("shift", torch.tensor(_SHIFT).view(1, 3, 1, 1))
        self.register_buffer("scale", torch.tensor(_SCALE).view(1, 3, 1, 1))
        for p in self.parameters():
            p.requires_grad_(False)

    def load_vgg16(self, state):
        """torchvision `vgg16` state dict (`features.<i>.weight|bias`; classifier entries are ignored)."""
        for conv, i in zip(self.convs, _TV_CONV_INDEX):
            conv.weight.copy_(state[f"features.{i}.weight"])
            conv.bias.copy_(state[f"features.{i}.bias"])
        return self

    def load_linear_heads(self, state):
        """LPIPS v0.1 `vgg.pth`: `lin<l>.model.1.weight` [1, C_l, 1, 1]."""
        for l, lin in enumerate(self.lins):
            lin.copy_(state[f"lin{l}.model.1.weight"])
        return self

    def features(self, x):
        """the five taps (relu1_2 ... relu5_3) of an already shifted / scaled batch [n,3,h,w]."""
        taps, k = [], 0
        for b, (n, _) in enumerate(_BLOCKS):
            if b > 0:
                x = F.max_pool2d(x, kernel_size=2, stride=2)
            for _ in range(n):
                x = F.relu(self.convs[k](x))
                k += 1
            taps.append(x)
        return taps

    @torch.no_grad()
    def forward(self, im0, im1, normalize=True):
        """im0, im1: [3,h,w] or [n,3,h,w]; `normalize` maps [0,1] inputs to [-1,1] first.  -> distance per image [n]."""
        if im0.dim() == 3:
            im0, im1 = im0[None], im1[None]
        if normalize:
            im0, im1 = 2.0 * im0 - 1.0, 2.0 * im1 - 1.0
        f0 = self.features((im0 - self.shift) / self.scale)
        f1 = self.features((im1 - self.shift) / self.scale)
        total = 0.0
        for a, b, lin in zip(f0, f1, self.lins):
            a = a / (a.pow(2).sum(dim=1, keepdim=True).sqrt() + 1e-10)
            b = b / (b.pow(2).sum(dim=1, keepdim=True).sqrt() + 1e-10)
            total = total + ((a - b).pow(2) * lin).sum(dim=1, keepdim=True).mean(dim=(2, 3))[:, 0]
        return total


def _weight_files():
    dirs = [os.environ.get("POB_LPIPS_DIR"), os.path.join(torch.hub.get_dir(), "checkpoints")]
    for d in [x for x in dirs if x]:
        vgg = [os.path.join(d, n) for n in ("vgg16-397923af.pth", "vgg16.pth") if os.path.isfile(os.path.join(d, n))]
        lin = [os.path.join(d, n) for n in ("vgg.pth", "lpips_vgg.pth") if os.path.isfile(os.path.join(d, n))]
        if vgg and lin:
            return vgg[0], lin[0]
    return None


def load_lpips(device="cpu"):
    """-> callable(gt [3,h,w] in [0,1], im [3,h,w] in [0,1]) -> float, or None when no weights can be found."""
    try:
        import lpips                                   # the reference's own dependency, when present
        net = lpips.LPIPS(net="vgg").eval().to(device)
        return lambda gt, im: float(net(gt, im, normalize=True).item())
    except ImportError:
        pass
    files = _weight_files()
    if files is None:
        return None
    net = LPIPSVGG()
    net.load_vgg16(torch.load(files[0], map_location="cpu"))
    net.load_linear_heads(torch.load(files[1], map_location="cpu"))
    net = net.eval().to(device)
    return lambda gt, im: float(net(gt, im, normalize=True)[0])
