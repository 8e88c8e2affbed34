"""Drop-in module path of the reference (`python -m nerf_sh.train`, README.md:58-67): thin shims over plenoctree_b200.nerf_sh."""
