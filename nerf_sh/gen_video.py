"""`python -m nerf_sh.gen_video` (reference README): the same flags, served by plenoctree_b200.nerf_sh.gen_video."""
import runpy

if __name__ == "__main__":
    runpy.run_module("plenoctree_b200.nerf_sh.gen_video", run_name="__main__", alter_sys=True)
else:
    from plenoctree_b200.nerf_sh.gen_video import *  # noqa: F401,F403
    from plenoctree_b200.nerf_sh.gen_video import main  # noqa: F401
