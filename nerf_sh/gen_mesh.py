"""`python -m nerf_sh.gen_mesh` (reference README): the same flags, served by plenoctree_b200.nerf_sh.gen_mesh."""
import runpy

if __name__ == "__main__":
    runpy.run_module("plenoctree_b200.nerf_sh.gen_mesh", run_name="__main__", alter_sys=True)
else:
    from plenoctree_b200.nerf_sh.gen_mesh import *  # noqa: F401,F403
    from plenoctree_b200.nerf_sh.gen_mesh import main  # noqa: F401
