"""`python -m octree.optimization` (reference README): the same flags and config files, served by plenoctree_b200.octree.optimization."""
import runpy

if __name__ == "__main__":
    runpy.run_module("plenoctree_b200.octree.optimization", run_name="__main__", alter_sys=True)
else:
    from plenoctree_b200.octree.optimization import *  # noqa: F401,F403
    from plenoctree_b200.octree.optimization import main  # noqa: F401
