"""`python -m octree.compression` (reference README): the same arguments, served by plenoctree_b200.octree.compression."""
from plenoctree_b200.octree.compression import *  # noqa: F401,F403
from plenoctree_b200.octree.compression import main

if __name__ == "__main__":
    raise SystemExit(main())
