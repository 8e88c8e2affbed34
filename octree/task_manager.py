"""`python -m octree.task_manager` (reference README): the same task files, served by plenoctree_b200.octree.task_manager."""
import sys

from plenoctree_b200.octree.task_manager import *  # noqa: F401,F403
from plenoctree_b200.octree.task_manager import main

if __name__ == "__main__":
    sys.exit(main())
