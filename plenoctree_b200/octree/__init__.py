"""PlenOctree side of the path: svox-compatible N3Tree / VolumeRenderer over the CUDA library, and mirrors of
octree/extraction.py and octree/optimization.py (§8f ranks 2-3)."""
from .n3tree import DataFormat, N3Tree, N3TreeView  # noqa: F401
from .renderer import Rays, VolumeRenderer  # noqa: F401
