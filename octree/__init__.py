"""Drop-in module path of the reference (`python -m octree.extraction`, README.md:120-135): thin shims over plenoctree_b200.octree."""
