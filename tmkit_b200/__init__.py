"""tmkit_b200 - B200-native batched motion-planning validity checking (scene-graph FK +
collision query) behind the Amino-shaped C ABI TMKit drives.

Package contents: csrc/ (host C scene graph + sm_100a CUDA kernels -> libtmkcl.so),
amino.py (ctypes mirror of the C ABI), scenes.py / workloads.py (benchmark fixtures and
synthetic inputs).  The CPU oracle lives outside the package, in oracle/, and is test
infrastructure only.
"""
__all__ = ["amino", "scenes", "workloads"]
