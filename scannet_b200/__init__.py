"""scannet_b200 — B200-native implementation of the ScanNet toolkit's per-frame hot path
(TSDF voxel-block fusion of .sens streams, Segmentator mesh over-segmentation, SensReader I/O).

The work is done by hand-written sm_100a CUDA kernels in ``csrc/`` behind the C ABI declared in
``include/scannet_b200.h``; these Python modules are thin ctypes mirrors used by tests/bench."""
from ._lib import LIB_PATH, ScnError  # noqa: F401
