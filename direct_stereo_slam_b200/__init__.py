"""dslam-b200: the photometric Gauss-Newton / Scan-Context hot path of IRVLab/direct_stereo_slam on B200 (sm_100a).

The compute lives in libdslam_b200.so (hand-written CUDA behind the C ABI of include/dslam_b200.h); this package is the
thin host-side mirror of the reference's interface.  Importing `direct_stereo_slam_b200.api` loads the library and raises
if it has not been built — there is no fallback path.
"""
from . import _lib  # noqa: F401

__all__ = ["api", "synthetic", "_lib"]
