"""lld_slam_b200 — B200-native point+line bundle adjustment and descriptor matching (hot path of LLD-SLAM).

The product is the CUDA library lld_slam_b200/csrc/liblldba.so behind the C-ABI of include/lldba.h.
This package holds its sources (csrc/), the C++ shim mirroring the reference's classes (host/), and a
ctypes harness (capi.py, api.py) plus the seeded synthetic generator (synth.py) used by tests and bench.
"""
from . import capi, synth  # noqa: F401
from .api import (ba_local, ba_global, pose_opt, sbp_frame, sbp_mappoints, line_match,  # noqa: F401
                  descriptor_distance)
