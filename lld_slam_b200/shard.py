"""Host-side partitioning of the hot path over ranks (one process per GPU).

* independent units (local-BA windows, frames, frame pairs) shard with no collective: `unit_range`;
* the global BA partitions LANDMARKS in contiguous blocks with every keyframe replicated: `landmark_shard` builds the
  per-rank view of an `lld_ba_problem` dict exactly as `lld_ba_global` does internally (bounds come from the library's own
  `lld_ba_shard_bounds`, so the test of this helper is a test of the product's partition).  The per-landmark Schur loop
  being partitioned is Thirdparty/g2o/g2o/core/block_solver.hpp:381-432 (independent iterations, sum into Hschur).
"""
from __future__ import annotations

import numpy as np

from . import capi


def unit_range(n_units: int, rank: int, world: int):
    """contiguous block of independent units owned by `rank` (windows / frames / pairs)"""
    return (n_units * rank) // world, (n_units * (rank + 1)) // world


def landmark_bounds(n_pt: int, n_ln: int, rank: int, world: int):
    out = np.zeros(4, np.int32)
    capi.load_library().dll.lld_ba_shard_bounds(n_pt, n_ln, rank, world, out.ctypes.data_as(capi.c_i32p))
    return tuple(int(x) for x in out)


def landmark_shard(p: dict, rank: int, world: int) -> dict:
    """rank-local global-BA problem: all keyframes, points [plo, phi) and lines [llo, lhi) with re-based CSR offsets"""
    assert int(p["n_win"]) == 1
    n_pt, n_ln = int(p["pt_off"][-1]), int(p["ln_off"][-1])
    plo, phi, llo, lhi = landmark_bounds(n_pt, n_ln, rank, world)
    pe0, pe1 = int(p["pt_obs_off"][plo]), int(p["pt_obs_off"][phi])
    lc0, lc1 = int(p["ln_obs_off"][llo]), int(p["ln_obs_off"][lhi])
    s = dict(p)
    s["pt_off"] = np.array([0, phi - plo], np.int32)
    s["ln_off"] = np.array([0, lhi - llo], np.int32)
    s["pt_xyz"] = np.ascontiguousarray(p["pt_xyz"][plo:phi])
    s["pt_obs_off"] = (p["pt_obs_off"][plo:phi + 1] - pe0).astype(np.int32)
    for k in ("pt_obs_kf", "pt_obs_uvr", "pt_obs_info"):
        s[k] = np.ascontiguousarray(p[k][pe0:pe1])
    s["ln_x0_dir"] = np.ascontiguousarray(p["ln_x0_dir"][llo:lhi])
    s["ln_obs_off"] = (p["ln_obs_off"][llo:lhi + 1] - lc0).astype(np.int32)
    for k in ("ln_obs_kf", "ln_obs_left", "ln_obs_right", "ln_obs_info", "ln_obs_stereo"):
        s[k] = np.ascontiguousarray(p[k][lc0:lc1])
    return s
