#!/usr/bin/env python
"""How well-determined are the final poses of the cfg5 global BA?  Runs the CPU oracle twice on the same problem, the
second time with the landmark ORDER permuted (mathematically the same problem: only the order of the floating-point sums
in H_pp, S and chi2 changes), and reports how far the oracle moves against itself.  Test infrastructure (oracle only).

  python tools/gba_spread.py --kf 1500 --pts 300000 --lines 60000 --iters 10 --out profiles/r2_gba_oracle_spread.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lld_slam_b200 import api, synth  # noqa: E402


def permute_landmarks(p, seed):
    """same problem, points and lines in a shuffled order (observations keep their order inside a landmark)"""
    rng = np.random.default_rng(seed)
    q = dict(p)

    def perm(off_key, lm_keys, obs_keys, n):
        order = rng.permutation(n)
        off = p[off_key]
        cnt = (off[1:] - off[:-1])[order]
        new_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
        idx = np.concatenate([np.arange(off[i], off[i + 1]) for i in order]) if n else np.zeros(0, np.int64)
        for k in lm_keys:
            q[k] = np.ascontiguousarray(p[k][order])
        for k in obs_keys:
            q[k] = np.ascontiguousarray(p[k][idx])
        q[off_key] = new_off
        return order

    op = perm("pt_obs_off", ["pt_xyz"], ["pt_obs_kf", "pt_obs_uvr", "pt_obs_info"], int(p["pt_off"][-1]))
    ol = perm("ln_obs_off", ["ln_x0_dir"], ["ln_obs_kf", "ln_obs_left", "ln_obs_right", "ln_obs_info", "ln_obs_stereo"],
              int(p["ln_off"][-1]))
    return q, op, ol


def reverse_keyframes(p):
    """same problem, keyframes renumbered in reverse (k -> n-1-k): the reduced camera system is eliminated from the other
    end of the chain, i.e. a different pivot order of the LDL^T (edges keep their order inside a landmark)"""
    q = dict(p)
    n = int(p["kf_off"][-1])
    for k in ("kf_Tcw", "kf_fixed", "kf_intr", "kf_line_cam"):
        q[k] = np.ascontiguousarray(p[k][::-1])
    q["pt_obs_kf"] = (n - 1 - p["pt_obs_kf"]).astype(np.int32)
    q["ln_obs_kf"] = (n - 1 - p["ln_obs_kf"]).astype(np.int32)
    return q


def rot_angle(Ta, Tb):
    Ra = Ta[:, :9].reshape(-1, 3, 3); Rb = Tb[:, :9].reshape(-1, 3, 3)
    D = np.einsum("nji,njk->nik", Ra, Rb) - np.eye(3)          # Ra^T Rb - I ~ [w]x : |w| = ||.||_F / sqrt(2), no arccos cancellation
    return np.sqrt((D * D).sum(axis=(1, 2)) / 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kf", type=int, default=1500)
    ap.add_argument("--pts", type=int, default=300000)
    ap.add_argument("--lines", type=int, default=60000)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--robust", action="store_true")
    ap.add_argument("--reverse-kf", action="store_true", help="second run = reversed keyframe numbering instead of permuted landmarks")
    ap.add_argument("--out", default="")
    ap.add_argument("--npz", default="", help="per-keyframe differences and both final pose sets")
    a = ap.parse_args()
    p = synth.make_global_ba(a.kf, a.pts, a.lines, synth.seed_for(5), robust_points=a.robust)
    t0 = time.time()
    o1 = api.ba_global(p, a.iters, impl="oracle")
    t1 = time.time()
    if a.reverse_kf:
        q = reverse_keyframes(p)
        o2 = dict(api.ba_global(q, a.iters, impl="oracle"))
        o2["kf_Tcw"] = o2["kf_Tcw"][::-1]
        op = np.arange(int(p["pt_off"][-1]))
    else:
        q, op, ol = permute_landmarks(p, 7)
        o2 = api.ba_global(q, a.iters, impl="oracle")
    dT = np.abs(o1["kf_Tcw"][:, 9:] - o2["kf_Tcw"][:, 9:]).max(axis=1)
    dR = rot_angle(o1["kf_Tcw"], o2["kf_Tcw"])
    dX = np.abs(o1["pt_xyz"][op] - o2["pt_xyz"]).max()
    rel = np.abs(o1["chi2_log"] - o2["chi2_log"]) / np.maximum(np.abs(o1["chi2_log"]), 1e-9)
    res = dict(variant="reversed keyframe numbering" if a.reverse_kf else "permuted landmark order", shape=[a.kf, a.pts, a.lines], iters=a.iters, robust=bool(a.robust),
               oracle_seconds=round(t1 - t0, 2),
               pose_t_max_m=float(dT.max()), pose_t_median_m=float(np.median(dT)), pose_t_argmax_kf=int(dT.argmax()),
               pose_rot_max_rad=float(dR.max()), point_max_m=float(dX),
               chi2_rel_max=float(rel.max()), chi2_final=[float(o1["chi2_log"][0, int(o1["n_iter_done"][0, 0])]),
                                                          float(o2["chi2_log"][0, int(o2["n_iter_done"][0, 0])])],
               same_trials=bool(np.array_equal(o1["trials_log"], o2["trials_log"])),
               pose_t_by_kf_decile=[float(x) for x in np.quantile(dT, np.linspace(0, 1, 11))])
    if a.npz:
        np.savez_compressed(a.npz, dT=dT, dR=dR, kf1=o1["kf_Tcw"], kf2=o2["kf_Tcw"])
    print(json.dumps(res))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
