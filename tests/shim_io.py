"""Binary exchange format between the Python tests and tests/hostcheck/shim_run.cpp (test infrastructure):
a sequence of records  [u32 name length][name][u8 type 'd' | 'f' | 'i' | 'b'][u64 byte count][bytes]."""
import struct

import numpy as np

from lld_slam_b200 import synth

_T = {"d": np.float64, "f": np.float32, "i": np.int32, "b": np.uint8}


def write(path, arrays):
    with open(path, "wb") as f:
        for k, v in arrays.items():
            a = np.ascontiguousarray(v)
            t = {np.dtype(np.float64): "d", np.dtype(np.float32): "f", np.dtype(np.int32): "i", np.dtype(np.uint8): "b"}[a.dtype]
            name = k.encode()
            f.write(struct.pack("<I", len(name))); f.write(name); f.write(t.encode()); f.write(struct.pack("<Q", a.nbytes)); f.write(a.tobytes())


def read(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            h = f.read(4)
            if len(h) < 4:
                break
            (nl,) = struct.unpack("<I", h)
            name = f.read(nl).decode()
            t = f.read(1).decode()
            (nb,) = struct.unpack("<Q", f.read(8))
            out[name] = np.frombuffer(f.read(nb), dtype=_T[t]).copy()
    return out


def ba_dump(p, gamma=1.0, n_iter=10, nLoopKF=0):
    """single-window lld_ba_problem field dict -> shim_run input"""
    assert int(p["n_win"]) == 1
    d = {k: p[k] for k in ("kf_Tcw", "kf_fixed", "kf_intr", "kf_line_cam", "pt_xyz", "pt_obs_off", "pt_obs_kf", "pt_obs_uvr", "pt_obs_info",
                           "ln_x0_dir", "ln_obs_off", "ln_obs_kf", "ln_obs_left", "ln_obs_right", "ln_obs_info", "ln_obs_stereo")}
    for k in ("delta_pt_mono", "delta_pt_stereo", "delta_ln_mono", "delta_ln_stereo", "robust_points", "ln_endpoints_normalized"):
        d[k] = np.array([float(p[k])], np.float64)
    d["gamma"] = np.array([gamma], np.float64)
    d["n_iter"] = np.array([float(n_iter)], np.float64)
    d["nLoopKF"] = np.array([float(nLoopKF)], np.float64)
    d["inv_level_sigma2"] = np.ascontiguousarray(synth.inv_level_sigma2(), np.float32)
    return d


def pose_dump(p, gamma=0.5):
    d = {k: p[k] for k in ("Tcw", "intr", "pt_off", "pt_xw", "pt_uvr", "pt_info", "ln_off", "ln_x0_dir", "ln_left", "ln_right", "ln_info")}
    d["gamma"] = np.array([gamma], np.float64)
    d["n_inliers_slot"] = np.zeros(int(p["n_frames"]), np.int32)
    d["inv_level_sigma2"] = np.ascontiguousarray(synth.inv_level_sigma2(), np.float32)
    return d
