"""The C++ shim (lld_slam_b200/host/lld_shim.h) executed end to end on POD mirrors of the reference's objects.

tests/hostcheck/shim_run.cpp builds KeyFrame / MapPoint / MapLine / Frame graphs from a synthetic flattened problem, checks
that the shim's own flattening reproduces that problem array by array (ordering: std::map<KeyFrame*> iteration, proj_map by
keyframe id, local-then-fixed keyframes; float narrowing; entry-point constants), runs the reference-named entry point and
dumps what it wrote back.  On a CPU-only box the driver links against the oracle (same C-ABI, lldo_ prefix) — that covers
the host logic; the `gpu` tests link the same driver against liblldba.so.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import shim_io  # noqa: E402
from lld_slam_b200 import api, synth  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "hostcheck", "shim_run.cpp")


def _build(tmp_path, oracle):
    exe = str(tmp_path / ("shim_run_oracle" if oracle else "shim_run_gpu"))
    if oracle:
        libdir, lib, extra = os.path.join(ROOT, "oracle"), "lld_oracle", ["-DLLD_SHIM_ORACLE"]
    else:
        libdir, lib, extra = os.path.join(ROOT, "lld_slam_b200", "csrc"), "lldba", []
    r = subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", *extra, SRC, "-o", exe, f"-L{libdir}", f"-l{lib}", f"-Wl,-rpath,{libdir}"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


def _run(exe, mode, dump, tmp_path):
    fin, fout = str(tmp_path / f"{mode}_in.bin"), str(tmp_path / f"{mode}_out.bin")
    shim_io.write(fin, dump)
    r = subprocess.run([exe, mode, fin, fout], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, f"shim_run {mode} rc={r.returncode}: {r.stdout}"
    return shim_io.read(fout)


def _check_local(exe, tmp_path, impl, ctx=None):
    rng = np.random.default_rng(5)
    p = synth.batch_ba([synth.make_ba_window(6, 300, 60, rng, n_fixed_extra=2)], "local")
    got = _run(exe, "local", shim_io.ba_dump(p, gamma=1.0), tmp_path)
    ref = api.ba_local(p, 5, 15, impl=impl, ctx=ctx)
    local = p["kf_fixed"] == 0
    local[0] = True                                # keyframe 0 is fixed by id but belongs to lLocalKeyFrames: SetPose is called on it
    T = got["kf_Tcw"].reshape(-1, 12)
    assert np.array_equal(T[local], ref["kf_Tcw"][local].astype(np.float32))
    assert np.array_equal(T[~local], p["kf_Tcw"][~local].astype(np.float32))         # lFixedCameras are not written back
    assert np.array_equal(got["pt_xyz"].reshape(-1, 3), ref["pt_xyz"].astype(np.float32))
    keep = ref["ln_removed"] == 0
    L = got["ln_x0_dir"].reshape(-1, 6)
    assert np.array_equal(L[keep], ref["ln_x0_dir"][keep]) and np.array_equal(L[~keep], p["ln_x0_dir"][~keep])
    # vToErase = (keyframe id, point id) of every flagged observation, in edge order
    e = np.nonzero(ref["pt_obs_bad"])[0]
    pt = np.searchsorted(p["pt_obs_off"], e, side="right") - 1
    assert np.array_equal(got["vToErase"].reshape(-1, 2), np.stack([p["pt_obs_kf"][e], pt], 1).astype(np.int32))
    exp = []
    for ln in range(int(p["ln_off"][-1])):
        if ref["ln_removed"][ln]:
            continue
        for c in range(int(p["ln_obs_off"][ln]), int(p["ln_obs_off"][ln + 1])):
            for s in range(2):
                if ref["ln_obs_bad"][c, s]:
                    exp.append((int(p["ln_obs_kf"][c]), ln))
    assert np.array_equal(got["vToEraseLines"].reshape(-1, 2), np.array(exp, np.int32).reshape(-1, 2))
    assert ref["pt_obs_bad"].sum() > 0 and ref["ln_removed"].sum() > 0


def _lines_with_4_observations(p):
    """BundleAdjustment skips map lines with fewer than 4 observations (src/Optimizer.cc:473): keep only those, so that the
    shim's flattening can be compared one to one"""
    q = dict(p)
    off = p["ln_obs_off"].astype(np.int64)
    cnt = off[1:] - off[:-1]
    keep = np.nonzero(cnt >= 4)[0]
    idx = np.concatenate([np.arange(off[i], off[i + 1]) for i in keep])
    q["ln_x0_dir"] = np.ascontiguousarray(p["ln_x0_dir"][keep])
    for k in ("ln_obs_kf", "ln_obs_left", "ln_obs_right", "ln_obs_info", "ln_obs_stereo"):
        q[k] = np.ascontiguousarray(p[k][idx])
    q["ln_obs_off"] = np.concatenate([[0], np.cumsum(cnt[keep])]).astype(np.int32)
    q["ln_off"] = np.array([0, len(keep)], np.int32)
    return q


def _check_global(exe, tmp_path, impl, ctx=None, loop_kf=0):
    p = _lines_with_4_observations(synth.make_global_ba(30, 1500, 300, 13, robust_points=True))
    got = _run(exe, "global", shim_io.ba_dump(p, n_iter=8, nLoopKF=loop_kf), tmp_path)
    ref = api.ba_global(p, 8, impl=impl, ctx=ctx)
    kT, kP = ("kf_TcwGBA", "pt_xyzGBA") if loop_kf else ("kf_Tcw", "pt_xyz")
    assert np.array_equal(got[kT].reshape(-1, 12), ref["kf_Tcw"].astype(np.float32))
    assert np.array_equal(got[kP].reshape(-1, 3), ref["pt_xyz"].astype(np.float32))
    assert np.array_equal(got["ln_x0_dir"].reshape(-1, 6), ref["ln_x0_dir"])
    if loop_kf:   # the live state is untouched when the result goes to the *GBA shadow fields (src/Optimizer.cc:505-539)
        assert np.array_equal(got["kf_Tcw"].reshape(-1, 12), p["kf_Tcw"].astype(np.float32))


def _check_pose(exe, tmp_path, impl, ctx=None):
    p = synth.make_pose_batch(5, 250, 50, 29)
    # the shim reproduces the reference's vnStereoLines[line id] indexing (src/Optimizer.cc:894-898): one flag is pushed per
    # EDGE but read per LINE; build the same selector for the flattened comparison run
    q = dict(p)
    gate = np.zeros_like(p["ln_gate_stereo"])
    for f in range(int(p["n_frames"])):
        a, b = int(p["ln_off"][f]), int(p["ln_off"][f + 1])
        st = p["ln_stereo"][a:b]
        per_edge = []
        for s in st:
            per_edge += [bool(s)] * (2 if s else 1)
        for j in range(b - a):
            gate[a + j, :] = per_edge[j] if j < len(per_edge) else st[j]
    q["ln_gate_stereo"] = gate
    got = _run(exe, "pose", shim_io.pose_dump(p), tmp_path)
    ref = api.pose_opt(q, impl=impl, ctx=ctx)
    assert np.array_equal(got["Tcw"].reshape(-1, 12), ref["Tcw"].astype(np.float32))
    assert np.array_equal(got["pt_outlier"], ref["pt_outlier"]) and np.array_equal(got["ln_outlier"], ref["ln_outlier"])
    assert np.array_equal(got["n_inliers"], ref["n_inliers"])


def test_shim_local_ba_on_host(tmp_path):
    _check_local(_build(tmp_path, True), tmp_path, "oracle")


@pytest.mark.parametrize("loop_kf", [0, 7])
def test_shim_global_ba_on_host(tmp_path, loop_kf):
    _check_global(_build(tmp_path, True), tmp_path, "oracle", loop_kf=loop_kf)


def test_shim_pose_optimization_on_host(tmp_path):
    _check_pose(_build(tmp_path, True), tmp_path, "oracle")


@pytest.mark.gpu
def test_shim_entry_points_on_gpu(tmp_path, gpu_ctx):
    """the same driver linked against liblldba.so: flatten -> lld_* on the device -> write-back, against the library called
    directly through the C-ABI on the flattened problem (bit-identical) and, through that, against the oracle"""
    exe = _build(tmp_path, False)
    _check_local(exe, tmp_path, "gpu", gpu_ctx)
    _check_global(exe, tmp_path, "gpu", gpu_ctx)
    _check_pose(exe, tmp_path, "gpu", gpu_ctx)
