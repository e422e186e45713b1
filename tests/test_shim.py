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


def _fuse_scene(seed, n_kp=900, n_mp=700):
    """one keyframe looking down +z, keypoints = projections of random 3D points (70 % stereo), candidate map points = those 3D
    points moved by a few centimetres (+ unrelated ones, bad ones, null entries, points already in the keyframe, points behind
    the camera, outside the distance range or seen from behind)"""
    rng = np.random.default_rng(seed)
    f32 = np.float32
    g = synth.frame_geom()
    sf = np.asarray(g["scale_factors"], f32)
    nl = len(sf)
    yaw = 0.1
    R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    t = np.array([0.3, -0.1, 0.5])
    Tcw = np.concatenate([R.reshape(-1), t]).astype(f32)
    Xc = np.stack([rng.uniform(-8, 8, n_kp), rng.uniform(-2.5, 2.5, n_kp), rng.uniform(4, 30, n_kp)], 1)
    Xw = (R.T @ (Xc - t).T).T
    u = g["fx"] * Xc[:, 0] / Xc[:, 2] + g["cx"]; v = g["fy"] * Xc[:, 1] / Xc[:, 2] + g["cy"]
    keep = (u > g["min_x"] + 5) & (u < g["max_x"] - 5) & (v > g["min_y"] + 5) & (v < g["max_y"] - 5)
    Xw, Xc, u, v = Xw[keep], Xc[keep], u[keep], v[keep]
    N = len(u)
    kp_oct = rng.integers(0, nl, N).astype(np.int32)
    stereo = rng.random(N) < 0.7
    kp_ur = np.where(stereo, u - g["bf"] / Xc[:, 2], -1.0).astype(f32)
    kp_desc = rng.integers(0, 256, (N, 32), dtype=np.uint8)
    kp_mp_nobs = np.where(rng.random(N) < 0.3, rng.integers(0, 6, N), -1).astype(np.int32)
    src = rng.integers(0, N, n_mp)
    rel = rng.random(n_mp) < 0.75
    pos = np.where(rel[:, None], Xw[src] + rng.normal(0, 0.02, (n_mp, 3)), rng.uniform(-10, 10, (n_mp, 3)) + np.array([0, 0, 12.0]))
    Ow = -R.T @ t
    dist = np.linalg.norm(pos - Ow, axis=1)
    lvl = np.where(rel, kp_oct[src], rng.integers(0, nl, n_mp))
    # mfMaxDistance such that PredictScale lands on (about) the keypoint's level: ratio = max / dist = 1.2^level
    maxd = dist * sf[0] ** 0 * (1.2 ** (lvl - 0.3))
    mind = maxd / 1.2 ** (nl - 1)
    normal = (Ow - pos) / dist[:, None] * -1.0            # viewing direction from the camera: PO.Pn = dist
    back = rng.random(n_mp) < 0.05
    normal[back] *= -1                                     # seen from behind: rejected
    far = rng.random(n_mp) < 0.05
    maxd[far] = dist[far] * 0.5                            # outside the scale-invariance range
    desc = np.where(rel[:, None], synth._flip_bits(kp_desc[src][None], 0.07, rng)[0], rng.integers(0, 256, (n_mp, 32), dtype=np.uint8)).astype(np.uint8)
    return dict(
        Tcw=Tcw, intr=np.array([g["fx"], g["fy"], g["cx"], g["cy"], g["bf"]], f32),
        bounds=np.array([g["min_x"], g["max_x"], g["min_y"], g["max_y"]], f32), scale_factors=sf,
        inv_level_sigma2=np.ascontiguousarray(synth.inv_level_sigma2()[:nl], f32), log_scale_factor=np.array([np.log(f32(1.2))], f32),
        kp_xy=np.ascontiguousarray(np.stack([u, v], 1), f32), kp_octave=kp_oct, kp_uright=kp_ur, kp_desc=np.ascontiguousarray(kp_desc),
        kp_mp_nobs=kp_mp_nobs,
        mp_pos=np.ascontiguousarray(pos, f32), mp_normal=np.ascontiguousarray(normal, f32),
        mp_minmax=np.ascontiguousarray(np.stack([mind, maxd], 1), f32), mp_desc=np.ascontiguousarray(desc),
        mp_bad=(rng.random(n_mp) < 0.03).astype(np.uint8), mp_in_kf=(rng.random(n_mp) < 0.03).astype(np.uint8),
        mp_null=(rng.random(n_mp) < 0.02).astype(np.uint8), mp_nobs=rng.integers(0, 6, n_mp).astype(np.int32),
        th=np.array([3.0], f32)), g


def _fuse_expected(d, g, impl, ctx):
    """the reference's Fuse from numpy: float32 projection with double accumulation (cv::Mat products), the tests of
    src/ORBmatcher.cc:851-880, MapPoint::PredictScale, then lld_kf_search through the C-ABI and the sequential Replace / Add rule"""
    f32, f64 = np.float32, np.float64
    T = d["Tcw"]; R = T[:9].reshape(3, 3).astype(f64); t = T[9:].astype(f64)
    M = len(d["mp_nobs"])
    Ow = (-(R.T @ t)).astype(f32)
    valid = np.zeros(M, np.uint8); proj = np.zeros((M, 3), f32); lvl = np.zeros(M, np.int32)
    fx, fy, cx, cy, bf = (f32(x) for x in d["intr"])
    minx, maxx, miny, maxy = (f32(x) for x in d["bounds"])
    nl = len(d["scale_factors"])
    for i in range(M):
        if d["mp_null"][i] or d["mp_bad"][i] or d["mp_in_kf"][i]:
            continue
        X = d["mp_pos"][i]
        pc = (R @ X.astype(f64) + t).astype(f32)
        if pc[2] < 0:
            continue
        invz = f32(1) / pc[2]
        x = pc[0] * invz; y = pc[1] * invz
        u = fx * x + cx; v = fy * y + cy
        if not (u >= minx and u < maxx and v >= miny and v < maxy):
            continue
        ur = u - bf * invz
        PO = X - Ow
        dist = f32(np.sqrt(np.sum(PO.astype(f64) ** 2)))
        mind, maxd = d["mp_minmax"][i]
        if dist < f32(0.8) * mind or dist > f32(1.2) * maxd:
            continue
        if float(np.dot(PO.astype(f64), d["mp_normal"][i].astype(f64))) < 0.5 * float(dist):
            continue
        ratio = maxd / dist
        ns = int(np.ceil(np.log(f32(ratio)) / d["log_scale_factor"][0]))
        lvl[i] = min(max(ns, 0), nl - 1)
        valid[i] = 1; proj[i] = (u, v, ur)
    N = len(d["kp_octave"])
    p = dict(n_pairs=1, geom=g, th=float(d["th"][0]), th_low=50, chi2_gate=1, sequential_claims=0,
             inv_level_sigma2=np.pad(d["inv_level_sigma2"], (0, 8 - nl)),
             kp_off=np.array([0, N], np.int32), kp_xy=d["kp_xy"], kp_octave=d["kp_octave"].astype(np.uint8), kp_uright=d["kp_uright"],
             kp_desc=d["kp_desc"], kp_claimed=np.zeros(N, np.uint8),
             mp_off=np.array([0, M], np.int32), mp_valid=valid, mp_proj=proj, mp_level=lvl, mp_desc=d["mp_desc"])
    r = api.kf_search(p, impl=impl, ctx=ctx)
    kf_mp = np.where(d["kp_mp_nobs"] >= 0, 100000 + np.arange(N), -1).astype(np.int32)
    kf_nobs = d["kp_mp_nobs"].copy()
    acts = []
    for i in range(M):
        b = int(r["best_idx"][i])
        if b < 0:
            continue
        if kf_mp[b] >= 0:
            in_kf_nobs = int(kf_nobs[b])
            acts.append((i, b, 2 if in_kf_nobs > int(d["mp_nobs"][i]) else 1))
        else:
            kf_mp[b] = i; kf_nobs[b] = int(d["mp_nobs"][i]) + 1
            acts.append((i, b, 0))
    return acts, kf_mp, int(valid.sum())


def _check_fuse(exe, tmp_path, impl, ctx=None):
    d, g = _fuse_scene(77)
    got = _run(exe, "fuse", d, tmp_path)
    acts, kf_mp, n_valid = _fuse_expected(d, g, impl, ctx)
    assert n_valid > 300 and len(acts) > 150 and len({k for _, _, k in acts}) == 3        # all three outcomes occur
    assert int(got["n_fused"][0]) == len(acts)
    assert np.array_equal(got["act_mp"], [a[0] for a in acts]) and np.array_equal(got["act_idx"], [a[1] for a in acts])
    assert np.array_equal(got["act_kind"], [a[2] for a in acts])
    assert np.array_equal(got["kf_mp"], kf_mp)


def test_shim_fuse_on_host(tmp_path):
    """ORBmatcher::Fuse through the shim (projection and gates on the host, lld_kf_search, sequential Replace / Add) against the
    same function assembled in numpy around the C-ABI call"""
    _check_fuse(_build(tmp_path, True), tmp_path, "oracle")


def _check_tri(exe, tmp_path, impl, ctx=None):
    """SearchForTriangulation through the shim (epipole from the two poses, std::map feature vectors flattened in node order)
    against lld_tri_search called on the flat problem with the epipole computed here the same way"""
    p = synth.make_tri_search_batch(1, 900, 91)
    rng = np.random.default_rng(3)
    f32, f64 = np.float32, np.float64
    g = synth.frame_geom()
    # any two poses: the matcher only sees F12 and the epipole; F12 is taken from the synthetic problem, the epipole must
    # come out of the shim's own arithmetic on these poses
    T1 = np.concatenate([np.eye(3).reshape(-1), [0.1, -0.2, 0.3]]).astype(f32)
    yaw = 0.05
    R2 = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    T2 = np.concatenate([R2.reshape(-1), [0.9, 0.02, 0.4]]).astype(f32)
    R1 = T1[:9].reshape(3, 3).astype(f64); t1 = T1[9:].astype(f64)
    Cw = (-(R1.T @ t1)).astype(f32)
    C2 = (T2[:9].reshape(3, 3).astype(f64) @ Cw.astype(f64) + T2[9:].astype(f64)).astype(f32)
    invz = f32(1.0) / C2[2]
    ep = np.array([f32(g["fx"]) * C2[0] * invz + f32(g["cx"]), f32(g["fy"]) * C2[1] * invz + f32(g["cy"])], f32)
    nl = int(p["n_levels"])
    d = dict(Tcw1=T1, Tcw2=T2, intr=np.array([g["fx"], g["fy"], g["cx"], g["cy"], g["bf"]], f32),
             scale_factors=p["scale_factors"][:nl], level_sigma2=p["level_sigma2"][:nl], F12=p["F12"][0],
             check_orientation=np.array([p["check_orientation"]], np.uint8), only_stereo=np.array([p["only_stereo"]], np.uint8))
    for k in (1, 2):
        d[f"kp_xy{k}"] = p[f"kp{k}_xy"]; d[f"kp_angle{k}"] = p[f"kp{k}_angle"]; d[f"kp_uright{k}"] = p[f"kp{k}_uright"]
        d[f"kp_has_mp{k}"] = p[f"kp{k}_has_mp"]; d[f"kp_desc{k}"] = p[f"kp{k}_desc"]
        d[f"kp_octave{k}"] = p["kp2_octave"] if k == 2 else np.zeros(len(p["kp1_angle"]), np.uint8)
        d[f"fv_node{k}"] = p[f"fv{k}_node"]; d[f"fv_idx_off{k}"] = p[f"fv{k}_idx_off"]; d[f"fv_idx{k}"] = p[f"fv{k}_idx"]
    got = _run(exe, "tri", d, tmp_path)
    assert np.array_equal(got["epipole"], ep)
    q = dict(p); q["epipole"] = ep.reshape(1, 2)
    ref = api.tri_search(q, impl=impl, ctx=ctx)
    idx1 = np.nonzero(ref["match12"] >= 0)[0]
    assert int(got["n_matches"][0]) == int(ref["n_matches"][0]) == len(idx1) and len(idx1) > 150
    assert np.array_equal(got["pairs"].reshape(-1, 2), np.stack([idx1, ref["match12"][idx1]], 1))


def _check_bow(exe, tmp_path, impl, ctx=None):
    """SearchByBoW(KeyFrame*, KeyFrame*) through the shim against lld_bow_search on the flat problem"""
    p = synth.make_bow_search_batch(1, 900, 93, n_nodes=60, strict_th=1, nn_ratio=0.75)
    f32 = np.float32
    g = synth.frame_geom()
    T = np.concatenate([np.eye(3).reshape(-1), [0, 0, 0]]).astype(f32)
    n = len(p["kp1_angle"])
    d = dict(Tcw1=T, Tcw2=T, intr=np.array([g["fx"], g["fy"], g["cx"], g["cy"], g["bf"]], f32),
             scale_factors=np.asarray(g["scale_factors"], f32), level_sigma2=np.asarray(g["scale_factors"], f32) ** 2,
             F12=np.zeros(9, f32), check_orientation=np.array([p["check_orientation"]], np.uint8), only_stereo=np.zeros(1, np.uint8),
             nn_ratio=np.array([p["nn_ratio"]], f32))
    for k in (1, 2):
        d[f"kp_xy{k}"] = np.zeros((n, 2), f32); d[f"kp_angle{k}"] = p[f"kp{k}_angle"]; d[f"kp_uright{k}"] = np.full(n, -1, f32)
        d[f"kp_has_mp{k}"] = p[f"kp{k}_valid"]; d[f"kp_desc{k}"] = p[f"kp{k}_desc"]; d[f"kp_octave{k}"] = np.zeros(n, np.uint8)
        d[f"fv_node{k}"] = p[f"fv{k}_node"]; d[f"fv_idx_off{k}"] = p[f"fv{k}_idx_off"]; d[f"fv_idx{k}"] = p[f"fv{k}_idx"]
    got = _run(exe, "bow", d, tmp_path)
    ref = api.bow_search(p, impl=impl, ctx=ctx)
    assert int(got["n_matches"][0]) == int(ref["n_matches"][0]) > 100
    assert np.array_equal(got["match12"], ref["match12"])       # the driver's map points carry the keypoint index as their id


def test_shim_search_by_bow_on_host(tmp_path):
    _check_bow(_build(tmp_path, True), tmp_path, "oracle")


def test_shim_search_for_triangulation_on_host(tmp_path):
    _check_tri(_build(tmp_path, True), tmp_path, "oracle")


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
    _check_fuse(exe, tmp_path, "gpu", gpu_ctx)
    _check_tri(exe, tmp_path, "gpu", gpu_ctx)
    _check_bow(exe, tmp_path, "gpu", gpu_ctx)
