"""Python-side callers of the C-ABI (include/lldba.h): allocate outputs, fill the ctypes structs, call.

`impl` selects the library:  "gpu" = the product (CUDA, needs a Context), "oracle" = the CPU oracle
(tests / cpu_baseline only).  The product path never routes through the oracle.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional

import numpy as np

from . import capi


def _lib(impl: str):
    if impl == "gpu":
        return capi.load_library()
    if impl == "oracle":
        return capi.load_oracle()
    raise ValueError(impl)


def _handle(impl: str, ctx: Optional[capi.Context]):
    if impl == "gpu":
        if ctx is None:
            raise ValueError("the CUDA library needs a Context (lld_ctx_create)")
        return ctx.handle
    return C.c_void_p()


def _check(impl, ctx, rc, what):
    if rc != 0:
        if impl == "gpu" and ctx is not None:
            ctx.check(rc, what)
        raise RuntimeError(f"{what} failed with {rc}")


def _stop_ptr(stop):
    if stop is None:
        return None, None
    arr = np.ascontiguousarray(stop, dtype=np.uint8)
    return arr.ctypes.data_as(capi.c_u8p), arr


def _ba_outputs(p: Dict[str, Any], log_stride: int):
    n_kf = int(p["kf_off"][-1]); n_pt = int(p["pt_off"][-1]); n_ln = int(p["ln_off"][-1])
    n_pe = int(p["pt_obs_off"][-1]); n_lc = int(p["ln_obs_off"][-1]); nw = int(p["n_win"])
    return dict(
        kf_Tcw=np.zeros((n_kf, 12)), pt_xyz=np.zeros((n_pt, 3)), ln_x0_dir=np.zeros((n_ln, 6)),
        pt_obs_bad=np.zeros(n_pe, np.uint8), ln_obs_bad=np.zeros((n_lc, 2), np.uint8), ln_removed=np.zeros(n_ln, np.uint8),
        log_stride=log_stride,
        chi2_log=np.zeros((nw, log_stride)), lambda_log=np.zeros((nw, log_stride)),
        trials_log=np.zeros((nw, log_stride), np.int32), n_iter_done=np.zeros((nw, 2), np.int32),
    )


def ba_local(p: Dict[str, Any], its1: int = 5, its2: int = 15, *, impl: str = "gpu", ctx=None, stop=None):
    lib = _lib(impl)
    prob, keep = capi.fill_struct(capi.BaProblem, p)
    out = _ba_outputs(p, its1 + its2 + 2)
    res, keep2 = capi.fill_struct(capi.BaResult, out)
    sp, sk = _stop_ptr(stop)
    rc = lib.ba_local(_handle(impl, ctx), C.byref(prob), its1, its2, sp, C.byref(res))
    _check(impl, ctx, rc, "lld_ba_local")
    return out


def ba_global(p: Dict[str, Any], n_iter: int = 10, *, impl: str = "gpu", ctx=None, stop=None):
    lib = _lib(impl)
    prob, keep = capi.fill_struct(capi.BaProblem, p)
    out = _ba_outputs(p, n_iter + 2)
    res, keep2 = capi.fill_struct(capi.BaResult, out)
    sp, sk = _stop_ptr(stop)
    rc = lib.ba_global(_handle(impl, ctx), C.byref(prob), n_iter, sp, C.byref(res))
    _check(impl, ctx, rc, "lld_ba_global")
    return out


def pose_opt(p: Dict[str, Any], *, impl: str = "gpu", ctx=None):
    lib = _lib(impl)
    prob, keep = capi.fill_struct(capi.PoseProblem, p)
    F = int(p["n_frames"])
    out = dict(Tcw=np.zeros((F, 12)), pt_outlier=np.zeros(int(p["pt_off"][-1]), np.uint8),
               ln_outlier=np.zeros(int(p["ln_off"][-1]), np.uint8), n_inliers=np.zeros(F, np.int32),
               chi2_final=np.zeros(F))
    res, keep2 = capi.fill_struct(capi.PoseResult, out)
    rc = lib.pose_opt(_handle(impl, ctx), C.byref(prob), C.byref(res))
    _check(impl, ctx, rc, "lld_pose_opt")
    return out


def _sbp_out(n_cur, n_q, n_pairs):
    return dict(match=np.full(n_cur, -1, np.int32), n_matches=np.zeros(n_pairs, np.int32),
                best_idx=np.full(n_q, -1, np.int32), best_dist=np.full(n_q, 256, np.int32))


def sbp_frame(p: Dict[str, Any], *, impl: str = "gpu", ctx=None):
    lib = _lib(impl)
    geom, gk = capi.make_geom(p["geom"])
    f = dict(p); f["geom"] = geom
    prob, keep = capi.fill_struct(capi.SbpFrameProblem, f)
    out = _sbp_out(int(p["cur_off"][-1]), int(p["last_off"][-1]), int(p["n_pairs"]))
    res, keep2 = capi.fill_struct(capi.SbpResult, out)
    rc = lib.sbp_frame(_handle(impl, ctx), C.byref(prob), C.byref(res))
    _check(impl, ctx, rc, "lld_sbp_frame")
    return out


def sbp_mappoints(p: Dict[str, Any], *, impl: str = "gpu", ctx=None):
    lib = _lib(impl)
    geom, gk = capi.make_geom(p["geom"])
    f = dict(p); f["geom"] = geom
    prob, keep = capi.fill_struct(capi.SbpMpProblem, f)
    out = _sbp_out(int(p["cur_off"][-1]), int(p["mp_off"][-1]), int(p["n_pairs"]))
    res, keep2 = capi.fill_struct(capi.SbpResult, out)
    rc = lib.sbp_mappoints(_handle(impl, ctx), C.byref(prob), C.byref(res))
    _check(impl, ctx, rc, "lld_sbp_mappoints")
    return out


def line_match(p: Dict[str, Any], *, impl: str = "gpu", ctx=None):
    lib = _lib(impl)
    prob, keep = capi.fill_struct(capi.LineMatchProblem, p)
    n_left = int(p["left_off"][-1])
    out = dict(match=np.full(n_left, -1, np.int32), dist=np.full(n_left, np.inf, np.float32))
    res, keep2 = capi.fill_struct(capi.LineMatchResult, out)
    rc = lib.line_match(_handle(impl, ctx), C.byref(prob), C.byref(res))
    _check(impl, ctx, rc, "lld_line_match")
    return out


def descriptor_distance(a: np.ndarray, b: np.ndarray, *, impl: str = "gpu") -> int:
    lib = _lib(impl)
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    assert a.size == 32 and b.size == 32
    return int(lib.descriptor_distance(a.ctypes.data_as(capi.c_u8p), b.ctypes.data_as(capi.c_u8p)))
