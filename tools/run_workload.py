#!/usr/bin/env python
"""Runs one resident workload a few times (for ncu captures; never a bench number)."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lld_slam_b200 import api, capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("workload", choices=["ba", "match", "pose", "single", "lines"])
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--windows", type=int, default=64)
ap.add_argument("--pairs", type=int, default=1024)
ap.add_argument("--frames", type=int, default=512)
ap.add_argument("--time", action="store_true", help="also print a CUDA-event timing (not under ncu)")
a = ap.parse_args()
lib = capi.load_library()
d = bench.bind_resident(lib)
ctx = capi.Context(0)
if a.workload in ("ba", "single"):
    p = bench.make_batch(a.windows, 7) if a.workload == "ba" else synth.make_local_ba_batch(1, 10, 5000, 1000, 7)
    prob, keep = capi.fill_struct(capi.BaProblem, p)
    ctx.check(d.lld_ba_upload(ctx.handle, C.byref(prob), 0, 22), "upload")
    for _ in range(a.reps):
        ctx.check(d.lld_ba_run_local(ctx.handle, 5, 15, None), "run")
    d.lld_ba_sync(ctx.handle)
elif a.workload == "match":
    m = synth.make_sbp_frame_batch(a.pairs, 2000, 3)
    geom, gk = capi.make_geom(m["geom"])
    f = dict(m); f["geom"] = geom
    f.setdefault("th_high", 0); f.setdefault("allow_negative_depth", 0)
    prob, keep = capi.fill_struct(capi.SbpFrameProblem, f)
    ctx.check(d.lld_sbp_frame_upload(ctx.handle, C.byref(prob)), "upload")
    n = C.c_int()
    for _ in range(a.reps):
        ctx.check(d.lld_sbp_run(ctx.handle, C.byref(n)), "run")
    d.lld_ba_sync(ctx.handle)
    if a.time:
        d.lld_ctx_event_record(ctx.handle, 0)
        for _ in range(20):
            ctx.check(d.lld_sbp_run(ctx.handle, C.byref(n)), "run")
        d.lld_ctx_event_record(ctx.handle, 1)
        d.lld_ba_sync(ctx.handle)
        print("match ms per batch:", float(d.lld_ctx_event_elapsed_ms(ctx.handle)) / 20)
elif a.workload == "pose":
    pz = synth.make_pose_batch(a.frames, 1500, 300, 3)
    prob, keep = capi.fill_struct(capi.PoseProblem, pz)
    ctx.check(d.lld_pose_upload(ctx.handle, C.byref(prob)), "upload")
    for _ in range(a.reps):
        ctx.check(d.lld_pose_run(ctx.handle), "run")
    d.lld_ba_sync(ctx.handle)
else:
    lm = synth.make_line_match_batch(a.pairs, 500, 64, 3)
    for _ in range(a.reps):
        api.line_match(lm, impl="gpu", ctx=ctx)
print("done", a.workload)
