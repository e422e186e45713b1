#!/usr/bin/env python
"""Latency probe: one local-BA window at the north_star target shape (10 KF / 5k points / 1k lines), schedule 5+15.
Prints ms per call (CUDA events on the library stream) and, with --profile, per-kernel event times of one call."""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lld_slam_b200 import api, capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kf", type=int, default=10)
ap.add_argument("--pts", type=int, default=5000)
ap.add_argument("--lines", type=int, default=1000)
ap.add_argument("--windows", type=int, default=1)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--profile", action="store_true")
a = ap.parse_args()
lib = capi.load_library()
d = bench.bind_resident(lib)
ctx = capi.Context(0)
p = synth.make_local_ba_batch(a.windows, a.kf, a.pts, a.lines, synth.seed_for(1) + 5)
prob, keep = capi.fill_struct(capi.BaProblem, p)
out = api._ba_outputs(p, 22)
res, keep2 = capi.fill_struct(capi.BaResult, out)
ctx.check(d.lld_ba_upload(ctx.handle, C.byref(prob), 0, 22), "upload")
for _ in range(3):
    ctx.check(d.lld_ba_run_local(ctx.handle, 5, 15, None), "run")
d.lld_ba_sync(ctx.handle)
d.lld_ctx_event_record(ctx.handle, 0)
for _ in range(a.reps):
    ctx.check(d.lld_ba_run_local(ctx.handle, 5, 15, None), "run")
d.lld_ctx_event_record(ctx.handle, 1)
d.lld_ba_sync(ctx.handle)
ms = float(d.lld_ctx_event_elapsed_ms(ctx.handle)) / a.reps
ctx.check(d.lld_ba_download(ctx.handle, C.byref(prob), C.byref(res), 1), "download")
its, trials = int(out["n_iter_done"].sum()), int(out["trials_log"].sum())
r = {"ms_per_call": ms, "iters": its, "trials": trials, "lm_iters_per_sec": its / (ms * 1e-3), "us_per_trial": 1e3 * ms / max(trials, 1)}
if a.profile:
    d.lld_ctx_profile(ctx.handle, 1)
    ctx.check(d.lld_ba_run_local(ctx.handle, 5, 15, None), "run")
    prof = bench.profile_report(d, ctx)
    d.lld_ctx_profile(ctx.handle, 0)
    r["kernels_us_per_launch"] = {k: round(1e3 * v["ms"] / max(v["n"], 1), 2) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
print(json.dumps(r))
