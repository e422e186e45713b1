#!/usr/bin/env python
"""Wall-clock of the call a SLAM thread makes: lld_ba_local on ONE window (10 KF / 5k points / 1k lines by default) with
pinned host buffers, a different window (new structure) on every call.  LLD_TIMING=1 prints the host-stage timeline."""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lld_slam_b200 import api, capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kf", type=int, default=10)
ap.add_argument("--pts", type=int, default=5000)
ap.add_argument("--lines", type=int, default=1000)
ap.add_argument("--reps", type=int, default=12)
a = ap.parse_args()
lib = capi.load_library()
ctx = capi.Context(0)
probs = []
for i in range(4):   # four different windows, used round robin: the topology cache never hits
    p = synth.make_local_ba_batch(1, a.kf, a.pts, a.lines, synth.seed_for(1) + 50 + i)
    pp, k = bench.pinned_problem(p)
    prob, k2 = capi.fill_struct(capi.BaProblem, pp)
    out = api._ba_outputs(p, 22)
    outp, k3 = bench.pinned_problem(out)
    res, k4 = capi.fill_struct(capi.BaResult, outp)
    probs.append((prob, res, outp, (k, k2, k3, k4, pp)))
ts = []
for i in range(a.reps + 4):
    prob, res, outp, _ = probs[i % 4]
    t = time.perf_counter()
    ctx.check(lib.ba_local(ctx.handle, C.byref(prob), 5, 15, None, C.byref(res)), "ba_local")
    ts.append(1e3 * (time.perf_counter() - t))
ts = sorted(ts[4:])
print(json.dumps({"ms_per_call_median": ts[len(ts) // 2], "ms_min": ts[0], "ms_max": ts[-1], "device_ms_last": ctx.last_timing(),
                  "iters": int(outp["n_iter_done"].sum())}))
