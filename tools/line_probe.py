#!/usr/bin/env python
"""Timing probe of the stereo line matcher (1024 pairs x 500 x 500, D = 64): device ms + per-kernel event times."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lld_slam_b200 import api, capi, synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lib = capi.load_library(); d = bench.bind_resident(lib); ctx = capi.Context(0)
lm = synth.make_line_match_batch(P, 500, 64, 9)
for _ in range(2):
    api.line_match(lm, impl="gpu", ctx=ctx)
ts = []
for _ in range(5):
    api.line_match(lm, impl="gpu", ctx=ctx); ts.append(ctx.last_timing()[1])
d.lld_ctx_profile(ctx.handle, 1)
g = api.line_match(lm, impl="gpu", ctx=ctx)
prof = bench.profile_report(d, ctx)
d.lld_ctx_profile(ctx.handle, 0)
r = {"pairs": P, "ms": float(np.median(ts)), "kernels_ms": {k: round(v["ms"], 3) for k, v in prof.items()}, "matched": int((g["match"] >= 0).sum())}
if "--check" in sys.argv:
    sub = synth.make_line_match_batch(6, 500, 64, 9)
    gg = api.line_match(sub, impl="gpu", ctx=ctx); oo = api.line_match(sub, impl="oracle")
    same = gg["match"] == oo["match"]; fin = np.isfinite(oo["dist"]) & same
    r["check"] = {"mismatch": int((~same).sum()), "max_abs_dist_err": float(np.abs(gg["dist"][fin] - oo["dist"][fin]).max())}
print(json.dumps(r))
