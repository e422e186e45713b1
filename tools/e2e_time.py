import os, time, ctypes as C, numpy as np
os.environ["LLD_TIMING"]="1"
import sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lld_slam_b200 import api, capi, synth
lib = capi.load_library(); ctx = capi.Context(0)
p = bench.make_batch(64, 5)
pp, k = bench.pinned_problem(p)
prob, k2 = capi.fill_struct(capi.BaProblem, pp)
out = api._ba_outputs(p, 22); outp, k3 = bench.pinned_problem(out); res, k4 = capi.fill_struct(capi.BaResult, outp)
for i in range(3):
    t=time.perf_counter(); ctx.check(lib.ba_local(ctx.handle, C.byref(prob), 5, 15, None, C.byref(res)), "x"); print("total ms", 1e3*(time.perf_counter()-t), ctx.last_timing())
