"""how close to the chi2 gate are the outlier flags that differ between the GPU and the oracle at configs[3] size?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from lld_slam_b200 import api, capi, synth
import test_gpu_parity as t
ctx = capi.Context(0)
for seed in (synth.seed_for(4), synth.seed_for(4) + 1, synth.seed_for(4) + 2):
    p16 = synth.make_local_ba_batch(16, 20, 5000, 1000, seed)
    g = api.ba_local(p16, 5, 15, impl="gpu", ctx=ctx)
    o = api.ba_local(p16, 5, 15, impl="oracle")
    flip = np.nonzero(g["pt_obs_bad"] != o["pt_obs_bad"])[0]
    out = []
    for e in flip:
        c2, th = t._point_edge_chi2(p16, g, int(e))
        c2o, _ = t._point_edge_chi2(p16, o, int(e))
        out.append((int(e), abs(c2 - th) / th, abs(c2o - th) / th))
    print(seed, "flips", len(flip), out, "ln flips", int((g["ln_obs_bad"] != o["ln_obs_bad"]).sum()))
