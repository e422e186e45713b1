import numpy as np, sys
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lld_slam_b200 import api, capi, synth
ctx = capi.Context(0)
for seed in (1234, 1235, 1236, 1237):
    p = synth.make_local_ba_batch(2, 6, 300, 60, seed)
    g = api.ba_local(p, 5, 15, impl="gpu", ctx=ctx)
    o = api.ba_local(p, 5, 15, impl="oracle")
    rel = np.abs(g["chi2_log"] - o["chi2_log"]) / np.maximum(np.abs(o["chi2_log"]), 1e-12)
    print(seed, "iters", g["n_iter_done"].tolist(), o["n_iter_done"].tolist(), "max rel", rel.max())
    np.set_printoptions(linewidth=250, precision=3)
    print(" rel per it w0", rel.reshape(2, -1)[0][:22])
    print(" rel per it w1", rel.reshape(2, -1)[1][:22])
    print(" flags equal", np.array_equal(g["pt_obs_bad"], o["pt_obs_bad"]), "pose diff", np.abs(g["kf_Tcw"] - o["kf_Tcw"]).max(), "trials eq", np.array_equal(g["trials_log"], o["trials_log"]))
