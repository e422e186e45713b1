"""Generates tests/golden/oracle_golden.json from the CPU oracle on small seeded inputs.

The reference ships no tests / golden vectors and cannot be built here (no Eigen / OpenCV / LBDMOD), so these
fixtures pin the ORACLE against regressions; they were produced by `python tests/golden/make_golden.py` at the commit
that introduced them, after the oracle had passed its numeric-Jacobian / numpy / cv2 cross-checks."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lld_slam_b200 import api, synth  # noqa: E402

g = {}
seed = 4242
p = synth.make_local_ba_batch(1, 5, 120, 30, seed)
o = api.ba_local(p, 5, 15, impl="oracle")
g["ba_local"] = dict(seed=seed, n_iter_done=o["n_iter_done"].tolist(), chi2_log=o["chi2_log"][0].tolist(),
                     n_pt_bad=int(o["pt_obs_bad"].sum()), n_ln_bad=int(o["ln_obs_bad"].sum()), kf_Tcw=o["kf_Tcw"].tolist())
q = synth.make_pose_batch(3, 80, 20, seed + 1)
r = api.pose_opt(q, impl="oracle")
g["pose"] = dict(seed=seed + 1, n_inliers=r["n_inliers"].tolist(), Tcw=r["Tcw"].tolist())
m = synth.make_sbp_frame_batch(2, 300, seed + 2)
s = api.sbp_frame(m, impl="oracle")
g["sbp"] = dict(seed=seed + 2, n_matches=s["n_matches"].tolist(), dist_sum=int(s["best_dist"][s["best_idx"] >= 0].sum()),
                match=s["match"].tolist())
lm = synth.make_line_match_batch(2, 60, 64, seed + 3)
t = api.line_match(lm, impl="oracle")
g["lines"] = dict(seed=seed + 3, match=t["match"].tolist())
json.dump(g, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json"), "w"))
print("written")
