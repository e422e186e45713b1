#!/usr/bin/env python
"""Timing of the SURVEY §8(f) kernels (ComputeStereoMatches, relocalisation / map-point SearchByProjection, temporal line
association, distinctive descriptors): device compute time from the library's own CUDA events (host buffers in, results
out through the public call), wall clock of the same call, and the CPU oracle on the same or a smaller sample.
Prints one JSON object; results are compared with the oracle where the sample is the same."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from lld_slam_b200 import api, capi, synth

ctx = capi.Context(0)
out = {}


def timed(fn, reps=5):
    fn()
    ms, wall = [], []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); wall.append((time.perf_counter() - t0) * 1e3); ms.append(ctx.last_timing()[1])
    return r, float(np.median(ms)), float(np.median(wall))


def cpu(fn):
    t0 = time.perf_counter(); r = fn(); return r, (time.perf_counter() - t0) * 1e3


# ---- Frame::ComputeStereoMatches, KITTI-sized frames
nf = 32
frames = [synth.make_stereo_frame(100 + s, n_kp=2000, rows=376, cols=1241, n_levels=8) for s in range(nf)]
p = synth.batch_stereo(frames)
g, ms, wall = timed(lambda: api.stereo_matches(p, impl="gpu", ctx=ctx))
p4 = synth.batch_stereo(frames[:4])
o, tc = cpu(lambda: api.stereo_matches(p4, impl="oracle"))
out["compute_stereo_matches"] = {"frames": nf, "keypoints": 2000, "image": "1241x376, 8 levels", "device_ms": ms, "call_ms": wall,
                                 "frames_per_s_device": nf / ms * 1e3, "cpu_oracle_ms_per_frame": tc / 4,
                                 "bit_exact_on_sample": bool(np.array_equal(api.stereo_matches(p4, impl="gpu", ctx=ctx)["uright"], o["uright"]))}

# ---- relocalisation variant of SearchByProjection
pr = synth.make_sbp_frame_batch(256, 2000, 777, th=10.0)
pr["mono"] = 1
pr["cur_uright"] = np.full_like(pr["cur_uright"], -1.0)
pr["last_has_obs"] = np.ones_like(pr["last_has_obs"])
pr["th_high"] = 100; pr["allow_negative_depth"] = 1
g, ms, wall = timed(lambda: api.sbp_frame(pr, impl="gpu", ctx=ctx))
out["sbp_relocalisation"] = {"pairs": 256, "keypoints": 2000, "device_ms": ms, "call_ms": wall, "queries_per_s_device": 256 * 2000 / ms * 1e3}

# ---- map points -> frame
pm = synth.make_sbp_mp_batch(256, 2000, 1500, 5)
g, ms, wall = timed(lambda: api.sbp_mappoints(pm, impl="gpu", ctx=ctx))
o, tc = cpu(lambda: api.sbp_mappoints(synth.make_sbp_mp_batch(8, 2000, 1500, 5), impl="oracle"))
out["sbp_mappoints"] = {"frames": 256, "keypoints": 2000, "map_points": 1500, "device_ms": ms, "call_ms": wall,
                        "queries_per_s_device": 256 * 1500 / ms * 1e3, "cpu_oracle_ms_per_frame": tc / 8}

# ---- keyframe searches (Fuse with the reprojection gate; SearchByProjection(KeyFrame*, Scw) with sequential claims)
for name, gate, claims in (("fuse_search", 1, 0), ("sim3_search_by_projection", 0, 1)):
    pk = synth.make_kf_search_batch(256, 2000, 1500, 9, chi2_gate=gate, sequential_claims=claims)
    g, ms, wall = timed(lambda: api.kf_search(pk, impl="gpu", ctx=ctx))
    o, tc = cpu(lambda: api.kf_search(synth.make_kf_search_batch(8, 2000, 1500, 9, chi2_gate=gate, sequential_claims=claims), impl="oracle"))
    out[name] = {"keyframes": 256, "keypoints": 2000, "map_points": 1500, "device_ms": ms, "call_ms": wall,
                 "queries_per_s_device": 256 * 1500 / ms * 1e3, "cpu_oracle_ms_per_keyframe": tc / 8}

# ---- SearchForTriangulation
pt = synth.make_tri_search_batch(256, 2000, 13)
g, ms, wall = timed(lambda: api.tri_search(pt, impl="gpu", ctx=ctx))
o, tc = cpu(lambda: api.tri_search(synth.make_tri_search_batch(8, 2000, 13), impl="oracle"))
out["search_for_triangulation"] = {"keyframe_pairs": 256, "keypoints": 2000, "vocabulary_nodes": 300, "device_ms": ms, "call_ms": wall,
                                   "pairs_per_s_device": 256 / ms * 1e3, "cpu_oracle_ms_per_pair": tc / 8}

# ---- SearchByBoW (keyframe -> frame)
pb = synth.make_bow_search_batch(256, 2000, 17, n_nodes=300)
g, ms, wall = timed(lambda: api.bow_search(pb, impl="gpu", ctx=ctx))
o, tc = cpu(lambda: api.bow_search(synth.make_bow_search_batch(8, 2000, 17, n_nodes=300), impl="oracle"))
out["search_by_bow"] = {"pairs": 256, "keypoints": 2000, "vocabulary_nodes": 300, "device_ms": ms, "call_ms": wall,
                        "pairs_per_s_device": 256 / ms * 1e3, "cpu_oracle_ms_per_pair": tc / 8}

# ---- temporal line association
pa = synth.make_line_assoc_batch(256, 300, 250, 64, 21, n_cand=40)
g, ms, wall = timed(lambda: api.line_associate(pa, impl="gpu", ctx=ctx))
pa8 = synth.make_line_assoc_batch(8, 300, 250, 64, 21, n_cand=40)
o, tc = cpu(lambda: api.line_associate(pa8, impl="oracle"))
out["line_association"] = {"frames": 256, "map_lines": 300, "current_lines": 250, "candidates": 40, "device_ms": ms, "call_ms": wall,
                           "frames_per_s_device": 256 / ms * 1e3, "cpu_oracle_ms_per_frame": tc / 8}

# ---- distinctive descriptors
import test_cpu_oracle as tco
off, desc = tco._medoid_landmarks(11, n_lm=100000, max_obs=40)
g, ms, wall = timed(lambda: api.medoid_orb(off, desc, impl="gpu", ctx=ctx))
off8, desc8 = tco._medoid_landmarks(11, n_lm=5000, max_obs=40)
o, tc = cpu(lambda: api.medoid_orb(off8, desc8, impl="oracle"))
out["medoid_orb"] = {"map_points": 100000, "observations": int(off[-1]), "device_ms": ms, "call_ms": wall,
                     "map_points_per_s_device": 100000 / ms * 1e3, "cpu_oracle_us_per_point": tc / 5000 * 1e3}
off, desc = tco._medoid_landmarks(12, n_lm=20000, max_obs=30, dim=64)
g, ms, wall = timed(lambda: api.medoid_float(off, desc, impl="gpu", ctx=ctx))
out["medoid_float"] = {"map_lines": 20000, "observations": int(off[-1]), "dim": 64, "device_ms": ms, "call_ms": wall,
                       "map_lines_per_s_device": 20000 / ms * 1e3}
print(json.dumps(out))
