#!/usr/bin/env python
"""Global BA on 1..N GPUs (one process per GPU, launched with torchrun): landmarks sharded, reduced camera system
all-reduced over NCCL inside liblldba.  Prints one JSON line on rank 0.  --check compares with the CPU oracle."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lld_slam_b200 import api, capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kf", type=int, default=1500)
ap.add_argument("--pts", type=int, default=300000)
ap.add_argument("--lines", type=int, default=60000)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--check", action="store_true")
ap.add_argument("--robust", action="store_true")
ap.add_argument("--profile", action="store_true")
ap.add_argument("--mono", action="store_true", help="all point observations monocular: no float invz in the residuals (conditioning experiments)")
a = ap.parse_args()

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
lib = capi.load_library()
ctx = capi.Context(lr)
if world > 1:
    import torch
    idb = np.zeros(128, np.uint8)
    if rank == 0:
        assert lib.dll.lld_comm_unique_id(idb.ctypes.data_as(capi.c_u8p)) == 0
    t = torch.from_numpy(idb).cuda()
    dist.broadcast(t, 0)
    idb = t.cpu().numpy()
    ctx.check(lib.dll.lld_comm_init(ctx.handle, world, rank, idb.ctypes.data_as(capi.c_u8p)), "comm_init")
p = synth.make_global_ba(a.kf, a.pts, a.lines, synth.seed_for(5), robust_points=a.robust)
if a.mono:
    p["pt_obs_uvr"] = p["pt_obs_uvr"].copy()
    p["pt_obs_uvr"][:, 2] = -1.0
n_pe = int(p["pt_obs_off"][-1]); n_lc = int(p["ln_obs_off"][-1])
g = None
times = []
for rep in range(a.reps):
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    g = api.ba_global(p, a.iters, impl="gpu", ctx=ctx)
    if dist is not None:
        dist.barrier()
    times.append(time.perf_counter() - t0)
h2d, comp, d2h = ctx.last_timing()
prof = None
if a.profile:
    import bench
    d = bench.bind_resident(lib)
    d.lld_ctx_profile(ctx.handle, 1)
    api.ba_global(p, a.iters, impl="gpu", ctx=ctx)
    prof = bench.profile_report(d, ctx)
    d.lld_ctx_profile(ctx.handle, 0)
res = {"mono": bool(a.mono), "n_gpus": world, "kf": a.kf, "pts": a.pts, "lines": a.lines, "point_edges": n_pe, "line_cells": n_lc,
       "iters_done": int(g["n_iter_done"][0, 0]), "trials": int(g["trials_log"].sum()),
       "wall_s_best": min(times), "device_compute_ms": comp, "chi2_first": float(g["chi2_log"][0, 0]),
       "chi2_last": float(g["chi2_log"][0, int(g["n_iter_done"][0, 0])])}
if dist is not None:
    import torch
    tc = torch.tensor([comp], device=f"cuda:{lr}", dtype=torch.float64)
    dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    res["device_compute_ms"] = float(tc[0])
    # landmark results live on their owner rank only (zeros elsewhere): sum = gather
    for k in ("pt_xyz", "ln_x0_dir"):
        t = torch.from_numpy(g[k]).cuda()
        dist.all_reduce(t)
        g[k] = t.cpu().numpy()
res["lm_iters_per_sec"] = res["iters_done"] / (res["device_compute_ms"] * 1e-3)
if a.check and rank == 0:
    o = api.ba_global(p, a.iters, impl="oracle")
    rel = np.abs(g["chi2_log"] - o["chi2_log"]) / np.maximum(np.abs(o["chi2_log"]), 1e-9)
    res["check"] = {"iters_equal": bool(np.array_equal(g["n_iter_done"], o["n_iter_done"])),
                    "trials_equal": bool(np.array_equal(g["trials_log"], o["trials_log"])),
                    "chi2_rel_max": float(rel.max()),
                    "pose_t_max": float(np.abs(g["kf_Tcw"][:, 9:] - o["kf_Tcw"][:, 9:]).max()),
                    "pose_R_max": float(np.abs(g["kf_Tcw"][:, :9] - o["kf_Tcw"][:, :9]).max()),
                    "pt_max": float(np.abs(g["pt_xyz"] - o["pt_xyz"]).max())}
    dT = np.abs(g["kf_Tcw"][:, 9:] - o["kf_Tcw"][:, 9:]).max(axis=1)
    res["check"]["pose_t_deciles"] = [float(x) for x in np.quantile(dT, np.linspace(0, 1, 11))]
    res["check"]["pose_t_worst_kf"] = [int(k) for k in np.argsort(-dT)[:8]]
    os.makedirs("gpurun_out", exist_ok=True)
    if world == 1:   # GPU against itself with the landmark order permuted: the amplitude of the weak mode under pure reordering
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from gba_spread import permute_landmarks
        q, op, ol = permute_landmarks(p, 7)
        g2 = api.ba_global(q, a.iters, impl="gpu", ctx=ctx)
        dS = np.abs(g["kf_Tcw"][:, 9:] - g2["kf_Tcw"][:, 9:]).max(axis=1)
        res["check"]["gpu_self_spread_deciles"] = [float(x) for x in np.quantile(dS, np.linspace(0, 1, 11))]
        res["check"]["ratio_to_gpu_self_spread_deciles"] = [float(x) for x in np.quantile(dT / np.maximum(dS, 1e-12), np.linspace(0, 1, 11))]
    np.savez_compressed(f"gpurun_out/gba_state_n{world}{'_mono' if a.mono else ''}.npz", g_kf=g["kf_Tcw"], o_kf=o["kf_Tcw"], g_chi2=g["chi2_log"], o_chi2=o["chi2_log"],
                        g_lambda=g["lambda_log"], o_lambda=o["lambda_log"])
if prof is not None:
    res["kernels_us_per_launch"] = {k: round(1e3 * v["ms"] / max(v["n"], 1), 1) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
if rank == 0:
    print(json.dumps(res))
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
ctx.close()
