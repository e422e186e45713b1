"""N>1 host-side logic on CPU: two processes over the gloo backend (no GPU).

What is checked is what the multi-GPU global BA relies on: the library's landmark partition covers every
landmark exactly once, keyframes are replicated, and quantities the device path all-reduces are additive over the
partition — the initial robust chi2 (sum over edges, sparse_optimizer.cpp:100-114) of the per-rank shards, summed with
`all_reduce`, equals the chi2 of the whole problem.  The CPU oracle is the checker here, never the product.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lld_slam_b200 import api, shard, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = synth.make_global_ba(12, 600, 120, 41, robust_points=True)
        s = shard.landmark_shard(p, rank, world)
        # coverage bookkeeping: owned landmark / edge counts add up to the whole problem
        cnt = torch.tensor([int(s["pt_off"][-1]), int(s["ln_off"][-1]), int(s["pt_obs_off"][-1]), int(s["ln_obs_off"][-1])],
                           dtype=torch.int64)
        dist.all_reduce(cnt)
        # additive part of the reduced system: initial robust chi2 of the rank's own landmarks (all keyframes replicated)
        o = api.ba_global(s, 1, impl="oracle")
        chi = torch.tensor([o["chi2_log"][0, 0]], dtype=torch.float64)
        dist.all_reduce(chi)
        # independent units: window ranges partition [0, n)
        lo, hi = shard.unit_range(513, rank, world)
        rng = torch.tensor([hi - lo], dtype=torch.int64)
        dist.all_reduce(rng)
        if rank == 0:
            full = api.ba_global(p, 1, impl="oracle")
            q.put(dict(cnt=cnt.tolist(), chi=float(chi[0]), chi_full=float(full["chi2_log"][0, 0]), units=int(rng[0]),
                       want=[int(p["pt_off"][-1]), int(p["ln_off"][-1]), int(p["pt_obs_off"][-1]), int(p["ln_obs_off"][-1])]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_landmark_partition_and_additive_chi2_world2(built):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = q.get(timeout=240)
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    assert res["cnt"] == res["want"]
    assert res["units"] == 513
    assert abs(res["chi"] - res["chi_full"]) <= 1e-12 * abs(res["chi_full"])


def test_shard_bounds_cover_without_overlap(built):
    for world in (1, 2, 3, 8):
        pts, lns = [], []
        for r in range(world):
            plo, phi, llo, lhi = shard.landmark_bounds(1001, 77, r, world)
            pts += list(range(plo, phi))
            lns += list(range(llo, lhi))
        assert pts == list(range(1001)) and lns == list(range(77))


def test_landmark_shard_views_are_consistent(built):
    p = synth.make_global_ba(10, 300, 50, 3)
    tot_e = 0
    for r in range(3):
        s = shard.landmark_shard(p, r, 3)
        assert s["kf_Tcw"] is p["kf_Tcw"]  # keyframes replicated, not copied
        assert s["pt_obs_off"][0] == 0 and s["pt_obs_off"][-1] == len(s["pt_obs_kf"])
        assert s["ln_obs_off"][0] == 0 and s["ln_obs_off"][-1] == len(s["ln_obs_kf"])
        tot_e += len(s["pt_obs_kf"])
    assert tot_e == int(p["pt_obs_off"][-1])
