"""BASELINE.json configs[4]: localBA 50 KF x 20k pts x 150k obs, landmarks split over the ranks, one
NCCL all-reduce of the reduced camera system per LM iteration.

    torchrun --nproc-per-node N scripts/ba_sharded_nccl.py        (N = 1: unsharded reference timing)

Rank 0 checks the sharded result against the unsharded solve and prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ov2slam_b200 import api, synth  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ncam, npts, nobs = 50, 20000, 150000
small = os.environ.get("OV2_BA_SMALL")
if small:
    ncam, npts, nobs = 20, 3000, 18000
pb = synth.make_ba_problem(5, ncam, npts, nobs)
ctx = api.Context(local)
clone = lambda d: {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in d.items()}
reps = 5
if world == 1:
    opt = api.Optimizer(ctx)
    opt.local_ba(clone(pb))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        res, _ = opt.local_ba(clone(pb))
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"config": f"C5 localBA {ncam} KF x {npts} pts x {nobs} obs", "n_gpus": 1, "ms_per_solve": 1e3 * dt,
                      "solves_per_s": 1 / dt, "lm_iterations": res["iters_robust"] + res["iters_refine"], "final_cost": res["final_cost"]}))
else:
    shards = api.partition_ba_problem(pb, world)
    cb = api.make_torch_allreduce(dist, torch)
    mine = clone(shards[rank][0])
    api.local_ba_sharded(ctx, clone(shards[rank][0]), cb, rank)      # warm-up
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        mine = clone(shards[rank][0])
        res, flags = api.local_ba_sharded(ctx, mine, cb, rank)
    torch.cuda.synchronize(); dist.barrier(); dt = (time.perf_counter() - t0) / reps
    if rank == 0:
        ref = clone(pb)
        rres, _ = api.Optimizer(ctx).local_ba(ref)
        pose_err = float(np.abs(mine["pose"] - ref["pose"]).max())
        invd_err = float(np.abs(mine["lm_invdepth"] - ref["lm_invdepth"][shards[0][1]]).max())
        print(json.dumps({"config": f"C5 localBA {ncam} KF x {npts} pts x {nobs} obs, landmarks split {world}-way, NCCL allreduce",
                          "n_gpus": world, "ms_per_solve": 1e3 * dt, "solves_per_s": 1 / dt,
                          "lm_iterations": res["iters_robust"] + res["iters_refine"],
                          "same_iterations_as_unsharded": (res["iters_robust"], res["iters_refine"]) == (rres["iters_robust"], rres["iters_refine"]),
                          "max_pose_diff_vs_unsharded": pose_err, "max_invdepth_diff_vs_unsharded": invd_err,
                          "final_cost": res["final_cost"], "unsharded_final_cost": rres["final_cost"]}))
    dist.barrier()
    dist.destroy_process_group()
