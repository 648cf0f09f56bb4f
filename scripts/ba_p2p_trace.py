"""Per-phase device time of the sharded (peer-memory) local-BA solve: run under torchrun, rank 0's trace goes to stderr.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/ba_p2p_trace.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
if rank == 0:
    os.environ["OV2_BA_TRACE"] = "1"
from ov2slam_b200 import api, synth  # noqa: E402

torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = api.Context(local)
pb = synth.make_ba_problem(5, 50, 20000, 150000)
clone = lambda d: {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in d.items()}
shards = api.partition_ba_problem(pb, world)
solver = api.ShardedOptimizer(ctx, dist, torch, rank, world)
for barrier in (True, False):
    ts = []
    for it in range(6):
        mine = clone(shards[rank][0])
        if barrier:
            dist.barrier()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        res, _ = solver.local_ba(mine)
        ts.append(time.perf_counter() - t0)
    if rank == 0:
        print(f"world {world} barrier-before-solve {barrier}: ms per solve {[round(1e3 * t, 3) for t in ts]}", file=sys.stderr, flush=True)
dist.barrier()
