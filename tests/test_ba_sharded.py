"""Multi-GPU localBA: landmark partitioning + one sum-allreduce of the reduced
camera system per LM iteration.

CPU: the partition is a partition, and the allreduce callback sums across a 2-process gloo group.
GPU (one device): two shards run concurrently on two contexts with a loop-back "fake NCCL"
allreduce (thread barrier + sum) and must reproduce the unsharded solve."""
import os
import threading

import numpy as np
import pytest

from ov2slam_b200 import api, synth


def test_partition_is_a_balanced_partition():
    pb = synth.make_ba_problem(7, 8, 500, 3000)
    for world in (2, 3, 8):
        shards = api.partition_ba_problem(pb, world)
        all_l = np.concatenate([s[1] for s in shards])
        all_o = np.concatenate([s[2] for s in shards])
        assert sorted(all_l.tolist()) == list(range(500)) and sorted(all_o.tolist()) == list(range(3000))
        loads = [len(s[2]) for s in shards]
        assert max(loads) - min(loads) <= 0.05 * 3000 / world + 8
        for sh, lms, obs in shards:
            assert (np.diff(sh["obs_lm"]) >= 0).all()                      # still CSR by landmark
            assert np.array_equal(lms[sh["obs_lm"]], pb["obs_lm"][obs])     # remap is consistent
            assert np.array_equal(sh["obs_px"], pb["obs_px"][obs])
            assert np.array_equal(sh["pose"], pb["pose"])


def _gloo_worker(rank, world, port, q):
    import ctypes as C
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ov2slam_b200 import api as A
    A.load()
    cb = A.make_torch_allreduce(dist, torch)
    buf = np.arange(10, dtype=np.float64) * (rank + 1)
    rc = cb(None, buf.ctypes.data, 10, None)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, rc, buf.tolist()))


def test_allreduce_callback_sums_over_gloo_group():
    import multiprocessing as mp
    from ov2slam_b200 import build
    build.build()                                           # before the workers start: on a fresh checkout both would compile it at once
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctxmp.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    for rank, rc, buf in out:
        assert rc == 0
        assert buf == (np.arange(10) * 3.0).tolist()       # 1x + 2x summed on both ranks


@pytest.mark.gpu
@pytest.mark.parametrize("world,empty", [(2, False), (3, False), (3, True)])
def test_sharded_solve_matches_unsharded(ctx, world, empty):
    """The round-1 callback path (ov2_localba_solve_sharded).  With `empty` the last rank owns no landmark and no observation:
    it must still take part in every all-reduce (zero-block launches made it return early and hang the others)."""
    import ctypes as C
    import torch
    pb = synth.make_ba_problem(31, 12, 1500, 9000)
    ref = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in pb.items()}
    rres, rflags = api.Optimizer(ctx).local_ba(ref)
    shards = api.partition_ba_problem(pb, world - 1 if empty else world)
    if empty:
        e = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in shards[0][0].items()}
        for k in ("lm_anchor_cam", "lm_anchor_px", "lm_invdepth", "obs_cam", "obs_lm", "obs_px"):
            e[k] = e[k][:0].copy()
        shards.append((e, np.zeros(0, np.int64), np.zeros(0, np.int64)))
    ctxs = [api.Context(0) for _ in range(world)]
    barrier = threading.Barrier(world)
    slots = [None] * world
    total = [None]

    def make_cb(rank):
        def fn(user, buf, count, stream):
            t = torch.as_tensor(api._CudaView(buf, count), device="cuda")
            torch.cuda.synchronize()
            slots[rank] = t
            barrier.wait()
            if rank == 0:
                total[0] = torch.stack([s.clone() for s in slots]).sum(0)
            barrier.wait()
            t.copy_(total[0])
            torch.cuda.synchronize()
            barrier.wait()
            return 0
        return api.ALLREDUCE_FN(fn)

    results = [None] * world
    errs = []

    def run(rank):
        try:
            results[rank] = api.local_ba_sharded(ctxs[rank], shards[rank][0], make_cb(rank), rank)
        except Exception as e:  # pragma: no cover
            errs.append(e)
            barrier.abort()

    ths = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=120)
    assert not errs, errs
    invd = np.zeros(len(pb["lm_invdepth"]))
    flags = np.zeros(len(pb["obs_cam"]), np.uint8)
    for r, (sh, lms, obs) in enumerate(shards):
        res, fl = results[r]
        assert res["iters_robust"] == rres["iters_robust"] and res["iters_refine"] == rres["iters_refine"]
        assert abs(res["final_cost"] - rres["final_cost"]) <= 1e-9 * max(1.0, rres["final_cost"])
        assert np.abs(sh["pose"] - ref["pose"]).max() <= 1e-7
        invd[lms] = sh["lm_invdepth"]
        flags[obs] = fl
    assert np.abs(invd - ref["lm_invdepth"]).max() <= 1e-7
    assert (flags != rflags).sum() <= 2
    for c in ctxs:
        c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world,empty", [(2, False), (3, False), (3, True)])
def test_p2p_sharded_solve_matches_unsharded(ctx, monkeypatch, world, empty):
    """(All ranks share ONE GPU here, so their persistent kernels must be co-resident: 24 CTAs each.)
    ov2_localba_solve_p2p: `world` ranks as contexts of ONE process on one GPU (communicators connected with direct
    pointers); every rank's persistent kernel sums the partial reduced systems out of the other ranks' exchange buffers.
    Must reproduce the unsharded solve; with `empty` the last rank owns no landmark at all (an empty shard
    used to fail with a zero-block launch and hang the others)."""
    monkeypatch.setenv("OV2_BA_CTAS", "24")
    pb = synth.make_ba_problem(31, 12, 1500, 9000)
    ref = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in pb.items()}
    rres, rflags = api.Optimizer(ctx).local_ba(ref)
    shards = api.partition_ba_problem(pb, world - 1 if empty else world)
    if empty:
        e = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in shards[0][0].items()}
        for k in ("lm_anchor_cam", "lm_anchor_px", "lm_invdepth", "obs_cam", "obs_lm", "obs_px"):
            e[k] = e[k][:0].copy()
        shards.append((e, np.zeros(0, np.int64), np.zeros(0, np.int64)))
    ctxs = [api.Context(0) for _ in range(world)]
    for c in ctxs:
        # every context allocates its arenas / pinned staging on its first solve; do that BEFORE the concurrent solves: a
        # cudaMalloc / cudaHostAlloc issued while another rank's kernel spins in a rank barrier on the SAME device can block
        # behind that kernel (ranks on different devices, the real configuration, do not share this hazard)
        api.Optimizer(c).local_ba({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in pb.items()})
    comms = [api.BaComm(ctxs[r], r, world) for r in range(world)]
    api.BaComm.connect_local(comms)
    for rep in range(2):                       # twice: epochs of consecutive launches must not collide
        work = [{k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in sh[0].items()} for sh in shards]
        results = [None] * world
        errs = []

        def run(rank):
            try:
                results[rank] = comms[rank].local_ba(work[rank])
            except Exception as ex:  # pragma: no cover
                errs.append((rank, ex))

        ths = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for t in ths:
            t.start()
        for t in ths:
            t.join(timeout=180)
        assert not errs, errs
        invd = np.zeros(len(pb["lm_invdepth"]))
        flags = np.zeros(len(pb["obs_cam"]), np.uint8)
        for r, (sh, lms, obs) in enumerate(shards):
            res, fl = results[r]
            assert (res["iters_robust"], res["iters_refine"]) == (rres["iters_robust"], rres["iters_refine"]), (r, res)
            assert abs(res["final_cost"] - rres["final_cost"]) <= 1e-9 * max(1.0, rres["final_cost"])
            assert np.abs(work[r]["pose"] - ref["pose"]).max() <= 1e-7
            assert np.array_equal(work[r]["pose"], work[0]["pose"])          # bit-identical decisions on every rank
            invd[lms] = work[r]["lm_invdepth"]
            flags[obs] = fl
        assert np.abs(invd - ref["lm_invdepth"]).max() <= 1e-7
        assert (flags != rflags).sum() <= 2
    for c in comms:
        c.close()
    for c in ctxs:
        c.close()
