"""N>1 path on CPU: two gloo ranks shard a batch of independent input sets with bench.shard_range,
evaluate their shard (here with the C oracle standing in for the device, as the checker) and only
exchange timing/bookkeeping, exactly like bench.py does (gloo): no collective touches witness data."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bench
from oracle import cref
from tests import util


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = util.golden_graph("circuit5_poseidon")
    cg = cref.CGraph(data)
    lo, hi = bench.shard_range(B, rank, world)
    inp = bench.synth_inputs("circuit5_poseidon", lo, hi, cg.n_inputs, {"a": (1, 1)}, seed=9)    # this rank's rows of the job's batch
    out = cg.evaluate_batch(inp)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                   # the only cross-rank traffic: max of the step time
    n = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(n)
    ret[rank] = (lo, hi, out.tobytes(), float(t.item()), int(n.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_covers_the_batch_once():
    B, world = 37, 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), B, ret), nprocs=world, join=True)
        ret = dict(ret)
    assert [ret[r][:2] for r in range(world)] == [(0, 18), (18, 37)]
    assert all(ret[r][3] == 2.0 and ret[r][4] == B for r in range(world))
    cg = cref.CGraph(util.golden_graph("circuit5_poseidon"))
    full = cg.evaluate_batch(bench.synth_inputs("circuit5_poseidon", 0, B, cg.n_inputs, {"a": (1, 1)}, seed=9))
    assert ret[0][2] + ret[1][2] == full.tobytes()


def test_rows_do_not_depend_on_the_slice_and_launch_plans_cover_the_shard():
    a = bench.synth_inputs("circuit5_poseidon", 0, 20000, 2, {"a": (1, 1)}, seed=9)
    b = bench.synth_inputs("circuit5_poseidon", 8000, 17000, 2, {"a": (1, 1)}, seed=9)
    assert (a[8000:17000] == b).all() and not (a[0] == a[1]).all()
    for n, mx in ((262144, 75776), (32768, 75776), (100, 75776), (65536, 4736), (0, 4736)):
        plan = bench.launch_plan(n, mx, 148)
        assert sum(hi - lo for lo, hi in plan) == n and all(0 < hi - lo <= max(mx, 148 * 32) for lo, hi in plan)
        assert all(plan[k][1] == plan[k + 1][0] for k in range(len(plan) - 1))
        assert all((hi - lo) % (148 * 32) == 0 for lo, hi in plan[:-1])
    assert bench.launch_plan(262144, 80000, 148) == [(0, 75776), (75776, 151552), (151552, 227328), (227328, 262144)]


def test_shard_range_properties():
    for n in (0, 1, 7, 262144):
        for world in (1, 2, 4, 8):
            r = [bench.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
