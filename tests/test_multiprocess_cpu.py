"""World-size-2 test of the multi-GPU path's host logic on CPU (gloo): lane sharding + final gather.

Each rank integrates its shard (with the CPU oracle standing in for the GPU: no CUDA here), the final states are
all_gather'ed, and the result must be BIT-IDENTICAL to the single-process run of the whole batch — the property
test/ensemble_propagate.cpp:413-431 pins for the reference's ensemble mode ("partitioning does not change results").
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_lanes, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import heyoka_b200 as hb
    import oracle
    from common import outer_ss_batch_state, sys_outer_ss
    from heyoka_b200.ensemble import gather_lanes, lane_shard

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = hb.Program(sys_outer_ss(), high_accuracy=True)
    st = outer_ss_batch_state(n_lanes)
    b, e = lane_shard(n_lanes, rank, world)
    o = oracle.OracleIntegrator(P, st[:, b:e], e - b, mode=oracle.FMA)
    o.propagate_until(3.0)
    full = gather_lanes(torch.from_numpy(o.state), n_lanes)
    steps = gather_lanes(torch.from_numpy(o.n_steps.astype(np.int64)), n_lanes)
    np.save(os.path.join(out_dir, "state_%d.npy" % rank), full.numpy())
    np.save(os.path.join(out_dir, "steps_%d.npy" % rank), steps.numpy())
    dist.destroy_process_group()


def test_lane_shard_partition():
    from heyoka_b200.ensemble import all_shards, lane_shard
    for n in (1, 7, 8, 13, 1 << 20):
        for w in (1, 2, 3, 8):
            sh = all_shards(n, w)
            assert sh[0][0] == 0 and sh[-1][1] == n
            assert all(sh[i][1] == sh[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in sh) - min(e - b for b, e in sh) <= 1
            assert lane_shard(n, w - 1, w) == sh[-1]


def test_two_rank_sharded_run_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import heyoka_b200 as hb
    import oracle
    from common import outer_ss_batch_state, sys_outer_ss

    n_lanes, world = 13, 2  # odd: the shards differ by one lane
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_lanes, str(tmp_path)), nprocs=world, join=True)

    P = hb.Program(sys_outer_ss(), high_accuracy=True)
    o = oracle.OracleIntegrator(P, outer_ss_batch_state(n_lanes), n_lanes, mode=oracle.FMA)
    o.propagate_until(3.0)
    for r in range(world):
        st = np.load(os.path.join(str(tmp_path), "state_%d.npy" % r))
        ns = np.load(os.path.join(str(tmp_path), "steps_%d.npy" % r))
        assert np.array_equal(st, o.state)          # bit-identical, on every rank
        assert np.array_equal(ns, o.n_steps.astype(np.int64))
