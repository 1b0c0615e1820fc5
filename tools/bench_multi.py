"""Multi-GPU measurements through the product's own sharded batch (hy_batch_create_multi(): one process, one shard and
one host thread per GPU) — BASELINE.json configs[2] (model::nbody N = 32, 65,536 randomised ICs over the GPUs of one
box) and the strong-scaling line of the headline workload (outer Solar System, 2^20 lanes IN TOTAL over the GPUs).
One JSON line per measurement; host buffers in and out are part of every timed call (upload, propagate_until on all
shards concurrently with the reference's global exits across shards, download of state / times / results).

    python tools/bench_multi.py [n32] [s6strong] [s6weak]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import heyoka_b200 as hb  # noqa: E402
from common import nbody32_batch_state, outer_ss_batch_state, sys_nbody32, sys_outer_ss  # noqa: E402


def run(name, P, st, t_final, devices, reps=3):
    batch = st.shape[1]
    b = hb.Batch(P, batch, devices)
    z = np.zeros(batch)
    tf = np.full(batch, t_final)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        b.upload(st, None, z, z)
        b.propagate_until(tf)
        state, t_hi, t_lo, last_h = b.download()
        oc, mn, mx, ns = b.prop_res()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    assert np.all(t_hi == t_final) and np.all(np.isfinite(state))
    n_steps = int(ns.sum())
    print(json.dumps({"config": name, "n_gpus": max(b.n_shards, 1), "lanes_total": batch, "kernel": b.kernel_info()["tape"],
                      "seconds_e2e": best, "lane_steps": n_steps, "lane_steps_per_s_e2e": n_steps / best,
                      "h2d_bytes": int(st.nbytes + 3 * 8 * batch), "d2h_bytes": int(st.nbytes + 3 * 8 * batch + 32 * batch)}),
          flush=True)
    return state


def main():
    which = sys.argv[1:] or ["n32", "s6strong"]
    n_dev = hb.lib.hy_device_count()
    devs = list(range(n_dev))
    if "n32" in which:
        st = nbody32_batch_state(65536)
        many = run("model::nbody N=32, 65,536 ICs sharded over %d GPU(s), propagate_until(1)" % n_dev,
                   hb.Program(sys_nbody32()), st, 1.0, devs if n_dev > 1 else -1)
        if "check" in which:
            one = run("idem, first 8192 lanes on one GPU", hb.Program(sys_nbody32()), st[:, :8192].copy(), 1.0, -1, reps=1)
            assert np.array_equal(one, many[:, :8192])
    if "s6strong" in which:
        P = hb.Program(sys_outer_ss(), high_accuracy=True)
        st = outer_ss_batch_state(1 << 20)
        run("outer_ss 6-body, 2^20 lanes IN TOTAL over %d GPU(s) (strong scaling), propagate_until(20 yr)" % n_dev, P, st,
            20.0, devs if n_dev > 1 else -1)
    if "s6weak" in which:
        P = hb.Program(sys_outer_ss(), high_accuracy=True)
        st = outer_ss_batch_state((1 << 20) * max(n_dev, 1) // 2)
        run("outer_ss 6-body, 2^19 lanes per GPU over %d GPU(s), propagate_until(20 yr)" % n_dev, P, st, 20.0,
            devs if n_dev > 1 else -1)


if __name__ == "__main__":
    main()
