"""Dry run of GPU test bodies without a GPU: heyoka_b200.taylor_adaptive_batch is replaced by the oracle-backed front end
(tests/oracle.py OracleEventIntegrator: the product's Python front end over the CPU oracle) with a dummy non-terminal
event that never triggers and does not change the step size (value 0.5, derivatives ~1e-300), so that the front end
takes its host loops. Catches Python-level mistakes and wrong expectations in tests that could not be run on hardware.

    python tools/dry_run_late_gpu_tests.py test_tutorial_adaptive_gpu test_more_tutorials_gpu

Not everything can be dry-run: continuous output and the sharded batches need the device library, and tests that
expect an integrator WITHOUT events cannot take the dummy event.
"""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import heyoka_b200 as hb  # noqa: E402
import oracle  # noqa: E402
import test_zz_gpu_late_additions as late  # noqa: E402

real = hb.taylor_adaptive_batch


def fake(sys_, st, n, **kw):
    kw.pop("device", None)
    x = hb.make_vars("x")[0]
    kw.setdefault("nt_events", []).append(hb.nt_event_batch(0.5 + 1e-300 * x, lambda ta, t, d, i: None))
    hb.taylor_adaptive_batch = real
    try:
        return oracle.OracleEventIntegrator(sys_, st, n, **kw)
    finally:
        hb.taylor_adaptive_batch = fake


if __name__ == "__main__":
    hb.taylor_adaptive_batch = fake
    bad = 0
    for name in sys.argv[1:] or ["test_tutorial_adaptive_gpu", "test_more_tutorials_gpu", "test_tutorial_events_golden_gpu",
                                 "test_step_callback_must_not_alter_the_time_gpu"]:
        try:
            getattr(late, name)()
            print(name, "OK")
        except Exception:  # noqa: BLE001
            traceback.print_exc(limit=4)
            print(name, "FAILED")
            bad += 1
    sys.exit(1 if bad else 0)
