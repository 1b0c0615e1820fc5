"""Extract the reference's printed known answers into JSON fixtures.

Run HERE (the container that has /root/reference); the GPU box only sees the committed JSON.

Sources (bluescarni/heyoka @ 9c91f71):
  doc/tut_batch_mode.rst:160-340   output of tutorial/batch_mode.cpp:52-131 (batch 4, order 20,
                                   x' = v, v' = cos(t) - alpha v - sin(x), alpha = par[0])
  README.md:118-139                scalar pendulum x(10), v(10)
"""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

FLOAT = r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?"


def braces_to_list(txt):
    txt = txt.replace("{", "[").replace("}", "]")
    # "20." / ".5" are not valid JSON numbers.
    txt = re.sub(r"(\d)\.(?=[^\d])", r"\1.0", txt)
    return json.loads(txt)


def main():
    rst = open(os.path.join(REF, "doc", "tut_batch_mode.rst")).read()

    # Step results: "Batch index i: (taylor_outcome::xxx, h)".
    step_blocks = re.findall(r"((?:   Batch index \d: \(taylor_outcome::\w+, " + FLOAT + r"\)\n)+)", rst)
    steps = []
    for blk in step_blocks:
        rows = re.findall(r"taylor_outcome::(\w+), (" + FLOAT + r")\)", blk)
        steps.append([{"outcome": oc, "h": float(h)} for oc, h in rows])

    prop_blocks = re.findall(
        r"((?:   Batch index \d: \(taylor_outcome::\w+, " + FLOAT + ", " + FLOAT + r", \d+\)\n)+)", rst)
    props = []
    for blk in prop_blocks:
        rows = re.findall(r"taylor_outcome::(\w+), (" + FLOAT + "), (" + FLOAT + r"), (\d+)\)", blk)
        props.append([{"outcome": oc, "min_h": float(a), "max_h": float(b), "n_steps": int(n)} for oc, a, b, n in rows])

    states = [braces_to_list(m) for m in re.findall(r"State array:\n((?:   .*\n)+?)\n", rst)]
    states = [s for s in states if len(s) == 2 and len(s[0]) == 4][-4:]
    times = [braces_to_list(m) for m in re.findall(r"Time array:\n   (\{.*\})\n", rst)]

    tc_txt = re.search(r"Array of Taylor coefficients:\n((?:   .*\n)+)", rst).group(1)
    tc = braces_to_list(tc_txt)

    out = {
        "source": "doc/tut_batch_mode.rst (output of tutorial/batch_mode.cpp)",
        "system": "x' = v, v' = cos(t) - par[0]*v - sin(x)",
        "batch_size": 4,
        "x0": [0.01, 0.02, 0.03, 0.04],
        "v0": [1.85, 1.86, 1.87, 1.88],
        "alpha": [0.10, 0.11, 0.12, 0.13],
        "first_step": steps[0],
        "clamped_step_limits": [0.010, 0.011, 0.012, 0.013],
        "clamped_step": steps[1],
        "propagate_for": {"delta_ts": [10., 11., 12., 13.], "res": props[0]},
        "propagate_until": {"ts": [20., 21., 22., 23.], "res": props[1]},
        # states/times printed after: first step, clamped step, propagate_for, propagate_until.
        # (the very first "State array" of the tutorial is the initial condition and lives in tutorial/, not here)
        "states": states,
        "times": times,
        "tc_after_final_step": tc,
    }
    assert len(out["first_step"]) == 4 and len(out["clamped_step"]) == 4
    assert len(states) == 4 and len(times) == 4, (len(states), len(times))
    assert len(tc) == 2 and len(tc[0]) == 21 and len(tc[0][0]) == 4
    with open(os.path.join(HERE, "tut_batch_mode.json"), "w") as f:
        json.dump(out, f, indent=1)

    readme = open(os.path.join(REF, "README.md")).read()
    m = re.search(r"x\(10\) = (" + FLOAT + r")\s*\n\s*y?v?\(10\) = (" + FLOAT + ")", readme)
    if m is None:
        m = re.search(r"(" + FLOAT + r")[^\d]+?(" + FLOAT + ")", readme[readme.index("x(10)"):])
    with open(os.path.join(HERE, "readme_pendulum.json"), "w") as f:
        json.dump({"source": "README.md:118-139", "system": "x' = v, v' = -9.8 sin(x)", "x0": 0.05, "v0": 0.025,
                   "t": 10.0, "x": float(m.group(1)), "v": float(m.group(2))}, f, indent=1)


if __name__ == "__main__":
    main()
