"""Extract the reference's printed known answers into JSON fixtures.

Run HERE (the container that has /root/reference); the GPU box only sees the committed JSON.

Sources (bluescarni/heyoka @ 9c91f71):
  doc/tut_batch_mode.rst:160-340   output of tutorial/batch_mode.cpp:52-131 (batch 4, order 20,
                                   x' = v, v' = cos(t) - alpha v - sin(x), alpha = par[0])
  README.md:118-139                scalar pendulum x(10), v(10)
  doc/tut_adaptive.rst:94-325      output of tutorial/adaptive_basic.cpp (scalar pendulum: single steps forwards /
                                   backwards / clamped, propagate_for(5), propagate_until(20), back to 0, propagate_grid)
  doc/tut_d_output.rst:64-203      output of tutorial/d_output.cpp (dense output after one step, continuous output of
                                   propagate_until(10): 48 steps, samples)
  doc/tut_events.rst:138-400       output of tutorial/event_basic.cpp (non-terminal events: zero-velocity times of the
                                   pendulum to 16 digits, direction filter, two events in chronological order; terminal
                                   event toggling a damping parameter: propagate_grid output to 16 digits)
  doc/tut_ensemble.rst:118-139     output of tutorial/ensemble.cpp (member 9 of ensemble_propagate_until(20): final state
                                   to 17 digits, 124 steps, min / max step)
  doc/tut_param.rst:60-128         output of tutorial/pendulum_param.cpp (runtime parameters: one period of the pendulum
                                   for g = 9.8 and g = 3.72)
  doc/tut_nonauto.rst:62-88        output of tutorial/forced_damped_pendulum.cpp (time-dependent right-hand side: x after
                                   every 2 time units, 25 values)
  doc/tut_adaptive_custom.rst:38-63 output of tutorial/adaptive_opt.cpp (tol = 1e-9: order 12, state after 0 -> 10 -> 0)
The scalar integrators of those tutorials are batch integrators of size 1 here.
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

    tutorials()


def console_blocks(rst):
    """The `.. code-block:: console` blocks of a page, in order, as lists of stripped lines."""
    out = []
    for m in re.finditer(r"\.\. code-block:: console\n\n((?:(?:   .*)?\n)+)", rst):
        out.append([ln.strip() for ln in m.group(1).splitlines() if ln.strip()])
    return out


def kv(lines):
    """'Key : value' lines -> dict of strings."""
    d = {}
    for ln in lines:
        if ":" in ln:
            k, v = ln.split(":", 1)
            d.setdefault(k.strip(), []).append(v.strip())
    return d


def tutorials():
    # ---- tut_adaptive.rst ----
    blocks = console_blocks(open(os.path.join(REF, "doc", "tut_adaptive.rst")).read())
    b = [kv(x) for x in blocks]
    step1 = next(x for x in b if x.get("Timestep") == ["0.216053"])
    back = next(x for x in b if x.get("Timestep") == ["-0.213123"])
    clamped = next(x for x in b if x.get("Timestep") == ["0.01", "-0.02"])
    props = [x for x in b if "Num. of steps" in x]
    grid_lines = next(bl for bl in blocks if bl and bl[0].startswith("x(0.4)"))
    out = {
        "source": "doc/tut_adaptive.rst (output of tutorial/adaptive_basic.cpp)",
        "system": "x' = v, v' = -9.8 sin(x)", "x0": 0.05, "v0": 0.025, "order": int(step1["Taylor order"][0]),
        "first_step": {"outcome": step1["Outcome"][0].split("::")[1], "h": float(step1["Timestep"][0]),
                       "time": float(step1["Time"][0]), "state": json.loads(step1["State"][0])},
        "step_backward": {"outcome": back["Outcome"][0].split("::")[1], "h": float(back["Timestep"][0])},
        "clamped_steps": [{"limit": float(h), "outcome": oc.split("::")[1], "h": float(h)}
                          for oc, h in zip(clamped["Outcome"], clamped["Timestep"])],
        # propagate_for(5) and propagate_until(20) share a block; propagate_until(0) follows with the final state.
        "propagate": [{"outcome": oc.split("::")[1], "min_h": float(a), "max_h": float(bb), "n_steps": int(n), "time": float(t)}
                      for x in props for oc, a, bb, n, t in zip(x["Outcome"], x["Min. timestep"], x["Max. timestep"],
                                                               x["Num. of steps"], x["Current time"])],
        "state_back_at_0": json.loads(next(x for x in props if "State" in x)["State"][0]),
        "grid": {"times": [0.1 * i for i in range(11)], "index": 4,
                 "x": float(grid_lines[0].split("=")[1]), "v": float(grid_lines[1].split("=")[1])},
    }
    assert [r["n_steps"] for r in out["propagate"]] == [24, 72, 97], out["propagate"]
    with open(os.path.join(HERE, "tut_adaptive.json"), "w") as f:
        json.dump(out, f, indent=1)

    # ---- tut_d_output.rst ----
    blocks = console_blocks(open(os.path.join(REF, "doc", "tut_d_output.rst")).read())
    d01 = next(bl for bl in blocks if bl[0].startswith("x(0.1)"))
    nst = next(kv(bl) for bl in blocks if any(ln.startswith("N of steps") for ln in bl))
    samples = next(bl for bl in blocks if bl[0].startswith("time=0,"))
    out = {
        "source": "doc/tut_d_output.rst (output of tutorial/d_output.cpp)",
        "system": "x' = v, v' = -9.8 sin(x)", "x0": 0.05, "v0": 0.025,
        "d_output_at_0.1": [float(d01[0].split("=")[1]), float(d01[1].split("=")[1])],
        "c_output": {"t_final": 10.0, "n_steps": int(nst["N of steps"][0]),
                     "samples": [[float(v.split("=")[1]) for v in ln.split(",")] for ln in samples]},
    }
    assert out["c_output"]["n_steps"] == 48 and len(out["c_output"]["samples"]) == 6
    with open(os.path.join(HERE, "tut_d_output.json"), "w") as f:
        json.dump(out, f, indent=1)

    # ---- tut_events.rst ----
    blocks = console_blocks(open(os.path.join(REF, "doc", "tut_events.rst")).read())
    times = [[float(ln.split(":")[1]) for ln in bl] for bl in blocks if bl[0].startswith("Event detection time")]
    xs = next([float(ln.split(":")[1]) for ln in bl] for bl in blocks if bl[0].startswith("Value of x when v is zero"))
    multi = next([(int(re.match(r"Event (\d)", ln).group(1)), float(ln.split("t=")[1])) for ln in bl]
                 for bl in blocks if bl[0].startswith("Event 0 triggering"))
    term = next(kv(bl) for bl in blocks if bl[0].startswith("Integration outcome"))
    grid = next(bl for bl in blocks if bl[0].startswith("[-0.0297"))
    out = {
        "source": "doc/tut_events.rst (output of tutorial/event_basic.cpp)",
        "system": "x' = v, v' = -9.8 sin(x)", "x0": -0.05, "v0": 0.0, "t_final": 5.0,
        "nt_zero_velocity": {"times": times[0], "x_at_events": xs},
        "nt_zero_velocity_positive_direction": {"times": times[1]},
        "nt_two_events": {"second_event": "v*v - 1e-12", "sequence": [{"event": e, "t": t} for e, t in multi]},
        "terminal_damping_toggle": {
            "system": "x' = v, v' = -9.8 sin(x) - par[0] v", "x0": 0.05, "v0": 0.025,
            "first_stop_event_index": int(term["Event index"][0]),
            "grid": [float(i) for i in range(1, 11)],
            "grid_output": [json.loads(ln) for ln in grid if ln.startswith("[")],
            "final_time": float(next(ln for ln in grid if ln.startswith("Final time")).split(":")[1]),
        },
    }
    assert len(times[0]) == 5 and len(times[1]) == 3 and len(multi) == 14 and len(out["terminal_damping_toggle"]["grid_output"]) == 10
    with open(os.path.join(HERE, "tut_events.json"), "w") as f:
        json.dump(out, f, indent=1)

    more_tutorials()


def more_tutorials():
    def state_of(block):
        return json.loads(kv(block)["State"][0])

    # ---- tut_ensemble.rst ----
    blocks = console_blocks(open(os.path.join(REF, "doc", "tut_ensemble.rst")).read())
    st = next(bl for bl in blocks if any(ln.startswith("State") for ln in bl))
    res = kv(next(bl for bl in blocks if bl[0].startswith("Integration outcome")))
    mn, mx = res["Min/max timesteps"][0].split("/")
    out = {
        "source": "doc/tut_ensemble.rst (output of tutorial/ensemble.cpp)",
        "system": "x' = v, v' = -9.8 sin(x)", "n_iter": 10, "t_final": 20.0,
        "ics": [[0.05 + i / 100., 0.025 + i / 100.] for i in range(10)],
        "member": 9, "state": state_of(st), "time": float(kv(st)["Time"][0]),
        "outcome": res["Integration outcome"][0].split("::")[1], "min_h": float(mn), "max_h": float(mx),
        "n_steps": int(res["N of timesteps"][0]),
    }
    assert out["n_steps"] == 124 and out["time"] == 20.0
    with open(os.path.join(HERE, "tut_ensemble.json"), "w") as f:
        json.dump(out, f, indent=1)

    # ---- tut_param.rst ----
    blocks = [bl for bl in console_blocks(open(os.path.join(REF, "doc", "tut_param.rst")).read())
              if any(ln.startswith("State") for ln in bl)]
    out = {
        "source": "doc/tut_param.rst (output of tutorial/pendulum_param.cpp)",
        "system": "x' = v, v' = -par[0] / par[1] * sin(x)", "x0": 0.05, "v0": 0.0,
        "runs": [{"pars": json.loads(kv(bl)["Parameters"][0]), "t_final": float(kv(bl)["Time"][0]), "state": state_of(bl)}
                 for bl in blocks[1:]],
        "order": int(kv(blocks[0])["Taylor order"][0]),
    }
    assert len(out["runs"]) == 2 and out["runs"][1]["pars"][0] == 3.72
    with open(os.path.join(HERE, "tut_param.json"), "w") as f:
        json.dump(out, f, indent=1)

    # ---- tut_nonauto.rst ----
    blocks = console_blocks(open(os.path.join(REF, "doc", "tut_nonauto.rst")).read())
    xs = [float(ln.split("=")[1]) for ln in next(bl for bl in blocks if bl[0].startswith("x = "))]
    out = {
        "source": "doc/tut_nonauto.rst (output of tutorial/forced_damped_pendulum.cpp)",
        "system": "x' = v, v' = cos(t) - 0.1 v - sin(x)", "x0": 0.0, "v0": 1.85, "delta_t": 2.0, "x": xs,
    }
    assert len(xs) == 25
    with open(os.path.join(HERE, "tut_nonauto.json"), "w") as f:
        json.dump(out, f, indent=1)

    # ---- tut_adaptive_custom.rst ----
    blocks = [bl for bl in console_blocks(open(os.path.join(REF, "doc", "tut_adaptive_custom.rst")).read())
              if any(ln.startswith("State") for ln in bl)]
    out = {
        "source": "doc/tut_adaptive_custom.rst (output of tutorial/adaptive_opt.cpp)",
        "system": "x' = v, v' = -9.8 sin(x)", "x0": 0.05, "v0": 0.025, "tol": 1e-9,
        "order": int(kv(blocks[0])["Taylor order"][0]), "times": [10.0, 0.0], "state_back_at_0": state_of(blocks[1]),
    }
    assert out["order"] == 12
    with open(os.path.join(HERE, "tut_adaptive_custom.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
