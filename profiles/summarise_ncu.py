#!/usr/bin/env python3
"""Summarise an `ncu --set full` report into the small text files committed under profiles/.

    python profiles/summarise_ncu.py gpurun_out/X.ncu-rep --name r1_coop_final --kind smem --lane-steps 1234567

Writes profiles/<name>_summary.csv (selected metrics of every profiled launch) and, when --lane-steps is given,
records DRAM bytes / lane-steps (+ FP64 pipe utilisation, instructions per lane-step) of the first launch in
profiles/r2_traffic.json under <kind> (read by bench.py for
roofline.traffic). Needs only the `ncu` CLI (no GPU)."""
import argparse
import csv
import io
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.per_cycle_active",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct", "smsp__issue_active.avg.per_cycle_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__maximum_warps_per_active_cycle_pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.strip().lower()
    mult = {"byte": 1, "bytes": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
    return v * mult


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--name", required=True)
    ap.add_argument("--kind", default=None, help="smem | hbm: key in r1_traffic.json")
    ap.add_argument("--lane-steps", type=float, default=0.0)
    ap.add_argument("--note", default="")
    ap.add_argument("--traffic-file", default="r2_traffic.json")
    a = ap.parse_args()

    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], check=True, capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    out_path = os.path.join(HERE, a.name + "_summary.csv")
    with open(out_path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["# source: %s %s" % (os.path.basename(a.report), a.note)])
        w.writerow(["launch", "kernel", "metric", "unit", "value"])
        for li, r in enumerate(data):
            kname = r[col["Kernel Name"]]
            for m in KEEP:
                if m in col:
                    w.writerow([li, kname[:60], m, units[col[m]], r[col[m]]])
    print("wrote", out_path)

    if a.kind and a.lane_steps > 0 and data:
        r = data[0]
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        tpath = os.path.join(HERE, a.traffic_file)
        t = {}
        if os.path.exists(tpath):
            with open(tpath) as f:
                t = json.load(f)
        def num(name):
            try:
                return float(r[col[name]].replace(",", ""))
            except (KeyError, ValueError):
                return None

        insts = num("smsp__inst_executed.sum")
        t[a.kind] = {"dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "lane_steps": a.lane_steps,
                     "bytes_per_lane_step": (rd + wr) / a.lane_steps, "kernel": r[col["Kernel Name"]][:80],
                     "fp64_pipe_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                     "issue_active": num("smsp__issue_active.avg.per_cycle_active"),
                     "warp_inst_per_lane_step": None if insts is None else insts / a.lane_steps,
                     "kernel_ms": num("gpu__time_duration.sum"),
                     "source": "profiles/%s_summary.csv (%s)" % (a.name, os.path.basename(a.report)), "note": a.note}
        with open(tpath, "w") as f:
            json.dump(t, f, indent=1)
        print("updated", tpath, t[a.kind]["bytes_per_lane_step"], "B/lane-step")


if __name__ == "__main__":
    main()
