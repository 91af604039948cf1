#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small CSV + markdown table for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_name [positions_per_launch [executed.json]]

With the 4th argument the executed-work figures of the LAST launch are also written as JSON
(profiles/k_map_executed.json is what bench.py reads for roofline.executed / roofline.traffic).
"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread",
    "launch__occupancy_limit_registers",
    "launch__grid_size", "launch__block_size",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__cycles_elapsed.avg",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_branch_targets_threads_divergent.sum",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio" ,
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    npos = float(sys.argv[3]) if len(sys.argv) > 3 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    name_i = idx.get("Kernel Name")
    lines = ["metric,unit," + ",".join(f"launch{k}" for k in range(len(data)))]
    md = [f"ncu --set full summary of `{rep}`", "",
          "kernels: " + ", ".join(r[name_i] for r in data) if name_i is not None else "", "",
          "| metric | unit | " + " | ".join(f"launch {k}" for k in range(len(data))) + " |",
          "|---|---|" + "---|" * len(data)]
    vals = {}
    for k in KEYS:
        if k in idx:
            v = [r[idx[k]] for r in data]
            vals[k] = v
            lines.append(f"{k},{units[idx[k]]}," + ",".join(v))
            md.append(f"| {k} | {units[idx[k]]} | " + " | ".join(v) + " |")
    if npos:
        try:
            f = lambda k: float(vals[k][-1].replace(",", ""))
            dur_unit = units[idx["gpu__time_duration.sum"]]
            dur = f("gpu__time_duration.sum") * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(dur_unit, 1e-3)
            cyc = f("smsp__cycles_elapsed.avg")           # .sum.per_cycle_elapsed x elapsed cycles = totals
            dfma = f("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") * cyc
            dmul = f("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed") * cyc
            dadd = f("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed") * cyc
            flop = 2 * dfma + dmul + dadd
            md += ["", f"positions per launch: {npos:.0f}",
                   f"executed FP64 thread instructions per position: DFMA {dfma / npos:.0f}, DMUL {dmul / npos:.0f}, "
                   f"DADD {dadd / npos:.0f} (total {(dfma + dmul + dadd) / npos:.0f})",
                   f"executed flop per position (2 DFMA + DMUL + DADD): {flop / npos:.0f}",
                   f"executed FP64 rate under ncu: {flop / dur / 1e12:.2f} TFLOP/s; "
                   f"positions/s under ncu: {npos / dur:.3e}"]
            if len(sys.argv) > 4:
                def tobytes(k):
                    u = units[idx[k]].lower()
                    return f(k) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
                js = {"kernel": data[-1][name_i] if name_i is not None else None, "positions_per_launch": npos,
                      "fp64_inst_per_position": (dfma + dmul + dadd) / npos,
                      "dfma_dmul_dadd_per_position": [dfma / npos, dmul / npos, dadd / npos],
                      "flop_per_position": flop / npos,
                      "fp64_pipe_pct": f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                      "registers_per_thread": f("launch__registers_per_thread"),
                      "dram_bytes_per_launch": tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"),
                      "duration_ms_under_ncu": dur * 1e3, "source": out + ".md"}
                json.dump(js, open(sys.argv[4], "w"), indent=1)
        except Exception as e:  # noqa
            md.append(f"(derived numbers unavailable: {e})")
    open(out + ".csv", "w").write("\n".join(lines) + "\n")
    open(out + ".md", "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
