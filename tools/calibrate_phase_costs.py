"""FP64 instructions per counted event of k_map<4>: per-phase ncu attribution / event counts of the same workload.

    python tools/calibrate_phase_costs.py PHASES.json EVENTS.json NCU_SUMMARY.csv OUT.json

PHASES.json: tools/ncu_phase_table.py; EVENTS.json: tools/count_events.py (same sources, same workload);
NCU_SUMMARY.csv: tools/ncu_summary.py (pipe %, DRAM bytes).  bench.py multiplies these costs with the event counts it
measures in the run (roofline.executed); the source hash ties the file to the kernels it was measured on."""
import csv
import json
import sys

# which counted event pays for which phase region (tools/ncu_phase_table.py; //@phase markers in csrc/mlm_core.cuh)
PHASE_EVENT = {
    "polynomial": "positions", "kernel_body": "positions", "track_extrapolate": "positions", "track_vieta": "positions",
    "track_other": "positions", "store": "positions", "coords": "positions", "unattributed": "positions",
    "cold_glue": "positions",
    "lens_newton": "lens_evals", "track_image_newton": "lens_evals",
    "track_nonphys_newton": "newton_evals", "track_nonphys_check": "nonphys_checks",
    "track_quadratic": "quadratics", "watch_fp32": "quadratics", "track_polish": "watch_near",
    "cold_solve": "cold_solves", "classify": "classify_calls",
    "image_wk": "images", "image_terms": "images", "image_sums": "images",
}


def main():
    phases = json.load(open(sys.argv[1]))
    events = json.load(open(sys.argv[2]))
    summ = {r[0]: r for r in csv.reader(open(sys.argv[3]))} if len(sys.argv) > 4 and sys.argv[3] != "-" else {}
    out_path = sys.argv[4] if len(sys.argv) > 4 else sys.argv[3]
    per_pos = events["per_position"]
    cost = {}
    for ph, t in phases["per_position"].items():
        e = PHASE_EVENT.get(ph, "positions")
        c = cost.setdefault(e, {"fp64_inst": 0.0, "flop": 0.0, "phases": []})
        cnt = per_pos[e]
        if cnt <= 0:
            e, cnt = "positions", 1.0
            c = cost.setdefault(e, {"fp64_inst": 0.0, "flop": 0.0, "phases": []})
        c["fp64_inst"] += t["fp64_thread"] / cnt
        c["flop"] += (2 * t["dfma"] + t["dmul"] + t["dadd"]) / cnt
        c["phases"].append(ph)
    tot = phases["total_per_position"]

    def num(key):
        try:
            return float(summ[key][-1])
        except Exception:
            return None
    dr, dw = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    unit_r = summ.get("dram__bytes_read.sum", [None, ""])[1] if summ else ""
    unit_w = summ.get("dram__bytes_write.sum", [None, ""])[1] if summ else ""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = None
    if dr is not None and dw is not None:
        dram = dr * scale.get(unit_r, 1.0) + dw * scale.get(unit_w, 1.0)
    res = {"kernel": "k_map<4>", "workload": events["workload"], "strip": events["strip"], "source_hash": events["source_hash"],
           "profile": phases["report"], "cost_per_event": cost, "events_per_position": per_pos,
           "ncu_fp64_inst_per_position": tot["fp64_thread"], "ncu_flop_per_position": tot["flop"],
           "ncu_dfma_dmul_dadd_per_position": [tot["dfma"], tot["dmul"], tot["dadd"]],
           "avg_active_lanes_fp64": tot["avg_active_lanes_fp64"],
           "ncu_fp64_pipe_pct": num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
           "ncu_duration_ms": num("gpu__time_duration.sum"), "dram_bytes_per_launch": dram,
           "model_check_fp64_inst_per_position": sum(per_pos[e] * c["fp64_inst"] for e, c in cost.items())}
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "cost_per_event"}, indent=1))
    for e, c in cost.items():
        print(f"{e:18s} {per_pos[e]:10.4f} /pos x {c['fp64_inst']:9.2f} FP64 = {per_pos[e] * c['fp64_inst']:8.1f}   {c['phases']}")


if __name__ == "__main__":
    main()
