#!/usr/bin/env python
"""Benchmark of the binary-lens A0/A2/A4 hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C4|C3|C2|C5]

One "step" = one pass of the hot path over the whole synthetic workload: for the headline
configuration C4 an 8192 x 8192 magnification map (s=1.0, q=0.5, rho=1e-3, Gamma=0.5,
centre-of-mass window centre (-0.08, 0), half-width 0.8), row-sharded over the N ranks
(strong scaling), followed for N>1 by the gather of the results to rank 0 over NCCL.

`value`   : whole-job A4 magnifications/s with everything resident on the device (zeta is
            generated in the kernel, results stay in HBM), CUDA-event timed, max over ranks.
`e2e`     : the same metric through the public host API (microlensing_b200.magnification_map):
            host buffers, D2H of all results inside the timed region (wall clock).
`roofline`: FP64-pipe roofline.  achieved = value x W / 1e12 TFLOP/s with the ALGORITHMIC work W
            per magnification of SURVEY.md §8(d) (12106 + 4068 f5 flop, f5 = measured fraction of
            5-image positions); peak = DFMA-chain microbenchmark measured live on this GPU
            (MEASURED_PEAKS.json has no FP64 entry; nominal 37.2 TFLOP/s also stated).
`cpu_baseline`: the oracle's per-position replay of the reference (NumPy, np.roots) on all host
            cores, on a bounded seeded sample of the same grid.
`--impl reference` times only that CPU path (the reference is pure Python; the oracle port is
what can travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "A4 magnifications/sec (FP64, 8192^2 map)"
UNIT = "mags/s"
NOMINAL_FP64_TFLOPS = 37.2        # 148 SM x 64 FMA lanes x 2 x 1.965 GHz

CONFIGS = {
    # name: dict(kind, s, q, rho, Gamma, ...)   windows: SURVEY.md §8(d)
    "C4": dict(kind="map", s=1.0, q=0.5, rho=1e-3, Gamma=0.5, cx=-0.08, cy=0.0, half=0.80, nx=8192, ny=8192),
    "C3": dict(kind="map", s=0.8, q=0.1, rho=1e-3, Gamma=0.5, cx=0.11, cy=0.0, half=0.75, nx=4096, ny=4096),
    "C2": dict(kind="lightcurve", s=1.0, q=1e-3, rho=1e-3, Gamma=0.5, u0=0.01, alpha=0.35, tau0=-2.0, tau1=2.0,
               T=1000000),
    "C5": dict(kind="models", M=10000, T=2000, seed=20240517),
}


def work_per_mag(f5: float) -> float:
    """Algorithmic FP64 flop per A4 magnification."""
    return 12106.0 + 4068.0 * f5


def map_window(c):
    nx, ny = c["nx"], c["ny"]
    dx, dy = 2 * c["half"] / nx, 2 * c["half"] / ny
    return c["cx"] - c["half"], c["cy"] - c["half"], dx, dy


# ------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference's per-position path), all host cores
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    s, q, rho, G, zeta = args
    from oracle import cassan_oracle as oracle
    out = np.empty((len(zeta), 4))
    for i, z in enumerate(zeta):
        out[i] = oracle.point_a4(s, q, rho, G, z)
    return out


def sample_positions(cfg, n, seed=12345):
    """Seeded uniform subsample of the workload's own grid -> (flat indices, primary-frame zeta)."""
    from microlensing_b200.batched import map_coordinates, lightcurve_coordinates
    rng = np.random.default_rng(seed)
    if cfg["kind"] == "map":
        x0, y0, dx, dy = map_window(cfg)
        idx = rng.integers(0, cfg["nx"] * cfg["ny"], size=n)
        r, c = np.divmod(idx, cfg["nx"])
        s, q = cfg["s"], cfg["q"]
        x = (x0 + (c + 0.5) * dx) + (-(s * q / (1. + q)))
        y = y0 + (r + 0.5) * dy
        return idx, x + 1j * y
    if cfg["kind"] == "lightcurve":
        idx = np.sort(rng.integers(0, cfg["T"], size=n))
        dtau = (cfg["tau1"] - cfg["tau0"]) / cfg["T"]
        z = lightcurve_coordinates(cfg["s"], cfg["q"], cfg["u0"], cfg["alpha"], cfg["tau0"], dtau, cfg["T"])
        return idx, z[idx]
    raise ValueError(cfg["kind"])


def run_cpu_baseline(cfg, per_core=6000, cores=None):
    """Time the oracle's per-position replay on `cores` processes; returns dict + sample results."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    n = per_core * cores
    idx, zeta = sample_positions(cfg, n)
    chunks = np.array_split(zeta, cores * 4)
    args = [(cfg["s"], cfg["q"], cfg["rho"], cfg["Gamma"], ch) for ch in chunks]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(cfg["s"], cfg["q"], cfg["rho"], cfg["Gamma"], zeta[:8])] * cores)   # warm the workers
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, args)
        dt = time.perf_counter() - t0
    res = np.concatenate(res)
    return dict(value=n / dt, unit=UNIT, cores=cores, kind="port", seconds=dt,
                sample=f"{n} seeded random positions of the same grid ({per_core} per core), oracle per-position "
                       f"replay of multipoles.py:181-202 (np.roots + NumPy recursion), multiprocessing.Pool({cores})"), \
        idx, zeta, res


# ------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": the upper half of the samples (idle samples before/after the region are dropped)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:]
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------
def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    per_core = max(200, args.ref_per_core)
    vals = []
    for it in range(args.warmup + args.steps):
        b, _, _, _ = run_cpu_baseline(cfg, per_core=per_core, cores=cores)
        if it >= args.warmup:
            vals.append(b)
    v = float(np.mean([b["value"] for b in vals]))
    sec = float(np.mean([b["seconds"] for b in vals]))
    n = per_core * cores
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg, args, extra={"step": f"bounded sample of {n} positions of the workload per step"}),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": vals[-1]["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is pure Python/NumPy (no compiled sources): timed through oracle/cassan_oracle.py, its "
                "restatement that travels to the GPU box; extrapolated full-workload time "
                f"{workload_size(cfg) / v / 3600.0:.2f} h on {cores} cores",
    }
    print(json.dumps(line))
    return 0


def workload_size(cfg):
    if cfg["kind"] == "map":
        return cfg["nx"] * cfg["ny"]
    if cfg["kind"] == "lightcurve":
        return cfg["T"]
    return cfg["M"] * cfg["T"]


def workload_config(cfg, args, extra=None):
    d = {"workload": f"{args.config}: " + (
        f"{cfg['nx']}x{cfg['ny']} magnification map s={cfg['s']} q={cfg['q']} rho={cfg['rho']} Gamma={cfg['Gamma']} "
        f"window centre ({cfg['cx']},{cfg['cy']}) half-width {cfg['half']} (centre-of-mass frame), row-sharded"
        if cfg["kind"] == "map" else str({k: v for k, v in cfg.items()})),
        "positions": workload_size(cfg), "order": 4,
        "l2": "no inputs (zeta generated in-kernel); outputs 26 B/position are written once and exceed the 126 MB L2"}
    if extra:
        d.update(extra)
    return d


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def bench_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; microlensing_b200 has no CPU fallback")
    cfg = CONFIGS[args.config]
    if cfg["kind"] == "models":
        return bench_models(args)
    if args.nx:
        cfg = dict(cfg, nx=args.nx, ny=args.nx)
    base = None
    if rank == 0 and world == 1 and cfg["kind"] == "map":
        # CPU baseline first: the worker processes are forked before any CUDA context exists
        base, idx, zeta, ref = run_cpu_baseline(cfg, per_core=args.cpu_per_core)
    torch.cuda.set_device(local)
    os.environ["MLM_DEVICE"] = str(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    import microlensing_b200 as mb
    from microlensing_b200 import _capi
    from microlensing_b200.sharding import ShardedMap, shard_range

    if cfg["kind"] != "map":
        raise SystemExit("bench.py times map configurations (C4 headline, C3); C2/C5 are parity-test cases")
    nx, ny = cfg["nx"], cfg["ny"]
    x0, y0, dx, dy = map_window(cfg)
    s, q, rho, G = cfg["s"], cfg["q"], cfg["rho"], cfg["Gamma"]
    ctx = _capi.context(local)
    dev = torch.device("cuda", local)
    sm = ShardedMap(nx, ny, rank, world, device=dev, nchunks=args.chunks or None,
                    gather=(args.gather if world > 1 else "none"), layout=args.layout)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(gather=True):
        sm.compute(s, q, rho, G, x0, y0, dx, dy, order=4, gather=gather)

    sampler = ClockSampler(local) if rank == 0 else None
    # ---- device-resident, whole job (kernel + gather for N>1) --------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    if sampler:
        sampler.start()
        time.sleep(0.3)
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = ctx.launch_count() - l0
    ms = e0.elapsed_time(e1) / args.steps
    # ---- kernel only (no gather), for the roofline of the dominant kernel -----------------------
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        step(gather=False)
    k1.record()
    barrier()
    ms_kernel = k0.elapsed_time(k1) / args.steps
    if sampler:
        clocks = sampler.stop()
    t = torch.tensor([ms, ms_kernel], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_kernel = float(t[0]), float(t[1])
    n_total = nx * ny
    value = n_total / (ms * 1e-3)

    # ---- same job storing / gathering only what an A4 map needs (A4, nimg, flags: 10 B/position) ------
    sm_a4 = ShardedMap(nx, ny, rank, world, device=dev, nchunks=args.chunks or None,
                       gather=(args.gather if world > 1 else "none"), outputs=("A4", "nimg", "flags"),
                       layout=args.layout)
    for _ in range(2):
        sm_a4.compute(s, q, rho, G, x0, y0, dx, dy, order=4)
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(args.steps):
        sm_a4.compute(s, q, rho, G, x0, y0, dx, dy, order=4)
    a1.record()
    barrier()
    t4 = torch.tensor([a0.elapsed_time(a1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t4, op=dist.ReduceOp.MAX)
    ms_a4 = float(t4[0])
    sm_a4.close()
    del sm_a4

    # image-count histogram of this run (for the algorithmic work per magnification)
    if world > 1 and sm.gather == "p2p":
        # the shards were written straight into rank 0's full map
        fv = sm.full_views()
        nim, flg = (fv["nimg"].reshape(-1), fv["flags"].reshape(-1)) if rank == 0 else (None, None)
    else:
        nim, flg = sm.local["nimg"][: sm.plan.nrows * nx], sm.local["flags"][: sm.plan.nrows * nx]
    cnt5 = torch.zeros(3, dtype=torch.float64, device=dev)
    if nim is not None:
        cnt5 = torch.tensor([float((nim == 5).sum()), float(nim.numel()), float((flg != 0).sum())],
                            dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(cnt5)
    f5 = float(cnt5[0] / cnt5[1])
    flag_any = float(cnt5[2])

    # ---- end to end through the public host API ---------------------------------------------------
    lo, hi = shard_range(ny, rank, world, sm.align)
    out = mb.alloc_outputs((hi - lo, nx), device=local, pinned=True)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, nx, ny, row0=lo, nrows=hi - lo, out=out, device=local)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, nx, ny, row0=lo, nrows=hi - lo, out=out, device=local)
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te[0])
    out4 = mb.alloc_outputs((hi - lo, nx), device=local, pinned=True, outputs=("A4", "nimg", "flags"))
    mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, nx, ny, row0=lo, nrows=hi - lo, out=out4, device=local)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, nx, ny, row0=lo, nrows=hi - lo, out=out4, device=local)
    torch.cuda.synchronize()
    te4 = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te4, op=dist.ReduceOp.MAX)
    t_e2e4 = float(te4[0])

    if rank != 0:
        if world > 1:
            sm.close()
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (k_map<4>) -----------------------------------------------------
    peak_meas = ctx.fp64_peak_tflops()
    W = work_per_mag(f5)
    per_launch_positions = n_total / world / len(sm.plan.chunks)
    kernel_rate = n_total / (ms_kernel * 1e-3)            # whole job, kernels only
    achieved = kernel_rate / world * W / 1e12             # per GPU
    info = ctx.kernel_info()
    ex = _executed_profile() if args.config == "C4" and not args.nx else None
    executed = None
    if ex:
        ex_tflops = kernel_rate / world * ex["flop_per_position"] / 1e12
        inst_rate = kernel_rate / world * ex["fp64_inst_per_position"]
        nominal_issue = 148 * 64 * 1.965e9                 # FP64 thread-instructions/s: 64 lanes per SM
        probe = ctx.fp64_probe(2, 4)                       # 3-distinct-operand DFMA/DMUL/DADD mix, 16 warps/SM
        executed = {"flop_per_mag": ex["flop_per_position"], "fp64_inst_per_mag": ex["fp64_inst_per_position"],
                    "tflops": ex_tflops, "frac": ex_tflops / peak_meas,
                    "fp64_inst_per_s": inst_rate, "frac_of_nominal_issue_rate": inst_rate / nominal_issue,
                    "operand_limited_issue_rate": probe, "frac_of_operand_limited_rate": inst_rate / probe,
                    "fp64_pipe_pct_ncu": ex.get("fp64_pipe_pct"), "source": ex.get("source"),
                    "note": "operand_limited_issue_rate = mlm_fp64_probe(mode 2): independent DFMA/DMUL/DADD (11:4:1) "
                            "whose three source operands are distinct registers sustain only ~0.73 of the nominal "
                            "FP64 issue rate on B200 (register-file operand bandwidth); the constant-operand DFMA chain "
                            "behind `peak` reaches 0.92"}
    roofline = {
        "bound": "fp64", "achieved": achieved, "peak": peak_meas, "unit": "TFLOP/s", "frac": achieved / peak_meas,
        "traffic": (ex or {}).get("dram_bytes_per_launch"), "executed": executed,
        "kernel": "k_map<4>", "peak_source": "DFMA-chain microbenchmark (mlm_fp64_peak) measured live on this GPU; "
                                             "MEASURED_PEAKS.json has no FP64 entry",
        "peak_nominal": NOMINAL_FP64_TFLOPS, "frac_of_nominal": achieved / NOMINAL_FP64_TFLOPS,
        "algorithmic_flop_per_mag": W, "f5": f5, "positions_per_launch": per_launch_positions,
        "kernel_ms_per_step": ms_kernel, "algorithmic_bytes_per_mag": 26,
        "hbm_frac": (kernel_rate / world * 26 / 1e9) / _hbm_peak(),
        "registers_per_thread": info["k_map4_regs"], "local_bytes_per_thread": info["k_map4_local_bytes"],
        "note": "achieved/frac use the ALGORITHMIC flop count of SURVEY.md §8(d) (the reference's formulae, 8 Aberth "
                "sweeps nominal) as the contract asks; the kernel reaches the same numbers with far less arithmetic "
                "(warm-started root tracking, closed-form mu2/mu4), so frac exceeds 1: `executed` is the ncu-counted "
                "FP64 work (2 DFMA + DMUL + DADD per magnification, profiles/) at the measured rate -- the fraction "
                "of the FP64 pipe actually in use",
    }
    # ---- CPU baseline + parity on the same sample ---------------------------------------------------
    base_out = None
    if base is not None:
        g = mb.magnification_points(s, q, rho, G, zeta, device=local)
        ok = g["nimg"] == ref[:, 0]
        with np.errstate(all="ignore"):
            rel = np.max(np.abs(np.stack([g["A0"], g["A2"], g["A4"]], 1) / ref[:, 1:4] - 1.0), axis=1)
        clean = ok & (g["flags"] == 0)
        parity = {"n": int(len(zeta)), "nimg_mismatch": int((~ok).sum()), "rel_gt_1e-9": int((rel[ok] > 1e-9).sum()),
                  "max_rel": float(np.nanmax(rel[ok])), "flagged": int((g["flags"] != 0).sum()),
                  "rel_gt_1e-9_unflagged": int((rel[clean] > 1e-9).sum()), "max_rel_unflagged": float(np.nanmax(rel[clean])),
                  "note": "flagged = near-critical / series-not-converging / ambiguous-residual positions (MLM_FLAG_*), "
                          "where the reference's own rounding is amplified (SURVEY.md §8(c) rule 3)"}
        base_out = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        base_out["parity_vs_gpu_on_sample"] = parity
        base_out["full_workload_extrapolated_hours"] = n_total / base["value"] / 3600.0

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(cfg, args, extra={"parallelism": f"rows/{world} ({sm.layout} strips of {sm.align} rows)", "chunks_per_rank": len(sm.plan.chunks),
                                                    "gather": {"p2p": "fused: the map kernel stores its rows straight into rank 0's HBM "
                                                                      "over NVLink (CUDA IPC peer mapping), 1-element NCCL "
                                                                      "all-reduce as completion fence",
                                                               "nccl": "NCCL gather to rank 0, chunk-pipelined",
                                                               "none": "none"}[sm.gather]}),
        "roofline": roofline, "cpu_baseline": base_out,
        "e2e": {"value": n_total / t_e2e, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": int(26 * n_total), "ms_per_step": t_e2e * 1e3, "steps": e2e_steps,
                "api": "microlensing_b200.magnification_map (host buffers, pinned; D2H inside the call)"},
        "gpu_launches": int(launches), "clocks": clocks,
        "kernel_only": {"value": kernel_rate, "ms_per_step": ms_kernel},
        "a4_outputs_only": {"note": "same job, same kernel arithmetic, but only A4 + nimg + flags (10 B/position instead "
                                    "of 26) are stored, gathered to rank 0 (value) or copied to the host (e2e); "
                                    "informational -- the headline `value` and `e2e` move all five arrays",
                            "value": n_total / (ms_a4 * 1e-3), "ms_per_step": ms_a4,
                            "e2e": {"value": n_total / t_e2e4, "ms_per_step": t_e2e4 * 1e3,
                                    "d2h_bytes_per_step": int(10 * n_total)}},
        "flagged_positions": flag_any,
    }
    if world == 1 and not args.no_other:
        line["other_configs"] = other_configs(mb, dev)
    print(json.dumps(line))
    if world > 1:
        sm.close()
        dist.barrier()
        dist.destroy_process_group()
    return 0


def other_configs(mb, dev, reps=3):
    """Device-resident kernel timings of the other BASELINE.json configurations (parity-test cases, not
    the headline): CUDA events around `reps` launches after one warm-up, outputs resident in HBM."""
    import torch
    from microlensing_b200.batched import map_device, models_device, points_device, lightcurve_coordinates
    st = torch.cuda.current_stream().cuda_stream
    out = {}

    def bufs(n):
        return ([torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
                + [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)])

    def timed(fn, n, label, extra=None):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[label] = dict(positions=n, ms=ms, mags_per_s=n / (ms * 1e-3), **(extra or {}))

    # C1: the drop-in functions on the example() case (single position, 5 images): wall-clock latency per call
    from microlensing_b200.multipoles import multipoles as mp
    s1, q1, rho1, G1 = 1.7, 0.2, 0.01, 0.0
    zeta0 = complex(-.5, 0.) - complex(s1 * q1 / (1. + q1), 0.)
    z0 = mp._solvelenseq(s1, q1, zeta0)
    Wk = np.zeros((7, 5), dtype=np.complex128)
    for k in range(2, 7):
        Wk[k] = (-1) ** (k - 1) * float(np.prod(np.arange(1, k))) / (1 + q1) * (z0 ** -k + q1 * (z0 + s1) ** -k)
    lat = {}
    for name, fn in (("_solvelenseq", lambda: mp._solvelenseq(s1, q1, zeta0)),
                     ("hexadecapole(Wk)", lambda: mp.hexadecapole(Wk, rho1, G1)),
                     ("magnification_points(1 position)", lambda: mb.magnification_points(s1, q1, rho1, G1, zeta0))):
        for _ in range(20):
            fn()
        t0 = time.perf_counter()
        for _ in range(300):
            fn()
        lat[name] = (time.perf_counter() - t0) / 300 * 1e6
    out["C1_dropin_latency_us_per_call"] = dict(lat, note="host call -> H2D -> kernel(s) -> D2H -> return, one position; "
                                                "the reference needs ~450 us for the same position on one core")

    c = CONFIGS["C3"]
    x0, y0, dx, dy = map_window(c)
    n = c["nx"] * c["ny"]
    b = bufs(n)
    from microlensing_b200.sharding import auto_strip
    ctx = mb._capi.context(dev.index)
    old_strip, strip3 = ctx.map_strip(), auto_strip(c["nx"], c["ny"])
    ctx.set_map_strip(strip3)                      # a 4096^2 map is a quarter of the headline: shorter strips (sharding.auto_strip)
    try:
        timed(lambda: map_device(c["s"], c["q"], c["rho"], c["Gamma"], x0, y0, dx, dy, c["nx"], 0, c["ny"], *b, stream=st),
              n, "C3_map_4096", {"kernel": "k_map<4>", "strip_rows": strip3, "f5": None})
    finally:
        ctx.set_map_strip(old_strip)
    out["C3_map_4096"]["f5"] = float((b[3] == 5).float().mean())

    c = CONFIGS["C2"]
    T = c["T"]
    dtau = (c["tau1"] - c["tau0"]) / T
    b = bufs(T)
    par = [torch.tensor([c[k]], dtype=torch.float64, device=dev) for k in ("s", "q", "rho", "Gamma", "u0", "alpha")]
    timed(lambda: models_device(*par, 1, c["tau0"], dtau, T, *b, stream=st), T, "C2_lightcurve_1e6_tracked",
          {"kernel": "k_models<4> (in-kernel trajectory, warm-started segments)"})
    zeta = torch.from_numpy(lightcurve_coordinates(c["s"], c["q"], c["u0"], c["alpha"], c["tau0"], dtau, T)).to(dev)
    timed(lambda: points_device(c["s"], c["q"], c["rho"], c["Gamma"], zeta, T, *b, stream=st), T,
          "C2_lightcurve_1e6_points", {"kernel": "k_points<4> (arbitrary positions, cold solve each)"})

    c = CONFIGS["C5"]
    rng = np.random.default_rng(c["seed"])
    M, T = c["M"], c["T"]
    sv = np.exp(rng.uniform(np.log(0.5), np.log(2.), M))
    qv = np.exp(rng.uniform(np.log(1e-4), np.log(1.), M))
    rv = np.exp(rng.uniform(np.log(1e-4), np.log(1e-2), M))
    gv = np.full(M, 0.5)
    uv, av = rng.uniform(-.5, .5, M), rng.uniform(0, 2 * np.pi, M)
    par = [torch.from_numpy(a).to(dev) for a in (sv, qv, rv, gv, uv, av)]
    b = bufs(M * T)
    timed(lambda: models_device(*par, M, -2.0, 4.0 / T, T, *b, stream=st), M * T, "C5_models_1e4x2000",
          {"kernel": "k_models<4>", "f5": None})
    out["C5_models_1e4x2000"]["f5"] = float((b[3] == 5).float().mean())
    out["C5_models_1e4x2000"]["noconv_flags"] = int((b[4] & 1).sum())
    # the same grid with the light-curve fit fused in: 8 doubles per model leave the GPU
    from microlensing_b200.batched import models_chi2_device
    flux_h = 2.0 * b[2][:T].cpu().numpy() + 0.5
    sig = np.full(T, 0.01)
    flux_d, w_d = torch.from_numpy(flux_h).to(dev), torch.from_numpy(1.0 / sig ** 2).to(dev)
    stats = torch.empty(M * 8, dtype=torch.float64, device=dev)
    timed(lambda: models_chi2_device(*par, M, -2.0, 4.0 / T, T, flux_d, w_d, stats, stream=st), M * T,
          "C5_models_1e4x2000_chi2", {"kernel": "k_models_chi2<4> (fit fused: 8 doubles per model out)"})
    t0 = time.perf_counter()
    for _ in range(reps):
        r = mb.models_chi2(sv, qv, rv, gv, uv, av, -2.0, 4.0 / T, flux_h, sig)
    dt = (time.perf_counter() - t0) / reps
    out["C5_models_1e4x2000_chi2"]["e2e_host_api"] = {"ms": dt * 1e3, "mags_per_s": M * T / dt,
                                                      "h2d_bytes": int(8 * (6 * M + 2 * T)), "d2h_bytes": int(64 * M),
                                                      "best_model": int(np.argmin(r["chi2"]))}
    return out


def bench_models(args):
    """--config C5: the 1e4 x 2000 model grid, models sharded over the ranks, results gathered to rank 0
    by peer stores (magnifications) or as fused fit statistics.  A parity-test configuration of
    BASELINE.json, timed for information (the headline is the C4 map)."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ["MLM_DEVICE"] = str(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    import microlensing_b200 as mb
    from microlensing_b200 import _capi
    from microlensing_b200.sharding import ShardedModels, shard_range
    c = CONFIGS["C5"]
    rng = np.random.default_rng(c["seed"])
    M, T = c["M"], c["T"]
    par = (np.exp(rng.uniform(np.log(0.5), np.log(2.), M)), np.exp(rng.uniform(np.log(1e-4), np.log(1.), M)),
           np.exp(rng.uniform(np.log(1e-4), np.log(1e-2), M)), np.full(M, 0.5), rng.uniform(-.5, .5, M),
           rng.uniform(0, 2 * np.pi, M))
    dev = torch.device("cuda", local)
    ctx = _capi.context(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(obj):
        for _ in range(args.warmup):
            obj.compute()
        barrier()
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            obj.compute()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), ctx.launch_count() - l0

    sm = ShardedModels(par, -2.0, 4.0 / T, T, rank, world, device=dev)
    ms, launches = timed(sm)
    flux = None
    if rank == 0:
        flux = 2.0 * sm.full["A4"][0].cpu().numpy() + 0.5
    box = [flux]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    flux = box[0]
    sm.close()
    sf = ShardedModels(par, -2.0, 4.0 / T, T, rank, world, device=dev, fit=(flux, np.full(T, 0.01)))
    ms_fit, _ = timed(sf)
    best = int(np.argmin(sf.result_host()["chi2"])) if rank == 0 else None
    sf.close()
    # end to end through the host API on this rank's models (26 B per epoch back over PCIe)
    lo, hi = shard_range(M, rank, world)
    sl = [a[lo:hi] for a in par]
    out = mb.alloc_outputs((hi - lo, T), device=local, pinned=True)
    mb.magnification_models(*sl, -2.0, 4.0 / T, T, out=out, device=local)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        mb.magnification_models(*sl, -2.0, 4.0 / T, T, out=out, device=local)
    te = torch.tensor([(time.perf_counter() - t0) / 3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    if rank == 0:
        n = M * T
        print(json.dumps({
            "metric": "A4 magnifications/sec (FP64, model grid 1e4 x 2000)", "value": n / (ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C5: batched fit grid {M} models x {T} epochs, rng seed {c['seed']}, log-uniform "
                                   "s in [0.5,2], q in [1e-4,1], rho in [1e-4,1e-2], Gamma 0.5; models sharded "
                                   f"contiguously over {world} rank(s), results stored into rank 0 by the kernels",
                       "positions": n, "order": 4},
            "fused_fit": {"value": n / (ms_fit * 1e-3), "ms_per_step": ms_fit, "bytes_to_rank0_per_model": 64,
                          "best_model": best},
            "e2e": {"value": n / float(te[0]), "unit": UNIT, "h2d_bytes_per_step": int(48 * M),
                    "d2h_bytes_per_step": int(26 * n), "ms_per_step": float(te[0]) * 1e3,
                    "api": "microlensing_b200.magnification_models"},
            "gpu_launches": int(launches)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _executed_profile():
    """FP64 instructions per magnification actually executed by k_map<4> on the headline map,
    counted by ncu (tools/ncu_summary.py writes this file next to the committed profile)."""
    try:
        with open(os.path.join(ROOT, "profiles", "k_map_executed.json")) as f:
            return json.load(f)
    except Exception:
        return None


def _hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0     # fallback stated in B200_PROFILING.md


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C4", choices=list(CONFIGS))
    ap.add_argument("--nx", type=int, default=0, help="override map size (debug only; invalidates the headline)")
    ap.add_argument("--chunks", type=int, default=0, help="row chunks per rank (0: 1, or 4 with --gather nccl)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"], help="N>1: how the shards reach rank 0")
    ap.add_argument("--layout", default=None, choices=["block", "cyclic"],
                    help="N>1: contiguous row blocks or strips dealt cyclically (default: cyclic with p2p)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-other", action="store_true", help="skip the kernel timings of C2/C3/C5")
    ap.add_argument("--cpu-per-core", type=int, default=6000)
    ap.add_argument("--ref-per-core", type=int, default=6000)
    args = ap.parse_args()
    if args.impl == "reference":
        return bench_reference(args)
    return bench_b200(args)


if __name__ == "__main__":
    sys.exit(main())
