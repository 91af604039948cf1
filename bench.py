#!/usr/bin/env python
"""Benchmark of the binary-lens A0/A2/A4 hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C4|C3|C5]

One "step" = one pass of the hot path over the whole synthetic workload: for the headline configuration C4 an
8192 x 8192 magnification map (s=1.0, q=0.5, rho=1e-3, Gamma=0.5, centre-of-mass window centre (-0.08, 0), half-width
0.8), strips of rows dealt to the N ranks (strong scaling), followed for N>1 by the gather of all five result arrays
(26 B/position) into rank 0's HBM.

`value`   : whole-job A4 magnifications/s with everything resident on the device (zeta is generated in the kernel,
            results stay in HBM -- for N>1 in rank 0's HBM), CUDA-event timed, max over ranks.
`e2e`     : the same metric through the public host API: N=1 `magnification_map` (host buffers, D2H of all results
            inside the call), N>1 `magnification_map_sharded` (every rank's D2H lands in ONE shared host map).  The
            measured D2H ceiling of the box for the same bytes is reported next to it.
`roofline`: FP64 pipe.  achieved = FP64 flop the kernel EXECUTES per magnification x rate (per GPU); the per-position
            executed work is measured in the run: event counters of the counting build (libmlmultipoles_stats.so,
            same sources with -DMLM_STATS) on this workload x FP64 instructions per event (profiles/k_map_phase_costs.json,
            from the per-phase ncu attribution of the kernels with the recorded source hash; flagged `stale` when the
            sources changed since).  peak = DFMA-chain microbenchmark measured live on this GPU.  The ALGORITHMIC work
            of SURVEY.md §8(d) (the reference's formulation) is reported as `algorithmic_equivalent`.
`parity`  : seeded samples of C2, C3, C4, C5 computed by the product path vs the oracle (image counts, 1e-9, flags,
            40-digit adjudication of the disputes).
`cpu_baseline`: the oracle's per-position replay of the reference (NumPy, np.roots) on all host cores, on a bounded
            seeded sample of the same grid.
`--impl reference` times only that CPU path (the reference is pure Python; the oracle port is what can travel to the
GPU box).
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
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "A4 magnifications/sec (FP64, 8192^2 map)"
UNIT = "mags/s"
NOMINAL_FP64_TFLOPS = 37.2        # 148 SM x 64 FMA lanes x 2 x 1.965 GHz
KERNEL_PARAM_BYTES = 640          # upper bound of one k_map parameter block (LensTable 288 + Deal 152 + scalars/pointers)

CONFIGS = {
    # name: dict(kind, s, q, rho, Gamma, ...)   windows: SURVEY.md §8(d)
    "C4": dict(kind="map", s=1.0, q=0.5, rho=1e-3, Gamma=0.5, cx=-0.08, cy=0.0, half=0.80, nx=8192, ny=8192),
    "C3": dict(kind="map", s=0.8, q=0.1, rho=1e-3, Gamma=0.5, cx=0.11, cy=0.0, half=0.75, nx=4096, ny=4096),
    "C2": dict(kind="lightcurve", s=1.0, q=1e-3, rho=1e-3, Gamma=0.5, u0=0.01, alpha=0.35, tau0=-2.0, tau1=2.0,
               T=1000000),
    "C5": dict(kind="models", M=10000, T=2000, seed=20240517),
}


def work_per_mag(f5: float) -> float:
    """Algorithmic FP64 flop per A4 magnification in the reference's formulation."""
    return 12106.0 + 4068.0 * f5


def map_window(c):
    nx, ny = c["nx"], c["ny"]
    dx, dy = 2 * c["half"] / nx, 2 * c["half"] / ny
    return c["cx"] - c["half"], c["cy"] - c["half"], dx, dy


def c5_params():
    c = CONFIGS["C5"]
    rng = np.random.default_rng(c["seed"])
    M = c["M"]
    return (np.exp(rng.uniform(np.log(0.5), np.log(2.), M)), np.exp(rng.uniform(np.log(1e-4), np.log(1.), M)),
            np.exp(rng.uniform(np.log(1e-4), np.log(1e-2), M)), np.full(M, 0.5), rng.uniform(-.5, .5, M),
            rng.uniform(0, 2 * np.pi, M))


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
    from microlensing_b200.batched import lightcurve_coordinates
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


def run_cpu_baseline(cfg, per_core=20000, cores=None):
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
    cfg = CONFIGS[args.config if CONFIGS[args.config]["kind"] == "map" else "C4"]
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
# parity blocks (product path vs oracle on seeded samples)
# ------------------------------------------------------------------------------------------
def parity_block(got_n, got_A, got_flags, ref, par, zeta, what, max_adjudicate=96):
    """ref = (nimg, B0, B2, B4, minJ, residuals) of the oracle on the same positions.  Rules: tests/parity.py."""
    import parity
    from oracle.truth_mp import truth_point
    nb, B0, B2, B4, minJ, res = ref
    ref_A = np.stack([np.ravel(B0), np.ravel(B2), np.ravel(B4)], 1)
    mask = parity.mask_from_reference(ref_A, res, np.ravel(minJ)) if res is not None else None
    t0 = time.perf_counter()
    rep = parity.compare(np.ravel(got_n), got_A, np.ravel(nb), ref_A, mask,
                         truth_fn=lambda i: truth_point(*par[i], zeta[i]), got_flags=np.ravel(got_flags),
                         max_adjudicate=max_adjudicate, strict=False)
    fl = np.ravel(got_flags).astype(int)
    out = {"what": what, "n": rep["n"], "nimg_mismatch": rep["nimg_mismatch"],
           "nimg_mismatch_rate": rep["nimg_mismatch"] / max(rep["n"], 1),
           "of_which_flagged_extra": rep.get("nimg_mismatch_flagged_extra"),
           "of_which_reference_ambiguous": rep.get("nimg_mismatch_ref_ambiguous"),
           "nimg_mismatch_unaccounted": rep.get("nimg_mismatch_unaccounted"),
           "disputes_outside_rule3_mask": rep["unmasked_disputes"],
           "disputes_outside_rule3_mask_rate": rep["unmasked_disputes"] / max(rep["n"], 1),
           "rel_gt_1e-9": rep["rel_gt_tol"], "rel_gt_1e-9_unflagged_unmasked": rep.get("rel_gt_tol_unflagged_unmasked"),
           "masked_by_reference_side_rule3": rep["masked"], "flagged_any": int((fl != 0).sum()),
           "flagged_extra": rep.get("flagged_extra"), "max_rel_undisputed": rep["max_rel_clean"],
           "adjudicated": rep["adjudicated"], "adjudication_failures": len(rep["adjudication_failures"]),
           "failed_rules": rep.get("failed", []), "seconds": time.perf_counter() - t0}
    if "adjudicated_subset_of" in rep:
        out["adjudicated_subset_of"] = rep["adjudicated_subset_of"]
    return out


def oracle_batch(s, q, rho, G, zeta):
    from oracle import cassan_oracle as oracle
    nb, B0, B2, B4, _, _, minJ, res = oracle.batch_a4(s, q, rho, G, zeta, return_extra=True)
    return nb, B0, B2, B4, minJ, res


def parity_samples(mb, local, n_sample=20000):
    """C2, C3, C5 through the product path (full-size C2/C3, a seeded subset of the C5 models), seeded samples vs oracle."""
    out = {}
    rng = np.random.default_rng(777)
    # C3: the full 4096^2 map through the host API, sampled pixels
    c = CONFIGS["C3"]
    x0, y0, dx, dy = map_window(c)
    m = mb.magnification_map(c["s"], c["q"], c["rho"], c["Gamma"], x0, y0, dx, dy, c["nx"], c["ny"], device=local)
    idx, zeta = sample_positions(c, n_sample, seed=31)
    got_A = np.stack([m[k].ravel()[idx] for k in ("A0", "A2", "A4")], 1)
    par = [(c["s"], c["q"], c["rho"], c["Gamma"])] * len(idx)
    out["C3"] = parity_block(m["nimg"].ravel()[idx], got_A, m["flags"].ravel()[idx],
                             oracle_batch(c["s"], c["q"], c["rho"], c["Gamma"], zeta), par, zeta,
                             f"{len(idx)} seeded pixels of the full 4096^2 map (magnification_map, k_map<4>)")
    del m
    # C2: the full 1e6-epoch light curve through both entry points
    c = CONFIGS["C2"]
    T = c["T"]
    dtau = (c["tau1"] - c["tau0"]) / T
    idx, zeta = sample_positions(c, n_sample, seed=32)
    ref = oracle_batch(c["s"], c["q"], c["rho"], c["Gamma"], zeta)
    par = [(c["s"], c["q"], c["rho"], c["Gamma"])] * len(idx)
    mm = mb.magnification_models([c["s"]], [c["q"]], [c["rho"]], [c["Gamma"]], [c["u0"]], [c["alpha"]], c["tau0"], dtau, T,
                                 device=local)
    out["C2_models"] = parity_block(mm["nimg"][0][idx], np.stack([mm[k][0][idx] for k in ("A0", "A2", "A4")], 1),
                                    mm["flags"][0][idx], ref, par, zeta,
                                    f"{len(idx)} seeded epochs of the full 1e6-epoch light curve (magnification_models, k_models<4>)")
    zfull = mb.lightcurve_coordinates(c["s"], c["q"], c["u0"], c["alpha"], c["tau0"], dtau, T)
    pp = mb.magnification_points(c["s"], c["q"], c["rho"], c["Gamma"], zfull, device=local)
    out["C2_points"] = parity_block(pp["nimg"][idx], np.stack([pp[k][idx] for k in ("A0", "A2", "A4")], 1), pp["flags"][idx],
                                    ref, par, zeta, "same epochs through magnification_points (k_points<4>)")
    del mm, pp
    # C5: 24 seeded models of the grid, all 2000 epochs each
    c = CONFIGS["C5"]
    sv, qv, rv, gv, uv, av = c5_params()
    pick = np.sort(rng.choice(c["M"], 24, replace=False))
    T = c["T"]
    r = mb.magnification_models(sv[pick], qv[pick], rv[pick], gv[pick], uv[pick], av[pick], -2.0, 4.0 / T, T, device=local)
    refs, zs, par = [], [], []
    for j, mi in enumerate(pick):
        z = mb.lightcurve_coordinates(sv[mi], qv[mi], uv[mi], av[mi], -2.0, 4.0 / T, T)
        refs.append(oracle_batch(sv[mi], qv[mi], rv[mi], gv[mi], z))
        zs.append(z)
        par += [(sv[mi], qv[mi], rv[mi], gv[mi])] * T
    ref = tuple(np.concatenate([np.asarray(x[i]) for x in refs]) for i in range(6))
    out["C5"] = parity_block(r["nimg"].ravel(), np.stack([r[k].ravel() for k in ("A0", "A2", "A4")], 1), r["flags"].ravel(),
                             ref, par, np.concatenate(zs),
                             f"24 seeded models x {T} epochs of the 1e4-model grid (magnification_models, k_models<4>); "
                             f"q of the models: {', '.join(f'{x:.1e}' for x in qv[pick])}")
    return out


# ------------------------------------------------------------------------------------------
# executed work of the map kernel, measured in the run
# ------------------------------------------------------------------------------------------
def measure_executed(local, cfg, strip):
    """Event counts of k_map<4> on this workload (counting build) x FP64 instructions per event (calibration file)."""
    import torch
    from microlensing_b200 import _capi
    from microlensing_b200.build import source_hash
    try:
        with open(os.path.join(ROOT, "profiles", "k_map_phase_costs.json")) as f:
            cal = json.load(f)
    except Exception as e:                              # noqa: BLE001
        return {"error": f"no calibration file: {e}"}
    sctx = _capi.context(local, stats=True)
    nx, ny = cfg["nx"], cfg["ny"]
    x0, y0, dx, dy = map_window(cfg)
    n = nx * ny
    dev = torch.device("cuda", local)
    bufs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)] + \
           [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
    sctx.stats_read(reset=True)
    sctx.check(sctx.lib.mlm_map(sctx.handle, cfg["s"], cfg["q"], cfg["rho"], cfg["Gamma"], x0, y0, dx, dy, nx, 0, ny, 4,
                                int(strip), *[_capi.ptr(b) for b in bufs], _capi.ptr(torch.cuda.current_stream().cuda_stream)),
               "mlm_map (counting build)")
    ev = sctx.stats_read(reset=True)
    npos = ev["positions"][0]
    per_pos = {k: v[0] / npos for k, v in ev.items()}
    lanes = {k: (v[0] / v[1] if v[1] else None) for k, v in ev.items()}
    inst = flop = 0.0
    phases = {}
    for e, c in cal["cost_per_event"].items():
        cnt = per_pos.get(e, 0.0)
        phases[e] = {"events_per_position": cnt, "fp64_inst_per_event": c["fp64_inst"], "fp64_inst_per_position": cnt * c["fp64_inst"],
                     "active_lanes": lanes.get(e)}
        inst += cnt * c["fp64_inst"]
        flop += cnt * c["flop"]
    cur = source_hash()
    return {"fp64_inst_per_mag": inst, "flop_per_mag": flop, "phases": phases, "positions_counted": npos,
            "calibration": {"file": "profiles/k_map_phase_costs.json", "source_hash": cal.get("source_hash"),
                            "ncu_fp64_inst_per_position": cal.get("ncu_fp64_inst_per_position"),
                            "ncu_fp64_pipe_pct": cal.get("ncu_fp64_pipe_pct"), "profile": cal.get("profile"),
                            "dram_bytes_per_launch": cal.get("dram_bytes_per_launch")},
            "source_hash_now": cur, "stale": cur != cal.get("source_hash"),
            "avg_active_lanes_fp64": cal.get("avg_active_lanes_fp64"),
            "source": "event counters of libmlmultipoles_stats.so (-DMLM_STATS) on this run's workload x FP64 thread "
                      "instructions per event from the per-phase ncu attribution (tools/ncu_phase_table.py, "
                      "tools/calibrate_phase_costs.py)"}


def _hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0     # fallback stated in B200_PROFILING.md


def d2h_ceiling(ctx, dev, nbytes, barrier, world, reps=3):
    """Plain cudaMemcpyAsync of this rank's share of the result bytes, HBM -> page-locked host memory, all ranks at
    once: what the box allows for the e2e leg (max over ranks, best of reps)."""
    import torch
    import torch.distributed as dist
    src = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
    dst = ctx.pinned_empty(max(nbytes, 1), np.uint8)
    st = torch.cuda.current_stream().cuda_stream
    best = None
    for r in range(reps + 1):
        barrier()
        t0 = time.perf_counter()
        ctx.memcpy_async(dst.ctypes.data, src.data_ptr(), nbytes, st)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if r > 0:
            best = float(t[0]) if best is None else min(best, float(t[0]))
    del src, dst
    return best


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
    if cfg["kind"] != "map":
        raise SystemExit("bench.py times map configurations (C4 headline, C3) and --config C5; C2 is a parity/other_configs case")
    if args.nx:
        cfg = dict(cfg, nx=args.nx, ny=args.nx)
    base = None
    if rank == 0 and world == 1 and not args.no_cpu:
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
    from microlensing_b200.batched import map_device
    from microlensing_b200.sharding import ShardedMap, SharedHostMap, magnification_map_sharded

    nx, ny = cfg["nx"], cfg["ny"]
    x0, y0, dx, dy = map_window(cfg)
    s, q, rho, G = cfg["s"], cfg["q"], cfg["rho"], cfg["Gamma"]
    ctx = _capi.context(local)
    dev = torch.device("cuda", local)
    n_total = nx * ny
    sm = ShardedMap(nx, ny, rank, world, device=dev, nchunks=args.chunks or None,
                    gather=(args.gather if world > 1 else "none"), layout=args.layout if world > 1 else "block",
                    root_share=args.root_share, align=args.strip or None)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        """CUDA events on the launching stream around `steps` calls, barrier + synchronize on both sides, max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), ctx.launch_count() - l0

    def wall(fn, steps, warmup=1):
        """host wall clock around `steps` blocking calls (host APIs), barrier on both sides, max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        t = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t[0])

    sampler = ClockSampler(local) if rank == 0 else None
    # ---- device-resident, whole job (kernel + gather to rank 0 for N>1) --------------------------
    for _ in range(args.warmup):
        sm.compute(s, q, rho, G, x0, y0, dx, dy, order=4)
    barrier()
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ms, launches = timed(lambda: sm.compute(s, q, rho, G, x0, y0, dx, dy, order=4), args.steps, 0)
    # ---- kernels only: same strips, results stay in each rank's own HBM (nothing crosses NVLink) --
    ms_kernel, _ = timed(lambda: sm.compute(s, q, rho, G, x0, y0, dx, dy, order=4, local_only=True), args.steps, 1)
    clocks = sampler.stop() if sampler else None
    value = n_total / (ms * 1e-3)

    # ---- N>1: rank 0 recomputes seeded strips on its own GPU and compares bytes with the gathered map --
    sharded_vs_single = None
    if world > 1:
        sm.compute(s, q, rho, G, x0, y0, dx, dy, order=4)
        barrier()
        if rank == 0:
            fv = sm.full_views()
            L = sm.align
            nstrips = -(-ny // L)
            ks = sorted(set([0, nstrips - 1] + [int(k) for k in np.random.default_rng(2024).integers(0, nstrips, 22)]))
            bad = 0
            st = torch.cuda.current_stream().cuda_stream
            for k in ks:
                rows = min(L, ny - k * L)
                b = [torch.empty(rows * nx, dtype=torch.float64, device=dev) for _ in range(3)] + \
                    [torch.empty(rows * nx, dtype=torch.uint8, device=dev) for _ in range(2)]
                map_device(s, q, rho, G, x0, y0, dx, dy, nx, k * L, rows, *b, stream=st, device=local, strip=L)
                torch.cuda.synchronize()
                for key, t_ in zip(("A0", "A2", "A4", "nimg", "flags"), b):
                    got = fv[key][k * L:k * L + rows].reshape(-1).view(torch.uint8)
                    bad += int((got != t_.view(torch.uint8)).sum())
            sharded_vs_single = {"strips": len(ks), "rows": int(sum(min(L, ny - k * L) for k in ks)), "strip_rows": L,
                                 "bytes_compared": int(sum(min(L, ny - k * L) for k in ks) * nx * 26),
                                 "mismatched_bytes": bad,
                                 "how": "rank 0 recomputed these strips alone (mlm_map, same strip length) and compared all "
                                        "five arrays byte for byte with the map gathered from the ranks"}
        barrier()

    # ---- same job storing / gathering only what an A4 map needs (A4, nimg, flags: 10 B/position) ------
    sm_a4 = ShardedMap(nx, ny, rank, world, device=dev, nchunks=args.chunks or None,
                       gather=(args.gather if world > 1 else "none"), outputs=("A4", "nimg", "flags"),
                       layout=args.layout if world > 1 else "block", align=args.strip or None)
    ms_a4, _ = timed(lambda: sm_a4.compute(s, q, rho, G, x0, y0, dx, dy, order=4), args.steps, 2)
    a4_share = sm_a4.root_share
    sm_a4.close()
    del sm_a4

    # image-count histogram of this run (for the algorithmic work per magnification)
    if world > 1 and sm.gather in ("p2p", "bulk"):
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
    e2e_steps = max(1, args.e2e_steps or args.steps)
    if world == 1:
        out = mb.alloc_outputs((ny, nx), device=local, pinned=True)
        l0 = ctx.launch_count()
        t_e2e = wall(lambda: mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, nx, ny, out=out, device=local), e2e_steps)
        e2e_launches = (ctx.launch_count() - l0) // (e2e_steps + 1)
        out4 = mb.alloc_outputs((ny, nx), device=local, pinned=True, outputs=("A4", "nimg", "flags"))
        t_e2e4 = wall(lambda: mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, nx, ny, out=out4, device=local), e2e_steps)
        del out4
        api = "microlensing_b200.magnification_map (host buffers, pinned; D2H inside the call)"
        e2e_strip = None
    else:
        sh = SharedHostMap(nx, ny, rank, world, device=local)
        l0 = ctx.launch_count()
        t_e2e = wall(lambda: magnification_map_sharded(s, q, rho, G, x0, y0, dx, dy, nx, ny, rank, world, device=local, shared=sh),
                     e2e_steps)
        e2e_launches = (ctx.launch_count() - l0) // (e2e_steps + 1)
        e2e_strip = sh.strip
        # the assembled host map on rank 0 equals the device-resident gathered map (same strip length) byte for byte
        e2e_check = None
        if rank == 0 and sm.align == sh.strip:
            fv = sm.full_views()
            rows = np.random.default_rng(5).integers(0, ny, 64)
            badb = 0
            for key in ("A0", "A2", "A4", "nimg", "flags"):
                dev_rows = fv[key][torch.as_tensor(rows, device=dev)].cpu().numpy()
                badb += int((dev_rows.view(np.uint8) != np.ascontiguousarray(sh.arrays[key][rows]).view(np.uint8)).sum())
            e2e_check = {"rows": 64, "mismatched_bytes_vs_device_gather": badb}
        sh.close()
        sh4 = SharedHostMap(nx, ny, rank, world, device=local, outputs=("A4", "nimg", "flags"))
        t_e2e4 = wall(lambda: magnification_map_sharded(s, q, rho, G, x0, y0, dx, dy, nx, ny, rank, world, device=local,
                                                        shared=sh4, outputs=("A4", "nimg", "flags")), e2e_steps)
        sh4.close()
        api = "microlensing_b200.sharding.magnification_map_sharded (every rank's D2H lands in one shared, page-locked host map)"
    share_bytes = 26 * (n_total // world)
    t_ceiling = d2h_ceiling(ctx, dev, share_bytes, barrier, world)
    # e2e leg with real inputs: C2 light curve positions through magnification_points (16 B in, 26 B out per position)
    e2e_points = None
    if world == 1:
        c2 = CONFIGS["C2"]
        zl = mb.lightcurve_coordinates(c2["s"], c2["q"], c2["u0"], c2["alpha"], c2["tau0"], (c2["tau1"] - c2["tau0"]) / c2["T"], c2["T"])
        zp = ctx.pinned_empty(zl.shape, np.complex128)
        zp[...] = zl
        outp = mb.alloc_outputs(zl.shape, device=local, pinned=True)
        tp = wall(lambda: mb.magnification_points(c2["s"], c2["q"], c2["rho"], c2["Gamma"], zp, out=outp, device=local), e2e_steps)
        e2e_points = {"value": c2["T"] / tp, "unit": UNIT, "ms_per_step": tp * 1e3, "steps": e2e_steps,
                      "h2d_bytes_per_step": int(16 * c2["T"]), "d2h_bytes_per_step": int(26 * c2["T"]),
                      "api": "microlensing_b200.magnification_points on the C2 light curve (1e6 positions from pinned host memory)"}
        del zp, outp

    # ---- C5 (model grid) at N>1: sharded, full outputs and fused fit, with its own sharded-vs-single check --
    other = None
    if not args.no_other:
        other = other_configs(mb, dev, rank, world, barrier)

    if rank != 0:
        if world > 1:
            sm.close()
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (k_map<4>) -----------------------------------------------------
    peak_meas = ctx.fp64_peak_tflops()
    W = work_per_mag(f5)
    kernel_rate = n_total / (ms_kernel * 1e-3)            # whole job, kernels only
    # the launch that sets the kernel-only time is the one of the rank with the largest share of the strips
    max_share = max(len(x) for x in sm.sel_all) / sm.period if world > 1 else 1.0
    launch_rate = max_share * n_total / (ms_kernel * 1e-3)  # positions/s of that GPU
    info = ctx.kernel_info()
    ex = measure_executed(local, cfg, sm.align)
    roofline = {"bound": "fp64", "unit": "TFLOP/s", "peak": peak_meas, "kernel": "k_map<4>",
                "peak_source": "DFMA-chain microbenchmark (mlm_fp64_peak) measured live on this GPU; MEASURED_PEAKS.json has no FP64 entry",
                "peak_nominal": NOMINAL_FP64_TFLOPS}
    if ex and "error" not in ex:
        ex_tflops = launch_rate * ex["flop_per_mag"] / 1e12
        inst_rate = launch_rate * ex["fp64_inst_per_mag"]
        nominal_issue = 148 * 64 * 1.965e9                 # FP64 thread-instructions/s: 64 lanes per SM
        probe = ctx.fp64_probe(2, 4)                       # 3-distinct-operand DFMA/DMUL/DADD mix, 16 warps/SM
        lanes = ex.get("avg_active_lanes_fp64") or 32.0
        clk = (clocks or {}).get("sm_mhz") or 1965.0
        roofline.update({
            "achieved": ex_tflops, "frac": ex_tflops / peak_meas, "frac_of_nominal": ex_tflops / NOMINAL_FP64_TFLOPS,
            "executed": dict(ex, tflops=ex_tflops, fp64_inst_per_s=inst_rate,
                             frac_of_nominal_issue_rate=inst_rate / nominal_issue,
                             operand_limited_issue_rate=probe, frac_of_operand_limited_rate=inst_rate / probe,
                             fp64_pipe_pct_model=100.0 * inst_rate / lanes * 32.0 / (148 * 64 * clk * 1e6)),
        })
    else:
        roofline.update({"achieved": None, "frac": None, "executed": ex})
    alg = launch_rate * W / 1e12
    roofline.update({
        "traffic": ((ex or {}).get("calibration") or {}).get("dram_bytes_per_launch"),
        "algorithmic_equivalent": {"flop_per_mag": W, "f5": f5, "tflops_equiv": alg, "ratio_to_peak": alg / peak_meas,
                                   "note": "SURVEY.md §8(d): flop the REFERENCE's formulation needs per magnification (8 Aberth sweeps "
                                           "nominal + the 20-factor a_kl recursion) x the measured rate; the kernel reaches the same "
                                           "numbers with several times less arithmetic (root/image tracking, closed operator form), so "
                                           "this is not a machine fraction"},
        "limiter": "register-file operand bandwidth of the FP64 pipe: a DFMA with three distinct register operands issues "
                   "every 3 cycles, not 2 (measured: operand_limited_issue_rate), and every other instruction (the FP32 "
                   "watch, addressing, selects) takes read slots too -- frac counts FP64 flop only, "
                   "frac_of_operand_limited_rate is the machine view (DESIGN.md §5)",
        "positions_per_launch": max_share * n_total, "kernel_ms_per_step": ms_kernel, "algorithmic_bytes_per_mag": 26,
        "launch": "the rank with the largest share of the strips (rank 0 of a weighted deal)" if world > 1 else "the whole map",
        "hbm_frac": (launch_rate * 26 / 1e9) / _hbm_peak(),
        "registers_per_thread": info["k_map4_regs"], "local_bytes_per_thread": info["k_map4_local_bytes"]})
    # ---- CPU baseline + parity on seeded samples ---------------------------------------------------------
    base_out, parity_out = None, None
    if base is not None:
        # the C4 sample against the map the e2e leg just produced (the product path being timed)
        got_A = np.stack([out[k].ravel()[idx] for k in ("A0", "A2", "A4")], 1)
        par = [(s, q, rho, G)] * len(idx)
        ref_full = oracle_batch(s, q, rho, G, zeta)       # vectorised replay: minJ / residuals for rule 3's mask
        replay_equals_batch = bool(np.array_equal(ref_full[0], ref[:, 0].astype(np.uint8)))
        parity_out = {"C4": parity_block(out["nimg"].ravel()[idx], got_A, out["flags"].ravel()[idx],
                                         (ref[:, 0], ref[:, 1], ref[:, 2], ref[:, 3], ref_full[4], ref_full[5]), par, zeta,
                                         f"{len(idx)} seeded pixels of the full 8192^2 map of the e2e leg (magnification_map, k_map<4>) "
                                         "vs the timed per-position replay")}
        parity_out["C4"]["oracle_replay_counts_equal_vectorised_oracle"] = replay_equals_batch
        if not args.no_other:
            parity_out.update(parity_samples(mb, local))
        base_out = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        base_out["full_workload_extrapolated_hours"] = n_total / base["value"] / 3600.0

    gather_txt = {"p2p": "fused: the map kernel stores its strips straight into rank 0's HBM over NVLink (CUDA IPC peer "
                         "mapping), 1-element NCCL all-reduce as completion fence",
                  "bulk": "ranks > 0 compute into local HBM chunk by chunk and push the strips into rank 0's HBM with 2-D "
                          "bulk peer copies on a copy stream",
                  "nccl": "NCCL gather to rank 0, chunk-pipelined", "none": "none"}[sm.gather]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(cfg, args, extra={
            "parallelism": f"strips of {sm.align} rows dealt to {world} rank(s), layout {sm.layout}"
                           + (f", rank 0 owns {len(sm.sel_all[0])} of every {sm.period} strips (share {len(sm.sel_all[0]) / sm.period:.3f}): "
                              "its results do not cross NVLink" if world > 1 else ""),
            "gather": gather_txt, "bytes_gathered_per_position": sm.bytes_per_position if world > 1 else 0}),
        "roofline": roofline, "cpu_baseline": base_out, "parity": parity_out,
        "e2e": {"value": n_total / t_e2e, "unit": UNIT,
                "h2d_bytes_per_step": int(KERNEL_PARAM_BYTES * max(e2e_launches, 1)),
                "d2h_bytes_per_step": int(26 * n_total), "ms_per_step": t_e2e * 1e3, "steps": e2e_steps, "api": api,
                "kernel_launches_per_step": int(e2e_launches), "strip_rows": e2e_strip,
                "h2d_note": "the map takes no input arrays: zeta is generated in the kernel; the only host->device bytes are "
                            f"the kernel parameter blocks (<= {KERNEL_PARAM_BYTES} B per launch); e2e_points is the leg with real inputs",
                "d2h_ceiling": {"ms": t_ceiling * 1e3, "aggregate_GBs": 26 * (n_total // world) * world / t_ceiling / 1e9,
                                "value_at_ceiling": n_total / t_ceiling, "frac_of_ceiling": t_ceiling / t_e2e,
                                "how": f"plain cudaMemcpyAsync of each rank's {share_bytes} result bytes HBM -> pinned host, "
                                       f"{world} rank(s) at once, measured in this run"}},
        "e2e_points": e2e_points,
        "gpu_launches": int(launches), "clocks": clocks,
        "kernel_only": {"value": kernel_rate, "ms_per_step": ms_kernel,
                        "note": "same strips per rank, results stored in each rank's OWN HBM (no peer traffic, no fence): "
                                "the sharded map left resident where it was computed"},
        "a4_outputs_only": {"note": "same job, same kernel arithmetic, but only A4 + nimg + flags (10 B/position instead "
                                    "of 26) are stored, gathered to rank 0 (value) or copied to the host (e2e); "
                                    "informational -- the headline `value` and `e2e` move all five arrays",
                            "value": n_total / (ms_a4 * 1e-3), "ms_per_step": ms_a4, "root_share": a4_share,
                            "e2e": {"value": n_total / t_e2e4, "ms_per_step": t_e2e4 * 1e3,
                                    "d2h_bytes_per_step": int(10 * n_total)}},
        "flagged_positions": flag_any,
    }
    if world > 1:
        line["sharded_vs_single"] = sharded_vs_single
        line["e2e"]["host_map_vs_device_gather"] = e2e_check
        line["gather_limits"] = {"root_share": sm.root_share, "ingress_GBs_achieved": (1.0 - len(sm.sel_all[0]) / sm.period) * 26 * n_total / (ms * 1e-3) / 1e9,
                                 "note": "all results of the peers enter ONE GPU: measured ceilings on this pool 790 GB/s (bulk peer copies), "
                                         "~660 GB/s (kernel peer stores), profiles/r2_ceilings_n8.json"}
    if other is not None:
        line["other_configs"] = other
    print(json.dumps(line))
    if world > 1:
        sm.close()
        dist.barrier()
        dist.destroy_process_group()
    return 0


def time_models_sharded(par, T, rank, world, dev, steps, warmup, barrier, fit=None, outputs=None, check_models=12):
    """C5 through ShardedModels: CUDA-event time per step (max over ranks) + rank 0's byte comparison of seeded
    models with a single-GPU recomputation."""
    import torch
    import torch.distributed as dist
    from microlensing_b200 import _capi
    from microlensing_b200.batched import models_device, models_chi2_device
    from microlensing_b200.sharding import ShardedModels
    ctx = _capi.context(dev.index)
    sm = ShardedModels(par, -2.0, 4.0 / T, T, rank, world, device=dev, fit=fit, outputs=outputs)
    for _ in range(warmup):
        sm.compute()
    barrier()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sm.compute()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res = {"ms": float(t[0]), "launches_per_step": (ctx.launch_count() - l0) // steps, "segment": sm.segment}
    extra = {}
    if rank == 0:
        M = len(par[0])
        pick = np.sort(np.random.default_rng(99).choice(M, check_models, replace=False))
        st = torch.cuda.current_stream().cuda_stream
        bad = 0
        p1 = [torch.from_numpy(np.ascontiguousarray(a[pick])).to(dev) for a in par]
        if fit is None:
            b = [torch.empty(len(pick) * T, dtype=torch.float64, device=dev) for _ in range(3)] + \
                [torch.empty(len(pick) * T, dtype=torch.uint8, device=dev) for _ in range(2)]
            models_device(*p1, len(pick), -2.0, 4.0 / T, T, *b, stream=st, device=dev.index, segment=sm.segment)
            torch.cuda.synchronize()
            for key, t_ in zip(("A0", "A2", "A4", "nimg", "flags"), b):
                if key in sm.outputs:
                    got = sm.full[key][torch.as_tensor(pick, device=dev)].reshape(-1).view(torch.uint8)
                    bad += int((got != t_.view(torch.uint8)).sum())
            extra["f5"] = float((sm.full["nimg"] == 5).float().mean()) if "nimg" in sm.outputs else None
            extra["flux_of_model_0"] = (2.0 * sm.full["A4"][0].cpu().numpy() + 0.5) if "A4" in sm.outputs else None
        else:
            stats = torch.empty(len(pick) * 8, dtype=torch.float64, device=dev)
            models_chi2_device(*p1, len(pick), -2.0, 4.0 / T, T, sm.fit[0], sm.fit[1], stats, stream=st, device=dev.index,
                               segment=sm.segment)
            torch.cuda.synchronize()
            got = sm.full["stats"][torch.as_tensor(pick, device=dev)].reshape(-1).view(torch.uint8)
            bad += int((got != stats.view(torch.uint8)).sum())
            extra["best_model"] = int(torch.argmin(sm.full["stats"][:, 0]))
        res["sharded_vs_single"] = {"models": int(len(pick)), "mismatched_bytes": bad,
                                    "how": "rank 0 recomputed these models alone and compared the bytes gathered from the ranks"}
    barrier()
    sm.close()
    return res, extra


def other_configs(mb, dev, rank, world, barrier, reps=3):
    """Kernel timings of the other BASELINE.json configurations (parity-test cases, not the headline).
    N = 1: C1 latency, C3 map, C2 light curve (ordered / shuffled), C5 grid (full outputs, fused fit, host API).
    N > 1: C5 model-sharded (full outputs gathered to rank 0, fused fit), each with a sharded-vs-single byte check."""
    import torch
    out = {}
    c = CONFIGS["C5"]
    M, T = c["M"], c["T"]
    par = c5_params()
    n = M * T
    r1, ex1 = time_models_sharded(par, T, rank, world, dev, reps, 2, barrier)
    flux = ex1.get("flux_of_model_0")
    if world > 1:
        import torch.distributed as dist
        box = [flux]
        dist.broadcast_object_list(box, src=0)
        flux = box[0]
    r2, ex2 = time_models_sharded(par, T, rank, world, dev, reps, 2, barrier, fit=(flux, np.full(T, 0.01)))
    if rank == 0:
        out["C5_models_1e4x2000"] = dict(positions=n, ms=r1["ms"], mags_per_s=n / (r1["ms"] * 1e-3), f5=ex1.get("f5"),
                                         kernel="k_models<4>", segment_epochs=r1.get("segment"),
                                         sharded_vs_single=r1.get("sharded_vs_single"),
                                         gather="26 B/epoch to rank 0 (bulk peer copies of contiguous model blocks)" if world > 1 else "none",
                                         n_gpus=world)
        out["C5_models_1e4x2000_chi2"] = dict(positions=n, ms=r2["ms"], mags_per_s=n / (r2["ms"] * 1e-3),
                                              kernel="k_models_chi2<4> (fit fused: 8 doubles per model out)",
                                              best_model=ex2.get("best_model"), sharded_vs_single=r2.get("sharded_vs_single"),
                                              n_gpus=world)
    if world > 1 or rank != 0:
        return out if rank == 0 else None

    from microlensing_b200.batched import map_device, models_device, points_device, lightcurve_coordinates
    st = torch.cuda.current_stream().cuda_stream

    def bufs(k):
        return ([torch.empty(k, dtype=torch.float64, device=dev) for _ in range(3)]
                + [torch.empty(k, dtype=torch.uint8, device=dev) for _ in range(2)])

    def timed(fn, k, label, extra=None):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[label] = dict(positions=k, ms=ms, mags_per_s=k / (ms * 1e-3), **(extra or {}))

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
    k = c["nx"] * c["ny"]
    b = bufs(k)
    from microlensing_b200.sharding import auto_strip
    strip3 = auto_strip(c["nx"], c["ny"])            # a 4096^2 map is a quarter of the headline: shorter strips
    timed(lambda: map_device(c["s"], c["q"], c["rho"], c["Gamma"], x0, y0, dx, dy, c["nx"], 0, c["ny"], *b, stream=st,
                             strip=strip3), k, "C3_map_4096", {"kernel": "k_map<4>", "strip_rows": strip3, "f5": None})
    out["C3_map_4096"]["f5"] = float((b[3] == 5).float().mean())

    c = CONFIGS["C2"]
    T2 = c["T"]
    dtau = (c["tau1"] - c["tau0"]) / T2
    b = bufs(T2)
    par2 = [torch.tensor([c[key]], dtype=torch.float64, device=dev) for key in ("s", "q", "rho", "Gamma", "u0", "alpha")]
    timed(lambda: models_device(*par2, 1, c["tau0"], dtau, T2, *b, stream=st), T2, "C2_lightcurve_1e6_tracked",
          {"kernel": "k_models<4> (in-kernel trajectory, warm-started segments)"})
    zh = lightcurve_coordinates(c["s"], c["q"], c["u0"], c["alpha"], c["tau0"], dtau, T2)
    zeta = torch.from_numpy(zh).to(dev)
    timed(lambda: points_device(c["s"], c["q"], c["rho"], c["Gamma"], zeta, T2, *b, stream=st), T2,
          "C2_lightcurve_1e6_points", {"kernel": "k_points<4> (ordered list: tracked along consecutive entries)"})
    zsh = torch.from_numpy(zh[np.random.default_rng(1).permutation(T2)]).to(dev)
    timed(lambda: points_device(c["s"], c["q"], c["rho"], c["Gamma"], zsh, T2, *b, stream=st, unordered=True), T2,
          "C2_lightcurve_1e6_points_shuffled",
          {"kernel": "k_points_direct<4> + k_points_list<4> (mlm_points_unordered: every entry solved on its own from "
                     "three single-lens image guesses, root finder on the compacted rest)"})
    timed(lambda: points_device(c["s"], c["q"], c["rho"], c["Gamma"], zsh, T2, *b, stream=st), T2,
          "C2_lightcurve_1e6_points_shuffled_tracking_kernel",
          {"kernel": "k_points<4> on the same shuffled list (every entry a cold solve with the root finder, divergent lanes)"})
    # C5 through the host API with the fit fused (8 doubles per model leave the GPU)
    sv, qv, rv, gv, uv, av = par
    sig = np.full(T, 0.01)
    t0 = time.perf_counter()
    for _ in range(reps):
        r = mb.models_chi2(sv, qv, rv, gv, uv, av, -2.0, 4.0 / T, flux, sig)
    dt = (time.perf_counter() - t0) / reps
    out["C5_models_1e4x2000_chi2"]["e2e_host_api"] = {"ms": dt * 1e3, "mags_per_s": M * T / dt,
                                                      "h2d_bytes": int(8 * (6 * M + 2 * T)), "d2h_bytes": int(64 * M),
                                                      "best_model": int(np.argmin(r["chi2"]))}
    return out


def bench_models(args):
    """--config C5: the 1e4 x 2000 model grid, models sharded over the ranks, results gathered to rank 0
    (magnifications: bulk peer copies; or the fused fit statistics).  A parity-test configuration of
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
    from microlensing_b200.sharding import shard_range
    c = CONFIGS["C5"]
    M, T = c["M"], c["T"]
    par = c5_params()
    dev = torch.device("cuda", local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    r1, ex1 = time_models_sharded(par, T, rank, world, dev, args.steps, args.warmup, barrier)
    box = [ex1.get("flux_of_model_0")]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    r2, ex2 = time_models_sharded(par, T, rank, world, dev, args.steps, args.warmup, barrier, fit=(box[0], np.full(T, 0.01)))
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
            "metric": "A4 magnifications/sec (FP64, model grid 1e4 x 2000)", "value": n / (r1["ms"] * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r1["ms"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C5: batched fit grid {M} models x {T} epochs, rng seed {c['seed']}, log-uniform "
                                   "s in [0.5,2], q in [1e-4,1], rho in [1e-4,1e-2], Gamma 0.5; models sharded "
                                   f"contiguously over {world} rank(s), results gathered into rank 0's HBM",
                       "positions": n, "order": 4},
            "sharded_vs_single": r1.get("sharded_vs_single"),
            "fused_fit": {"value": n / (r2["ms"] * 1e-3), "ms_per_step": r2["ms"], "bytes_to_rank0_per_model": 64,
                          "best_model": ex2.get("best_model"), "sharded_vs_single": r2.get("sharded_vs_single")},
            "e2e": {"value": n / float(te[0]), "unit": UNIT, "h2d_bytes_per_step": int(48 * M),
                    "d2h_bytes_per_step": int(26 * n), "ms_per_step": float(te[0]) * 1e3,
                    "api": "microlensing_b200.magnification_models"},
            "gpu_launches": int(r1["launches_per_step"] * args.steps)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C4", choices=list(CONFIGS))
    ap.add_argument("--nx", type=int, default=0, help="override map size (debug only; invalidates the headline)")
    ap.add_argument("--chunks", type=int, default=0, help="row chunks per rank (0: 1, or 4 with --gather nccl/bulk)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "bulk", "nccl"], help="N>1: how the shards reach rank 0")
    ap.add_argument("--layout", default=None, choices=["block", "cyclic", "weighted"],
                    help="N>1: contiguous row blocks, strips dealt cyclically, or dealt with a larger share for the "
                         "gather root (default: weighted with p2p/bulk)")
    ap.add_argument("--root-share", type=float, default=None, help="N>1, weighted layout: share of the strips rank 0 owns")
    ap.add_argument("--strip", type=int, default=0, help="rows per strip (0: auto from the per-rank size)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the e2e legs (0: --steps)")
    ap.add_argument("--no-other", action="store_true", help="skip the other configurations (C2/C3/C5 timings and parity)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline and the parity blocks (profiling runs)")
    ap.add_argument("--cpu-per-core", type=int, default=20000)
    ap.add_argument("--ref-per-core", type=int, default=20000)
    args = ap.parse_args()
    if args.impl == "reference":
        return bench_reference(args)
    return bench_b200(args)


if __name__ == "__main__":
    sys.exit(main())
