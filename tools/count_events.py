"""Event counters of a tracking kernel on one of the bench workloads, from the counting build
(libmlmultipoles_stats.so = the product sources with -DMLM_STATS):

    python tools/count_events.py [C4|C3] [strip] > events.json

Prints {event: [lanes, warp-level occurrences]} + per-position figures.  Input of tools/calibrate_phase_costs.py."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from microlensing_b200 import _capi  # noqa: E402
from microlensing_b200.build import source_hash  # noqa: E402
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
strip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfg = bench.CONFIGS[name]
nx, ny = cfg["nx"], cfg["ny"]
x0, y0, dx, dy = bench.map_window(cfg)
dev = torch.device("cuda", 0)
n = nx * ny
b = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
sctx = _capi.context(0, stats=True)
sctx.stats_read(reset=True)
sctx.check(sctx.lib.mlm_map(sctx.handle, cfg["s"], cfg["q"], cfg["rho"], cfg["Gamma"], x0, y0, dx, dy, nx, 0, ny, 4, strip,
                            *[_capi.ptr(t) for t in b], _capi.ptr(torch.cuda.current_stream().cuda_stream)), "mlm_map")
ev = sctx.stats_read(reset=True)
npos = ev["positions"][0]
print(json.dumps({"workload": name, "strip": strip or sctx.auto_strip(nx, ny), "source_hash": source_hash(), "events": ev,
                  "per_position": {k: v[0] / npos for k, v in ev.items()},
                  "active_lanes": {k: (v[0] / v[1] if v[1] else None) for k, v in ev.items()},
                  "f5": float((b[3] == 5).float().mean())}, indent=1))
