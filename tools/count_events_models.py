"""Event counters (counting build) of k_models<4> on the C5 grid (1e4 models x 2000 epochs): per-position figures and
active lanes per event.   python tools/count_events_models.py [M] [segment] > events.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from microlensing_b200 import _capi
import bench
M = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
seg = int(sys.argv[2]) if len(sys.argv) > 2 else 0
p = bench.c5_params()
T = bench.CONFIGS["C5"]["T"]
dev = torch.device("cuda", 0)
par = [torch.from_numpy(np.ascontiguousarray(a[:M])).to(dev) for a in p]
n = M * T
b = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
sctx = _capi.context(0, stats=True)
seg = seg or sctx.auto_segment(M, T)
sctx.stats_read(reset=True)
sctx.check(sctx.lib.mlm_models(sctx.handle, *[_capi.ptr(t) for t in par], M, -2.0, 4.0 / T, T, 4, seg, *[_capi.ptr(t) for t in b],
                               _capi.ptr(torch.cuda.current_stream().cuda_stream)), "mlm_models")
ev = sctx.stats_read(reset=True)
npos = ev["positions"][0]
print(json.dumps({"workload": "C5", "M": M, "segment": seg, "events": ev,
                  "per_position": {k: round(v[0] / npos, 4) for k, v in ev.items()},
                  "active_lanes": {k: (round(v[0] / v[1], 1) if v[1] else None) for k, v in ev.items()},
                  "f5": float((b[3] == 5).float().mean())}, indent=1))
