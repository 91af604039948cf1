"""Event counters (counting build) of the point-list kernels on the C2 light curve: ordered, shuffled through the tracking
kernel (a cold solve per entry), shuffled through mlm_points_unordered (direct path), and random 2-D positions through both."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from microlensing_b200 import _capi
from microlensing_b200.batched import lightcurve_coordinates
import bench
c = bench.CONFIGS["C2"]; T = c["T"]; dtau = (c["tau1"] - c["tau0"]) / T
zh = lightcurve_coordinates(c["s"], c["q"], c["u0"], c["alpha"], c["tau0"], dtau, T)
zs = zh[np.random.default_rng(1).permutation(T)]
zr = np.random.default_rng(2).uniform(-2, 2, T) + 1j * np.random.default_rng(3).uniform(-2, 2, T)
dev = torch.device("cuda", 0)
b = [torch.empty(T, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.empty(T, dtype=torch.uint8, device=dev) for _ in range(2)]
sctx = _capi.context(0, stats=True)
st = torch.cuda.current_stream().cuda_stream
out = {}
for name, z, fn in (("ordered", zh, "mlm_points"), ("shuffled_plain", zs, "mlm_points"), ("shuffled_unordered", zs, "mlm_points_unordered"),
                    ("random_plain", zr, "mlm_points"), ("random_unordered", zr, "mlm_points_unordered")):
    zt = torch.from_numpy(z).to(dev)
    sctx.stats_read(reset=True)
    sctx.check(getattr(sctx.lib, fn)(sctx.handle, c["s"], c["q"], c["rho"], c["Gamma"], _capi.ptr(zt), T, 4, *[_capi.ptr(t) for t in b], _capi.ptr(st)), fn)
    ev = sctx.stats_read(reset=True)
    n = ev["positions"][0]
    out[name] = {"per_position": {k: round(v[0] / n, 4) for k, v in ev.items()}, "lanes": {k: (round(v[0] / v[1], 1) if v[1] else None) for k, v in ev.items()}}
    print(name, out[name]["per_position"], "\n    lanes", out[name]["lanes"], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_events_points.json"), "w"), indent=1)
