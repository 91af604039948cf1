"""Minimal driver for ncu: k_points<4> on the C2 light curve (1e6 ordered epochs: tracked) and on the same
positions shuffled (cold solve per entry).   python tools/profile_points.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import microlensing_b200 as mb
from microlensing_b200.batched import points_device
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
s, q, rho, G = 1.0, 1e-3, 1e-3, 0.5
T = 1000000
zeta = mb.lightcurve_coordinates(s, q, 0.01, 0.35, -2.0, 4.0 / T, T)
dev = torch.device("cuda", 0)
b = [torch.empty(T, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.empty(T, dtype=torch.uint8, device=dev) for _ in range(2)]
st = torch.cuda.current_stream().cuda_stream
for name, z in (("ordered", zeta), ("shuffled", zeta[np.random.default_rng(0).permutation(T)])):
    zd = torch.from_numpy(np.ascontiguousarray(z)).to(dev)
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        points_device(s, q, rho, G, zd, T, *b, stream=st)
        e1.record(); torch.cuda.synchronize()
        print(f"{name} rep {r}: {e0.elapsed_time(e1):.3f} ms  {T / e0.elapsed_time(e1) / 1e6:.2f} G mags/s")
