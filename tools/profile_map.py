"""Minimal driver for ncu: the headline map kernel (k_map<4>) on a row block of the C4 map.

    python tools/profile_map.py [rows] [reps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import microlensing_b200 as mb  # noqa: E402
from microlensing_b200.batched import map_device  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
nx = 8192
s, q, rho, G = 1.0, 0.5, 1e-3, 0.5
half = 0.8
x0, y0, dx, dy = -0.08 - half, -half, 2 * half / nx, 2 * half / nx
n = rows * nx
dev = torch.device("cuda", 0)
A = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
ni = torch.empty(n, dtype=torch.uint8, device=dev)
fl = torch.empty(n, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
row0 = (8192 - rows) // 2
for r in range(reps):
    e0.record()
    map_device(s, q, rho, G, x0, y0, dx, dy, nx, row0, rows, A[0], A[1], A[2], ni, fl,
               stream=torch.cuda.current_stream().cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rep {r}: {ms:.3f} ms  {n / ms / 1e6:.1f} M mags/s  f5={(ni == 5).float().mean().item():.4f}")
