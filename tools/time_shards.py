"""Experiment: kernel time of row shards of the C4 map on ONE GPU (isolates tail / ordering effects of the
row-sharded launches from NVLink effects).   MLM_MAP_STRIP=.. python tools/time_shards.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import microlensing_b200 as mb
from microlensing_b200.batched import map_device
nx = 8192
s, q, rho, G = 1.0, 0.5, 1e-3, 0.5
half = 0.8
x0, y0, dx, dy = -0.08 - half, -half, 2 * half / nx, 2 * half / nx
dev = torch.device("cuda", 0)
n = nx * nx
A = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
ni = torch.empty(n, dtype=torch.uint8, device=dev); fl = torch.empty(n, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
def t(row0, rows, reps=4):
    best = 1e9
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        map_device(s, q, rho, G, x0, y0, dx, dy, nx, row0, rows, A[0], A[1], A[2], ni, fl, stream=st)
        e1.record(); torch.cuda.synchronize()
        if r: best = min(best, e0.elapsed_time(e1))
    return best
full = t(0, 8192)
print("strip", mb._capi.context().lib.mlm_map_strip(mb._capi.context().handle), "full map %.3f ms" % full)
for world in (2, 4, 8):
    rows = 8192 // world
    ts = [t(r * rows, rows) for r in range(world)]
    print(f"world {world}: ideal {full / world:.3f}  shards " + " ".join(f"{x:.3f}" for x in ts) + f"  max/ideal {max(ts) / (full / world):.3f}")
from microlensing_b200.batched import map_strips_device
def tc(phase, step, reps=4):
    best = 1e9
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        map_strips_device(s, q, rho, G, x0, y0, dx, dy, nx, 0, 8192, phase, step, A[0], A[1], A[2], ni, fl, stream=st)
        e1.record(); torch.cuda.synchronize()
        if r: best = min(best, e0.elapsed_time(e1))
    return best
for world in (2, 4, 8):
    ts = [tc(r, world) for r in range(world)]
    print(f"cyclic world {world}: ideal {full / world:.3f}  shards " + " ".join(f"{x:.3f}" for x in ts) + f"  max/ideal {max(ts) / (full / world):.3f}")
