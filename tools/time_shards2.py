import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
ctx = mb._capi.context()
def t(row0, rows, reps=4):
    best = 1e9
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        map_device(s, q, rho, G, x0, y0, dx, dy, nx, row0, rows, A[0], A[1], A[2], ni, fl, stream=st)
        e1.record(); torch.cuda.synchronize()
        if r: best = min(best, e0.elapsed_time(e1))
    return best
for strip in (int(x) for x in os.environ.get("STRIPS", "64,32,16").split(",")):
    ctx.set_map_strip(strip)
    full = t(0, 8192)
    print(f"strip {strip}: full {full:.3f} | halves {t(0,4096):.3f} {t(4096,4096):.3f} | quarters " + " ".join(f"{t(r*2048,2048):.3f}" for r in range(4)) + " | eighths " + " ".join(f"{t(r*1024,1024):.3f}" for r in range(8)))
# outer band only (no caustic: uniform 3-image work): rows 0..1023 vs 8 x that
ctx.set_map_strip(64)
print("uniform band rows 0-1023:", t(0,1024), " rows 0-2047:", t(0,2048), " rows 0-511:", t(0,512))
