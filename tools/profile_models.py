"""Minimal driver for ncu: the model-grid kernel (k_models<4>) on the C5 grid (1e4 models x 2000 epochs).
    python tools/profile_models.py [M] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from microlensing_b200.batched import models_device
M = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
T = 2000
rng = np.random.default_rng(20240517)
sv = np.exp(rng.uniform(np.log(0.5), np.log(2.), M)); qv = np.exp(rng.uniform(np.log(1e-4), np.log(1.), M))
rv = np.exp(rng.uniform(np.log(1e-4), np.log(1e-2), M)); gv = np.full(M, 0.5)
uv, av = rng.uniform(-.5, .5, M), rng.uniform(0, 2 * np.pi, M)
dev = torch.device("cuda", 0)
par = [torch.from_numpy(a).to(dev) for a in (sv, qv, rv, gv, uv, av)]
n = M * T
b = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
st = torch.cuda.current_stream().cuda_stream
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    models_device(*par, M, -2.0, 4.0 / T, T, *b, stream=st)
    e1.record(); torch.cuda.synchronize()
    print(f"rep {r}: {e0.elapsed_time(e1):.3f} ms  {n / e0.elapsed_time(e1) / 1e6:.2f} G mags/s  f5={(b[3] == 5).float().mean().item():.4f}")
