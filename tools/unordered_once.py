"""One call of the unordered-list path on the shuffled C2 light curve and on random 2-D positions (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from microlensing_b200.batched import points_device, lightcurve_coordinates
c = bench.CONFIGS["C2"]; T = c["T"]; dtau = (c["tau1"] - c["tau0"]) / T
zh = lightcurve_coordinates(c["s"], c["q"], c["u0"], c["alpha"], c["tau0"], dtau, T)
zs = zh[np.random.default_rng(1).permutation(T)]
zr = np.random.default_rng(2).uniform(-2, 2, T) + 1j * np.random.default_rng(3).uniform(-2, 2, T)
dev = torch.device("cuda", 0)
b = [torch.empty(T, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.empty(T, dtype=torch.uint8, device=dev) for _ in range(2)]
st = torch.cuda.current_stream().cuda_stream
for z in (zs, zr):
    zt = torch.from_numpy(z).to(dev)
    for rep in range(2):
        points_device(c["s"], c["q"], c["rho"], c["Gamma"], zt, T, *b, stream=st, unordered=True)
    torch.cuda.synchronize()
print("done")
