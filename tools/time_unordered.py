"""Timing of unordered point lists: plain mlm_points (tracking kernel, a cold solve per entry) vs mlm_points_unordered
(default: three-image guesses + root finder on the compacted rest; MLM_UNORDERED=sorted: Hilbert-sorted tracking)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from microlensing_b200.batched import points_device, lightcurve_coordinates
import bench
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream


def timed(par, z, unordered, reps=5):
    T = len(z)
    b = [torch.empty(T, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.empty(T, dtype=torch.uint8, device=dev) for _ in range(2)]
    zt = torch.from_numpy(np.ascontiguousarray(z)).to(dev)
    for _ in range(2): points_device(*par, zt, T, *b, stream=st, unordered=unordered)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): points_device(*par, zt, T, *b, stream=st, unordered=unordered)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, [x.cpu().numpy().copy() for x in b]


c = bench.CONFIGS["C2"]; T = c["T"]; dtau = (c["tau1"] - c["tau0"]) / T
par2 = (c["s"], c["q"], c["rho"], c["Gamma"])
zh = lightcurve_coordinates(c["s"], c["q"], c["u0"], c["alpha"], c["tau0"], dtau, T)
rng = np.random.default_rng(1)
cases = [("C2 shuffled", par2, zh[rng.permutation(T)]),
         ("random 2D q=1e-3", par2, rng.uniform(-2, 2, T) + 1j * rng.uniform(-2, 2, T)),
         ("C4 window random", (1.0, 0.5, 1e-3, 0.5), (-0.08 + rng.uniform(-.8, .8, T)) + 1j * rng.uniform(-.8, .8, T) - 0.5 / 1.5),
         ("C3 window random", (0.8, 0.1, 1e-3, 0.5), (0.11 + rng.uniform(-.75, .75, T)) + 1j * rng.uniform(-.75, .75, T) - 0.8 * 0.1 / 1.1)]
t0, r0 = timed(par2, zh, False)
print(f"C2 ordered (tracked): {t0:.3f} ms")
for name, par, z in cases:
    t1, r1 = timed(par, z, False); t2, r2 = timed(par, z, True)
    ok = (r1[4] == 0) & (r2[4] == 0)
    print(f"{name}: plain {t1:.3f} ms | unordered[{os.environ.get('MLM_UNORDERED', 'direct')}] {t2:.3f} ms | nimg equal {np.array_equal(r1[3], r2[3])} "
          f"max rel A4 (unflagged) {np.abs(r2[2][ok] / r1[2][ok] - 1).max():.2e} flags differ {(r1[4] != r2[4]).sum()}")
