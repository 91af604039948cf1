# quick GPU check of a kernel change: parity subset + k_map timing (strips 64/32/16) + event counts
python -m pytest tests -m gpu -x -q -k "golden or map_vs_batch or bitwise or flag_semantics or lightcurve_C2 or models_C5 or full_size or uneven or chunking" 2>&1 | tail -4
for st in 64 32 16; do echo strip $st; MLM_MAP_STRIP=$st python tools/profile_map.py 8192 4 2>&1 | tail -1; done
echo "no anchors:"; for st in 64 16; do MLM_MAP_ANCHOR=0 MLM_MAP_STRIP=$st python tools/profile_map.py 8192 3 2>&1 | tail -1; done
python tools/count_events.py C4 > gpurun_out/quick_events_c4.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/quick_events_c4.json')); print({k: round(v,4) for k,v in d['per_position'].items()})"
python - <<'PY'
import numpy as np, torch, time, sys
sys.path.insert(0, '.')
import microlensing_b200 as mb
from microlensing_b200.batched import points_device, lightcurve_coordinates
import bench
c = bench.CONFIGS["C2"]; T = c["T"]; dtau = (c["tau1"] - c["tau0"]) / T
zh = lightcurve_coordinates(c["s"], c["q"], c["u0"], c["alpha"], c["tau0"], dtau, T)
perm = np.random.default_rng(1).permutation(T)
dev = torch.device("cuda", 0)
b = [torch.empty(T, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.empty(T, dtype=torch.uint8, device=dev) for _ in range(2)]
st = torch.cuda.current_stream().cuda_stream
def timed(z, unordered, reps=5):
    zt = torch.from_numpy(z).to(dev)
    for _ in range(2): points_device(c["s"], c["q"], c["rho"], c["Gamma"], zt, T, *b, stream=st, unordered=unordered)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): points_device(c["s"], c["q"], c["rho"], c["Gamma"], zt, T, *b, stream=st, unordered=unordered)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, [x.cpu().numpy().copy() for x in b]
t0, r0 = timed(zh, False); t1, r1 = timed(zh[perm], False); t2, r2 = timed(zh[perm], True)
zr = (np.random.default_rng(2).uniform(-2, 2, T) + 1j * np.random.default_rng(3).uniform(-2, 2, T))
t3, r3 = timed(zr, False); t4, r4 = timed(zr, True)
print(f"C2 ordered {t0:.3f} ms | shuffled plain {t1:.3f} ms | shuffled sorted {t2:.3f} ms | random 2D plain {t3:.3f} ms, sorted {t4:.3f} ms")
ok = (r1[4] == 0) & (r2[4] == 0)
print("shuffled: nimg equal", np.array_equal(r1[3], r2[3]), "max rel A4 (unflagged)", np.abs(r2[2][ok] / r1[2][ok] - 1).max(), "vs ordered nimg", np.array_equal(r0[3][perm], r2[3]))
ok = (r3[4] == 0) & (r4[4] == 0)
print("random: nimg equal", np.array_equal(r3[3], r4[3]), int((r3[3] != r4[3]).sum()), "max rel A4", np.abs(r4[2][ok] / r3[2][ok] - 1).max())
PY
