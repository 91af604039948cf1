"""Host-API map timings for smaller maps: strip picked from the map size (default) vs pinned 64 rows."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import microlensing_b200 as mb
s, q, rho, G = 0.8, 0.1, 1e-3, 0.5
for n in (512, 1024, 2048, 4096):
    x0, y0, d = 0.11 - 0.75, -0.75, 1.5 / n
    out = mb.alloc_outputs((n, n), outputs=("A4",))
    for strip in (None, 64):
        for _ in range(2):
            mb.magnification_map(s, q, rho, G, x0, y0, d, d, n, n, out=out, strip=strip)
        t0 = time.perf_counter()
        for _ in range(5):
            mb.magnification_map(s, q, rho, G, x0, y0, d, d, n, n, out=out, strip=strip)
        dt = (time.perf_counter() - t0) / 5
        print(f"{n}^2 strip={strip}: {dt * 1e3:.3f} ms  {n * n / dt / 1e9:.2f} G mags/s (host API, A4 only)")
