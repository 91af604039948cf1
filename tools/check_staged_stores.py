"""The staged store path of k_map (one-byte arrays written a 128-byte row segment at a time, normally taken only when the
outputs live in another GPU's memory) forced onto LOCAL outputs and compared byte for byte with the plain stores.

    python tools/check_staged_stores.py            # runs itself twice (MLM_MAP_U8ROWS = 0 and 1) and compares
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (row0, nrows, strip, outputs, order, deal): windows starting inside a strip, partial last groups of four rows, NULL outputs
CASES = [(0, 1024, 64, "A0 A2 A4 nimg flags", 4, None), (3, 1021, 16, "A0 A2 A4 nimg flags", 4, None),
         (70, 257, 64, "A4 nimg flags", 4, None), (5, 38, 16, "A0 A4", 4, None), (128, 512, 16, "A0 A2 nimg", 2, None),
         (0, 2048, 16, "A0 A2 A4 nimg flags", 4, (13, (1, 5, 10)))]


def worker():
    import torch
    from microlensing_b200.batched import map_device, map_deal_device
    nx = 8192
    s, q, rho, G = 1.0, 0.5, 1e-3, 0.5
    half = 0.8
    x0, y0, dx, dy = -0.08 - half, -half + 3000 * 2 * half / nx, 2 * half / nx, 2 * half / nx
    dev = torch.device("cuda", 0)
    out = []
    for row0, nrows, strip, outs, order, deal in CASES:
        n = nrows * nx
        buf = {k: torch.full((n,), 7, dtype=torch.float64 if k[0] == "A" else torch.uint8, device=dev)
               for k in ("A0", "A2", "A4", "nimg", "flags")}
        arg = [buf[k] if k in outs.split() else None for k in ("A0", "A2", "A4", "nimg", "flags")]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(2):
            e0.record()
            if deal is None:
                map_device(s, q, rho, G, x0, y0, dx, dy, nx, row0, nrows, *arg, order=order, strip=strip)
            else:
                map_deal_device(s, q, rho, G, x0, y0, dx, dy, nx, row0, nrows, deal[0], deal[1], *arg, order=order, strip=strip)
            e1.record()
            torch.cuda.synchronize()
        h = hashlib.sha256()
        for k in ("A0", "A2", "A4", "nimg", "flags"):
            h.update(buf[k].cpu().numpy().tobytes())
        out.append({"case": [row0, nrows, strip, outs, order, deal is not None], "sha": h.hexdigest()[:16],
                    "ms": round(e0.elapsed_time(e1), 4)})
    print(json.dumps(out))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "worker":
        return worker()
    res = {}
    for mode in ("0", "1"):
        env = dict(os.environ, MLM_MAP_U8ROWS=mode)
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "worker"], env=env, capture_output=True, text=True)
        if p.returncode != 0:
            sys.stderr.write(p.stderr[-3000:])
            raise SystemExit(f"worker (staged = {mode}) failed")
        res[mode] = json.loads([l for l in p.stdout.splitlines() if l.startswith("[")][-1])
    bad = 0
    for a, b in zip(res["0"], res["1"]):
        same = a["sha"] == b["sha"]
        bad += not same
        print(a["case"], "plain", a["ms"], "ms | staged", b["ms"], "ms |", "identical" if same else "DIFFERENT")
    print("staged stores:", "all identical" if not bad else f"{bad} case(s) DIFFER")
    raise SystemExit(1 if bad else 0)


if __name__ == "__main__":
    main()
