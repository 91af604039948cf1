#!/usr/bin/env python
"""Per-phase instruction table of a kernel from an `ncu --set full --import-source on` report.

    python tools/ncu_phase_table.py REPORT.ncu-rep LIB.so KERNEL_REGEX OUT_PREFIX [positions] [launch_index]

ncu's SASS page has the executed count of every instruction, nvdisasm (-gi on the cubin of the SAME library, built
with -lineinfo) the chain of inlined call sites of every instruction.  Each instruction is attributed to the innermost
frame of its chain that lies in a PHASE region of the sources:
  * the body of a phase function (PHASE_FUNCS below), or
  * a sub-region opened by a `//@phase name` comment (until the next marker or the end of the enclosing function);
helpers (complex arithmetic, rcp, horner2, ...) are not phases, so their instructions go to the caller.
Output: OUT_PREFIX.md / .json with, per phase, warp-level instructions, thread-level instructions and the FP64 ones
(DFMA, DMUL, DADD, DSETP, DMNMX...) per source position.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "microlensing_b200", "csrc")
PHASE_FUNCS = {"lens_polynomial": "polynomial", "solve_roots": "cold_solve", "classify_roots": "classify",
               "lens_newton": "lens_newton", "nonphysical_pair_near": "watch_fp32", "image_terms_from_W": "image_terms", "sum_images": "image_sums",
               "advance_images": "track_other", "store_result": "store", "epoch_tau": "coords", "epoch_ratio": "coords"}
FP64_OPS = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX", "DFMA.RN", "DFMA.RM", "DFMA.RP")


def source_regions(path):
    """[(first line, last line, phase)] of a source file: phase functions and //@phase sub-regions."""
    lines = open(path).read().splitlines()
    funcs = []                                   # (start, end, name) of top-level functions
    i = 0
    while i < len(lines):
        m = re.match(r"^(?:MLM_HDI?|__device__ __forceinline__|__global__|static|inline)\b.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", lines[i])
        if lines[i].startswith("__global__"):                # kernels: the name may be on the next line
            m = re.search(r"\b(k_[a-z0-9_]+)\s*\(", lines[i]) or re.search(r"\b(k_[a-z0-9_]+)\s*\(", lines[i + 1])
        if m:
            j = i
            if lines[i].rstrip().endswith("}") and "{" in lines[i]:
                funcs.append((i + 1, i + 1, m.group(1)))
                i += 1
                continue
            while j < len(lines) and not lines[j].startswith("}"):
                j += 1
            funcs.append((i + 1, j + 1, m.group(1)))
            i = j + 1
        else:
            i += 1
    regions = []
    for a, b, name in funcs:
        marks = [(k + 1, re.search(r"//@phase\s+(\w+)", lines[k]).group(1)) for k in range(a - 1, b)
                 if "//@phase" in lines[k]]
        base = PHASE_FUNCS.get(name) or ("kernel_body" if name.startswith("k_") else None)
        if base is None and not marks:
            continue
        cur, cur_name = a, base
        for ln, nm in marks:
            if cur_name is not None and ln > cur:
                regions.append((cur, ln - 1, cur_name))
            cur, cur_name = ln, nm
        if cur_name is not None:
            regions.append((cur, b, cur_name))
    return regions


def load_regions():
    out = {}
    for f in ("mlm_core.cuh", "mlm_kernels.cu"):
        out[f] = source_regions(os.path.join(CSRC, f))
    return out


def phase_of(chain, regions):
    for file, line in chain:                                  # innermost first
        for a, b, name in regions.get(os.path.basename(file), ()):
            if a <= line <= b:
                return name
    return "unattributed"


def disasm(lib, kernel_re):
    """[(opcode text, chain)] of the first kernel whose mangled/demangled name matches."""
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
        cubins = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")]
        txt = subprocess.run(["nvdisasm", "-gi", "-c", cubins[0]], capture_output=True, text=True, check=True).stdout
    # split into functions
    out, cur, chain, pending, active = [], None, [], [], False
    for line in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            name = m.group(1)
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            active = bool(re.search(kernel_re, dem)) and not out
            if out and not active:
                pass
            cur = dem
            chain, pending = [], []
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
        if m:
            pending.append((m.group(1), int(m.group(2)), m.group(3), int(m.group(4)) if m.group(4) else None))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            if pending:
                # a block of annotations: first entries are inline frames (innermost first), the last is the outermost
                chain = []
                for f, ln, f2, ln2 in pending:
                    if not chain:
                        chain.append((f, ln))
                    if f2 is not None:
                        chain.append((f2, ln2))
                # dedupe consecutive duplicates
                ch = []
                for c in chain:
                    if not ch or ch[-1] != c:
                        ch.append(c)
                chain = ch
                pending = []
            out.append((int(m.group(1), 16), m.group(2).strip(), list(chain)))
    return out


def ncu_sass(rep, kernel_re, launch=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True, check=True).stdout
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(txt)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None and row[0] == "Address":
            cur["hdr"] = row
        elif cur is not None and cur["hdr"] is not None and row[0].startswith("0x"):
            cur["rows"].append(row)
    blocks = [b for b in blocks if re.search(kernel_re, b["name"])]
    b = blocks[launch if launch is not None else -1]
    h = {k: i for i, k in enumerate(b["hdr"])}
    base = int(b["rows"][0][0], 16)
    return [(int(r[0], 16) - base, r[h["Source"]].strip(), int(r[h["Instructions Executed"]]),
             int(r[h["Thread Instructions Executed"]]), int(r[h["Predicated-On Thread Instructions Executed"]]))
            for r in b["rows"]]


def main():
    rep, lib, kre, out = sys.argv[1:5]
    npos = float(sys.argv[5]) if len(sys.argv) > 5 else 8192.0 * 8192.0
    launch = int(sys.argv[6]) if len(sys.argv) > 6 else None
    regions = load_regions()
    dis = disasm(lib, kre)
    sass = ncu_sass(rep, kre, launch)
    if len(dis) != len(sass):
        raise SystemExit(f"instruction count differs: nvdisasm {len(dis)} vs ncu {len(sass)} -- not the profiled library?")
    table = {}
    mism = 0
    for (off, op, chain), (off2, op2, wi, ti, tp) in zip(dis, sass):
        if op.split()[0].lstrip("@!P0123456789 ") != op2.split()[0].lstrip("@!P0123456789 ") and off != off2:
            mism += 1
        ph = phase_of(chain, regions)
        t = table.setdefault(ph, dict(static=0, warp_inst=0, thread_inst=0, fp64_thread=0, dfma=0, dmul=0, dadd=0,
                                      fp64_warp=0, mufu64=0))
        mnem = re.sub(r"^@!?U?P\d+\s+", "", op).split()[0]
        t["static"] += 1
        t["warp_inst"] += wi
        t["thread_inst"] += tp
        root = mnem.split(".")[0]
        if root in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"):
            t["fp64_thread"] += tp
            t["fp64_warp"] += wi
            if root in ("DFMA", "DMUL", "DADD"):
                t[root.lower()] += tp
        if mnem.startswith("MUFU") and "64" in mnem:
            t["mufu64"] += tp
    if mism > len(dis) // 100:
        raise SystemExit(f"{mism} opcode mismatches between nvdisasm and ncu -- not the profiled library?")
    tot = {k: sum(t[k] for t in table.values()) for k in next(iter(table.values()))}
    res = {"report": os.path.basename(rep), "kernel": kre, "positions": npos,
           "per_position": {ph: {k: (v / npos if k != "static" else v) for k, v in t.items()} for ph, t in table.items()},
           "total_per_position": {k: (v / npos if k != "static" else v) for k, v in tot.items()}}
    res["total_per_position"]["flop"] = (2 * tot["dfma"] + tot["dmul"] + tot["dadd"]) / npos
    res["total_per_position"]["avg_active_lanes_fp64"] = tot["fp64_thread"] / max(tot["fp64_warp"], 1)
    with open(out + ".json", "w") as f:
        json.dump(res, f, indent=1)
    order = sorted(table, key=lambda p: -table[p]["fp64_thread"])
    md = [f"Per-phase executed instructions of `{kre}` per source position ({npos:.0f} positions; `{os.path.basename(rep)}`;",
          "thread-level, predicated-on; FP64 = DFMA+DMUL+DADD+DSETP+DMNMX; lanes = FP64 thread / (32 x FP64 warp instructions))", "",
          "| phase | static SASS | all thread instr | FP64 | DFMA | DMUL | DADD | share of FP64 | active lanes | warp instr x32 (all) | warp instr x32 (FP64) |", "|---|---|---|---|---|---|---|---|---|---|---|"]
    for ph in order + ["TOTAL"]:
        t = tot if ph == "TOTAL" else table[ph]
        md.append(f"| {ph} | {t['static']} | {t['thread_inst'] / npos:.1f} | {t['fp64_thread'] / npos:.1f} | {t['dfma'] / npos:.1f} | "
                  f"{t['dmul'] / npos:.1f} | {t['dadd'] / npos:.1f} | {100 * t['fp64_thread'] / max(tot['fp64_thread'], 1):.1f} % | "
                  f"{t['fp64_thread'] / max(t['fp64_warp'], 1):.1f} | {32 * t['warp_inst'] / npos:.1f} | {32 * t['fp64_warp'] / npos:.1f} |")
    with open(out + ".md", "w") as f:
        f.write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
