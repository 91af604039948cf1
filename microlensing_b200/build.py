"""Build libmlmultipoles.so in-tree with nvcc for sm_100a (no JIT cache, no torch arch list).

    python -m microlensing_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmlmultipoles.so")
STATS_LIB = os.path.join(HERE, "libmlmultipoles_stats.so")      # -DMLM_STATS: event counters, measurement only
SOURCES = ["mlm_kernels.cu"]
# mlmultipoles.h: the C ABI header.  In the repository csrc/mlmultipoles.h is a link to include/mlmultipoles.h;
# an installed package carries a copy (package-data), so `python -m microlensing_b200.build` works from a wheel.
HEADERS = ["mlm_core.cuh", "mlm_qkl_generated.cuh", "mlmultipoles.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xptxas", "-v", "-shared", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmlmultipoles.so cannot be built (there is no CPU fallback)")


def is_stale(lib: str = LIB) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HERE, "_version.py"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def _version() -> str:
    ns: dict = {}
    with open(os.path.join(HERE, "_version.py")) as f:
        exec(f.read(), ns)
    return ns["__version__"]


def build_library(force: bool = False, verbose: bool = False, stats: bool = False) -> str:
    """Compile the CUDA library (stats=True: the counting build libmlmultipoles_stats.so) if missing or older than
    its sources; return its path."""
    lib = STATS_LIB if stats else LIB
    if not force and not is_stale(lib):
        return lib
    cmd = [_nvcc()] + NVCC_FLAGS + (["-DMLM_STATS"] if stats else []) + \
        ["-I", CSRC, f'-DMLM_VERSION_STRING="{_version()}"', "-o", lib] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(HERE, "build_ptxas_stats.log" if stats else "build_ptxas.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if verbose or proc.returncode != 0:
        sys.stderr.write(log)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed building {os.path.basename(lib)}")
    return lib


def source_hash() -> str:
    """sha256 over the CUDA sources and headers: identifies the kernels a measurement under profiles/ belongs to."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted(SOURCES + HEADERS):
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


def ptxas_report() -> dict:
    """Registers / spill bytes per kernel parsed from the last build log (zero-spill gate)."""
    path = os.path.join(HERE, "build_ptxas.log")
    out: dict = {}
    if not os.path.exists(path):
        return out
    cur = None
    import re
    for line in open(path):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = m.group(1)
            out[cur] = {}
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and "stack" not in out[cur]:
            out[cur].update(stack=int(m.group(1)), spill_stores=int(m.group(2)), spill_loads=int(m.group(3)))
        m = re.search(r"Used (\d+) registers", line)
        if m:
            out[cur]["registers"] = int(m.group(1))
    return out


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv, stats=True))
    for k, v in ptxas_report().items():
        print(k, v)
