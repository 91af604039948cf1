"""Device-resident C4 map on N GPUs: time per step for the gather variants and root shares of ShardedMap.

  torchrun --nproc-per-node N tools/time_gather.py [--out gpurun_out/gather_nN.json] [--shares 0.2,0.25,...]

Per variant: CUDA events around `steps` compute() calls (barrier + synchronize on both sides, max over ranks), the
kernel-only time of the same deal (local_only=True: results stay in each rank's HBM), and a bitwise comparison of
seeded strips of rank 0's gathered map with a single-GPU recomputation.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--shares", default="auto")
    ap.add_argument("--gathers", default="p2p,bulk")
    ap.add_argument("--outputs", default="all")
    ap.add_argument("--nx", type=int, default=8192)
    ap.add_argument("--chunks", default="4")
    ap.add_argument("--strips", default="0", help="comma list of strip lengths (0: auto from the per-rank size)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ["MLM_DEVICE"] = str(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from microlensing_b200.sharding import ShardedMap
    from microlensing_b200.batched import map_device
    nx = ny = args.nx
    s, q, rho, G = 1.0, 0.5, 1e-3, 0.5
    half = 0.8
    x0, y0, dx, dy = -0.08 - half, -half, 2 * half / nx, 2 * half / ny
    outputs = None if args.outputs == "all" else tuple(args.outputs.split(","))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn):
        for _ in range(3):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def check(sm):
        """rank 0: seeded strips of the gathered map == the same strips computed here on one GPU"""
        if rank != 0:
            return None
        fv = sm.full_views()
        L = sm.align
        rng = np.random.default_rng(7)
        ks = sorted(set([0, ny // L - 1] + list(rng.integers(0, ny // L, 14))))
        bad = 0
        keys = [k for k in ("A0", "A2", "A4", "nimg", "flags") if k in sm.outputs]
        for k in ks:
            b = {key: torch.empty(L * nx, dtype=torch.float64 if key[0] == "A" else torch.uint8, device=dev) for key in keys}
            map_device(s, q, rho, G, x0, y0, dx, dy, nx, k * L, L, b.get("A0"), b.get("A2"), b.get("A4"), b.get("nimg"),
                       b.get("flags"), stream=torch.cuda.current_stream().cuda_stream, strip=L)
            torch.cuda.synchronize()
            for key in keys:
                got = fv[key][k * L:(k + 1) * L].reshape(-1).view(torch.uint8)
                bad += int((got != b[key].view(torch.uint8)).sum())
        return {"strips": len(ks), "rows": len(ks) * L, "mismatched_bytes": bad}

    res = []
    shares = [None if x == "auto" else float(x) for x in args.shares.split(",")]
    import itertools
    for gather in args.gathers.split(","):
        if world == 1:
            gather = "none"
        for share, strip in itertools.product(shares, [int(x) for x in args.strips.split(",")]):
            for nch in [int(c) for c in args.chunks.split(",")] if gather == "bulk" else [None]:
                sm = ShardedMap(nx, ny, rank, world, device=dev, gather=gather, outputs=outputs,
                                layout=None if world > 1 else "block", root_share=share, nchunks=nch, align=strip or None)
                if rank == 0:                 # poison: the check must see this run's data
                    fv = sm.full_views()
                    if fv is not None:
                        for v in fv.values():
                            v.fill_(float("nan") if v.dtype == torch.float64 else 255)
                barrier()
                ms = timed(lambda: sm.compute(s, q, rho, G, x0, y0, dx, dy))
                ms_local = timed(lambda: sm.compute(s, q, rho, G, x0, y0, dx, dy, local_only=True)) if world > 1 else ms
                chk = check(sm)
                row = dict(world=world, gather=gather, share_arg=share, root_share=sm.root_share, period=sm.period,
                           root_positions=len(sm.sel_all[0]), strip=sm.align, chunks=sm.nchunks, ms=ms, ms_kernel_only=ms_local,
                           mags_per_s=nx * ny / ms * 1e3, bytes_per_position=sm.bytes_per_position, check=chk)
                if rank == 0:
                    print(json.dumps(row), flush=True)
                res.append(row)
                sm.close()
                barrier()
            if world == 1:
                break
        if world == 1:
            break
    if rank == 0 and args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
