"""Transfer ceilings of the box that bound the multi-GPU map ("measure the ceiling"):

  torchrun --nproc-per-node N tools/measure_ceilings.py [--out gpurun_out/ceilings_nN.json]

 (1) concurrent D2H: every rank copies its 1/N share of the 8192^2 map results (26 B/pixel) from HBM to
     page-locked host memory at the same time; variants: own cudaHostAlloc buffer per process, and ONE POSIX
     shared-memory map that every process cudaHostRegister()s (what the sharded host API uses);
 (2) N-1 -> 1 peer copies: every rank > 0 pushes its share into rank 0's HBM (CUDA IPC mapping) with bulk
     cudaMemcpyAsync at the same time -- the ingress ceiling of a gather to one GPU.
Everything is timed between barriers with wall clock around stream synchronisations (max over ranks).
"""
import argparse
import json
import mmap
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--total-mb", type=float, default=26 * 8192 * 8192 / 1e6)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from microlensing_b200 import _capi
    ctx = _capi.context(local)
    rt = torch.cuda.cudart()
    share = int(args.total_mb * 1e6 / world) // 4096 * 4096
    src = torch.empty(share, dtype=torch.uint8, device=dev)
    src.fill_(rank + 1)
    res = {"world": world, "share_bytes": share}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, reps=args.reps):
        fn(); barrier()
        ts = []
        for _ in range(reps):
            barrier()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t[0]))
        return min(ts), float(np.median(ts))

    st = torch.cuda.current_stream().cuda_stream

    # (1a) own pinned buffer per process
    host = ctx.pinned_empty(share, np.uint8)
    hp = host.ctypes.data
    best, med = timed(lambda: ctx.memcpy_async(hp, src.data_ptr(), share, st))
    res["d2h_own_pinned"] = {"best_s": best, "median_s": med, "per_gpu_GBs": share / best / 1e9,
                             "aggregate_GBs": share * world / best / 1e9}
    # the same in 8 pieces on two streams (what the chunked host API does)
    s2 = [torch.cuda.Stream(), torch.cuda.Stream()]

    def chunked():
        n = 8
        c = share // n
        for k in range(n):
            ctx.memcpy_async(hp + k * c, src.data_ptr() + k * c, c, s2[k & 1].cuda_stream)
    best, med = timed(chunked)
    res["d2h_own_pinned_8chunks_2streams"] = {"best_s": best, "aggregate_GBs": share * world / best / 1e9}
    assert host[0] == rank + 1 and host[share - 1] == rank + 1
    # how the aggregate grows with the number of GPUs copying at once (same share each)
    res["d2h_own_pinned_by_active_gpus"] = {}
    k = 1
    while k <= world:
        best, med = timed(lambda: ctx.memcpy_async(hp, src.data_ptr(), share, st) if rank < k else None)
        res["d2h_own_pinned_by_active_gpus"][str(k)] = {"best_s": best, "aggregate_GBs": share * k / best / 1e9}
        k *= 2
    # H2D for completeness
    best, med = timed(lambda: ctx.memcpy_async(src.data_ptr(), hp, share, st))
    res["h2d_own_pinned"] = {"best_s": best, "aggregate_GBs": share * world / best / 1e9}
    src.fill_(rank + 1)
    if rank == 0:
        a = np.empty(share, np.uint8); a[:] = 1
        t0 = time.perf_counter(); np.copyto(a, host); dt = time.perf_counter() - t0
        res["host_memcpy_1thread_GBs"] = share / dt / 1e9
    del host

    # (1b) one shared-memory map registered by every process
    path = "/dev/shm/mlm_ceiling_%d" % int(os.environ.get("MASTER_PORT", "0"))
    total = share * world
    if rank == 0:
        with open(path, "wb") as f:
            f.truncate(total)
    barrier()
    fd = os.open(path, os.O_RDWR)
    mm = mmap.mmap(fd, total)
    arr = np.frombuffer(mm, dtype=np.uint8)
    base = arr.ctypes.data
    t0 = time.perf_counter()
    # each rank registers only its own slice (page aligned: share is a multiple of 4096)
    rc = rt.cudaHostRegister(base + rank * share, share, 1)      # cudaHostRegisterPortable
    res["host_register_own_slice_s"] = time.perf_counter() - t0
    res["host_register_rc"] = int(rc[0]) if isinstance(rc, tuple) else int(rc)
    mine = base + rank * share
    best, med = timed(lambda: ctx.memcpy_async(mine, src.data_ptr(), share, st))
    res["d2h_shared_shm_registered"] = {"best_s": best, "median_s": med, "per_gpu_GBs": share / best / 1e9,
                                        "aggregate_GBs": share * world / best / 1e9}
    barrier()
    if rank == 0:
        ok = all(arr[r * share] == r + 1 and arr[(r + 1) * share - 1] == r + 1 for r in range(world))
        res["shared_map_assembled_on_rank0"] = bool(ok)
    rt.cudaHostUnregister(mine)
    barrier()
    del arr
    mm.close(); os.close(fd)
    if rank == 0:
        os.unlink(path)

    # (2) N-1 -> 1 peer copies into rank 0's HBM
    if world > 1:
        box = [None]
        if rank == 0:
            dst = ctx.device_alloc(total)
            box[0] = ctx.ipc_export(dst)
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            dst = ctx.ipc_open(box[0])

        def push():
            if rank != 0:
                ctx.memcpy_async(dst + rank * share, src.data_ptr(), share, st)
        best, med = timed(push)
        res["peer_to_rank0_bulk"] = {"best_s": best, "median_s": med, "ingress_GBs": share * (world - 1) / best / 1e9}

        def push4():
            if rank != 0:
                c = share // 4
                for k in range(4):
                    ctx.memcpy_async(dst + rank * share + k * c, src.data_ptr() + k * c, c, s2[k & 1].cuda_stream)
        best, med = timed(push4)
        res["peer_to_rank0_bulk_4chunks_2streams"] = {"best_s": best, "ingress_GBs": share * (world - 1) / best / 1e9}
        # single peer -> rank 0 (one NVLink port pair)
        def push1():
            if rank == 1:
                ctx.memcpy_async(dst + rank * share, src.data_ptr(), share, st)
        best, med = timed(push1)
        res["peer_1_to_rank0_bulk"] = {"best_s": best, "GBs": share / best / 1e9}
        barrier()
        if rank != 0:
            ctx.ipc_close(dst)
        dist.barrier()
        if rank == 0:
            ctx.device_free(dst)
    if rank == 0:
        try:
            import subprocess
            res["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[-3000:]
            res["cpus"] = os.cpu_count()
        except Exception:
            pass
        print(json.dumps(res, indent=1))
        if args.out:
            with open(args.out, "w") as f:
                json.dump(res, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
