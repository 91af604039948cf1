"""Bandwidth of kernel stores into ANOTHER GPU's HBM by store pattern (VERDICT r1 #1c: why do the map kernel's peer
stores reach ~660 GB/s when bulk peer copies of the same bytes reach 790 GB/s?).

  torchrun --nproc-per-node N tools/peer_store_probe.py

Ranks > 0 store a full 8192^2 result set (or their 1/(N-1) share of the rows) into rank 0's buffer (CUDA IPC mapping)
with mlm_store_probe (no arithmetic): mode 0 = the map kernel's pattern, 1 = float64 arrays only, 2 = uint8 arrays
only, 3 = uint8 rows written 128 B at a time; also locally (own HBM) for reference.  CUDA events, max over ranks."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from microlensing_b200 import _capi
    from microlensing_b200.sharding import full_map_layout
    ctx = _capi.context(local)
    nx = ny = 8192
    layout, nbytes = full_map_layout(nx, ny)
    box = [None]
    base = None
    if rank == 0:
        base = ctx.device_alloc(nbytes)
        box[0] = ctx.ipc_export(base)
    dist.broadcast_object_list(box, src=0)
    if rank != 0:
        base = ctx.ipc_open(box[0])
    own = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    peers = world - 1
    # rows of this rank: ranks 1.. split the map evenly; rank 0 idles in the peer test
    rows = ny // peers if peers else ny
    r0 = (rank - 1) * rows if rank else 0
    st = torch.cuda.current_stream().cuda_stream
    bytes_mode = {0: 26, 1: 24, 2: 2, 3: 26}
    res = {}
    for target in ("peer", "local"):
        for mode in (0, 1, 2, 3):
            for rpb in (16, 64):
                def launch():
                    if target == "peer" and rank == 0:
                        return
                    b = base if target == "peer" else own.data_ptr()
                    ptr = [b + layout[k] + r0 * nx * (8 if k[0] == "A" else 1) for k in ("A0", "A2", "A4", "nimg", "flags")]
                    ctx.check(ctx.lib.mlm_store_probe(ctx.handle, mode, nx, rows, rpb, *[_capi.ptr(p) for p in ptr], _capi.ptr(st)),
                              "mlm_store_probe")
                for _ in range(2):
                    launch()
                torch.cuda.synchronize(); dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    launch()
                e1.record()
                torch.cuda.synchronize(); dist.barrier()
                t = torch.tensor([e0.elapsed_time(e1) / 5], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                total = bytes_mode[mode] * nx * rows * (peers if target == "peer" else 1)
                res[f"{target}_mode{mode}_rows{rpb}"] = {"ms": float(t[0]), "GBs": total / float(t[0]) / 1e6}
                if rank == 0:
                    print(target, "mode", mode, "rows/block", rpb, f"{float(t[0]):.3f} ms", f"{total / float(t[0]) / 1e6:.0f} GB/s"
                          + (" into rank 0" if target == "peer" else " per GPU"), flush=True)
    if rank == 0:
        with open(f"gpurun_out/r2_store_probe_n{world}.json", "w") as f:
            json.dump(res, f, indent=1)
    dist.barrier()
    if rank != 0:
        ctx.ipc_close(base)
    dist.barrier()
    if rank == 0:
        ctx.device_free(base)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
