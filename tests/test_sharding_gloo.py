"""N > 1 host logic on CPU: two processes, gloo backend.

The CUDA kernel is replaced by a stand-in that writes a deterministic function of the ABSOLUTE
pixel index, so what is checked is the plumbing around it: shard plan, strip alignment, chunk
loop, gather placement on rank 0 and the host assembly."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_kernel(s, q, rho, Gamma, x0, y0, dx, dy, nx, row0, nrows, views, order):
    """stand-in for mlm_map: value = f(absolute row, column), independent of the sharding"""
    import torch
    r = torch.arange(row0, row0 + nrows, dtype=torch.float64)[:, None]
    c = torch.arange(nx, dtype=torch.float64)[None, :]
    base = (r * nx + c).reshape(-1)
    views["A0"].copy_(base)
    views["A2"].copy_(base * 2 + s)
    views["A4"].copy_(base * 3 + q)
    views["nimg"].copy_((base % 3 * 2 + 1).to(torch.uint8))
    views["flags"].copy_((base % 7).to(torch.uint8))


def _worker(rank, world, port, nx, ny, nchunks, align, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from microlensing_b200.sharding import ShardedMap, shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        sm = ShardedMap(nx, ny, rank, world, device="cpu", nchunks=nchunks, align=align, gather="nccl",
                        compute_fn=_fake_kernel)
        lo, hi = shard_range(ny, rank, world, align)
        assert (sm.plan.row0, sm.plan.nrows) == (lo, hi - lo)
        sm.compute(1.5, 0.25, 1e-3, 0.5, 0.0, 0.0, 1.0, 1.0)
        full = sm.assemble_host()
        # timing protocol of bench.py: max over ranks of a per-rank number
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t[0]) == world
        if rank == 0:
            np.savez(os.path.join(out_dir, "full.npz"), **full)
        else:
            assert full is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nx,ny,nchunks,align", [(16, 200, 3, 8), (8, 37, 4, 1), (32, 128, 1, 64)])
def test_two_rank_gather_places_every_row(tmp_path, nx, ny, nchunks, align):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, nx, ny, nchunks, align, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "full.npz")
    base = np.arange(nx * ny, dtype=np.float64).reshape(ny, nx)
    assert np.array_equal(got["A0"], base)
    assert np.array_equal(got["A2"], base * 2 + 1.5)
    assert np.array_equal(got["A4"], base * 3 + 0.25)
    assert np.array_equal(got["nimg"], (base % 3 * 2 + 1).astype(np.uint8))
    assert np.array_equal(got["flags"], (base % 7).astype(np.uint8))


def test_full_map_layout_offsets():
    from microlensing_b200.sharding import full_map_layout, shard_range, KEYS
    nx, ny = 8192, 8192
    layout, nbytes = full_map_layout(nx, ny)
    assert nbytes == 26 * nx * ny                       # already 256-byte aligned at this size
    ends = 0
    for key, item in KEYS:
        assert layout[key] % 256 == 0 and layout[key] >= ends
        ends = layout[key] + nx * ny * item
    # the row windows of the 8 ranks tile every array exactly
    for key, item in KEYS:
        pos = layout[key]
        for r in range(8):
            lo, hi = shard_range(ny, r, 8, 64)
            assert layout[key] + lo * nx * item == pos
            pos += (hi - lo) * nx * item
        assert pos == layout[key] + nx * ny * item
    layout, nbytes = full_map_layout(3, 5)
    assert [layout[k] for k, _ in KEYS] == [0, 256, 512, 768, 1024] and nbytes == 1280


def _shm_worker(rank, world, port, nx, ny, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from microlensing_b200.sharding import SharedHostMap, shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        sh = SharedHostMap(nx, ny, rank, world, outputs=("A4", "nimg"), strip=8, register=False)
        assert (sh.lo, sh.hi) == shard_range(ny, rank, world, 8) and set(sh.arrays) == {"A4", "nimg"}
        assert not os.path.exists(sh.path)                       # unlinked once every rank has mapped it
        rows = sh.own_rows()
        r = np.arange(sh.lo, sh.hi)[:, None] * nx + np.arange(nx)[None, :]
        rows["A4"][...] = r * 0.5
        rows["nimg"][...] = r % 5
        dist.barrier()
        if rank == 0:                                            # rank 0 sees every rank's rows in ITS mapping
            np.savez(os.path.join(out_dir, "shm.npz"), **{k: v.copy() for k, v in sh.arrays.items()})
        sh.close()
    finally:
        dist.destroy_process_group()


def test_shared_host_map_is_one_map_for_all_ranks(tmp_path):
    """SharedHostMap: the POSIX shared-memory result map of the sharded host API (CPU: without page-locking)."""
    import torch.multiprocessing as mp
    nx, ny = 24, 50
    mp.spawn(_shm_worker, args=(2, _free_port(), nx, ny, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "shm.npz")
    base = np.arange(nx * ny).reshape(ny, nx)
    assert np.array_equal(got["A4"], base * 0.5) and np.array_equal(got["nimg"], (base % 5).astype(np.uint8))
