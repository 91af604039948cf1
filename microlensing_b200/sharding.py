"""Row / model sharding across the GPUs of one box and the final gather.

Every source position is independent (reference multipoles.py:184), so there is no exchange
step on the data path: maps shard by contiguous row blocks, model grids by contiguous model
ranges.  The only collective is the final gather of the SoA outputs to rank 0
(26 B/position: A0, A2, A4 float64 + nimg, flags uint8) over NCCL / NVLink, issued per row
chunk on a side stream so that it overlaps the kernel of the next chunk.

torch / torch.distributed are plumbing here (device buffers, streams, NCCL); the arithmetic is
in libmlmultipoles.so.
"""
from __future__ import annotations

from dataclasses import dataclass

__all__ = ["shard_range", "ShardPlan", "plan_map", "BYTES_PER_POSITION", "ShardedMap"]

BYTES_PER_POSITION = 26           # 3 x float64 + 2 x uint8


MAP_STRIP = 64                    # default rows per warm-start strip of the map kernel (mlm_map_strip)


def shard_range(n: int, rank: int, world: int, align: int = 1) -> tuple[int, int]:
    """Contiguous block [lo, hi) of n units for `rank`; the first ranks get the remainder.

    With align > 1 the boundaries are multiples of `align` (map shards start on a strip boundary
    of the map kernel, which makes sharded == unsharded bit for bit)."""
    if world <= 0 or not (0 <= rank < world) or n < 0 or align < 1:
        raise ValueError("bad shard arguments")
    nb = -(-n // align)
    base, extra = divmod(nb, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return min(lo * align, n), min(hi * align, n)


@dataclass(frozen=True)
class ShardPlan:
    rank: int
    world: int
    row0: int                     # first map row of this rank
    nrows: int                    # rows of this rank
    max_rows: int                 # rows of the largest shard (gather slots are padded to it)
    chunks: tuple                 # ((row offset within shard, rows), ...) pipelining units


def plan_map(ny: int, rank: int, world: int, nchunks: int = 4, align: int = MAP_STRIP) -> ShardPlan:
    lo, hi = shard_range(ny, rank, world, align)
    max_rows = shard_range(ny, 0, world, align)[1]
    n = hi - lo
    # chunk boundaries come from max_rows so that they are identical on every rank (the gather
    # needs equal sizes); a rank with fewer rows computes only the rows it owns in each chunk
    nchunks = max(1, min(nchunks, max_rows)) if max_rows else 1
    chunks = []
    for c in range(nchunks):
        a, b = shard_range(max_rows, c, nchunks, align)
        if b > a:
            chunks.append((a, b - a))
    return ShardPlan(rank, world, lo, n, max_rows, tuple(chunks))


class ShardedMap:
    """Device-resident, row-sharded magnification map with a chunk-pipelined NCCL gather.

    Each rank owns a byte buffer laid out [A0 | A2 | A4 | nimg | flags] for `max_rows` rows
    (padded so that every rank contributes the same size).  `compute()` enqueues the map
    kernel chunk by chunk on the compute stream; with world > 1, after every chunk the chunk's
    rows are gathered to rank 0 on a communication stream (torch.distributed.gather over NCCL),
    overlapping the next chunk's kernel.
    """

    def __init__(self, nx: int, ny: int, rank: int = 0, world: int = 1, device=None, nchunks: int = 4, group=None,
                 align: int = MAP_STRIP):
        import torch
        self.torch = torch
        self.nx, self.ny = nx, ny
        self.align = align
        self.plan = plan_map(ny, rank, world, nchunks, align)
        self.group = group
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        n_slot = self.plan.max_rows * nx
        self.n_slot = n_slot
        with torch.cuda.device(self.device):
            self.local = self._alloc(n_slot)
            self.comm_stream = torch.cuda.Stream() if world > 1 else None
            self.full = None
            if world > 1 and rank == 0:
                self.full = [self._alloc(n_slot) if r else self.local for r in range(world)]

    def _alloc(self, n):
        t = self.torch
        return {"A0": t.empty(n, dtype=t.float64, device=self.device),
                "A2": t.empty(n, dtype=t.float64, device=self.device),
                "A4": t.empty(n, dtype=t.float64, device=self.device),
                "nimg": t.empty(n, dtype=t.uint8, device=self.device),
                "flags": t.empty(n, dtype=t.uint8, device=self.device)}

    def compute(self, s, q, rho, Gamma, x0, y0, dx, dy, order=4, gather=True):
        """Enqueue this rank's shard (and the gather to rank 0 when world > 1)."""
        import torch.distributed as dist
        from .batched import map_device
        t = self.torch
        p, nx = self.plan, self.nx
        cur = t.cuda.current_stream(self.device)
        for (r_off, nr) in p.chunks:
            if nr == 0:
                continue
            o, cnt = r_off * nx, nr * nx
            buf = self.local
            own = max(0, min(r_off + nr, p.nrows) - r_off)      # rows of this chunk this rank owns
            if own:
                oc = own * nx
                map_device(s, q, rho, Gamma, x0, y0, dx, dy, nx, p.row0 + r_off, own,
                           buf["A0"][o:o + oc], buf["A2"][o:o + oc], buf["A4"][o:o + oc],
                           buf["nimg"][o:o + oc], buf["flags"][o:o + oc], order=order,
                           stream=cur.cuda_stream, device=self.device.index)
            if p.world > 1 and gather:
                self.comm_stream.wait_stream(cur)
                with t.cuda.stream(self.comm_stream):
                    for key in ("A0", "A2", "A4", "nimg", "flags"):
                        src = buf[key][o:o + cnt]
                        dst = [self.full[r][key][o:o + cnt] for r in range(p.world)] if p.rank == 0 else None
                        dist.gather(src, dst, dst=0, group=self.group)
        if p.world > 1 and gather:
            cur.wait_stream(self.comm_stream)

    def assemble_host(self):
        """Rank 0: the full (ny, nx) arrays on the host (numpy), from the gathered shards."""
        import numpy as np
        p, nx = self.plan, self.nx
        if p.rank != 0:
            return None
        out = {}
        for key in ("A0", "A2", "A4", "nimg", "flags"):
            parts = []
            for r in range(p.world):
                lo, hi = shard_range(self.ny, r, p.world, self.align)
                src = self.local if (p.world == 1 or r == 0) else self.full[r]
                parts.append(src[key][:(hi - lo) * nx].cpu().numpy().reshape(hi - lo, nx))
            out[key] = np.concatenate(parts, axis=0)
        return out
