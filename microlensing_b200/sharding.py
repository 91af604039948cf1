"""Row / model sharding across the GPUs of one box and the final gather.

Every source position is independent (reference multipoles.py:184), so there is no exchange
step on the data path: maps shard by strips of rows, model grids by contiguous model ranges.
The only collective is the final gather of the SoA outputs (26 B/position: A0, A2, A4 float64 +
nimg, flags uint8).  A gather to ONE GPU is bound by that GPU's NVLink ingress (measured on this
pool: 790 GB/s for bulk peer copies from 7 peers, ~660 GB/s for the map kernel's own peer stores;
profiles/r2_ceilings_n8.json) -- 26 B x 8192^2 x 7/8 = 1.53 GB per map, 1.9 ms, three times the
0.76 ms a 1/8 share takes to compute.  Hence

* the strips are DEALT with weights (`deal_rounds`): the gather root owns more strips than its peers,
  because its own results do not cross NVLink; the root share is chosen so that the root's compute
  time equals the time its peers' results need to arrive;
* gather="p2p" (default on CUDA, world > 1): rank 0 owns the full-map output buffers and shares
  them with the other processes (CUDA IPC, mlm_ipc_*); every rank passes pointers into that buffer
  to the map kernel, whose epilogue stores go straight into rank 0's HBM over NVLink while the
  kernel is still computing -- the gather is fused into the kernel, one launch per rank, and a
  1-element NCCL all-reduce on the compute stream is the completion fence;
* gather="bulk": the ranks > 0 compute into local HBM in a few row chunks and push each chunk's strips into
  rank 0's buffer with 2-D bulk peer copies (copy engines; one call per output array and deal position)
  while the next chunk computes;
* gather="nccl": each rank computes a contiguous row block chunk by chunk and torch.distributed
  gathers every chunk to rank 0 on a side stream (baseline).
* `SharedHostMap` / `magnification_map_sharded`: the result a user wants on the HOST.  One POSIX
  shared-memory map is mapped into every rank's process; each rank page-locks the rows it owns
  (mlm_host_register) and its D2H copies land directly in the map rank 0 returns -- no GPU-side gather, every
  GPU uses its own PCIe link (ceiling measured on this pool: 95 GB/s aggregate for 8 GPUs, 57 GB/s for one).

torch / torch.distributed are plumbing here (device buffers, streams, NCCL); the arithmetic is
in libmlmultipoles.so.
"""
from __future__ import annotations

from dataclasses import dataclass

__all__ = ["shard_range", "ShardPlan", "plan_map", "full_map_layout", "auto_strip", "BYTES_PER_POSITION", "KEYS", "ShardedMap",
           "ShardedModels", "deal_rounds", "root_share_model", "owned_strips", "SharedHostMap", "magnification_map_sharded"]

BYTES_PER_POSITION = 26           # 3 x float64 + 2 x uint8
KEYS = (("A0", 8), ("A2", 8), ("A4", 8), ("nimg", 1), ("flags", 1))     # output arrays and their item sizes


MAP_STRIP = 64                    # strip length of the stand-in kernels of the CPU tests


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


def auto_strip(nx: int, rows_per_rank: int, resident_threads: int = 148 * 16 * 32, waves: int = 12) -> int:
    """Strip length (power of two, 16..64) that leaves every rank about `waves` waves of work items:
    one work item = one column of one strip; a B200 keeps 148 SMs x 16 warps of them resident.
    Measured on the 8192-wide C4 map (tools/time_shards2.py): a launch of w waves of 64-row strips
    costs about ceil(w + 0.5) block durations, so 1/2, 1/4, 1/8 of the map run best with 32-, 16-,
    16-row strips (a shorter strip costs ~2000/L more FP64 instructions per pixel in cold solves).
    Same rule as mlm_auto_strip() of the C ABI."""
    strip = 64
    while strip > 16 and nx * max(rows_per_rank, 1) // strip < waves * resident_threads:
        strip //= 2
    return strip


def full_map_layout(nx: int, ny: int) -> tuple[dict, int]:
    """Byte offsets of the five full-map SoA arrays inside rank 0's shared buffer (each 256-byte
    aligned) and the total size: {"A0": off, ...}, nbytes."""
    off, out = 0, {}
    for key, item in KEYS:
        out[key] = off
        off += (nx * ny * item + 255) // 256 * 256
    return out, off


# ---- weighted deal of the strips over the ranks ----------------------------------------------------------
MAX_SEL = 32                      # MLM_DEAL_MAXSEL of the C ABI: positions one rank may own in a round


def root_share_model(world: int, bytes_per_position: float, rate: float = 1.17e10, ingress: float = 720e9,
                     latency_fraction: float = 0.0) -> float:
    """Share of the map the gather root should compute itself.

    Time of the root = f / rate per position; time for the peers' results to enter the root = (1 - f) x bytes /
    ingress; the peers compute (1 - f) / (world - 1) each.  The balanced share equalises the first two (never below
    the equal share 1 / world, never so large that the peers idle while the root still computes).
    rate = positions/s of one GPU on this map in the 16-row strips of a sharded run (1.17e10 on the C4 map; 1.2e10 in
    64-row strips), ingress = bytes/s into one GPU (measured: ~720 GB/s for the map kernel's peer stores with the
    one-byte arrays written 128 bytes at a time, 657 GB/s before that, 790 GB/s for bulk peer copies;
    profiles/r2_ceilings_n8.json, profiles/r2b_gather_sweep.md; measured optimum at 8 GPUs: 3 of every 10 strips, at 4 GPUs
    4 of every 13 -- 2.00 ms per C4 map against 2.16 ms with the equal deal and 2.05 ms with 3 of 9, tools/gpu_runs/run_r4b.sh).
    The coarser deal has to pay: the balanced share is returned when the model predicts a step at least 5 % shorter than
    with the equal deal (max of the root's compute and its peers' bytes), else the equal share."""
    if world <= 1:
        return 1.0
    a, b = 1.0 / rate, bytes_per_position / ingress
    f = (b * (1.0 + latency_fraction)) / (a + b)
    t_equal = max(a / world, (1.0 - 1.0 / world) * b)
    if f <= 1.0 / world or t_equal < 1.05 * a * f:
        return 1.0 / world
    return min(f, 0.9)


def deal_rounds(world: int, root_share: float | None = None, max_period: int = 64) -> tuple[int, tuple]:
    """(period, sel) with sel[r] = ascending positions of a round of `period` strips owned by rank r.

    Rank 0 gets `a` positions, every other rank `b`, with a / (a + (world-1) b) as close as possible to `root_share`
    (None or <= 1/world: the plain cyclic deal, period = world).  The root's positions are spread evenly over the
    round, the others fill the gaps in rank order, so every rank sees the same mix of cheap and expensive rows."""
    if world < 1:
        raise ValueError("world must be >= 1")
    if world == 1:
        return 1, ((0,),)
    if root_share is None or root_share <= 1.0 / world + 1e-9:
        return world, tuple((r,) for r in range(world))
    best = None
    for b in range(1, 9):
        for a in range(b, MAX_SEL + 1):
            period = a + (world - 1) * b
            if period > max_period:
                break
            err = abs(a / period - root_share)
            # the shortest round within 0.02 of the wanted share (short rounds = fine-grained balance), else the closest
            key = (0, period, err) if err <= 0.02 else (1, err, period)
            if best is None or key < best[0]:
                best = (key, a, b, period)
    _, a, b, period = best
    # positions of the root: evenly spread (Bresenham); the other positions go to ranks 1.. in turn
    root_pos = sorted({(i * period) // a for i in range(a)})
    assert len(root_pos) == a
    others = [p for p in range(period) if p not in set(root_pos)]
    sel = [tuple(root_pos)] + [tuple(others[r - 1::world - 1]) for r in range(1, world)]
    assert all(len(x) == b for x in sel[1:]) and sum(len(x) for x in sel) == period
    return period, tuple(sel)


def owned_strips(ny: int, strip: int, period: int, sel) -> list:
    """Strip indices (ascending) of an ny-row map owned under (period, sel)."""
    nstrips = -(-ny // strip)
    pos = set(sel)
    return [k for k in range(nstrips) if k % period in pos]


class _DevMem:
    """Raw device allocation seen through __cuda_array_interface__ (so torch can view it)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class ShardedMap:
    """Device-resident, strip-sharded magnification map with the gather to rank 0 (module docstring).

    `compute()` only enqueues work on the current stream of `device`; after a stream / device
    synchronisation rank 0 holds the whole map (`full_views()` / `assemble_host()`).
    `compute_fn` (default: the CUDA map kernel through the C ABI) and device="cpu" exist so that
    the N > 1 host logic (plan, offsets, chunk loop, gather placement) is testable with gloo.

    layout: "block" = one contiguous, strip-aligned row block per rank; "cyclic" = strip k goes to rank
    k mod world; "weighted" = `deal_rounds(world, root_share)` (root_share None: `root_share_model` for the
    gathered bytes per position).  cyclic / weighted need gather "p2p" or "bulk", where every rank addresses the
    full map.  For a given strip length (`align`) the result is bit-identical for every world size and layout.
    """

    def __init__(self, nx: int, ny: int, rank: int = 0, world: int = 1, device=None, nchunks: int | None = None,
                 group=None, align: int | None = None, gather: str | None = None, compute_fn=None,
                 outputs=None, layout: str | None = None, root_share: float | None = None, rate_hint: float = 1.17e10):
        import torch
        self.torch = torch
        self.nx, self.ny = nx, ny
        # strip length of the map kernel = alignment of the shards.  None: chosen from the per-rank
        # size (auto_strip) -- the same function of (nx, ny, world) on every rank.
        if align is None:
            align = auto_strip(nx, -(-ny // world)) if compute_fn is None else MAP_STRIP
        self.align = align
        self.group = group
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.on_cuda = self.device.type == "cuda"
        if gather is None:
            gather = "none" if world == 1 else ("p2p" if self.on_cuda else "nccl")
        if gather not in ("none", "p2p", "bulk", "nccl"):
            raise ValueError("gather must be 'none', 'p2p', 'bulk' or 'nccl'")
        self.gather = gather
        peer = gather in ("p2p", "bulk") and world > 1
        if layout is None:
            layout = "weighted" if peer else "block"
        if layout not in ("block", "cyclic", "weighted") or (layout != "block" and world > 1 and not peer) or \
                (layout == "block" and gather == "bulk"):
            raise ValueError("layout must be 'block', 'cyclic' or 'weighted' (the last two need gather='p2p'/'bulk'; "
                             "'bulk' needs a dealt layout)")
        self.layout = layout
        # subset of the five output arrays that is stored / gathered (None: all); the kernel takes
        # NULL for the others
        self.outputs = tuple(k for k, _ in KEYS) if outputs is None else tuple(outputs)
        if not self.outputs or set(self.outputs) - {k for k, _ in KEYS}:
            raise ValueError("outputs must be a non-empty subset of A0, A2, A4, nimg, flags")
        self.bytes_per_position = sum(item for k, item in KEYS if k in self.outputs)
        if layout == "weighted" and root_share is None:
            root_share = root_share_model(world, self.bytes_per_position, rate=rate_hint,
                                          ingress=780e9 if gather == "bulk" else 720e9,
                                          latency_fraction=0.12 if gather == "bulk" else 0.0)
        self.root_share = root_share if layout == "weighted" else (1.0 / world)
        self.period, self.sel_all = deal_rounds(world, self.root_share if layout == "weighted" else None)
        self.sel = self.sel_all[rank]
        if nchunks is None:
            nchunks = 4 if (world > 1 and gather in ("nccl", "bulk")) else 1   # chunks only pipeline a separate gather
        self.nchunks = nchunks
        self.plan = plan_map(ny, rank, world, nchunks, align)
        self.compute_fn = compute_fn
        n_slot = self.plan.max_rows * nx
        self.n_slot = n_slot
        self.full = None
        self.comm_stream = None
        self._ipc = None
        self.stage = None
        if peer and not self._setup_p2p():
            # peer mapping is not available on this box (no NVLink / IPC refused): every rank agreed to
            # fall back to the NCCL gather with contiguous blocks
            import warnings
            warnings.warn("ShardedMap: CUDA IPC peer mapping failed, falling back to gather='nccl'")
            gather = self.gather = "nccl"
            self.layout = "block"
            peer = False
            self.plan = plan_map(ny, rank, world, 4, align)
            n_slot = self.n_slot = self.plan.max_rows * nx
        if not peer:
            self.local = self._alloc(n_slot)
            if world > 1 and self.on_cuda:
                with torch.cuda.device(self.device):
                    self.comm_stream = torch.cuda.Stream()
            if world > 1 and rank == 0:
                self.full = [self._alloc(n_slot) if r else self.local for r in range(world)]
        elif gather == "bulk" and rank != 0:
            # local staging of this rank's strips, addressed like the full map (only the owned strips are written)
            self.stage = {key: torch.empty(nx * ny, dtype=torch.float64 if item == 8 else torch.uint8, device=self.device)
                          for key, item in KEYS if key in self.outputs}
            with torch.cuda.device(self.device):
                self.comm_stream = torch.cuda.Stream()

    # ---- buffers ---------------------------------------------------------------------------------
    def _alloc(self, n):
        t = self.torch
        return {key: t.empty(n, dtype=t.float64 if item == 8 else t.uint8, device=self.device) for key, item in KEYS}

    def _setup_p2p(self):
        """Rank 0 allocates the full-map buffer and shares it; `local` becomes this rank's row
        window INSIDE rank 0's memory (peer-mapped on ranks > 0)."""
        import torch.distributed as dist
        from . import _capi
        t, p = self.torch, self.plan
        ctx = _capi.context(self.device.index)
        layout, nbytes = full_map_layout(self.nx, self.ny)
        box, base, err = [None], None, None
        import os
        if os.environ.get("MLM_DISABLE_IPC"):          # test hook: behave like a box without peer mapping
            err = RuntimeError("MLM_DISABLE_IPC")
        try:
            if p.rank == 0 and err is None:
                base = ctx.device_alloc(nbytes)
                box[0] = ctx.ipc_export(base)
        except Exception as e:            # noqa: BLE001 -- reported collectively below
            err = e
        dist.broadcast_object_list(box, src=0, group=self.group)
        if p.rank != 0 and box[0] is not None:
            try:
                base = ctx.ipc_open(box[0])
            except Exception as e:        # noqa: BLE001
                err = e
        # the outcome is collective: either every rank has the mapping or none uses it
        okf = t.tensor([0.0 if (err is not None or base is None) else 1.0], device=self.device)
        dist.all_reduce(okf, op=dist.ReduceOp.MIN, group=self.group)
        if float(okf[0]) < 1.0:
            if base is not None:
                (ctx.device_free if p.rank == 0 else ctx.ipc_close)(base)
            return False
        self._ipc = (ctx, base, p.rank == 0)
        n = self.nx * self.ny
        lo = p.row0 * self.nx if self.layout == "block" else 0
        # launch targets: raw addresses of this rank's row window (block layout) or of the whole map
        # (dealt layouts) inside rank 0's buffer (peer-mapped on ranks > 0, where torch never touches
        # the memory)
        self.map_layout = layout
        self.local = {key: base + layout[key] + lo * item for key, item in KEYS}
        self.full_map = None
        if p.rank == 0:
            raw = t.as_tensor(_DevMem(base, nbytes), device=self.device)
            self.full_map = {key: raw[layout[key]: layout[key] + n * item].view(t.float64 if item == 8 else t.uint8)
                             for key, item in KEYS}
        self._fence = t.zeros(1, dtype=t.float32, device=self.device)
        return True

    def close(self):
        """Release the shared buffer (collective: every rank unmaps before rank 0 frees)."""
        if self._ipc is None:
            return
        import torch.distributed as dist
        ctx, base, owner = self._ipc
        self._ipc = None
        self.local = self.full_map = self.stage = None
        self.torch.cuda.synchronize(self.device)
        if not owner:
            ctx.ipc_close(base)
        dist.barrier(group=self.group)
        if owner:
            ctx.device_free(base)

    # ---- work --------------------------------------------------------------------------------------
    def _launch(self, s, q, rho, Gamma, x0, y0, dx, dy, row0, nrows, views, order, stream):
        if self.compute_fn is not None:
            self.compute_fn(s, q, rho, Gamma, x0, y0, dx, dy, self.nx, row0, nrows, views, order)
            return
        from .batched import map_device
        v = {k: (views[k] if k in self.outputs else None) for k, _ in KEYS}
        map_device(s, q, rho, Gamma, x0, y0, dx, dy, self.nx, row0, nrows, v["A0"], v["A2"], v["A4"],
                   v["nimg"], v["flags"], order=order, stream=stream, device=self.device.index, strip=self.align)

    def _launch_deal(self, s, q, rho, Gamma, x0, y0, dx, dy, row0, nrows, targets, order, stream):
        from .batched import map_deal_device
        v = {k: (targets[k] if k in self.outputs else None) for k, _ in KEYS}
        map_deal_device(s, q, rho, Gamma, x0, y0, dx, dy, self.nx, row0, nrows, self.period, self.sel, v["A0"], v["A2"],
                        v["A4"], v["nimg"], v["flags"], order=order, stream=stream, device=self.device.index,
                        strip=self.align)

    def compute(self, s, q, rho, Gamma, x0, y0, dx, dy, order=4, gather=True, local_only=False):
        """Enqueue this rank's shard (and the gather to rank 0 when world > 1).

        local_only=True: the same strips are computed but stored into this rank's OWN HBM (a scratch map) --
        the kernel-only time of the sharded job, nothing crosses NVLink."""
        import torch.distributed as dist
        t = self.torch
        p, nx = self.plan, self.nx
        cur = t.cuda.current_stream(self.device) if self.on_cuda else None
        stream = cur.cuda_stream if cur is not None else None
        if self.gather in ("p2p", "bulk") and p.world > 1:
            if local_only:
                tgt = self._scratch_targets()
                if self.layout == "block":
                    if p.nrows:
                        self._launch(s, q, rho, Gamma, x0, y0, dx, dy, p.row0, p.nrows,
                                     {k: tgt[k] + p.row0 * nx * item for k, item in KEYS if k in tgt}, order, stream)
                else:
                    self._launch_deal(s, q, rho, Gamma, x0, y0, dx, dy, 0, self.ny, tgt, order, stream)
                return
            if self.gather == "bulk" and p.rank != 0:
                self._compute_bulk(s, q, rho, Gamma, x0, y0, dx, dy, order, cur)
            elif self.layout == "block":
                # one launch; the kernel's stores ARE the gather (peer memory on ranks > 0)
                if p.nrows:
                    self._launch(s, q, rho, Gamma, x0, y0, dx, dy, p.row0, p.nrows, self.local, order, stream)
            else:
                self._launch_deal(s, q, rho, Gamma, x0, y0, dx, dy, 0, self.ny, self.local, order, stream)
            if gather:
                dist.all_reduce(self._fence, group=self.group)      # stream-ordered completion fence
            return
        for (r_off, nr) in p.chunks:
            if nr == 0:
                continue
            o, cnt = r_off * nx, nr * nx
            buf = self.local
            own = max(0, min(r_off + nr, p.nrows) - r_off)      # rows of this chunk this rank owns
            if own:
                oc = own * nx
                self._launch(s, q, rho, Gamma, x0, y0, dx, dy, p.row0 + r_off, own,
                             {k: v[o:o + oc] for k, v in buf.items()}, order, stream)
            if p.world > 1 and gather and self.gather == "nccl" and not local_only:
                if self.comm_stream is not None:
                    self.comm_stream.wait_stream(cur)
                with (t.cuda.stream(self.comm_stream) if self.comm_stream is not None else _null()):
                    for key in self.outputs:
                        src = buf[key][o:o + cnt]
                        dst = [self.full[r][key][o:o + cnt] for r in range(p.world)] if p.rank == 0 else None
                        dist.gather(src, dst, dst=0, group=self.group)
        if p.world > 1 and gather and self.comm_stream is not None and not local_only:
            cur.wait_stream(self.comm_stream)

    def _scratch_targets(self):
        """Full-map-sized LOCAL buffers (allocated on first use) for the kernel-only timing."""
        if getattr(self, "_scratch", None) is None:
            t = self.torch
            src = self.stage if self.stage is not None else None
            self._scratch = src if src is not None else {
                key: t.empty(self.nx * self.ny, dtype=t.float64 if item == 8 else t.uint8, device=self.device)
                for key, item in KEYS if key in self.outputs}
        return {k: v.data_ptr() for k, v in self._scratch.items()}

    def _compute_bulk(self, s, q, rho, Gamma, x0, y0, dx, dy, order, cur):
        """Ranks > 0 of gather='bulk': kernel per row chunk into local staging, then 2-D peer copies of the owned
        strips of that chunk into rank 0's buffer on the copy stream while the next chunk computes."""
        ctx, base, _ = self._ipc
        nx, L, P = self.nx, self.align, self.period
        rows_round = L * P
        nrounds = -(-self.ny // rows_round)
        tgt = {k: v.data_ptr() for k, v in self.stage.items()}
        cs = self.comm_stream
        for c in range(self.nchunks):
            ra, rb = shard_range(nrounds, c, self.nchunks)
            if rb == ra:
                continue
            row0, row1 = ra * rows_round, min(rb * rows_round, self.ny)
            # the kernel addresses its outputs relative to the first row of the launch
            win = {k: v + row0 * nx * item for k, item in KEYS if (v := tgt.get(k)) is not None}
            self._launch_deal(s, q, rho, Gamma, x0, y0, dx, dy, row0, row1 - row0, win,
                              order, cur.cuda_stream)
            cs.wait_stream(cur)
            for key, item in KEYS:
                if key not in self.stage:
                    continue
                sb = L * nx * item                                  # bytes of one strip of this array
                for pos in self.sel:
                    first = row0 + pos * L                          # first row of this position's strip in the chunk
                    if first >= row1:
                        continue
                    # strips first, first + rows_round, ... inside [row0, row1); the last one may be cut by the map end
                    cnt = (row1 - first + rows_round - 1) // rows_round
                    last_rows = min(L, self.ny - (first + (cnt - 1) * rows_round))
                    full_cnt = cnt if last_rows == L else cnt - 1
                    off = first * nx * item
                    if full_cnt > 0:
                        ctx.memcpy2d_async(base + self.map_layout[key] + off, rows_round * nx * item,
                                           self.stage[key].data_ptr() + off, rows_round * nx * item, sb, full_cnt,
                                           cs.cuda_stream)
                    if full_cnt < cnt:
                        off2 = (first + full_cnt * rows_round) * nx * item
                        ctx.memcpy_async(base + self.map_layout[key] + off2, self.stage[key].data_ptr() + off2,
                                         last_rows * nx * item, cs.cuda_stream)
        cur.wait_stream(cs)

    def full_views(self):
        """Rank 0: {key: (ny, nx) tensor} of the whole map without a copy (p2p / bulk / single rank), else None."""
        p = self.plan
        if p.rank != 0:
            return None
        if self.gather in ("p2p", "bulk") and p.world > 1:
            return {k: v.view(self.ny, self.nx) for k, v in self.full_map.items()}
        if p.world == 1:
            return {k: v[: self.ny * self.nx].view(self.ny, self.nx) for k, v in self.local.items()}
        return None

    def assemble_host(self):
        """Rank 0: the full (ny, nx) arrays on the host (numpy), from the gathered shards."""
        import numpy as np
        p, nx = self.plan, self.nx
        if p.rank != 0:
            return None
        fv = self.full_views()
        if fv is not None:
            return {k: v.cpu().numpy() for k, v in fv.items()}
        out = {}
        for key, _ in KEYS:
            parts = []
            for r in range(p.world):
                lo, hi = shard_range(self.ny, r, p.world, self.align)
                src = self.local if r == 0 else self.full[r]
                parts.append(src[key][:(hi - lo) * nx].cpu().numpy().reshape(hi - lo, nx))
            out[key] = np.concatenate(parts, axis=0)
        return out


# ---- host-resident result of a multi-GPU map ---------------------------------------------------------------
class SharedHostMap:
    """ONE (ny, nx) result map in POSIX shared memory, mapped into every rank's process (collective constructor over
    `group`; any torch.distributed backend).  Rank r owns the strip-aligned row block shard_range(ny, r, world, strip);
    it page-locks exactly those rows (mlm_host_register) so that mlm_map_host's D2H copies run at PCIe speed straight
    into the map.  `arrays` are numpy views of the whole map (valid on every rank; rank 0 is the one that returns them)."""

    def __init__(self, nx: int, ny: int, rank: int = 0, world: int = 1, group=None, device=None, outputs=None,
                 strip: int | None = None, register: bool = True):
        import mmap
        import os
        import numpy as np
        self.nx, self.ny, self.rank, self.world, self.group = nx, ny, rank, world, group
        self.outputs = tuple(k for k, _ in KEYS) if outputs is None else tuple(outputs)
        self.strip = auto_strip(nx, -(-ny // world)) if strip is None else int(strip)
        self.lo, self.hi = shard_range(ny, rank, world, self.strip)
        layout, nbytes = {}, 0
        for key, item in KEYS:
            if key in self.outputs:
                layout[key] = nbytes
                nbytes += (nx * ny * item + 4095) // 4096 * 4096
        self.layout, self.nbytes = layout, max(nbytes, 4096)
        box = [None]
        if rank == 0:
            self.path = f"/dev/shm/mlm_map_{os.getpid()}_{id(self) & 0xffffff:x}"
            fd = os.open(self.path, os.O_RDWR | os.O_CREAT | os.O_EXCL, 0o600)
            os.ftruncate(fd, self.nbytes)
            box[0] = self.path
        if world > 1:
            import torch.distributed as dist
            dist.broadcast_object_list(box, src=0, group=group)
        self.path = box[0]
        if rank != 0:
            fd = os.open(self.path, os.O_RDWR)
        self._mm = mmap.mmap(fd, self.nbytes)
        os.close(fd)
        raw = np.frombuffer(self._mm, dtype=np.uint8)
        self._base = raw.ctypes.data
        self.arrays = {key: raw[layout[key]: layout[key] + nx * ny * item].view(np.float64 if item == 8 else np.uint8)
                       .reshape(ny, nx) for key, item in KEYS if key in self.outputs}
        self._registered = []
        self._ctx = None
        if register and self.hi > self.lo:
            from . import _capi
            self._ctx = _capi.context(device)
            for key, item in KEYS:
                if key not in self.outputs:
                    continue
                a = self._base + layout[key] + self.lo * nx * item
                b = self._base + layout[key] + self.hi * nx * item
                a, b = a // 4096 * 4096, (b + 4095) // 4096 * 4096      # whole pages (first touch happens here)
                if self._registered and a < self._registered[-1][1]:     # adjacent arrays may share a page
                    a = self._registered[-1][1]
                if b > a:
                    self._ctx.host_register(a, b - a)
                    self._registered.append((a, b))
        if world > 1:
            import torch.distributed as dist
            dist.barrier(group=group)      # every rank has opened and mapped the file
        if rank == 0:
            os.unlink(self.path)           # the mappings keep the memory alive; nothing is left behind in /dev/shm
        if world > 1:
            dist.barrier(group=group)      # ... and nobody returns before the name is gone

    def own_rows(self):
        """{key: (hi-lo, nx) view} of the rows this rank owns."""
        return {k: v[self.lo:self.hi] for k, v in self.arrays.items()}

    def close(self):
        if self._mm is None:
            return
        for a, _ in self._registered:
            self._ctx.host_unregister(a)
        self._registered = []
        self.arrays = None
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)
        try:
            self._mm.close()
        except BufferError:                 # a caller still holds a view: the mapping dies with it
            pass
        self._mm = None


def magnification_map_sharded(s, q, rho, Gamma, x0, y0, dx, dy, nx, ny, rank=0, world=1, group=None, device=None,
                              order=4, outputs=None, shared: SharedHostMap | None = None):
    """The whole nx x ny map on the HOST, computed by `world` ranks (one GPU each; collective call).

    Every rank computes its strip-aligned row block with the host API (mlm_map_host: kernel chunks overlapped with
    the D2H copies) straight into the shared result map; no data crosses NVLink.  Returns the dict of (ny, nx) numpy
    arrays on rank 0 (views of the shared map; pass `shared` to reuse one across calls), None elsewhere.  The result
    is bit-identical to magnification_map(..., strip=shared.strip) on one GPU."""
    from .batched import magnification_map, Magnification
    own = shared is None
    sh = SharedHostMap(nx, ny, rank, world, group=group, device=device, outputs=outputs) if own else shared
    rows = sh.own_rows()
    if sh.hi > sh.lo:
        magnification_map(s, q, rho, Gamma, x0, y0, dx, dy, nx, ny, row0=sh.lo, nrows=sh.hi - sh.lo, order=order,
                          out=Magnification(rows), device=device, outputs=sh.outputs, strip=sh.strip)
    if world > 1:
        import torch.distributed as dist
        dist.barrier(group=group)
    res = Magnification(sh.arrays) if rank == 0 else None
    if own:
        keep = sh.arrays
        sh.close()
        if rank == 0:
            res = Magnification({k: v.copy() for k, v in keep.items()})   # the mapping is gone after close()
    return res


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class ShardedModels:
    """Model grid (M models x T epochs of a straight trajectory) sharded by contiguous model ranges over
    the ranks (BASELINE.json configs[4]), results on rank 0.

    fit=None : the M x T magnifications (26 B per epoch, or the `outputs` subset) reach rank 0's buffer as
               bulk peer copies of each rank's contiguous model block, chunk-pipelined behind the kernels
               (the light-curve kernel's stores are strided by the segment length: sent over NVLink
               directly they would travel 8 bytes at a time -- measured 2x SLOWER at 8 GPUs than 1).
    fit=(flux, sigma): the light-curve fit is fused into the kernel (mlm_models_chi2); only 8 doubles per
               model travel.
    A model's light curve does not depend on which rank computes it: the segment length of the kernel (`segment`,
    default mlm_auto_segment of the LARGEST share of the grid, like the strip length of a sharded map) is the same on
    every rank, so the result is bit-identical to the single-GPU one computed with that segment (`self.segment`).
    """

    STAT_KEYS = ("chi2", "Fs", "Fb", "SA", "SAA", "SAF", "n5", "nflag")

    def __init__(self, params, tau0: float = 0.0, dtau: float = 0.0, T: int = 0, rank: int = 0, world: int = 1,
                 device=None, group=None, outputs=None, fit=None, times=None, segment: int | None = None):
        import numpy as np
        import torch
        import torch.distributed as dist
        from . import _capi
        self.torch = torch
        self.rank, self.world, self.group = rank, world, group
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        # epochs: the uniform grid tau0 + t dtau, or observation `times` with params = (..., t0, tE)
        self.times = None
        if times is not None:
            th = np.ascontiguousarray(times, dtype=np.float64).ravel()
            T = th.size
            self.times = torch.from_numpy(th).to(self.device)
        self.tau0, self.dtau, self.T = float(tau0), float(dtau), int(T)
        arrs = [np.ascontiguousarray(np.asarray(a, dtype=np.float64)).ravel() for a in params]   # s,q,rho,Gamma,u0,alpha[,t0,tE]
        if len(arrs) != (6 if times is None else 8) or any(a.size != arrs[0].size for a in arrs):
            raise ValueError("params = (s, q, rho, Gamma, u0, alpha) [+ (t0, tE) with times=], arrays of equal length M")
        self.M = arrs[0].size
        self.lo, self.hi = shard_range(self.M, rank, world)
        self.par = [torch.from_numpy(a[self.lo:self.hi].copy()).to(self.device) for a in arrs]
        # epochs per tracking segment: the same on every rank (the bits of a light curve depend on it), chosen for the size
        # of ONE rank's launch (a 1/8 share in the whole grid's 32-epoch segments is 1.06 rounds of resident blocks)
        self.segment = int(segment) if segment else _capi.context(self.device.index).auto_segment(-(-self.M // world), self.T)
        self.outputs = tuple(k for k, _ in KEYS) if outputs is None else tuple(outputs)
        self.fit = None
        ctx = _capi.context(self.device.index)
        if fit is not None:
            flux = np.ascontiguousarray(fit[0], dtype=np.float64).ravel()
            sigma = np.ascontiguousarray(np.broadcast_to(np.asarray(fit[1], dtype=np.float64), flux.shape))
            if flux.size != self.T:
                raise ValueError("fit = (flux, sigma) with T entries")
            self.fit = (torch.from_numpy(flux).to(self.device), torch.from_numpy(1.0 / sigma ** 2).to(self.device))
            layout, nbytes = {"stats": 0}, self.M * 8 * 8
        else:
            layout, nbytes = full_map_layout(self.T, self.M)              # (M, T) arrays, row = model
        nbytes = max(nbytes, 256)
        # rank 0 owns the result buffer; the others map it (CUDA IPC) and address their model rows in it
        box = [None]
        if rank == 0:
            base = ctx.device_alloc(nbytes)
            if world > 1:
                box[0] = ctx.ipc_export(base)
        if world > 1:
            dist.broadcast_object_list(box, src=0, group=group)
            if rank != 0:
                base = ctx.ipc_open(box[0])
        self._buf = (ctx, base, rank == 0)
        self.layout, self.nbytes = layout, nbytes
        self.full = None
        if rank == 0:
            raw = torch.as_tensor(_DevMem(base, nbytes), device=self.device)
            if self.fit is not None:
                self.full = {"stats": raw[: self.M * 64].view(torch.float64).view(self.M, 8)}
            else:
                n = self.M * self.T
                self.full = {key: raw[layout[key]: layout[key] + n * item].view(torch.float64 if item == 8 else torch.uint8)
                             .view(self.M, self.T) for key, item in KEYS}
        self._fence = torch.zeros(1, dtype=torch.float32, device=self.device) if world > 1 else None
        # ranks > 0 compute into a local block and push it to rank 0 in `nchunks` bulk copies
        self.local = None
        self.comm_stream = None
        if rank != 0 and self.fit is None and self.hi > self.lo:
            n = (self.hi - self.lo) * self.T
            self.local = {key: torch.empty(n, dtype=torch.float64 if item == 8 else torch.uint8, device=self.device)
                          for key, item in KEYS if key in self.outputs}
            with torch.cuda.device(self.device):
                self.comm_stream = torch.cuda.Stream()
        # pipelining units of the bulk copies: a chunk must still be several ROUNDS of resident blocks (a block = 64
        # segments of one light curve, 8 blocks resident per SM), or the tail of every chunk costs more than the
        # overlap gains: 1250-model chunks of the C5 grid are 1.06 rounds and ran the 2-GPU job in 2.9 ms, one launch
        # of 5000 models + one copy in ~1.6 ms
        nseg = -(-self.T // max(self.segment, 1))
        blocks = (self.hi - self.lo) * -(-nseg // 64)
        resident = 8 * (torch.cuda.get_device_properties(self.device).multi_processor_count
                        if self.device.type == "cuda" else 148)
        self.nchunks = max(1, min(4, blocks // (4 * resident)))

    def compute(self, order=4):
        """Enqueue this rank's models on the current stream (results land in rank 0's buffer)."""
        import torch.distributed as dist
        from .batched import models_device, models_chi2_device, models_times_device, models_chi2_times_device
        ctx, base, _ = self._buf
        st = self.torch.cuda.current_stream(self.device).cuda_stream
        m = self.hi - self.lo
        dev = self.device.index

        def run(par, nm, A0, A2, A4, ni, fl):
            if self.times is None:
                models_device(*par, nm, self.tau0, self.dtau, self.T, A0, A2, A4, ni, fl, order=order, stream=st, device=dev,
                              segment=self.segment)
            else:
                models_times_device(*par, nm, self.times, self.T, A0, A2, A4, ni, fl, order=order, stream=st, device=dev,
                                    segment=self.segment)

        if m:
            if self.fit is not None:
                if self.times is None:
                    models_chi2_device(*self.par, m, self.tau0, self.dtau, self.T, self.fit[0], self.fit[1],
                                       base + self.lo * 64, order=order, stream=st, device=dev, segment=self.segment)
                else:
                    models_chi2_times_device(*self.par, m, self.times, self.T, self.fit[0], self.fit[1],
                                             base + self.lo * 64, order=order, stream=st, device=dev, segment=self.segment)
            elif self.local is None:
                ptr = {key: (base + self.layout[key] + self.lo * self.T * item if key in self.outputs else None)
                       for key, item in KEYS}
                run(self.par, m, ptr["A0"], ptr["A2"], ptr["A4"], ptr["nimg"], ptr["flags"])
            else:
                cur = self.torch.cuda.current_stream(self.device)
                for c in range(self.nchunks):
                    a, b = shard_range(m, c, self.nchunks)
                    if b == a:
                        continue
                    v = {key: (self.local[key][a * self.T:] if key in self.local else None) for key, _ in KEYS}
                    run([p[a:b] for p in self.par], b - a, v["A0"], v["A2"], v["A4"], v["nimg"], v["flags"])
                    self.comm_stream.wait_stream(cur)
                    for key, item in KEYS:
                        if key in self.local:
                            ctx.memcpy_async(base + self.layout[key] + (self.lo + a) * self.T * item,
                                             self.local[key].data_ptr() + a * self.T * item, (b - a) * self.T * item,
                                             self.comm_stream.cuda_stream)
                cur.wait_stream(self.comm_stream)
        if self.world > 1:
            dist.all_reduce(self._fence, group=self.group)          # stream-ordered completion fence

    def result_host(self):
        """Rank 0: numpy arrays ((M, T) magnifications or the per-model fit statistics); None elsewhere."""
        if self.rank != 0:
            return None
        self.torch.cuda.synchronize(self.device)
        if self.fit is not None:
            st = self.full["stats"].cpu().numpy()
            return {k: st[:, i].copy() for i, k in enumerate(self.STAT_KEYS)}
        return {k: v.cpu().numpy() for k, v in self.full.items() if k in self.outputs}

    def close(self):
        """Collective: every rank unmaps before rank 0 frees."""
        if self._buf is None:
            return
        import torch.distributed as dist
        ctx, base, owner = self._buf
        self._buf = None
        self.full = None
        self.torch.cuda.synchronize(self.device)
        if not owner:
            ctx.ipc_close(base)
        if self.world > 1:
            dist.barrier(group=self.group)
        if owner:
            ctx.device_free(base)
