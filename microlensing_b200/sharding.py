"""Row / model sharding across the GPUs of one box and the final gather.

Every source position is independent (reference multipoles.py:184), so there is no exchange
step on the data path: maps shard by contiguous row blocks, model grids by contiguous model
ranges.  The only collective is the final gather of the SoA outputs to rank 0
(26 B/position: A0, A2, A4 float64 + nimg, flags uint8) over NVLink.  Two implementations:

* gather="p2p" (default on CUDA, world > 1): rank 0 owns the full-map output buffers and shares
  them with the other processes (CUDA IPC, mlm_ipc_*); every rank passes pointers into that buffer
  to the map kernel, whose epilogue stores go straight into rank 0's HBM over NVLink while the
  kernel is still computing -- the gather is fused into the kernel, one launch per rank, and a
  1-element NCCL all-reduce on the compute stream is the completion fence.
* gather="nccl": each rank computes into a local buffer chunk by chunk and torch.distributed
  gathers every chunk to rank 0 on a side stream, overlapping the next chunk's kernel (baseline).

torch / torch.distributed are plumbing here (device buffers, streams, NCCL); the arithmetic is
in libmlmultipoles.so.
"""
from __future__ import annotations

from dataclasses import dataclass

__all__ = ["shard_range", "ShardPlan", "plan_map", "full_map_layout", "auto_strip", "BYTES_PER_POSITION", "KEYS", "ShardedMap",
           "ShardedModels"]

BYTES_PER_POSITION = 26           # 3 x float64 + 2 x uint8
KEYS = (("A0", 8), ("A2", 8), ("A4", 8), ("nimg", 1), ("flags", 1))     # output arrays and their item sizes


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


def auto_strip(nx: int, rows_per_rank: int, resident_threads: int = 148 * 16 * 32, waves: int = 12) -> int:
    """Strip length (power of two, 16..64) that leaves every rank about `waves` waves of work items:
    one work item = one column of one strip; a B200 keeps 148 SMs x 16 warps of them resident.
    Measured on the 8192-wide C4 map (tools/time_shards2.py): a launch of w waves of 64-row strips
    costs about ceil(w + 0.5) block durations, so 1/2, 1/4, 1/8 of the map run best with 32-, 16-,
    16-row strips (a shorter strip costs ~2000/L more FP64 instructions per pixel in cold solves)."""
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


class _DevMem:
    """Raw device allocation seen through __cuda_array_interface__ (so torch can view it)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class ShardedMap:
    """Device-resident, row-sharded magnification map with the gather to rank 0 (module docstring).

    `compute()` only enqueues work on the current stream of `device`; after a stream / device
    synchronisation rank 0 holds the whole map (`full_views()` / `assemble_host()`).
    `compute_fn` (default: the CUDA map kernel through the C ABI) and device="cpu" exist so that
    the N > 1 host logic (plan, offsets, chunk loop, gather placement) is testable with gloo.
    """

    def __init__(self, nx: int, ny: int, rank: int = 0, world: int = 1, device=None, nchunks: int | None = None,
                 group=None, align: int | None = None, gather: str | None = None, compute_fn=None,
                 outputs=None, layout: str | None = None):
        import torch
        self.torch = torch
        self.nx, self.ny = nx, ny
        # strip length of the map kernel = alignment of the shards.  None: chosen from the per-rank
        # size (auto_strip) -- the same function of (nx, ny, world) on every rank.  For a given strip
        # length the sharded map is bit-identical to the unsharded one (Context.set_map_strip).
        if align is None:
            align = auto_strip(nx, -(-ny // world)) if compute_fn is None else MAP_STRIP
        self.align = align
        self.group = group
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.on_cuda = self.device.type == "cuda"
        if gather is None:
            gather = "none" if world == 1 else ("p2p" if self.on_cuda else "nccl")
        if gather not in ("none", "p2p", "nccl"):
            raise ValueError("gather must be 'none', 'p2p' or 'nccl'")
        self.gather = gather
        # how the strips of the map are dealt to the ranks: "block" = one contiguous, strip-aligned row
        # block per rank; "cyclic" = strip k goes to rank k mod world (same mix of cheap and expensive
        # rows on every rank; needs the p2p gather, where every rank addresses the full map)
        if layout is None:
            layout = "cyclic" if (gather == "p2p" and world > 1) else "block"
        if layout not in ("block", "cyclic") or (layout == "cyclic" and world > 1 and gather != "p2p"):
            raise ValueError("layout must be 'block' or 'cyclic' (cyclic needs gather='p2p')")
        self.layout = layout
        if nchunks is None:
            nchunks = 4 if (world > 1 and gather == "nccl") else 1       # chunks only pipeline the NCCL gather
        self.plan = plan_map(ny, rank, world, nchunks, align)
        self.compute_fn = compute_fn
        # subset of the five output arrays that is stored / gathered (None: all); the kernel takes
        # NULL for the others
        self.outputs = tuple(k for k, _ in KEYS) if outputs is None else tuple(outputs)
        if not self.outputs or set(self.outputs) - {k for k, _ in KEYS}:
            raise ValueError("outputs must be a non-empty subset of A0, A2, A4, nimg, flags")
        n_slot = self.plan.max_rows * nx
        self.n_slot = n_slot
        self.full = None
        self.comm_stream = None
        self._ipc = None
        if gather == "p2p" and world > 1 and not self._setup_p2p():
            # peer mapping is not available on this box (no NVLink / IPC refused): every rank agreed to
            # fall back to the NCCL gather with contiguous blocks
            import warnings
            warnings.warn("ShardedMap: CUDA IPC peer mapping failed, falling back to gather='nccl'")
            gather = self.gather = "nccl"
            self.layout = "block"
            self.plan = plan_map(ny, rank, world, 4, align)
            n_slot = self.n_slot = self.plan.max_rows * nx
        if not (gather == "p2p" and world > 1):
            self.local = self._alloc(n_slot)
            if world > 1 and self.on_cuda:
                with torch.cuda.device(self.device):
                    self.comm_stream = torch.cuda.Stream()
            if world > 1 and rank == 0:
                self.full = [self._alloc(n_slot) if r else self.local for r in range(world)]

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
        # (cyclic layout) inside rank 0's buffer (peer-mapped on ranks > 0, where torch never touches
        # the memory)
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
        self.local = self.full_map = None
        self.torch.cuda.synchronize(self.device)
        if not owner:
            ctx.ipc_close(base)
        dist.barrier(group=self.group)
        if owner:
            ctx.device_free(base)

    # ---- work --------------------------------------------------------------------------------------
    def _set_strip(self):
        """Make the context's strip length ours; returns the previous one (restored after the launches
        are enqueued: the strip is read on the host at launch time)."""
        if self.compute_fn is not None:
            return None
        from . import _capi
        ctx = _capi.context(self.device.index)
        old = ctx.map_strip()
        if old != self.align:
            ctx.set_map_strip(self.align)
        return old

    def _restore_strip(self, old):
        if old is not None and old != self.align:
            from . import _capi
            _capi.context(self.device.index).set_map_strip(old)

    def _launch(self, s, q, rho, Gamma, x0, y0, dx, dy, row0, nrows, views, order, stream):
        if self.compute_fn is not None:
            self.compute_fn(s, q, rho, Gamma, x0, y0, dx, dy, self.nx, row0, nrows, views, order)
            return
        from .batched import map_device
        v = {k: (views[k] if k in self.outputs else None) for k, _ in KEYS}
        map_device(s, q, rho, Gamma, x0, y0, dx, dy, self.nx, row0, nrows, v["A0"], v["A2"], v["A4"],
                   v["nimg"], v["flags"], order=order, stream=stream, device=self.device.index)

    def compute(self, s, q, rho, Gamma, x0, y0, dx, dy, order=4, gather=True):
        """Enqueue this rank's shard (and the gather to rank 0 when world > 1)."""
        old = self._set_strip()
        try:
            self._compute(s, q, rho, Gamma, x0, y0, dx, dy, order, gather)
        finally:
            self._restore_strip(old)

    def _compute(self, s, q, rho, Gamma, x0, y0, dx, dy, order, gather):
        import torch.distributed as dist
        t = self.torch
        p, nx = self.plan, self.nx
        cur = t.cuda.current_stream(self.device) if self.on_cuda else None
        stream = cur.cuda_stream if cur is not None else None
        if self.gather == "p2p" and p.world > 1:
            # one launch; the kernel's stores ARE the gather (peer memory on ranks > 0)
            if self.layout == "cyclic":
                from .batched import map_strips_device
                v = {k: (self.local[k] if k in self.outputs else None) for k, _ in KEYS}
                map_strips_device(s, q, rho, Gamma, x0, y0, dx, dy, nx, 0, self.ny, p.rank, p.world, v["A0"], v["A2"],
                                  v["A4"], v["nimg"], v["flags"], order=order, stream=stream,
                                  device=self.device.index)
            elif p.nrows:
                self._launch(s, q, rho, Gamma, x0, y0, dx, dy, p.row0, p.nrows, self.local, order, stream)
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
            if p.world > 1 and gather and self.gather == "nccl":
                if self.comm_stream is not None:
                    self.comm_stream.wait_stream(cur)
                with (t.cuda.stream(self.comm_stream) if self.comm_stream is not None else _null()):
                    for key in self.outputs:
                        src = buf[key][o:o + cnt]
                        dst = [self.full[r][key][o:o + cnt] for r in range(p.world)] if p.rank == 0 else None
                        dist.gather(src, dst, dst=0, group=self.group)
        if p.world > 1 and gather and self.comm_stream is not None:
            cur.wait_stream(self.comm_stream)

    def full_views(self):
        """Rank 0: {key: (ny, nx) tensor} of the whole map without a copy (p2p / single rank), else None."""
        p = self.plan
        if p.rank != 0:
            return None
        if self.gather == "p2p" and p.world > 1:
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
    A model's light curve does not depend on which rank computes it (the segment length of the kernel
    is a function of T only), so the result is bit-identical to the single-GPU one.
    """

    STAT_KEYS = ("chi2", "Fs", "Fb", "SA", "SAA", "SAF", "n5", "nflag")

    def __init__(self, params, tau0: float = 0.0, dtau: float = 0.0, T: int = 0, rank: int = 0, world: int = 1,
                 device=None, group=None, outputs=None, fit=None, times=None):
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
        # pipelining units of the bulk copies: a chunk should still be a few waves of blocks (one block
        # = one model; ~1000 resident blocks per GPU), small launches pay a start-up/drain phase each
        self.nchunks = max(1, min(4, (self.hi - self.lo) // 1000))

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
                models_device(*par, nm, self.tau0, self.dtau, self.T, A0, A2, A4, ni, fl, order=order, stream=st, device=dev)
            else:
                models_times_device(*par, nm, self.times, self.T, A0, A2, A4, ni, fl, order=order, stream=st, device=dev)

        if m:
            if self.fit is not None:
                if self.times is None:
                    models_chi2_device(*self.par, m, self.tau0, self.dtau, self.T, self.fit[0], self.fit[1],
                                       base + self.lo * 64, order=order, stream=st, device=dev)
                else:
                    models_chi2_times_device(*self.par, m, self.times, self.T, self.fit[0], self.fit[1],
                                             base + self.lo * 64, order=order, stream=st, device=dev)
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
