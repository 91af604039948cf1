"""Batched entry points (NEW API: the reference has only an implied Python loop,
"... to be computed for all source center positions", multipoles.py:184).

Every function takes HOST data (numpy), calls the CUDA library through the C ABI
(`*_host` entry points: the host<->device copies happen inside the call, chunked and
overlapped with the kernels) and returns numpy arrays.  `*_device` variants take device
pointers (e.g. torch tensors' data_ptr()) and only enqueue work on a stream.
"""
from __future__ import annotations

import numpy as np

from . import _capi

__all__ = ["Magnification", "magnification_points", "magnification_map", "magnification_models", "models_chi2",
           "models_chi2_device", "models_times_device", "models_chi2_times_device",
           "map_coordinates", "lightcurve_coordinates", "lightcurve_coordinates_at", "alloc_outputs", "ALL_OUTPUTS", "map_device", "map_strips_device", "map_deal_device", "points_device",
           "models_device"]


class Magnification(dict):
    """dict with keys A0, A2, A4 (float64), nimg, flags (uint8); attribute access too."""
    __getattr__ = dict.__getitem__


_OUT_DTYPES = (("A0", np.float64), ("A2", np.float64), ("A4", np.float64), ("nimg", np.uint8), ("flags", np.uint8))
ALL_OUTPUTS = tuple(k for k, _ in _OUT_DTYPES)


def alloc_outputs(shape, device=None, pinned=True, outputs=ALL_OUTPUTS) -> Magnification:
    """Output buffers (page-locked by default so D2H overlaps with compute).  `outputs` selects
    which of A0, A2, A4, nimg, flags are wanted; the others are neither stored by the kernel nor
    copied back (the C ABI takes NULL for them)."""
    bad = set(outputs) - set(ALL_OUTPUTS)
    if bad or not outputs:
        raise ValueError(f"outputs must be a non-empty subset of {ALL_OUTPUTS}")
    ctx = _capi.context(device)
    mk = (lambda dt: ctx.pinned_empty(shape, dt)) if pinned else (lambda dt: np.empty(shape, dt))
    return Magnification({k: mk(dt) for k, dt in _OUT_DTYPES if k in outputs})


def _check_out(out, shape):
    if not any(k in out for k in ALL_OUTPUTS):
        raise ValueError(f"out must hold at least one of {ALL_OUTPUTS}")
    for k, dt in _OUT_DTYPES:
        a = out.get(k)
        if a is None:
            continue
        if a.shape != tuple(shape) or a.dtype != dt or not a.flags.c_contiguous:
            raise ValueError(f"out[{k!r}] must be C-contiguous {dt} of shape {tuple(shape)}")


def _out_ptrs(out):
    return [_capi.ptr(out.get(k)) for k in ALL_OUTPUTS]


def map_coordinates(s, q, x0, y0, dx, dy, nx, ny, row0=0, nrows=None):
    """Primary-frame zeta of every pixel, bit-identical to what the map kernel generates:
    x = x0 + (c + 0.5) dx, y = y0 + (r + 0.5) dy (centre of mass), zeta = x - s q/(1+q) + i y
    (multipoles.py:177-178)."""
    nrows = ny - row0 if nrows is None else nrows
    c = np.arange(nx, dtype=np.float64)
    r = np.arange(row0, row0 + nrows, dtype=np.float64)
    x = (x0 + (c + 0.5) * dx) + (-(s * q / (1. + q)))
    y = y0 + (r + 0.5) * dy
    return x[None, :] + 1j * y[:, None]


def lightcurve_coordinates_at(s, q, u0, alpha, times, t0=0.0, tE=1.0):
    """Primary-frame zeta at observation times: tau = (times - t0) / tE, zeta_cm = (tau + i u0) exp(i alpha),
    same arithmetic as the models kernel (mlm_models_times)."""
    tau = (np.asarray(times, dtype=np.float64) + (-t0)) / tE
    ca, sa = np.cos(alpha), np.sin(alpha)
    zx = (tau * ca + (-(u0 * sa))) + (-(s * q / (1. + q)))
    zy = tau * sa + u0 * ca
    return zx + 1j * zy


def lightcurve_coordinates(s, q, u0, alpha, tau0, dtau, T):
    """Primary-frame zeta of the epochs tau_t = tau0 + t dtau of a straight trajectory,
    zeta_cm = (tau + i u0) exp(i alpha), same arithmetic as the models kernel."""
    t = np.arange(T, dtype=np.float64)
    tau = tau0 + t * dtau
    ca, sa = np.cos(alpha), np.sin(alpha)
    zx = (tau * ca + (-(u0 * sa))) + (-(s * q / (1. + q)))
    zy = tau * sa + u0 * ca
    return zx + 1j * zy


def magnification_points(s, q, rho, Gamma, zeta, order=4, out=None, device=None,
                         outputs=ALL_OUTPUTS) -> Magnification:
    """A0/A2/A4, image count and flags at arbitrary primary-frame positions `zeta` (any shape).

    Per position this is the body of example(), multipoles.py:181-202.
    """
    ctx = _capi.context(device)
    shape = np.shape(zeta)
    zeta = np.ascontiguousarray(zeta, dtype=np.complex128).reshape(shape)   # ascontiguousarray makes 0-d into 1-d
    if out is None:
        out = alloc_outputs(shape, device, pinned=zeta.size >= (1 << 16), outputs=outputs)
    _check_out(out, shape)
    ctx.check(ctx.lib.mlm_points_host(ctx.handle, s, q, rho, Gamma, _capi.ptr(zeta), zeta.size, order,
                                      *_out_ptrs(out)), "mlm_points_host")
    return out


def magnification_map(s, q, rho, Gamma, x0, y0, dx, dy, nx, ny, row0=0, nrows=None, order=4, out=None,
                      device=None, outputs=ALL_OUTPUTS, strip=None) -> Magnification:
    """Rows [row0, row0+nrows) of an nx x ny magnification map (centre-of-mass window, pixel-centre
    sampling); coordinates are generated on the GPU, results land in `out` (nrows, nx).

    `strip` = rows a thread tracks before it starts afresh (a numerical parameter like a tile size, see
    mlm_map in include/mlmultipoles.h).  None: chosen from the size of the WHOLE map (sharding.auto_strip(nx, ny):
    64 rows for 8192^2, 16 for 4096^2 and smaller), so that row blocks of one map computed by separate calls --
    on strip boundaries -- are bit-identical to the map computed at once."""
    import os
    from .sharding import auto_strip
    ctx = _capi.context(device)
    nrows = ny - row0 if nrows is None else nrows
    shape = (nrows, nx)
    if out is None:
        out = alloc_outputs(shape, device, pinned=nrows * nx >= (1 << 16), outputs=outputs)
    _check_out(out, shape)
    if strip is None:
        strip = 0 if os.environ.get("MLM_MAP_STRIP") else auto_strip(nx, ny)      # the env var pins it (experiments)
    ctx.check(ctx.lib.mlm_map_host(ctx.handle, s, q, rho, Gamma, x0, y0, dx, dy, nx, row0, nrows, order, int(strip),
                                   *_out_ptrs(out)), "mlm_map_host")
    return out


def _model_arrays(what, s, q, rho, Gamma, u0, alpha, t0=None, tE=None):
    names = (s, q, rho, Gamma, u0, alpha) if t0 is None and tE is None else \
        (s, q, rho, Gamma, u0, alpha, 0.0 if t0 is None else t0, 1.0 if tE is None else tE)
    arrs = [np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), np.shape(s))) for a in names]
    if not (np.all(arrs[0] > 0) and np.all(arrs[1] > 0)):
        raise ValueError(f"{what}: s and q must be positive")
    if len(arrs) == 8 and not np.all(arrs[7] > 0):
        raise ValueError(f"{what}: tE must be positive")
    return arrs


def magnification_models(s, q, rho, Gamma, u0, alpha, tau0=None, dtau=None, T=None, order=4, out=None, device=None,
                         outputs=ALL_OUTPUTS, times=None, t0=None, tE=None, segment=0) -> Magnification:
    """Batched fit grid: M models (arrays s, q, rho, Gamma, u0, alpha of length M) x T epochs, either on the
    uniform grid tau = tau0 + t dtau, or at observation `times` (any spacing) with per-model (or scalar) time of
    closest approach t0 and Einstein time tE: tau = (times - t0) / tE.
    `segment` = epochs a thread tracks before it starts afresh (numerical parameter, see mlm_models in
    include/mlmultipoles.h; 0: chosen from M and T of this call -- parts of ONE grid computed by separate calls
    are bit-identical to the grid computed at once only if they pass the same segment)."""
    ctx = _capi.context(device)
    if times is None:
        if tau0 is None or dtau is None or T is None:
            raise ValueError("magnification_models: give (tau0, dtau, T) or times=")
        arrs = _model_arrays("magnification_models", s, q, rho, Gamma, u0, alpha)
    else:
        times = np.ascontiguousarray(times, dtype=np.float64).ravel()
        T = times.size
        arrs = _model_arrays("magnification_models", s, q, rho, Gamma, u0, alpha, 0.0 if t0 is None else t0,
                             1.0 if tE is None else tE)
    M = arrs[0].size
    shape = (M, T)
    if out is None:
        out = alloc_outputs(shape, device, pinned=M * T >= (1 << 16), outputs=outputs)
    _check_out(out, shape)
    if times is None:
        ctx.check(ctx.lib.mlm_models_host(ctx.handle, *[_capi.ptr(a) for a in arrs], M, tau0, dtau, T, order, int(segment),
                                          *_out_ptrs(out)), "mlm_models_host")
    else:
        ctx.check(ctx.lib.mlm_models_times_host(ctx.handle, *[_capi.ptr(a) for a in arrs], M, _capi.ptr(times), T, order,
                                                int(segment), *_out_ptrs(out)), "mlm_models_times_host")
    return out


def models_chi2(s, q, rho, Gamma, u0, alpha, tau0=None, dtau=None, flux=None, sigma=None, order=4, device=None,
                times=None, t0=None, tE=None, segment=0) -> dict:
    """Fit statistics of a grid of M models against ONE observed light curve: the model
    magnifications never leave the GPU.  `flux`, `sigma`: observed fluxes and their errors at the epochs
    tau_t = tau0 + t dtau, or at the observation `times` with per-model (or scalar) t0, tE:
    tau = (times - t0) / tE.  Returns arrays of length M: chi2 (minimised over the source and blend fluxes Fs, Fb
    of F = Fs A + Fb), Fs, Fb, the sums SA, SAA, SAF, and the counts n5 (5-image epochs), nflag."""
    ctx = _capi.context(device)
    if flux is None or sigma is None:
        raise ValueError("models_chi2: flux and sigma are required")
    if times is None:
        if tau0 is None or dtau is None:
            raise ValueError("models_chi2: give (tau0, dtau) or times=")
        arrs = _model_arrays("models_chi2", s, q, rho, Gamma, u0, alpha)
    else:
        times = np.ascontiguousarray(times, dtype=np.float64).ravel()
        arrs = _model_arrays("models_chi2", s, q, rho, Gamma, u0, alpha, 0.0 if t0 is None else t0,
                             1.0 if tE is None else tE)
    M = arrs[0].size
    flux = np.ascontiguousarray(flux, dtype=np.float64).ravel()
    if times is not None and times.size != flux.size:
        raise ValueError("models_chi2: times and flux must have the same length")
    sigma = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma, dtype=np.float64), flux.shape))
    if not np.all(sigma > 0):
        raise ValueError("models_chi2: sigma must be positive")
    w = 1.0 / (sigma * sigma)
    stats = np.empty((M, 8), dtype=np.float64)
    if times is None:
        ctx.check(ctx.lib.mlm_models_chi2_host(ctx.handle, *[_capi.ptr(a) for a in arrs], M, tau0, dtau, flux.size, order,
                                               int(segment), _capi.ptr(flux), _capi.ptr(w), _capi.ptr(stats)),
                  "mlm_models_chi2_host")
    else:
        ctx.check(ctx.lib.mlm_models_chi2_times_host(ctx.handle, *[_capi.ptr(a) for a in arrs], M, _capi.ptr(times),
                                                     flux.size, order, int(segment), _capi.ptr(flux), _capi.ptr(w),
                                                     _capi.ptr(stats)),
                  "mlm_models_chi2_times_host")
    keys = ("chi2", "Fs", "Fb", "SA", "SAA", "SAF", "n5", "nflag")
    return Magnification({k: stats[:, i].copy() for i, k in enumerate(keys)})


# ---- device-pointer variants (enqueue only; buffers e.g. torch CUDA tensors) -----------------
def _p(x):
    return _capi.ptr(x)


def points_device(s, q, rho, Gamma, zeta, n, A0, A2, A4, nimg=None, flags=None, order=4, stream=None, device=None,
                  unordered=False):
    """unordered=True: the list has no spatial order -- every entry is solved on its own (mlm_points_unordered: three
    single-lens image guesses + Newton on the lens equation, the root finder on the compacted rest)."""
    ctx = _capi.context(device)
    fn, name = (ctx.lib.mlm_points_unordered, "mlm_points_unordered") if unordered else (ctx.lib.mlm_points, "mlm_points")
    ctx.check(fn(ctx.handle, s, q, rho, Gamma, _p(zeta), n, order, _p(A0), _p(A2), _p(A4), _p(nimg), _p(flags), _p(stream)),
              name)


def map_device(s, q, rho, Gamma, x0, y0, dx, dy, nx, row0, nrows, A0, A2, A4, nimg=None, flags=None, order=4,
               stream=None, device=None, strip=0):
    """strip = 0: the library picks it from the size of this launch (mlm_auto_strip(nx, nrows))."""
    ctx = _capi.context(device)
    ctx.check(ctx.lib.mlm_map(ctx.handle, s, q, rho, Gamma, x0, y0, dx, dy, nx, row0, nrows, order, int(strip),
                              _p(A0), _p(A2), _p(A4), _p(nimg), _p(flags), _p(stream)), "mlm_map")


def map_strips_device(s, q, rho, Gamma, x0, y0, dx, dy, nx, row0, nrows, kphase, kstep, A0, A2, A4, nimg=None,
                      flags=None, order=4, stream=None, device=None, strip=0):
    """Strips k = kphase (mod kstep) of the rows [row0, row0+nrows) only (cyclic sharding over GPUs)."""
    ctx = _capi.context(device)
    ctx.check(ctx.lib.mlm_map_strips(ctx.handle, s, q, rho, Gamma, x0, y0, dx, dy, nx, row0, nrows, kphase, kstep,
                                     order, int(strip), _p(A0), _p(A2), _p(A4), _p(nimg), _p(flags), _p(stream)),
              "mlm_map_strips")


def map_deal_device(s, q, rho, Gamma, x0, y0, dx, dy, nx, row0, nrows, period, sel, A0, A2, A4, nimg=None,
                    flags=None, order=4, stream=None, device=None, strip=0):
    """Strips whose index modulo `period` is in `sel` (ascending) of the rows [row0, row0+nrows): this caller's
    share of a cyclic or weighted deal of the map over GPUs (mlm_map_deal)."""
    import ctypes
    ctx = _capi.context(device)
    arr = (ctypes.c_int32 * len(sel))(*[int(v) for v in sel])
    ctx.check(ctx.lib.mlm_map_deal(ctx.handle, s, q, rho, Gamma, x0, y0, dx, dy, nx, row0, nrows, int(period), len(sel),
                                   arr, order, int(strip), _p(A0), _p(A2), _p(A4), _p(nimg), _p(flags), _p(stream)),
              "mlm_map_deal")


def models_chi2_device(s, q, rho, Gamma, u0, alpha, M, tau0, dtau, T, flux, weight, stats, order=4, stream=None,
                       device=None, segment=0):
    ctx = _capi.context(device)
    ctx.check(ctx.lib.mlm_models_chi2(ctx.handle, _p(s), _p(q), _p(rho), _p(Gamma), _p(u0), _p(alpha), M, tau0, dtau, T,
                                      order, int(segment), _p(flux), _p(weight), _p(stats), _p(stream)), "mlm_models_chi2")


def models_times_device(s, q, rho, Gamma, u0, alpha, t0, tE, M, times, T, A0, A2, A4, nimg=None, flags=None, order=4,
                        stream=None, device=None, segment=0):
    """Model grid at observation times (device arrays; t0 / tE may be None = 0 / 1)."""
    ctx = _capi.context(device)
    ctx.check(ctx.lib.mlm_models_times(ctx.handle, _p(s), _p(q), _p(rho), _p(Gamma), _p(u0), _p(alpha), _p(t0), _p(tE), M,
                                       _p(times), T, order, int(segment), _p(A0), _p(A2), _p(A4), _p(nimg), _p(flags),
                                       _p(stream)),
              "mlm_models_times")


def models_chi2_times_device(s, q, rho, Gamma, u0, alpha, t0, tE, M, times, T, flux, weight, stats, order=4, stream=None,
                             device=None, segment=0):
    ctx = _capi.context(device)
    ctx.check(ctx.lib.mlm_models_chi2_times(ctx.handle, _p(s), _p(q), _p(rho), _p(Gamma), _p(u0), _p(alpha), _p(t0), _p(tE),
                                            M, _p(times), T, order, int(segment), _p(flux), _p(weight), _p(stats),
                                            _p(stream)),
              "mlm_models_chi2_times")


def models_device(s, q, rho, Gamma, u0, alpha, M, tau0, dtau, T, A0, A2, A4, nimg=None, flags=None, order=4,
                  stream=None, device=None, segment=0):
    """segment = 0: the library picks it from M and T of THIS call (mlm_auto_segment)."""
    ctx = _capi.context(device)
    ctx.check(ctx.lib.mlm_models(ctx.handle, _p(s), _p(q), _p(rho), _p(Gamma), _p(u0), _p(alpha), M, tau0, dtau, T,
                                 order, int(segment), _p(A0), _p(A2), _p(A4), _p(nimg), _p(flags), _p(stream)), "mlm_models")
