// mlm_kernels.cu -- CUDA kernels (sm_100a, FP64) and the C ABI of libmlmultipoles.so.
//
// One thread = one source position; everything between the 16-byte coalesced load of zeta
// (or the in-kernel generation of the map / light-curve coordinate) and the coalesced SoA
// stores of A0/A2/A4/nimg/flags lives in registers (mlm_core.cuh).  The workload is bound by
// the FP64 pipe (~4 k DFMA-class instructions per position vs 16-41 bytes of HBM traffic),
// so there is no shared-memory tiling and no tensor-core path; the per-(s,q) table sits in
// the constant bank (kernel parameter) for maps / point lists and in shared memory for the
// batched model grid.
//
// Reference interfaces replaced: see include/mlmultipoles.h.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "mlmultipoles.h"                               /* include/mlmultipoles.h; -I set by build.py */

// -DMLM_STATS (libmlmultipoles_stats.so, built next to the product library and used ONLY by bench.py / tools to
// count what a kernel executes): every MLM_COUNT(event) of mlm_core.cuh adds the number of active lanes to
// g_mlm_cnt[0][event] and 1 to g_mlm_cnt[1][event] (one atomic pair per warp-level event).
#define MLM_CNT_NAMES(X) X(positions) X(laguerre_evals) X(newton_evals) X(lens_evals) X(track_fallbacks) X(cold_solves) \
    X(classify_calls) X(classify_newton) X(nonphys_checks) X(images) X(quadratics) X(anchor_positions) X(watch_near)
enum {
#define X(n) MLM_CNT_##n,
    MLM_CNT_NAMES(X)
#undef X
    MLM_CNT_N
};
#ifdef MLM_STATS
__device__ unsigned long long g_mlm_cnt[2][MLM_CNT_N];
__host__ __device__ __forceinline__ void mlm_count(int idx) {
#if defined(__CUDA_ARCH__)
    const unsigned m = __activemask();
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1) {
        atomicAdd(&g_mlm_cnt[0][idx], (unsigned long long)__popc(m));
        atomicAdd(&g_mlm_cnt[1][idx], 1ULL);
    }
#else
    (void)idx;
#endif
}
#define MLM_COUNT(name) mlm_count(MLM_CNT_##name)
#endif
#include "mlm_core.cuh"
#include "mlm_qkl_generated.cuh"

using namespace mlm;

#define MLM_BLOCK 128
#ifndef MLM_MAP_MINBLOCKS
#define MLM_MAP_MINBLOCKS 4          /* k_map: <= 128 registers -> 4 blocks (16 warps) per SM, no spill */
#endif

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_result(const PointResult& r, int64_t i, double* __restrict__ A0,
                                             double* __restrict__ A2, double* __restrict__ A4,
                                             uint8_t* __restrict__ nimg, uint8_t* __restrict__ flags) {
    if (A0) A0[i] = r.A0;
    if (A2) A2[i] = r.A2;
    if (A4) A4[i] = r.A4;
    if (nimg) nimg[i] = (uint8_t)r.nimg;
    if (flags) flags[i] = (uint8_t)r.flags;
}

// K1a: list of positions, one lens.  A thread owns `seg` CONSECUTIVE list entries: when the list is a
// coherent sequence (a light curve along any trajectory, a scan line) it walks along it with
// warm-started root tracking, the extrapolation scaled by the ratio of successive source steps;
// entries that are far from their predecessor (> 0.02 Einstein radii), or whose tracking is
// rejected, are solved cold, so an unordered list costs one cold solve per entry.  seg is a function of n
// only; the 16-byte loads of a warp are seg entries apart (sector-granular, the path is compute-bound).
// (Lists WITHOUT order go through k_points_direct / k_points_list below.  Visiting them along a Hilbert curve through the
// source plane -- radix sort of curve keys, then this kernel along the sorted order -- was built and measured: slower
// than a cold solve per entry on every list tried except a shuffled light curve, and 2.3-4x slower than the direct path.)
template <int ORDER>
__global__ void __launch_bounds__(MLM_BLOCK, 4)
k_points(const __grid_constant__ LensTable T, const SourceFactors sf, const double2* __restrict__ zeta, int64_t n,
         int seg, double* __restrict__ A0, double* __restrict__ A2,
         double* __restrict__ A4, uint8_t* __restrict__ nimg, uint8_t* __restrict__ flags) {
    __shared__ cd s_roots[2][5][MLM_BLOCK];
    cd* const sz = &s_roots[0][0][threadIdx.x];
    cd* const szp = &s_roots[1][0][threadIdx.x];
    const int64_t i0 = ((int64_t)blockIdx.x * MLM_BLOCK + threadIdx.x) * seg;
    const int64_t i1 = (i0 + seg < n) ? i0 + seg : n;
    TrackState st;
    st.hist = 0; st.nimg = 0; st.extra = 0;
    cd zprev = mk(0.0, 0.0), dprev = mk(0.0, 0.0);
    // every lane runs `seg` iterations (entries past the end are skipped) and the warp RECONVERGES at the two
    // __syncwarp() of an iteration: without them lanes whose cold solve took fewer Laguerre iterations run ahead into
    // the next entries on their own and the warp ends up executing the loop body several times over (measured on
    // an unordered list: 12 of 32 lanes active)
    for (int j = 0; j < seg; ++j) {
        const int64_t iv = i0 + j;
        const bool act = iv < i1;
        PointResult r;
        r.flags = 0;
        int64_t i = 0;
        cd zt = mk(0.0, 0.0), d = mk(0.0, 0.0);
        __syncwarp();
        if (act) {
        i = iv;
        const double2 zz = __ldg(zeta + i);
        zt = mk(zz.x, zz.y);
        d = zt - zprev;
        const double d2 = norm2(d);
        double ratio = 1.0;
        // jump of more than 0.02 Einstein radii (or NaN): start afresh.  Tracking across larger steps often
        // works, but a lane whose attempt fails pays tracking AND cold solve while its warp waits: on an
        // unordered list even 3 % of chance neighbours (threshold 0.1) cost 1.6x.
        if (st.hist >= 1 && !(d2 <= 4e-4)) st.hist = 0;
        if (st.hist >= 2) {
            // projection of this step on the previous one; a sensible extrapolation only for 0 < ratio < 4
            ratio = re_mulc(d, dprev) * rcp(norm2(dprev));
            if (!(ratio > 0.0 && ratio < 4.0)) st.hist = 1;
        }
        MLM_COUNT(positions);
        advance_images(T, zt, sz, szp, nullptr, MLM_BLOCK, st, r.flags, ratio);
        }
        __syncwarp();
        if (act) {
            sum_images<ORDER>(T, sf, sz, MLM_BLOCK, st.nimg, r);
            store_result(r, i, A0, A2, A4, nimg, flags);
            zprev = zt;
            dprev = d;
        }
    }
}

// K1e: list of UNRELATED positions (mlm_points_unordered), one position per thread, nothing carried between entries --
// the result of an entry does not depend on the rest of the list.  First pass (k_points_direct): the three images that
// every binary-lens configuration has, by Newton on the lens equation from single-lens guesses, the other two roots of
// the quintic by Vieta (point_magnification_direct, ~1 300 FP64 instructions with the sums).  What this does not
// finish is set aside -- one atomicAdd per warp and list -- instead of being dealt with on the spot with most lanes
// waiting, and two more launches work through the COMPACTED lists with full warps:
//   list 1 (k_points_list<ORDER, 1>): three images found, but the other two roots are close to the lens equation and
//           need the classification (next to caustics: ~10 % of a planetary light curve);
//   list 0 (k_points_list<ORDER, 0>): the guesses did not converge on three distinct images (inside and next to
//           resonant caustics: ~1 % of a planetary light curve, half of a map window centred on a resonant caustic):
//           root finder (Laguerre + deflation + polish, ~4 600 warp-level instructions per entry).
// lists[0 .. n) is list 0, lists[n .. 2n) list 1; counters[0], counters[1] their lengths.
template <int ORDER>
__global__ void __launch_bounds__(MLM_BLOCK, 4)
k_points_direct(const __grid_constant__ LensTable T, const SourceFactors sf, const double2* __restrict__ zeta, int64_t n,
                uint32_t* __restrict__ lists, unsigned* __restrict__ counters, double* __restrict__ A0,
                double* __restrict__ A2, double* __restrict__ A4, uint8_t* __restrict__ nimg, uint8_t* __restrict__ flags) {
    __shared__ cd s_roots[5][MLM_BLOCK];
    cd* const sz = &s_roots[0][threadIdx.x];
    const int64_t i = (int64_t)blockIdx.x * MLM_BLOCK + threadIdx.x;
    const bool act = i < n;
    int status = MLM_GUESS_THREE;
    PointResult r;
    r.flags = 0;
    if (act) {
        const double2 zz = __ldg(zeta + i);
        MLM_COUNT(positions);
        status = point_magnification_direct<ORDER>(T, sf, mk(zz.x, zz.y), sz, MLM_BLOCK, r, false);
        if (status == MLM_GUESS_THREE) store_result(r, i, A0, A2, A4, nimg, flags);
    }
    __syncwarp();
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        const bool mine = status == (l ? MLM_GUESS_CLASSIFY : MLM_GUESS_FAILED);
        const unsigned m = __ballot_sync(0xffffffffu, mine);
        if (m) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(counters + l, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (mine) lists[(size_t)l * (size_t)n + base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
        }
    }
}

template <int ORDER, int MODE>
__global__ void __launch_bounds__(MLM_BLOCK, 4)
k_points_list(const __grid_constant__ LensTable T, const SourceFactors sf, const double2* __restrict__ zeta, int64_t n,
              const uint32_t* __restrict__ lists, const unsigned* __restrict__ counters, double* __restrict__ A0,
              double* __restrict__ A2, double* __restrict__ A4, uint8_t* __restrict__ nimg, uint8_t* __restrict__ flags) {
    __shared__ cd s_roots[5][MLM_BLOCK];
    cd* const sz = &s_roots[0][threadIdx.x];
    const unsigned cnt = __ldg(counters + MODE);
    const uint32_t* const list = lists + (size_t)MODE * (size_t)n;
    for (unsigned j0 = blockIdx.x * MLM_BLOCK; j0 < cnt; j0 += gridDim.x * MLM_BLOCK) {
        const unsigned j = j0 + threadIdx.x;
        __syncwarp();
        if (j >= cnt) continue;
        const int64_t i = (int64_t)__ldg(list + j);
        const double2 zz = __ldg(zeta + i);
        const cd zt = mk(zz.x, zz.y);
        PointResult r;
        r.flags = 0;
        MLM_COUNT(positions);
        if (MODE == 0 || point_magnification_direct<ORDER>(T, sf, zt, sz, MLM_BLOCK, r, true) == MLM_GUESS_FAILED) {
            r.flags = 0;
            TrackState st;
            st.hist = 0; st.nimg = 0; st.extra = 0;
            advance_images(T, zt, sz, sz, nullptr, MLM_BLOCK, st, r.flags);
            sum_images<ORDER>(T, sf, sz, MLM_BLOCK, st.nimg, r);
        }
        store_result(r, i, A0, A2, A4, nimg, flags);
    }
}

// K1b: magnification map, zeta generated in-kernel (no input traffic).
// blockIdx.x tiles the columns; blockIdx.y strides over STRIPS of `strip` consecutive rows
// (aligned to absolute row numbers, so any row-sharding on multiples of `strip` reproduces the
// single-GPU arithmetic bit for bit).  A thread walks down its column inside the strip: the
// first row is a cold solve (Laguerre + deflation), every further row warm-starts Newton from
// the roots of the row above (pixel pitch ~2e-4 theta_E) and falls back to the cold solve when
// the tracking check fails (next to caustics).  Stores are coalesced across the 128 columns.
//
// Which strips a launch owns is described by a DEAL: the strips of the map are dealt in rounds of
// `period` strips, and the launch owns the positions sel[0..nsel) of every round (period = 1, nsel = 1:
// all strips; period = world, sel = {rank}: cyclic sharding; unequal nsel per rank: weighted sharding,
// e.g. the gather root takes more strips because its results do not cross NVLink).
// ANCHORS: the roots at the first row of every owned strip, found by tracking ALONG that row (k_anchor, a small
// launch before k_map): a thread walks `aseg` consecutive columns of one anchor row, so only one pixel in `aseg` is a
// cold solve (Laguerre + deflation, ~3000 FP64 instructions against ~400 for a tracked step without the image sums).
// k_map then starts its strips from the stored roots.  Segments are aligned to absolute column numbers and anchor rows
// to absolute strip numbers, so sharding does not change a bit.  Layout: roots[(ii * 5 + k) * nx + col] (double2),
// meta[ii * nx + col] = flags | nimg << 8 | extra << 11 | valid << 13, ii = ownership-relative strip index of the launch.
#define MLM_ANCHOR_VALID 0x2000u
__global__ void __launch_bounds__(MLM_BLOCK, MLM_MAP_MINBLOCKS)
k_anchor(const __grid_constant__ LensTable T, double x0, double y0, double dx, double dy, double shift, int64_t nx,
         int64_t row0, int strip, const __grid_constant__ Deal deal, int64_t i_count, int aseg,
         double2* __restrict__ roots, uint16_t* __restrict__ meta) {
    __shared__ cd s_roots[3][5][MLM_BLOCK];
    cd* const sz = &s_roots[0][0][threadIdx.x];
    cd* const szp = &s_roots[1][0][threadIdx.x];
    cd* const szpp = &s_roots[2][0][threadIdx.x];
    const int64_t c0 = ((int64_t)blockIdx.x * MLM_BLOCK + threadIdx.x) * aseg;
    const int64_t c1 = (c0 + aseg < nx) ? c0 + aseg : nx;
    for (int64_t ii = blockIdx.y; ii < i_count; ii += gridDim.y) {
        const int64_t k = deal_strip(deal, deal.t0 + ii);
        const int64_t ra = (k * strip > row0) ? k * strip : row0;
        const double y = __dadd_rn(y0, __dmul_rn((double)ra + 0.5, dy));
        TrackState st;
        st.hist = 0; st.nimg = 0; st.extra = 0;
        for (int j0 = 0; j0 < aseg; ++j0) {
            const int64_t c = c0 + j0;
            __syncwarp();                                // whole warps take every column step together
            if (c >= c1) continue;
            const double x = __dadd_rn(__dadd_rn(x0, __dmul_rn((double)c + 0.5, dx)), -shift);
            unsigned fl = 0;
            MLM_COUNT(anchor_positions);
            advance_images(T, mk(x, y), sz, szp, szpp, MLM_BLOCK, st, fl);
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const cd zj = sz[j * MLM_BLOCK];
                roots[(ii * 5 + j) * nx + c] = make_double2(zj.x, zj.y);
            }
            meta[ii * nx + c] = (uint16_t)((fl & 0xffu) | ((unsigned)st.nimg << 8) | ((unsigned)(st.extra & 3) << 11) |
                                           (st.hist >= 1 ? MLM_ANCHOR_VALID : 0u));
        }
    }
}

template <int ORDER>
__global__ void __launch_bounds__(MLM_BLOCK, MLM_MAP_MINBLOCKS)
k_map(const __grid_constant__ LensTable T, const SourceFactors sf, double x0, double y0, double dx, double dy,
      double shift, int64_t nx, int64_t row0, int64_t nrows, int strip, const __grid_constant__ Deal deal,
      int64_t i_count, int64_t i_axis, const double2* __restrict__ aroots, const uint16_t* __restrict__ ameta,
      int u8rows, double* __restrict__ A0,
      double* __restrict__ A2, double* __restrict__ A4, uint8_t* __restrict__ nimg, uint8_t* __restrict__ flags) {
    // u8rows (outputs in ANOTHER GPU's memory): the one-byte arrays are staged in shared memory for four rows and
    // written 128 contiguous bytes per row by one warp instead of 32 bytes per warp and row -- 32-byte stores are what
    // holds the kernel's peer stores at ~670 GB/s over NVLink when 128-byte ones reach ~720 (tools/peer_store_probe.py)
    __shared__ uint8_t s_u8[2][4][MLM_BLOCK];
    // root store of this thread: current and previous roots, [k][thread] so that a warp touches
    // 512 contiguous bytes per k (conflict-free); keeps 20 doubles out of the register file
    __shared__ cd s_roots[3][5][MLM_BLOCK];
    cd* const sz = &s_roots[0][0][threadIdx.x];
    cd* const szp = &s_roots[1][0][threadIdx.x];
    cd* const szpp = &s_roots[2][0][threadIdx.x];
    const int64_t col = (int64_t)blockIdx.x * MLM_BLOCK + threadIdx.x;
    const bool act = col < nx;                         // lanes past the last column idle but keep the warp whole
    // x = x0 + (col + 0.5) dx with two roundings (no FMA), so that host code can reproduce
    // the coordinates bit for bit with the same expression.
    const double x = __dadd_rn(__dadd_rn(x0, __dmul_rn((double)col + 0.5, dx)), -shift);
    // The owned strips (ownership index t = deal.t0 + ii, ii = 0 .. i_count-1, strip k = (t / nsel) period +
    // sel[t % nsel]) are visited from the lens axis (y = 0, index i_axis) outwards, alternating sides: the
    // caustics -- the 5-image, i.e. expensive, pixels -- sit on the axis, so the heavy blocks are
    // scheduled first and the tail of the launch is made of cheap ones.  Pure scheduling order:
    // the arithmetic of a strip does not depend on it.
    const int64_t n_lo = i_axis, n_hi = i_count - 1 - i_axis, n_min = n_lo < n_hi ? n_lo : n_hi;
    for (int64_t j = blockIdx.y; j < i_count; j += gridDim.y) {
        int64_t ii;
        if (j <= 2 * n_min) ii = (j & 1) ? i_axis + (j + 1) / 2 : i_axis - j / 2;
        else ii = (n_hi > n_lo) ? i_axis + (j - n_min) : i_axis - (j - n_min);
        const int64_t k = deal_strip(deal, deal.t0 + ii);
        const int64_t ra = (k * strip > row0) ? k * strip : row0;
        const int64_t rb = ((k + 1) * strip < row0 + nrows) ? (k + 1) * strip : row0 + nrows;
        TrackState st;
        st.hist = 0; st.nimg = 0; st.extra = 0;
        // the strip starts from its anchor (roots found by k_anchor at exactly this pixel) when there is one
        unsigned anchor = 0;
        if (ameta != nullptr && act) {
            anchor = __ldg(ameta + ii * nx + col);
            if (anchor & MLM_ANCHOR_VALID) {
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const double2 zj = __ldg(aroots + (ii * 5 + j) * nx + col);
                    sz[j * MLM_BLOCK] = mk(zj.x, zj.y);
                }
                st.hist = 1; st.nimg = (int)((anchor >> 8) & 7u); st.extra = (int)((anchor >> 11) & 3u);
            }
        }
        for (int64_t r = ra; r < rb; ++r) {
            const double y = __dadd_rn(y0, __dmul_rn((double)r + 0.5, dy));
            const cd zeta = mk(x, y);
            PointResult res;
            res.flags = 0;
            __syncwarp();                                // the lanes of a warp take every row together
            if (act) {
                MLM_COUNT(positions);
                if (r == ra && (anchor & MLM_ANCHOR_VALID)) res.flags = anchor & 0xffu;     // the anchor IS this pixel's solution
                else advance_images(T, zeta, sz, szp, szpp, MLM_BLOCK, st, res.flags);
            }
            __syncwarp();
            if (act) {
                sum_images<ORDER>(T, sf, sz, MLM_BLOCK, st.nimg, res);
                store_result(res, (r - row0) * nx + col, A0, A2, A4, u8rows ? nullptr : nimg, u8rows ? nullptr : flags);
            }
            if (u8rows) {                                // block-uniform
                const int jr = (int)((r - ra) & 3);
                s_u8[0][jr][threadIdx.x] = (uint8_t)res.nimg;
                s_u8[1][jr][threadIdx.x] = (uint8_t)res.flags;
                if (jr == 3 || r == rb - 1) {
                    __syncthreads();
                    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
                    if (w <= jr) {                       // warp w writes row r - jr + w of this block's 128 columns
                        const int64_t base = (r - jr + w - row0) * nx + (int64_t)blockIdx.x * MLM_BLOCK;
                        if (nimg) ((uint32_t*)(nimg + base))[l] = ((const uint32_t*)s_u8[0][w])[l];
                        if (flags) ((uint32_t*)(flags + base))[l] = ((const uint32_t*)s_u8[1][w])[l];
                    }
                    __syncthreads();
                }
            }
        }
    }
}

// K1c: batched model grid / light curves.  Each thread owns a SEGMENT of `seg` consecutive epochs of a model's
// straight trajectory and walks along it with warm-started root tracking (cold solve at the segment start and
// wherever tracking is rejected), exactly like a map column in k_map.  A block of 64 threads serves G models at a
// time, P = 64 / G threads each (P = 64 when a light curve has 64 segments or more, else the next power of two >=
// the number of segments: short light curves of a big grid share a block instead of leaving lanes idle); the
// models' tables sit in shared memory.  seg is an argument of the call (like the strip length of a map): the bits
// of a light curve depend on it, not on how the grid is sharded over ranks.
#define MLM_MODELS_BLOCK 64
#ifndef MLM_MODELS_MINBLOCKS
#define MLM_MODELS_MINBLOCKS 8      /* k_models on a uniform grid: 128 registers -> 8 blocks (16 warps) per SM, no spill */
#endif
#ifndef MLM_CHI2_MINBLOCKS
#define MLM_CHI2_MINBLOCKS 6        /* observation times (four doubles more per thread) and k_models_chi2 (six accumulators more) */
#endif
#define MLM_MODELS_GMAX 16                            /* models per block: P >= 4 threads per model */

// Epochs of the light curves of a model grid: tau_t = tau0 + t dtau (uniform grid, times == NULL) or
// tau_t = (times[t] - t0[m]) / tE[m] (observation times; t0, tE per model, NULL = 0 and 1).
struct Epochs {
    double tau0, dtau;
    const double* times;
    const double* t0;
    const double* tE;
};

__device__ __forceinline__ double epoch_tau(const Epochs& ep, int64_t t, double t0m, double tEm) {
    if (ep.times) return __ddiv_rn(__dadd_rn(__ldg(ep.times + t), -t0m), tEm);
    return __dadd_rn(ep.tau0, __dmul_rn((double)t, ep.dtau));
}

// One epoch of a light-curve walk: unevenly spaced epochs scale the extrapolation by the ratio of successive
// steps in tau (the source moves at unit speed in tau), a step of more than 0.02 restarts cold.
__device__ __forceinline__ double epoch_ratio(const Epochs& ep, double tau, double& tau_prev, double& dt_prev,
                                              TrackState& st) {
    double ratio = 1.0;
    if (ep.times) {
        const double dt = tau - tau_prev;
        if (st.hist >= 1 && !(fabs(dt) <= 0.02)) st.hist = 0;
        if (st.hist >= 2) {
            ratio = dt / dt_prev;
            if (!(ratio > 0.0 && ratio < 4.0)) st.hist = 1;
        }
        tau_prev = tau;
        dt_prev = dt;
    }
    return ratio;
}

// per-model constants of a block's models (shared memory)
struct ModelSlot {
    LensTable T;
    SourceFactors sf;
    double ca, sa, shift;                              // cos(alpha), sin(alpha), s q / (1 + q)
};

__device__ __forceinline__ void stage_model(ModelSlot* slot, const double* s, const double* q, const double* rho,
                                            const double* Gamma, const double* alpha, int64_t m) {
    const double sm = s[m], qm = q[m];
    make_table(sm, qm, &slot->T);
    slot->sf = source_factors(rho[m], Gamma[m]);
    double sa, ca;
    sincos(alpha[m], &sa, &ca);
    slot->ca = ca; slot->sa = sa; slot->shift = sm * qm / (1.0 + qm);
}

template <int ORDER, bool TIMES>
__global__ void __launch_bounds__(MLM_MODELS_BLOCK, TIMES ? MLM_CHI2_MINBLOCKS : MLM_MODELS_MINBLOCKS)
k_models(const double* __restrict__ s, const double* __restrict__ q, const double* __restrict__ rho,
         const double* __restrict__ Gamma, const double* __restrict__ u0, const double* __restrict__ alpha,
         int64_t M, const __grid_constant__ Epochs ep, int64_t Tn, int seg, int P, double* __restrict__ A0,
         double* __restrict__ A2, double* __restrict__ A4, uint8_t* __restrict__ nimg, uint8_t* __restrict__ flags) {
    __shared__ ModelSlot slots[MLM_MODELS_GMAX];
    __shared__ cd s_roots[3][5][MLM_MODELS_BLOCK];
    cd* const sz = &s_roots[0][0][threadIdx.x];
    cd* const szp = &s_roots[1][0][threadIdx.x];
    cd* const szpp = &s_roots[2][0][threadIdx.x];
    const int64_t nseg = (Tn + seg - 1) / seg;
    const int G = MLM_MODELS_BLOCK / P;                // models per block
    const int g = threadIdx.x / P, lane = threadIdx.x % P;
    const int64_t nmb = (M + G - 1) / G;               // model groups
    for (int64_t mb = blockIdx.y; mb < nmb; mb += gridDim.y) {
        __syncthreads();
        if (threadIdx.x < G && mb * G + threadIdx.x < M) stage_model(&slots[threadIdx.x], s, q, rho, Gamma, alpha, mb * G + threadIdx.x);
        __syncthreads();
        const int64_t m = mb * G + g;
        const bool have_model = m < M;
        const ModelSlot& ms = slots[have_model ? g : 0];
        const double u = have_model ? u0[m] : 0.0;
        const double ur1 = __dmul_rn(u, ms.sa), ur0 = __dmul_rn(u, ms.ca);
        const double t0m = (TIMES && have_model && ep.t0) ? ep.t0[m] : 0.0, tEm = (TIMES && have_model && ep.tE) ? ep.tE[m] : 1.0;
        // block-uniform trip counts (idle lanes skip the bodies) so that the warps can reconverge at every epoch
        for (int64_t sbase = (int64_t)blockIdx.x * P; sbase < nseg; sbase += (int64_t)gridDim.x * P) {
            const int64_t sg = sbase + lane;
            const int64_t t0 = sg * seg;
            const int64_t t1 = (!have_model || sg >= nseg) ? t0 : ((t0 + seg < Tn) ? t0 + seg : Tn);
            TrackState st;
            st.hist = 0; st.nimg = 0; st.extra = 0;
            double tau_prev = 0.0, dt_prev = 1.0;
            for (int jt = 0; jt < seg; ++jt) {
                const int64_t t = t0 + jt;
                const bool act = t < t1;
                PointResult res;
                res.flags = 0;
                __syncwarp();                            // the lanes of a warp take every epoch together ...
                if (act) {
                    // compiler barrier: the table is re-read from shared memory where it is used instead of
                    // being hoisted out of the loop into ~40 live registers
                    asm volatile("" ::: "memory");
                    const double tau = TIMES ? epoch_tau(ep, t, t0m, tEm) : __dadd_rn(ep.tau0, __dmul_rn((double)t, ep.dtau));
                    const double ratio = TIMES ? epoch_ratio(ep, tau, tau_prev, dt_prev, st) : 1.0;
                    // zeta_cm = (tau + i u0) (cos a + i sin a), then shift to the primary frame
                    const double zx = __dadd_rn(__dadd_rn(__dmul_rn(tau, ms.ca), -ur1), -ms.shift);
                    const double zy = __dadd_rn(__dmul_rn(tau, ms.sa), ur0);
                    MLM_COUNT(positions);
                    advance_images(ms.T, mk(zx, zy), sz, szp, szpp, MLM_MODELS_BLOCK, st, res.flags, ratio);
                }
                __syncwarp();                            // ... and meet again before the image sums
                if (!act) continue;
                sum_images<ORDER>(ms.T, ms.sf, sz, MLM_MODELS_BLOCK, st.nimg, res);
                store_result(res, m * Tn + t, A0, A2, A4, nimg, flags);
            }
        }
    }
}

// K1d: model grid with the light-curve fit fused in.  The P threads of a model walk their
// segments of the model light curve exactly like k_models, but instead of storing the magnifications they
// accumulate the weighted sums of the linear flux fit F_t = Fs A_t + Fb against observed fluxes f_t with weights
// w_t = 1/sigma_t^2; a fixed-order tree reduction over the P threads makes the result deterministic.  Output per
// model (8 doubles):
//   chi2 (minimised over Fs, Fb), Fs, Fb, sum w A, sum w A^2, sum w A f, #epochs with 5 images, #flagged epochs
template <int ORDER, bool TIMES>
__global__ void __launch_bounds__(MLM_MODELS_BLOCK, MLM_CHI2_MINBLOCKS)
k_models_chi2(const double* __restrict__ s, const double* __restrict__ q, const double* __restrict__ rho,
              const double* __restrict__ Gamma, const double* __restrict__ u0, const double* __restrict__ alpha,
              int64_t M, const __grid_constant__ Epochs ep, int64_t Tn, int seg, int P, const double* __restrict__ flux,
              const double* __restrict__ weight, double* __restrict__ stats) {
    __shared__ ModelSlot slots[MLM_MODELS_GMAX];
    __shared__ cd s_roots[3][5][MLM_MODELS_BLOCK];
    __shared__ double s_red[8][MLM_MODELS_BLOCK];
    cd* const sz = &s_roots[0][0][threadIdx.x];
    cd* const szp = &s_roots[1][0][threadIdx.x];
    cd* const szpp = &s_roots[2][0][threadIdx.x];
    const int64_t nseg = (Tn + seg - 1) / seg;
    const int G = MLM_MODELS_BLOCK / P;
    const int g = threadIdx.x / P, lane = threadIdx.x % P;
    const int64_t nmb = (M + G - 1) / G;
    for (int64_t mb = blockIdx.x; mb < nmb; mb += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < G && mb * G + threadIdx.x < M) stage_model(&slots[threadIdx.x], s, q, rho, Gamma, alpha, mb * G + threadIdx.x);
        __syncthreads();
        const int64_t m = mb * G + g;
        double a_w = 0.0, a_wa = 0.0, a_waa = 0.0, a_wf = 0.0, a_waf = 0.0, a_wff = 0.0, a_n5 = 0.0, a_nf = 0.0;
        {
            const bool have_model = m < M;
            const ModelSlot& ms = slots[have_model ? g : 0];
            const double u = have_model ? u0[m] : 0.0;
            const double ur1 = __dmul_rn(u, ms.sa), ur0 = __dmul_rn(u, ms.ca);
            const double t0m = (TIMES && have_model && ep.t0) ? ep.t0[m] : 0.0, tEm = (TIMES && have_model && ep.tE) ? ep.tE[m] : 1.0;
            for (int64_t sbase = 0; sbase < nseg; sbase += P) {      // block-uniform trip counts, see k_models
                const int64_t sg = sbase + lane;
                const int64_t t0 = sg * seg;
                const int64_t t1 = (!have_model || sg >= nseg) ? t0 : ((t0 + seg < Tn) ? t0 + seg : Tn);
                TrackState st;
                st.hist = 0; st.nimg = 0; st.extra = 0;
                double tau_prev = 0.0, dt_prev = 1.0;
                for (int jt = 0; jt < seg; ++jt) {
                    const int64_t t = t0 + jt;
                    const bool act = t < t1;
                    PointResult res;
                    res.flags = 0;
                    __syncwarp();
                    if (act) {
                        asm volatile("" ::: "memory");
                        const double tau = TIMES ? epoch_tau(ep, t, t0m, tEm) : __dadd_rn(ep.tau0, __dmul_rn((double)t, ep.dtau));
                        const double ratio = TIMES ? epoch_ratio(ep, tau, tau_prev, dt_prev, st) : 1.0;
                        const double zx = __dadd_rn(__dadd_rn(__dmul_rn(tau, ms.ca), -ur1), -ms.shift);
                        const double zy = __dadd_rn(__dmul_rn(tau, ms.sa), ur0);
                        MLM_COUNT(positions);
                        advance_images(ms.T, mk(zx, zy), sz, szp, szpp, MLM_MODELS_BLOCK, st, res.flags, ratio);
                    }
                    __syncwarp();
                    if (!act) continue;
                    sum_images<ORDER>(ms.T, ms.sf, sz, MLM_MODELS_BLOCK, st.nimg, res);
                    const double A = (ORDER >= 4) ? res.A4 : res.A2;
                    const double w = __ldg(weight + t), f = __ldg(flux + t);
                    const double wa = w * A;
                    a_w += w; a_wa += wa; a_waa = fma(wa, A, a_waa);
                    a_wf = fma(w, f, a_wf); a_waf = fma(wa, f, a_waf); a_wff = fma(w * f, f, a_wff);
                    a_n5 += (res.nimg == 5) ? 1.0 : 0.0;
                    a_nf += (res.flags != 0) ? 1.0 : 0.0;
                }
            }
        }
        // fixed-order tree reduction over the P threads of every model of the block
        s_red[0][threadIdx.x] = a_w;  s_red[1][threadIdx.x] = a_wa;  s_red[2][threadIdx.x] = a_waa;
        s_red[3][threadIdx.x] = a_wf; s_red[4][threadIdx.x] = a_waf; s_red[5][threadIdx.x] = a_wff;
        s_red[6][threadIdx.x] = a_n5; s_red[7][threadIdx.x] = a_nf;
        __syncthreads();
        for (int o = P / 2; o > 0; o >>= 1) {
            if (lane < o) {
#pragma unroll
                for (int k = 0; k < 8; ++k) s_red[k][threadIdx.x] += s_red[k][threadIdx.x + o];
            }
            __syncthreads();
        }
        if (lane == 0 && m < M) {
            const int b = threadIdx.x;
            const double S = s_red[0][b], SA = s_red[1][b], SAA = s_red[2][b], SF = s_red[3][b], SAF = s_red[4][b],
                         SFF = s_red[5][b];
            const double det = S * SAA - SA * SA;
            const double Fs = (S * SAF - SA * SF) / det;
            const double Fb = (SAA * SF - SA * SAF) / det;
            double* o = stats + 8 * m;
            o[0] = SFF - Fs * SAF - Fb * SF;
            o[1] = Fs; o[2] = Fb; o[3] = SA; o[4] = SAA; o[5] = SAF; o[6] = s_red[6][b]; o[7] = s_red[7][b];
        }
    }
}

// K3: roots only (drop-in for _solvelenseq)
__global__ void __launch_bounds__(MLM_BLOCK)
k_solve(const __grid_constant__ LensTable T, const double2* __restrict__ zeta, int64_t n, double2* __restrict__ roots) {
    const int64_t i = (int64_t)blockIdx.x * MLM_BLOCK + threadIdx.x;
    if (i >= n) return;
    const double2 zz = __ldg(zeta + i);
    cd c[6], z[5];
    unsigned fl = 0;
    lens_polynomial(T, mk(zz.x, zz.y), c);
    solve_roots(c, T.s, z, fl);
#pragma unroll
    for (int k = 0; k < 5; ++k) roots[5 * i + k] = make_double2(z[k].x, z[k].y);
}

// K2: literal quadrupole / hexadecapole on a user-supplied (7, n) Wk table.
// Per-block partial sums (fixed tree order -> deterministic), reduced by k_from_wk_final.
template <int ORDER>
__global__ void __launch_bounds__(MLM_BLOCK)
k_from_wk(const double2* __restrict__ Wk, int64_t n, double h2, double h4, double* __restrict__ partial) {
    double a0 = 0.0, a2 = 0.0, a4 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * MLM_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * MLM_BLOCK) {
        const double2 w2 = __ldg(Wk + 2 * n + i), w3 = __ldg(Wk + 3 * n + i), w4 = __ldg(Wk + 4 * n + i);
        double2 w5 = make_double2(0.0, 0.0), w6 = w5;
        if (ORDER >= 4) { w5 = __ldg(Wk + 5 * n + i); w6 = __ldg(Wk + 6 * n + i); }
        const ImageTerms t = image_terms_literal<ORDER>(mk(w2.x, w2.y), mk(w3.x, w3.y), mk(w4.x, w4.y),
                                                        mk(w5.x, w5.y), mk(w6.x, w6.y));
        const double v2 = fma(h2, t.mu2, t.mu0);
        a0 += fabs(t.mu0);
        a2 += fabs(v2);
        if (ORDER >= 4) a4 += fabs(fma(h4, t.mu4, v2));
    }
    __shared__ double sh[3][MLM_BLOCK / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_down_sync(0xffffffffu, a0, o);
        a2 += __shfl_down_sync(0xffffffffu, a2, o);
        a4 += __shfl_down_sync(0xffffffffu, a4, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = a0; sh[1][w] = a2; sh[2][w] = a4; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0.0, t2 = 0.0, t4 = 0.0;
        for (int k = 0; k < MLM_BLOCK / 32; ++k) { t0 += sh[0][k]; t2 += sh[1][k]; t4 += sh[2][k]; }
        partial[3 * blockIdx.x + 0] = t0;
        partial[3 * blockIdx.x + 1] = t2;
        partial[3 * blockIdx.x + 2] = t4;
    }
}

__global__ void k_from_wk_final(const double* __restrict__ partial, int nblocks, int nout, double* __restrict__ out) {
    if (threadIdx.x < nout) {
        double t = 0.0;
        for (int b = 0; b < nblocks; ++b) t += partial[3 * b + threadIdx.x];
        out[threadIdx.x] = t;
    }
}

// _akl, elementwise
__global__ void __launch_bounds__(MLM_BLOCK)
k_akl(const double2* __restrict__ mu, const double2* __restrict__ W2, const double2* __restrict__ Q, int64_t n,
      double2* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * MLM_BLOCK + threadIdx.x;
    if (i >= n) return;
    const double2 m = __ldg(mu + i), w = __ldg(W2 + i), qq = __ldg(Q + i);
    const cd a = cmul(mk(m.x, m.y), akl(1.0, mk(w.x, w.y), mk(qq.x, qq.y)));
    out[i] = make_double2(a.x, a.y);
}

// Store-pattern probe (measurement only, tools/peer_store_probe.py): the store side of k_map without its arithmetic --
// a block of 128 threads owns 128 columns and walks down `nrows` rows.  mode 0: the five SoA arrays exactly like
// store_result (per warp and row 3 x 256 B + 2 x 32 B); 1: the three float64 arrays only; 2: the two uint8 arrays only;
// 3: five arrays, the uint8 ones transposed through shared memory so that one warp writes 128 contiguous bytes per row.
__global__ void __launch_bounds__(MLM_BLOCK) k_store_probe(int mode, int64_t nx, int64_t nrows, int rows_per_block,
                                                           double* __restrict__ A0, double* __restrict__ A2,
                                                           double* __restrict__ A4, uint8_t* __restrict__ nimg,
                                                           uint8_t* __restrict__ flags) {
    __shared__ uint8_t sb[2][4][MLM_BLOCK];
    const int64_t col = (int64_t)blockIdx.x * MLM_BLOCK + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = (r0 + rows_per_block < nrows) ? r0 + rows_per_block : nrows;
    if (col >= nx) return;                                   // nx is a multiple of 128 in the probe
    for (int64_t r = r0; r < r1; ++r) {
        const int64_t i = r * nx + col;
        const double v = (double)i;
        if (mode != 2) { A0[i] = v; A2[i] = v + 1.0; A4[i] = v + 2.0; }
        if (mode == 0 || mode == 2) { nimg[i] = (uint8_t)(i & 7); flags[i] = (uint8_t)(i & 3); }
        if (mode == 3) {
            const int j = (int)((r - r0) & 3);
            sb[0][j][threadIdx.x] = (uint8_t)(i & 7);
            sb[1][j][threadIdx.x] = (uint8_t)(i & 3);
            if (j == 3 || r == r1 - 1) {
                __syncthreads();
                // warp w writes row (r - j + w) of both arrays: 32 lanes x 4 bytes = 128 contiguous bytes
                const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
                if (w <= j) {
                    const int64_t base = (r - j + w) * nx + (int64_t)blockIdx.x * MLM_BLOCK;
                    ((uint32_t*)(nimg + base))[l] = ((const uint32_t*)sb[0][w])[l];
                    ((uint32_t*)(flags + base))[l] = ((const uint32_t*)sb[1][w])[l];
                }
                __syncthreads();
            }
        }
    }
}

// FP64 FMA throughput microbenchmark: 8 independent DFMA chains per thread.
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double sum = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (sum == 12345.678) out[0] = sum;               // never true; keeps the chains alive
}

// FP64-pipe probe with the operand pattern of real code: every instruction reads three DISTINCT
// register pairs; MODE 1: DFMA only, MODE 2: DFMA/DMUL/DADD mixed 11:4:1, about the mix of k_map.
template <int MODE>
__global__ void __launch_bounds__(128, 1) k_fp64_probe(double* out, int iters, const double* __restrict__ in) {
    double x[8], a[8], b[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { x[k] = in[k] + threadIdx.x; a[k] = in[8 + k]; b[k] = in[16 + k]; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (MODE == 1) x[k] = fma(x[k], a[(k + r) & 7], b[(k + 2 * r + 1) & 7]);
                else if (k < 5 || (k == 7 && (r & 1))) x[k] = fma(x[k], a[(k + r) & 7], b[(k + 2 * r + 1) & 7]);
                else if (k < 7) x[k] = x[k] * a[(k + r) & 7];
                else x[k] = x[k] + b[(k + r) & 7];
            }
        }
    }
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += x[k];
    if (sum == 12345.678) out[0] = sum;
}

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
#define MLM_HSTAGE_POINTS 4096                          /* results of <= this many positions travel as one block */
#define MLM_HSTAGE_BYTES (MLM_HSTAGE_POINTS * 26 + 4096)

struct mlm_context {
    int device = 0;
    cudaStream_t stream[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    void* dbuf = nullptr;                              // device scratch for *_host entry points
    size_t dbuf_bytes = 0;
    void* hstage = nullptr;                            // pinned staging for small results (one D2H instead of five)
    double* partial = nullptr;                         // from_wk partial sums
    int partial_blocks = 0;
    std::atomic<int64_t> launches{0};
    std::string err;
    int sm_count = 148;
    int map_strip_env = 0;                             // MLM_MAP_STRIP (experiments): overrides strip = 0 (auto)
    // Serialises the entry points that use the context's own scratch (every *_host function, mlm_from_wk's
    // partial sums): two host threads may share a context.  Device-pointer entry points only enqueue on the
    // caller's stream and touch no context state.
    std::recursive_mutex mu;
    int anchor_segment = 16;                           // columns per thread of k_anchor (MLM_MAP_ANCHOR; 0: no anchors, cold strip starts)
    int u8rows_env = -1;                               // MLM_MAP_U8ROWS = 0/1 forces the row-wise one-byte stores off/on (experiments)
    int points_segment = 0;                            // consecutive list entries per thread of k_points (0: from n; MLM_POINTS_SEGMENT)
    int models_segment = 0;                            // epochs per warm-start segment of k_models (0: from T; MLM_SEGMENT)
};

static std::string g_err;
static std::mutex g_mu;

static int fail(mlm_context* ctx, int code, const char* what) {
    std::string msg = std::string(what);
    if (code > 0) msg += std::string(": ") + cudaGetErrorString((cudaError_t)code);
    std::lock_guard<std::mutex> lk(g_mu);
    if (ctx) ctx->err = msg;
    g_err = msg;
    return code;
}

#define CK(call)                                                       \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) return fail(ctx, (int)e_, #call);       \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

static int ensure_dbuf(mlm_context* ctx, size_t bytes) {
    if (bytes <= ctx->dbuf_bytes) return 0;
    if (ctx->dbuf) CK(cudaFree(ctx->dbuf));
    ctx->dbuf = nullptr;
    ctx->dbuf_bytes = 0;
    CK(cudaMalloc(&ctx->dbuf, bytes));
    ctx->dbuf_bytes = bytes;
    return 0;
}

// NULL is the CUDA default stream (what torch.cuda.current_stream().cuda_stream is unless the
// caller switched streams), like every CUDA library API.
static inline cudaStream_t pick(mlm_context* ctx, void* stream) {
    (void)ctx;
    return (cudaStream_t)stream;
}

static bool bad_lens(double s, double q) { return !(s > 0.0) || !(q > 0.0) || !std::isfinite(s) || !std::isfinite(q); }

extern "C" {

#ifndef MLM_VERSION_STRING
#define MLM_VERSION_STRING "0.4.0"                     /* build.py passes the package version */
#endif
const char* mlm_version(void) { return "microlensing_b200 " MLM_VERSION_STRING " (sm_100a)"; }

const char* mlm_last_error(mlm_context* ctx) {
    if (ctx) return ctx->err.c_str();
    std::lock_guard<std::mutex> lk(g_mu);
    return g_err.c_str();
}

int mlm_create(int device, mlm_context** out) {
    mlm_context* ctx = nullptr;
    if (!out) return fail(nullptr, MLM_EINVAL, "mlm_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return fail(nullptr, (int)e, "mlm_create: no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(nullptr, MLM_EINVAL, "mlm_create: bad device index");
    ctx = new (std::nothrow) mlm_context();
    if (!ctx) return fail(nullptr, MLM_ENOMEM, "mlm_create: out of host memory");
    ctx->device = device;
    DeviceGuard g(device);
    for (int i = 0; i < 2; ++i) {
        e = cudaStreamCreateWithFlags(&ctx->stream[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming);
        if (e != cudaSuccess) {
            int rc = fail(nullptr, (int)e, "mlm_create: stream/event creation");
            delete ctx;
            return rc;
        }
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (cudaHostAlloc(&ctx->hstage, MLM_HSTAGE_BYTES, cudaHostAllocPortable) != cudaSuccess) {
        ctx->hstage = nullptr;                         // optional: the general path works without it
        cudaGetLastError();
    }
    if (const char* e2 = getenv("MLM_MAP_STRIP")) {
        const int v = atoi(e2);
        if (v >= 1 && v <= 4096) ctx->map_strip_env = v;
    }
    if (const char* e6 = getenv("MLM_MAP_U8ROWS")) ctx->u8rows_env = atoi(e6) ? 1 : 0;
    if (const char* e5 = getenv("MLM_MAP_ANCHOR")) {
        const int v = atoi(e5);
        if (v >= 0 && v <= 4096) ctx->anchor_segment = v;
    }
    {   // the anchors live in stream-ordered allocations (cudaMallocAsync): keep the pool's memory across synchronisations
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    if (const char* e4 = getenv("MLM_POINTS_SEGMENT")) {
        const int v = atoi(e4);
        if (v >= 1 && v <= 65536) ctx->points_segment = v;
    }
    if (const char* e3 = getenv("MLM_SEGMENT")) {
        const int v = atoi(e3);
        if (v >= 1 && v <= 65536) ctx->models_segment = v;
    }
    *out = ctx;
    return 0;
}

int mlm_destroy(mlm_context* ctx) {
    if (!ctx) return 0;
    DeviceGuard g(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->dbuf) cudaFree(ctx->dbuf);
    if (ctx->hstage) cudaFreeHost(ctx->hstage);
    if (ctx->partial) cudaFree(ctx->partial);
    for (int i = 0; i < 2; ++i) {
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
        if (ctx->stream[i]) cudaStreamDestroy(ctx->stream[i]);
    }
    delete ctx;
    return 0;
}

int mlm_launch_count(mlm_context* ctx, int64_t* out) {
    if (!ctx || !out) return fail(ctx, MLM_EINVAL, "mlm_launch_count: NULL argument");
    *out = ctx->launches.load();
    return 0;
}

// Strip length for a map of nx x nrows pixels on one GPU: power of two in 16..64 that leaves about 12 waves of
// work items (one item = one column of one strip; sm_count x 16 warps of them are resident).  Measured on the
// 8192-wide C4 map: a launch of w waves of 64-row strips costs about ceil(w + 0.5) block durations, while a
// shorter strip costs ~2000/L more FP64 instructions per pixel in cold solves.
static int auto_strip(int sm_count, int64_t nx, int64_t nrows) {
    const int64_t resident = (int64_t)sm_count * 16 * 32;
    int strip = 64;
    while (strip > 16 && nx * (nrows > 0 ? nrows : 1) / strip < 12 * resident) strip /= 2;
    return strip;
}

int mlm_auto_strip(mlm_context* ctx, int64_t nx, int64_t nrows) {
    if (!ctx || nx < 0 || nrows < 0) return MLM_EINVAL;
    return auto_strip(ctx->sm_count, nx, nrows);
}

static int resolve_strip(mlm_context* ctx, int strip, int64_t nx, int64_t nrows) {
    if (strip > 0) return strip;
    if (ctx->map_strip_env > 0) return ctx->map_strip_env;
    return auto_strip(ctx->sm_count, nx, nrows);
}

int mlm_synchronize(mlm_context* ctx) {
    if (!ctx) return fail(ctx, MLM_EINVAL, "mlm_synchronize: NULL context");
    DeviceGuard g(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream[0]));
    CK(cudaStreamSynchronize(ctx->stream[1]));
    return 0;
}

int mlm_alloc_host(mlm_context* ctx, int64_t bytes, void** out) {
    if (!ctx || !out || bytes < 0) return fail(ctx, MLM_EINVAL, "mlm_alloc_host: bad argument");
    DeviceGuard g(ctx->device);
    *out = nullptr;
    if (bytes == 0) return 0;
    CK(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable));
    return 0;
}

int mlm_free_host(mlm_context* ctx, void* p) {
    if (!ctx) return fail(ctx, MLM_EINVAL, "mlm_free_host: NULL context");
    if (!p) return 0;
    DeviceGuard g(ctx->device);
    CK(cudaFreeHost(p));
    return 0;
}

// ---- device buffers that other processes of the box can map (multi-GPU gather without a copy) ----
int mlm_device_alloc(mlm_context* ctx, int64_t bytes, void** out) {
    if (!ctx || !out || bytes <= 0) return fail(ctx, MLM_EINVAL, "mlm_device_alloc: bad argument");
    DeviceGuard g(ctx->device);
    *out = nullptr;
    CK(cudaMalloc(out, (size_t)bytes));                // plain cudaMalloc: exportable with cudaIpcGetMemHandle
    return 0;
}

int mlm_device_free(mlm_context* ctx, void* p) {
    if (!ctx) return fail(ctx, MLM_EINVAL, "mlm_device_free: NULL context");
    if (!p) return 0;
    DeviceGuard g(ctx->device);
    CK(cudaFree(p));
    return 0;
}

int mlm_ipc_export(mlm_context* ctx, void* p, unsigned char* handle) {
    if (!ctx || !p || !handle) return fail(ctx, MLM_EINVAL, "mlm_ipc_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == MLM_IPC_HANDLE_BYTES, "handle size");
    DeviceGuard g(ctx->device);
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, p));
    memcpy(handle, &h, sizeof(h));
    return 0;
}

int mlm_ipc_open(mlm_context* ctx, const unsigned char* handle, void** out) {
    if (!ctx || !handle || !out) return fail(ctx, MLM_EINVAL, "mlm_ipc_open: NULL argument");
    DeviceGuard g(ctx->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    *out = nullptr;
    CK(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int mlm_memcpy_async(mlm_context* ctx, void* dst, const void* src, int64_t bytes, void* stream) {
    if (!ctx || bytes < 0 || (bytes > 0 && (!dst || !src))) return fail(ctx, MLM_EINVAL, "mlm_memcpy_async: bad argument");
    if (bytes == 0) return 0;
    DeviceGuard g(ctx->device);
    CK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, pick(ctx, stream)));   // UVA: peer-mapped dst ok
    return 0;
}

int mlm_memcpy2d_async(mlm_context* ctx, void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width,
                       int64_t height, void* stream) {
    if (!ctx || width < 0 || height < 0 || dpitch < width || spitch < width || ((width > 0 && height > 0) && (!dst || !src)))
        return fail(ctx, MLM_EINVAL, "mlm_memcpy2d_async: bad argument");
    if (width == 0 || height == 0) return 0;
    DeviceGuard g(ctx->device);
    CK(cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width, (size_t)height, cudaMemcpyDefault,
                         pick(ctx, stream)));
    return 0;
}

int mlm_host_register(mlm_context* ctx, void* p, int64_t bytes) {
    if (!ctx || !p || bytes <= 0) return fail(ctx, MLM_EINVAL, "mlm_host_register: bad argument");
    DeviceGuard g(ctx->device);
    CK(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable));
    return 0;
}

int mlm_host_unregister(mlm_context* ctx, void* p) {
    if (!ctx || !p) return fail(ctx, MLM_EINVAL, "mlm_host_unregister: bad argument");
    DeviceGuard g(ctx->device);
    CK(cudaHostUnregister(p));
    return 0;
}

int mlm_ipc_close(mlm_context* ctx, void* p) {
    if (!ctx) return fail(ctx, MLM_EINVAL, "mlm_ipc_close: NULL context");
    if (!p) return 0;
    DeviceGuard g(ctx->device);
    CK(cudaIpcCloseMemHandle(p));
    return 0;
}

// ---- fused path, device pointers --------------------------------------------------------------
int mlm_points(mlm_context* ctx, double s, double q, double rho, double Gamma, const double* zeta, int64_t n,
               int order, double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags, void* stream) {
    if (!ctx) return fail(ctx, MLM_EINVAL, "mlm_points: NULL context");
    if (n < 0 || (order != 2 && order != 4) || bad_lens(s, q)) return fail(ctx, MLM_EINVAL, "mlm_points: bad argument");
    if (n == 0) return 0;
    if (!zeta || !(A0 || A2 || A4 || nimg || flags)) return fail(ctx, MLM_EINVAL, "mlm_points: NULL buffer");
    DeviceGuard g(ctx->device);
    LensTable T;
    make_table(s, q, &T);
    const SourceFactors sf = source_factors(rho, Gamma);
    // consecutive entries per thread: short runs while the list is small (about three waves of resident
    // threads: sm_count x 4 blocks x 128; measured on 1e6 epochs: runs of 4-5 beat runs of 14), up to 32
    const int64_t resident = (int64_t)ctx->sm_count * 4 * MLM_BLOCK;
    int seg = ctx->points_segment > 0 ? ctx->points_segment : (int)((n + 3 * resident - 1) / (3 * resident));
    if (ctx->points_segment <= 0) seg = seg < 1 ? 1 : (seg > 32 ? 32 : seg);
    const int64_t nthreads = (n + seg - 1) / seg;
    const unsigned blocks = (unsigned)((nthreads + MLM_BLOCK - 1) / MLM_BLOCK);
    cudaStream_t st = pick(ctx, stream);
    if (order == 4)
        k_points<4><<<blocks, MLM_BLOCK, 0, st>>>(T, sf, (const double2*)zeta, n, seg, A0, A2, A4, nimg, flags);
    else
        k_points<2><<<blocks, MLM_BLOCK, 0, st>>>(T, sf, (const double2*)zeta, n, seg, A0, A2, A4, nimg, flags);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

int mlm_points_unordered(mlm_context* ctx, double s, double q, double rho, double Gamma, const double* zeta, int64_t n,
                         int order, double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags, void* stream) {
    if (!ctx) return fail(ctx, MLM_EINVAL, "mlm_points_unordered: NULL context");
    if (n < 0 || n > 0x7fffffffll || (order != 2 && order != 4) || bad_lens(s, q))
        return fail(ctx, MLM_EINVAL, "mlm_points_unordered: bad argument");
    if (n == 0) return 0;
    if (!zeta || !(A0 || A2 || A4 || nimg || flags)) return fail(ctx, MLM_EINVAL, "mlm_points_unordered: NULL buffer");
    if (n < 4096) return mlm_points(ctx, s, q, rho, Gamma, zeta, n, order, A0, A2, A4, nimg, flags, stream);
    DeviceGuard g(ctx->device);
    cudaStream_t st = pick(ctx, stream);
    // every entry on its own: three-image guesses, then the root finder on the compacted rest
    LensTable T;
    make_table(s, q, &T);
    const SourceFactors sf = source_factors(rho, Gamma);
    char* ws = nullptr;                                  // stream-ordered scratch: two counters | two work lists
    CK(cudaMallocAsync((void**)&ws, 256 + 2 * (size_t)n * sizeof(uint32_t), st));
    unsigned* counters = (unsigned*)ws;
    uint32_t* lists = (uint32_t*)(ws + 256);
    cudaError_t e = cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned), st);
    if (e == cudaSuccess) {
        const unsigned blocks = (unsigned)((n + MLM_BLOCK - 1) / MLM_BLOCK);
        const unsigned blocks2 = blocks < (unsigned)ctx->sm_count * 8u ? blocks : (unsigned)ctx->sm_count * 8u;
        const double2* z2 = (const double2*)zeta;
        if (order == 4) {
            k_points_direct<4><<<blocks, MLM_BLOCK, 0, st>>>(T, sf, z2, n, lists, counters, A0, A2, A4, nimg, flags);
            k_points_list<4, 1><<<blocks2, MLM_BLOCK, 0, st>>>(T, sf, z2, n, lists, counters, A0, A2, A4, nimg, flags);
            k_points_list<4, 0><<<blocks2, MLM_BLOCK, 0, st>>>(T, sf, z2, n, lists, counters, A0, A2, A4, nimg, flags);
        } else {
            k_points_direct<2><<<blocks, MLM_BLOCK, 0, st>>>(T, sf, z2, n, lists, counters, A0, A2, A4, nimg, flags);
            k_points_list<2, 1><<<blocks2, MLM_BLOCK, 0, st>>>(T, sf, z2, n, lists, counters, A0, A2, A4, nimg, flags);
            k_points_list<2, 0><<<blocks2, MLM_BLOCK, 0, st>>>(T, sf, z2, n, lists, counters, A0, A2, A4, nimg, flags);
        }
        ctx->launches += 3;
        e = cudaGetLastError();
    }
    cudaFreeAsync(ws, st);
    if (e != cudaSuccess) return fail(ctx, (int)e, "mlm_points_unordered");
    return 0;
}

int mlm_map_deal(mlm_context* ctx, double s, double q, double rho, double Gamma, double x0, double y0, double dx,
                 double dy, int64_t nx, int64_t row0, int64_t nrows, int64_t period, int nsel, const int32_t* sel,
                 int order, int strip, double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags, void* stream) {
    if (!ctx) return fail(ctx, MLM_EINVAL, "mlm_map: NULL context");
    if (nx < 0 || nrows < 0 || row0 < 0 || (order != 2 && order != 4) || bad_lens(s, q) || period < 1 || nsel < 1 ||
        nsel > MLM_DEAL_MAXSEL || nsel > period || !sel || strip < 0 || strip > 4096)
        return fail(ctx, MLM_EINVAL, "mlm_map: bad argument");
    Deal deal;
    deal.period = period;
    deal.nsel = nsel;
    for (int j = 0; j < MLM_DEAL_MAXSEL; ++j) deal.sel[j] = 0;
    for (int j = 0; j < nsel; ++j) {
        if (sel[j] < 0 || sel[j] >= period || (j > 0 && sel[j] <= sel[j - 1]))
            return fail(ctx, MLM_EINVAL, "mlm_map: sel must be ascending positions inside the period");
        deal.sel[j] = sel[j];
    }
    if (nx == 0 || nrows == 0) return 0;
    if (!(A0 || A2 || A4 || nimg || flags)) return fail(ctx, MLM_EINVAL, "mlm_map: NULL buffer");
    DeviceGuard g(ctx->device);
    LensTable T;
    make_table(s, q, &T);
    const SourceFactors sf = source_factors(rho, Gamma);
    const double shift = s * q / (1.0 + q);            // multipoles.py:178
    strip = resolve_strip(ctx, strip, nx, nrows);
    const int64_t k_first = row0 / strip, k_last = (row0 + nrows - 1) / strip;
    auto owned_from = [&](int64_t k) { return deal_owned_from(deal, k); };
    const int64_t t_first = owned_from(k_first), t_end = owned_from(k_last + 1);
    const int64_t i_count = t_end - t_first;
    if (i_count <= 0) return 0;
    deal.t0 = t_first;
    // owned strip nearest to the lens axis y = 0 (row index (0 - y0)/dy - 0.5)
    int64_t i_axis = 0;
    {
        const double r_axis = (dy != 0.0) ? (-y0 / dy - 0.5) : 0.0;
        double ka = std::floor(r_axis / (double)strip);
        if (!(ka >= (double)k_first)) ka = (double)k_first;            // also catches NaN
        if (ka > (double)k_last) ka = (double)k_last;
        int64_t ia = owned_from((int64_t)ka) - t_first;
        if (ia < 0) ia = 0;
        if (ia > i_count - 1) ia = i_count - 1;
        i_axis = ia;
    }
    dim3 grid((unsigned)((nx + MLM_BLOCK - 1) / MLM_BLOCK), (unsigned)(i_count < 65535 ? i_count : 65535));
    cudaStream_t st = pick(ctx, stream);
    // anchors of the owned strips (stream-ordered scratch: 82 bytes per anchor pixel); strips of one or two rows are
    // not worth it, and without the scratch the strips simply start cold
    double2* aroots = nullptr;
    uint16_t* ameta = nullptr;
    void* ws = nullptr;
    const int aseg = ctx->anchor_segment;
    if (aseg > 0 && strip >= 4 && nx >= 2 * (int64_t)aseg) {
        const size_t npix = (size_t)i_count * (size_t)nx;
        if (cudaMallocAsync(&ws, npix * 82 + 256, st) == cudaSuccess) {
            aroots = (double2*)ws;
            ameta = (uint16_t*)((char*)ws + npix * 80);
            const int64_t nseg = (nx + aseg - 1) / aseg;
            dim3 ga((unsigned)((nseg + MLM_BLOCK - 1) / MLM_BLOCK), grid.y);
            k_anchor<<<ga, MLM_BLOCK, 0, st>>>(T, x0, y0, dx, dy, shift, nx, row0, strip, deal, i_count, aseg, aroots, ameta);
            ctx->launches++;
        } else {
            cudaGetLastError();
            ws = nullptr;
        }
    }
    // one-byte outputs that live in another GPU's memory are written a 128-byte row segment at a time
    int u8rows = 0;
    if ((nimg || flags) && nx % MLM_BLOCK == 0) {
        cudaPointerAttributes pa;
        const void* p8 = nimg ? (const void*)nimg : (const void*)flags;
        if (cudaPointerGetAttributes(&pa, p8) == cudaSuccess && pa.type == cudaMemoryTypeDevice && pa.device != ctx->device &&
            (((uintptr_t)nimg | (uintptr_t)flags) & 3u) == 0)
            u8rows = 1;
        cudaGetLastError();
        if (ctx->u8rows_env >= 0) u8rows = ctx->u8rows_env && (((uintptr_t)nimg | (uintptr_t)flags) & 3u) == 0;
    }
    if (order == 4)
        k_map<4><<<grid, MLM_BLOCK, 0, st>>>(T, sf, x0, y0, dx, dy, shift, nx, row0, nrows, strip, deal, i_count, i_axis,
                                              aroots, ameta, u8rows, A0, A2, A4, nimg, flags);
    else
        k_map<2><<<grid, MLM_BLOCK, 0, st>>>(T, sf, x0, y0, dx, dy, shift, nx, row0, nrows, strip, deal, i_count, i_axis,
                                              aroots, ameta, u8rows, A0, A2, A4, nimg, flags);
    ctx->launches++;
    cudaError_t le = cudaGetLastError();
    if (ws) cudaFreeAsync(ws, st);
    if (le != cudaSuccess) return fail(ctx, (int)le, "k_map launch");
    return 0;
}

int mlm_map_strips(mlm_context* ctx, double s, double q, double rho, double Gamma, double x0, double y0, double dx,
                   double dy, int64_t nx, int64_t row0, int64_t nrows, int64_t kphase, int64_t kstep, int order, int strip,
                   double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags, void* stream) {
    if (kstep < 1 || kphase < 0 || kphase >= kstep) return fail(ctx, MLM_EINVAL, "mlm_map_strips: bad argument");
    const int32_t sel = (int32_t)kphase;
    return mlm_map_deal(ctx, s, q, rho, Gamma, x0, y0, dx, dy, nx, row0, nrows, kstep, 1, &sel, order, strip, A0, A2, A4,
                        nimg, flags, stream);
}

int mlm_map(mlm_context* ctx, double s, double q, double rho, double Gamma, double x0, double y0, double dx,
            double dy, int64_t nx, int64_t row0, int64_t nrows, int order, int strip, double* A0, double* A2, double* A4,
            uint8_t* nimg, uint8_t* flags, void* stream) {
    const int32_t sel = 0;
    return mlm_map_deal(ctx, s, q, rho, Gamma, x0, y0, dx, dy, nx, row0, nrows, 1, 1, &sel, order, strip, A0, A2, A4, nimg,
                        flags, stream);
}

static int models_segment(mlm_context* ctx, int seg, int64_t M, int64_t Tn);
static int models_threads_per_model(int64_t Tn, int seg, int64_t M, bool share_blocks);

static int launch_models(mlm_context* ctx, const char* who, const double* s, const double* q, const double* rho,
                         const double* Gamma, const double* u0, const double* alpha, int64_t M, const Epochs& ep,
                         int64_t Tn, int order, int seg, double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags,
                         void* stream) {
    if (!ctx) return fail(ctx, MLM_EINVAL, who);
    if (M < 0 || Tn < 0 || (order != 2 && order != 4) || seg < 0 || seg > 65536) return fail(ctx, MLM_EINVAL, who);
    if (M == 0 || Tn == 0) return 0;
    if (!s || !q || !rho || !Gamma || !u0 || !alpha || !(A0 || A2 || A4 || nimg || flags)) return fail(ctx, MLM_EINVAL, who);
    DeviceGuard g(ctx->device);
    seg = models_segment(ctx, seg, M, Tn);
    const int P = models_threads_per_model(Tn, seg, M, true);
    const int64_t nseg = (Tn + seg - 1) / seg;
    int64_t bx = (nseg + P - 1) / P;
    if (bx > 4096) bx = 4096;
    const int64_t nmb = (M + MLM_MODELS_BLOCK / P - 1) / (MLM_MODELS_BLOCK / P);
    dim3 grid((unsigned)bx, (unsigned)(nmb < 65535 ? nmb : 65535));
    cudaStream_t st = pick(ctx, stream);
    if (order == 4)
        (ep.times ? k_models<4, true> : k_models<4, false>)<<<grid, MLM_MODELS_BLOCK, 0, st>>>(s, q, rho, Gamma, u0, alpha, M, ep, Tn, seg, P, A0, A2, A4, nimg, flags);
    else
        (ep.times ? k_models<2, true> : k_models<2, false>)<<<grid, MLM_MODELS_BLOCK, 0, st>>>(s, q, rho, Gamma, u0, alpha, M, ep, Tn, seg, P, A0, A2, A4, nimg, flags);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

int mlm_models(mlm_context* ctx, const double* s, const double* q, const double* rho, const double* Gamma,
               const double* u0, const double* alpha, int64_t M, double tau0, double dtau, int64_t Tn, int order, int seg,
               double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags, void* stream) {
    const Epochs ep = {tau0, dtau, nullptr, nullptr, nullptr};
    return launch_models(ctx, "mlm_models: bad argument", s, q, rho, Gamma, u0, alpha, M, ep, Tn, order, seg, A0, A2, A4,
                         nimg, flags, stream);
}

int mlm_models_times(mlm_context* ctx, const double* s, const double* q, const double* rho, const double* Gamma,
                     const double* u0, const double* alpha, const double* t0, const double* tE, int64_t M,
                     const double* times, int64_t Tn, int order, int seg, double* A0, double* A2, double* A4, uint8_t* nimg,
                     uint8_t* flags, void* stream) {
    if (Tn > 0 && !times) return fail(ctx, MLM_EINVAL, "mlm_models_times: NULL times");
    const Epochs ep = {0.0, 0.0, times, t0, tE};
    return launch_models(ctx, "mlm_models_times: bad argument", s, q, rho, Gamma, u0, alpha, M, ep, Tn, order, seg, A0, A2,
                         A4, nimg, flags, stream);
}

// Epochs per warm-start segment for a grid of M models x Tn epochs.  A segment is a serial chain (one thread, ~7 us per
// epoch) and a block of 64 threads is the unit of scheduling, so a launch needs MANY more blocks than fit on the GPU at
// once or its tail dominates.  Measured on the C5 grid (2e7 positions; rounds = blocks / resident blocks): 128-epoch
// segments (2.1 rounds) 4.0 ms, 64 (4.2) 3.03 ms, 32 (8.4) 2.98 ms, 16 (17; eight times the cold solves of 128)
// 2.89 ms -- and 24 or 48 epochs, which leave a light curve with a number of segments just above a power of two
// (84, 42: the threads of a model are a power of two), 3.40 and 4.21 ms.  Rule: ~8 segments per resident thread
// (sm_count x MLM_MODELS_MINBLOCKS blocks x 64 threads) -- but not below 16 epochs (a cold solve costs ~3 tracked
// epochs) while 16-epoch segments still give two rounds -- snapped to Tn / 2^k, between 4 and 128 epochs.
static int auto_segment(int sm_count, int64_t M, int64_t Tn) {
    const int64_t resident = (int64_t)sm_count * MLM_MODELS_MINBLOCKS * MLM_MODELS_BLOCK;
    double target = (double)(M * Tn) / (8.0 * (double)resident);
    const double two_rounds = (double)(M * Tn) / (2.0 * (double)resident);
    if (target < 16.0) target = fmax(target, fmin(16.0, two_rounds));
    target = target < 4.0 ? 4.0 : (target > 128.0 ? 128.0 : target);
    int64_t best = 0;
    double best_err = 1e300;
    for (int k = 0; k < 40 && ((int64_t)1 << k) <= Tn; ++k) {
        const int64_t n = (int64_t)1 << k, sk = (Tn + n - 1) / n;      // 2^k segments of sk epochs
        if (sk > 128) continue;
        if (sk < 4) break;
        const double err = fabs(log((double)sk / target));
        if (err < best_err) { best_err = err; best = sk; }
    }
    if (best == 0) best = (int64_t)target < Tn ? (int64_t)target : Tn;   // Tn < 4 (or no admissible power of two)
    return (int)(best < 1 ? 1 : best);
}

static int models_segment(mlm_context* ctx, int seg, int64_t M, int64_t Tn) {
    if (seg > 0) return seg;
    if (ctx->models_segment > 0) return ctx->models_segment;          // MLM_SEGMENT (experiments)
    return auto_segment(ctx->sm_count, M, Tn);
}

// threads of a block that share one model: 64, or the next power of two >= the number of segments (>= 4)
static int models_threads_per_model(int64_t Tn, int seg, int64_t M, bool share_blocks) {
    const int64_t nseg = (Tn + seg - 1) / seg;
    int P = MLM_MODELS_BLOCK / MLM_MODELS_GMAX;
    while (P < MLM_MODELS_BLOCK && P < nseg) P *= 2;
    // k_models on a grid of many models: 8 threads per model, so that a warp holds 8 ADJACENT segments of each of 4 light
    // curves (the same phase of every light curve: similar numbers of Newton iterations and images) instead of 32 consecutive
    // segments of one (wings and peak in one warp).  C5: 3.07 -> 2.95 ms, 17.9 -> 20.8 active lanes in the Newton loop.
    // Pure scheduling: which thread walks a segment does not change its arithmetic.
    if (share_blocks && M >= 8 && P > 8) P = 8;
    if (const char* e = getenv("MLM_MODELS_P")) {        // experiments: threads of a block per model (4..64, power of two)
        const int v = atoi(e);
        if (v >= MLM_MODELS_BLOCK / MLM_MODELS_GMAX && v <= MLM_MODELS_BLOCK && (v & (v - 1)) == 0) P = v;
    }
    return P;
}

int mlm_auto_segment(mlm_context* ctx, int64_t M, int64_t Tn) {
    if (!ctx || M < 0 || Tn < 0) return MLM_EINVAL;
    return auto_segment(ctx->sm_count, M > 0 ? M : 1, Tn > 0 ? Tn : 1);
}

static int launch_models_chi2(mlm_context* ctx, const char* who, const double* s, const double* q, const double* rho,
                              const double* Gamma, const double* u0, const double* alpha, int64_t M, const Epochs& ep,
                              int64_t Tn, int order, int seg, const double* flux, const double* weight, double* stats,
                              void* stream) {
    if (!ctx) return fail(ctx, MLM_EINVAL, who);
    if (M < 0 || Tn < 0 || (order != 2 && order != 4) || seg < 0 || seg > 65536) return fail(ctx, MLM_EINVAL, who);
    if (M == 0) return 0;
    if (!s || !q || !rho || !Gamma || !u0 || !alpha || !stats || (Tn > 0 && (!flux || !weight))) return fail(ctx, MLM_EINVAL, who);
    DeviceGuard g(ctx->device);
    seg = models_segment(ctx, seg, M, Tn);
    const int P = models_threads_per_model(Tn, seg, M, false);
    const int64_t nmb = (M + MLM_MODELS_BLOCK / P - 1) / (MLM_MODELS_BLOCK / P);
    const unsigned blocks = (unsigned)(nmb < 1048576 ? nmb : 1048576);
    cudaStream_t st = pick(ctx, stream);
    if (order == 4)
        (ep.times ? k_models_chi2<4, true> : k_models_chi2<4, false>)<<<blocks, MLM_MODELS_BLOCK, 0, st>>>(s, q, rho, Gamma, u0, alpha, M, ep, Tn, seg, P, flux, weight, stats);
    else
        (ep.times ? k_models_chi2<2, true> : k_models_chi2<2, false>)<<<blocks, MLM_MODELS_BLOCK, 0, st>>>(s, q, rho, Gamma, u0, alpha, M, ep, Tn, seg, P, flux, weight, stats);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

int mlm_models_chi2(mlm_context* ctx, const double* s, const double* q, const double* rho, const double* Gamma,
                    const double* u0, const double* alpha, int64_t M, double tau0, double dtau, int64_t Tn, int order,
                    int seg, const double* flux, const double* weight, double* stats, void* stream) {
    const Epochs ep = {tau0, dtau, nullptr, nullptr, nullptr};
    return launch_models_chi2(ctx, "mlm_models_chi2: bad argument", s, q, rho, Gamma, u0, alpha, M, ep, Tn, order, seg,
                              flux, weight, stats, stream);
}

int mlm_models_chi2_times(mlm_context* ctx, const double* s, const double* q, const double* rho, const double* Gamma,
                          const double* u0, const double* alpha, const double* t0, const double* tE, int64_t M,
                          const double* times, int64_t Tn, int order, int seg, const double* flux, const double* weight,
                          double* stats, void* stream) {
    if (Tn > 0 && !times) return fail(ctx, MLM_EINVAL, "mlm_models_chi2_times: NULL times");
    const Epochs ep = {0.0, 0.0, times, t0, tE};
    return launch_models_chi2(ctx, "mlm_models_chi2_times: bad argument", s, q, rho, Gamma, u0, alpha, M, ep, Tn, order,
                              seg, flux, weight, stats, stream);
}

// host variant of both: npar = 6 (uniform epochs) or 8 (+ t0, tE) parameter arrays, times optional
static int models_chi2_host_impl(mlm_context* ctx, const char* who, const double* const* hp, int npar, int64_t M,
                                 double tau0, double dtau, const double* times, int64_t Tn, int order, int seg,
                                 const double* flux, const double* weight, double* stats) {
    if (!ctx) return fail(ctx, MLM_EINVAL, who);
    if (M < 0 || Tn < 0 || (order != 2 && order != 4)) return fail(ctx, MLM_EINVAL, who);
    if (M == 0) return 0;
    for (int i = 0; i < npar; ++i)
        if (!hp[i]) return fail(ctx, MLM_EINVAL, who);
    if (!stats || (Tn > 0 && (!flux || !weight))) return fail(ctx, MLM_EINVAL, who);
    DeviceGuard g(ctx->device);
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    const size_t mpad = ((size_t)M + 31) & ~(size_t)31, tpad = ((size_t)Tn + 31) & ~(size_t)31;
    int rc = ensure_dbuf(ctx, (mpad * 16 + tpad * 3) * sizeof(double));
    if (rc) return rc;
    double* par = (double*)ctx->dbuf;
    double* dflux = par + 8 * mpad;
    double* dw = dflux + tpad;
    double* dtimes = dw + tpad;
    double* dstats = dtimes + tpad;
    cudaStream_t st = ctx->stream[0];
    for (int i = 0; i < npar; ++i) CK(cudaMemcpyAsync(par + i * mpad, hp[i], M * 8, cudaMemcpyHostToDevice, st));
    if (Tn > 0) {
        CK(cudaMemcpyAsync(dflux, flux, Tn * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dw, weight, Tn * 8, cudaMemcpyHostToDevice, st));
        if (times) CK(cudaMemcpyAsync(dtimes, times, Tn * 8, cudaMemcpyHostToDevice, st));
    }
    const Epochs ep = {tau0, dtau, times ? dtimes : nullptr, npar == 8 ? par + 6 * mpad : nullptr,
                       npar == 8 ? par + 7 * mpad : nullptr};
    rc = launch_models_chi2(ctx, who, par, par + mpad, par + 2 * mpad, par + 3 * mpad, par + 4 * mpad, par + 5 * mpad, M,
                            ep, Tn, order, seg, dflux, dw, dstats, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(stats, dstats, (size_t)M * 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

int mlm_models_chi2_host(mlm_context* ctx, const double* s, const double* q, const double* rho, const double* Gamma,
                         const double* u0, const double* alpha, int64_t M, double tau0, double dtau, int64_t Tn,
                         int order, int seg, const double* flux, const double* weight, double* stats) {
    const double* hp[6] = {s, q, rho, Gamma, u0, alpha};
    return models_chi2_host_impl(ctx, "mlm_models_chi2_host: bad argument", hp, 6, M, tau0, dtau, nullptr, Tn, order, seg,
                                 flux, weight, stats);
}

int mlm_models_chi2_times_host(mlm_context* ctx, const double* s, const double* q, const double* rho,
                               const double* Gamma, const double* u0, const double* alpha, const double* t0,
                               const double* tE, int64_t M, const double* times, int64_t Tn, int order, int seg,
                               const double* flux, const double* weight, double* stats) {
    if (Tn > 0 && !times) return fail(ctx, MLM_EINVAL, "mlm_models_chi2_times_host: NULL times");
    const double* hp[8] = {s, q, rho, Gamma, u0, alpha, t0, tE};
    return models_chi2_host_impl(ctx, "mlm_models_chi2_times_host: bad argument", hp, 8, M, 0.0, 0.0, times, Tn, order,
                                 seg, flux, weight, stats);
}

int mlm_from_wk(mlm_context* ctx, const double* Wk, int64_t n, int order, double rho, double Gamma, double* out,
                void* stream) {
    if (!ctx) return fail(ctx, MLM_EINVAL, "mlm_from_wk: NULL context");
    if (n < 0 || (order != 2 && order != 4) || !out) return fail(ctx, MLM_EINVAL, "mlm_from_wk: bad argument");
    if (n > 0 && !Wk) return fail(ctx, MLM_EINVAL, "mlm_from_wk: NULL Wk");
    DeviceGuard g(ctx->device);
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    cudaStream_t st = pick(ctx, stream);
    const int nout = order == 4 ? 3 : 2;
    if (n == 0) {                                      // zero images -> zeros
        CK(cudaMemsetAsync(out, 0, nout * sizeof(double), st));
        return 0;
    }
    int64_t blocks = (n + MLM_BLOCK - 1) / MLM_BLOCK;
    if (blocks > 1184) blocks = 1184;                  // 8 x 148
    if (ctx->partial_blocks < blocks) {
        if (ctx->partial) CK(cudaFree(ctx->partial));
        ctx->partial = nullptr;
        ctx->partial_blocks = 0;
        CK(cudaMalloc(&ctx->partial, 1184 * 3 * sizeof(double)));
        ctx->partial_blocks = 1184;
    }
    const SourceFactors sf = source_factors(rho, Gamma);
    if (order == 4)
        k_from_wk<4><<<(unsigned)blocks, MLM_BLOCK, 0, st>>>((const double2*)Wk, n, sf.h2, sf.h4, ctx->partial);
    else
        k_from_wk<2><<<(unsigned)blocks, MLM_BLOCK, 0, st>>>((const double2*)Wk, n, sf.h2, sf.h4, ctx->partial);
    k_from_wk_final<<<1, 32, 0, st>>>(ctx->partial, (int)blocks, nout, out);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return 0;
}

int mlm_solve(mlm_context* ctx, double s, double q, const double* zeta, int64_t n, double* roots, void* stream) {
    if (!ctx) return fail(ctx, MLM_EINVAL, "mlm_solve: NULL context");
    if (n < 0 || bad_lens(s, q)) return fail(ctx, MLM_EINVAL, "mlm_solve: bad argument");
    if (n == 0) return 0;
    if (!zeta || !roots) return fail(ctx, MLM_EINVAL, "mlm_solve: NULL buffer");
    DeviceGuard g(ctx->device);
    LensTable T;
    make_table(s, q, &T);
    const unsigned blocks = (unsigned)((n + MLM_BLOCK - 1) / MLM_BLOCK);
    k_solve<<<blocks, MLM_BLOCK, 0, pick(ctx, stream)>>>(T, (const double2*)zeta, n, (double2*)roots);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

int mlm_akl(mlm_context* ctx, const double* mu, const double* W2, const double* Q, int64_t n, double* out,
            void* stream) {
    if (!ctx) return fail(ctx, MLM_EINVAL, "mlm_akl: NULL context");
    if (n < 0) return fail(ctx, MLM_EINVAL, "mlm_akl: bad argument");
    if (n == 0) return 0;
    if (!mu || !W2 || !Q || !out) return fail(ctx, MLM_EINVAL, "mlm_akl: NULL buffer");
    DeviceGuard g(ctx->device);
    const unsigned blocks = (unsigned)((n + MLM_BLOCK - 1) / MLM_BLOCK);
    k_akl<<<blocks, MLM_BLOCK, 0, pick(ctx, stream)>>>((const double2*)mu, (const double2*)W2, (const double2*)Q, n, (double2*)out);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

// ---- host-buffer entry points -------------------------------------------------------------------
// Layout of the device scratch for n positions: [zeta 16n | A0 8n | A2 8n | A4 8n | nimg n | flags n]
struct Scratch {
    double *zeta, *A0, *A2, *A4;
    uint8_t *nimg, *flags;
};
static int carve(mlm_context* ctx, int64_t n, bool with_zeta, Scratch* sc) {
    const size_t npad = ((size_t)n + 255) & ~(size_t)255;
    const size_t bytes = npad * ((with_zeta ? 16 : 0) + 24 + 2);
    int rc = ensure_dbuf(ctx, bytes);
    if (rc) return rc;
    char* p = (char*)ctx->dbuf;
    sc->zeta = (double*)p;
    if (with_zeta) p += npad * 16;
    sc->A0 = (double*)p; p += npad * 8;
    sc->A2 = (double*)p; p += npad * 8;
    sc->A4 = (double*)p; p += npad * 8;
    sc->nimg = (uint8_t*)p; p += npad;
    sc->flags = (uint8_t*)p;
    return 0;
}

// copy one chunk of results back on `st`
static int d2h_chunk(mlm_context* ctx, const Scratch& sc, int64_t off, int64_t cnt, double* A0, double* A2,
                     double* A4, uint8_t* nimg, uint8_t* flags, cudaStream_t st) {
    if (A0) CK(cudaMemcpyAsync(A0 + off, sc.A0 + off, cnt * 8, cudaMemcpyDeviceToHost, st));
    if (A2) CK(cudaMemcpyAsync(A2 + off, sc.A2 + off, cnt * 8, cudaMemcpyDeviceToHost, st));
    if (A4) CK(cudaMemcpyAsync(A4 + off, sc.A4 + off, cnt * 8, cudaMemcpyDeviceToHost, st));
    if (nimg) CK(cudaMemcpyAsync(nimg + off, sc.nimg + off, cnt, cudaMemcpyDeviceToHost, st));
    if (flags) CK(cudaMemcpyAsync(flags + off, sc.flags + off, cnt, cudaMemcpyDeviceToHost, st));
    return 0;
}

#define MLM_CHUNK_POINTS (int64_t(1) << 21)

// Does a host list have spatial order?  Fraction of sampled consecutive pairs within 0.02 Einstein radii.
static bool host_list_is_coherent(const double* zeta, int64_t n) {
    if (n < 8192) return true;                         // small lists: the plain path (no sort) is as fast
    const int64_t samples = 2048, stride = (n - 1) / samples;
    int64_t close = 0;
    for (int64_t k = 0; k < samples; ++k) {
        const int64_t i = k * stride;
        const double dx = zeta[2 * i + 2] - zeta[2 * i], dy = zeta[2 * i + 3] - zeta[2 * i + 1];
        if (dx * dx + dy * dy <= 4e-4) ++close;
    }
    return 2 * close >= samples;
}

int mlm_points_host(mlm_context* ctx, double s, double q, double rho, double Gamma, const double* zeta, int64_t n,
                    int order, double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags) {
    if (!ctx) return fail(ctx, MLM_EINVAL, "mlm_points_host: NULL context");
    if (n < 0 || (order != 2 && order != 4) || bad_lens(s, q))
        return fail(ctx, MLM_EINVAL, "mlm_points_host: bad argument");
    if (n == 0) return 0;
    if (!zeta || !(A0 || A2 || A4 || nimg || flags)) return fail(ctx, MLM_EINVAL, "mlm_points_host: NULL buffer");
    DeviceGuard g(ctx->device);
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    Scratch sc;
    int rc = carve(ctx, n, true, &sc);
    if (rc) return rc;
    if (n <= MLM_HSTAGE_POINTS && ctx->hstage) {
        // small call (example(), a handful of epochs): latency matters, not bandwidth -- one H2D, one kernel,
        // ONE D2H of the whole result block into pinned staging, then host copies into the caller's arrays
        cudaStream_t st = ctx->stream[0];
        const size_t npad = ((size_t)n + 255) & ~(size_t)255;
        CK(cudaMemcpyAsync(sc.zeta, zeta, n * 16, cudaMemcpyHostToDevice, st));
        rc = mlm_points(ctx, s, q, rho, Gamma, sc.zeta, n, order, sc.A0, sc.A2, sc.A4, sc.nimg, sc.flags, st);
        if (rc) return rc;
        CK(cudaMemcpyAsync(ctx->hstage, sc.A0, npad * 26, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const char* h = (const char*)ctx->hstage;
        if (A0) memcpy(A0, h, n * 8);
        if (A2) memcpy(A2, h + npad * 8, n * 8);
        if (A4) memcpy(A4, h + npad * 16, n * 8);
        if (nimg) memcpy(nimg, h + npad * 24, n);
        if (flags) memcpy(flags, h + npad * 25, n);
        return 0;
    }
    // chunks alternate between the two streams: H2D(k+1) and D2H(k-1) overlap kernel(k)
    int k = 0;
    const int64_t chunk = (n / 8 > MLM_CHUNK_POINTS) ? (n + 7) / 8 : MLM_CHUNK_POINTS;   // at most ~8 chunks
    // a list without spatial order is visited along a space-filling curve (each chunk sorted on the device)
    const bool coherent = host_list_is_coherent(zeta, n);
    for (int64_t off = 0; off < n; off += chunk, ++k) {
        const int64_t cnt = (n - off < chunk) ? (n - off) : chunk;
        cudaStream_t st = ctx->stream[k & 1];
        CK(cudaMemcpyAsync(sc.zeta + 2 * off, zeta + 2 * off, cnt * 16, cudaMemcpyHostToDevice, st));
        rc = (coherent ? mlm_points : mlm_points_unordered)(
                        ctx, s, q, rho, Gamma, sc.zeta + 2 * off, cnt, order, A0 ? sc.A0 + off : nullptr,
                        A2 ? sc.A2 + off : nullptr, A4 ? sc.A4 + off : nullptr, nimg ? sc.nimg + off : nullptr,
                        flags ? sc.flags + off : nullptr, st);
        if (rc) return rc;
        rc = d2h_chunk(ctx, sc, off, cnt, A0, A2, A4, nimg, flags, st);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(ctx->stream[0]));
    CK(cudaStreamSynchronize(ctx->stream[1]));
    return 0;
}

int mlm_map_host(mlm_context* ctx, double s, double q, double rho, double Gamma, double x0, double y0, double dx,
                 double dy, int64_t nx, int64_t row0, int64_t nrows, int order, int strip, double* A0, double* A2,
                 double* A4, uint8_t* nimg, uint8_t* flags) {
    if (!ctx) return fail(ctx, MLM_EINVAL, "mlm_map_host: NULL context");
    if (nx < 0 || nrows < 0 || (order != 2 && order != 4) || bad_lens(s, q) || strip < 0 || strip > 4096)
        return fail(ctx, MLM_EINVAL, "mlm_map_host: bad argument");
    if (nx == 0 || nrows == 0) return 0;
    if (!(A0 || A2 || A4 || nimg || flags)) return fail(ctx, MLM_EINVAL, "mlm_map_host: NULL buffer");
    DeviceGuard g(ctx->device);
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    const int64_t n = nx * nrows;
    Scratch sc;
    int rc = carve(ctx, n, false, &sc);
    if (rc) return rc;
    // pipelining unit: at least MLM_CHUNK_POINTS pixels, at most ~8 chunks per call (small launches
    // leave SMs idle at the tail), a whole number of strips so that chunking does not change a bit
    int64_t rows_per = MLM_CHUNK_POINTS / nx;
    if (rows_per < (nrows + 7) / 8) rows_per = (nrows + 7) / 8;
    if (rows_per < 1) rows_per = 1;
    const int64_t L = strip = resolve_strip(ctx, strip, nx, nrows);
    rows_per = (rows_per + L - 1) / L * L;
    // the first chunk also takes the partial strip up to the next strip boundary of the absolute row
    // numbering, so that every later chunk starts on one
    const int64_t first = (row0 % L) ? L - row0 % L : 0;
    int k = 0;
    for (int64_t r = 0; r < nrows; ++k) {
        int64_t nr = (k == 0 && first) ? first + rows_per : rows_per;
        if (nr > nrows - r) nr = nrows - r;
        const int64_t off = r * nx, cnt = nr * nx;
        cudaStream_t st = ctx->stream[k & 1];
        rc = mlm_map(ctx, s, q, rho, Gamma, x0, y0, dx, dy, nx, row0 + r, nr, order, strip, A0 ? sc.A0 + off : nullptr,
                     A2 ? sc.A2 + off : nullptr, A4 ? sc.A4 + off : nullptr, nimg ? sc.nimg + off : nullptr,
                     flags ? sc.flags + off : nullptr, st);
        if (rc) return rc;
        rc = d2h_chunk(ctx, sc, off, cnt, A0, A2, A4, nimg, flags, st);
        if (rc) return rc;
        r += nr;
    }
    CK(cudaStreamSynchronize(ctx->stream[0]));
    CK(cudaStreamSynchronize(ctx->stream[1]));
    return 0;
}

static int models_host_impl(mlm_context* ctx, const char* who, const double* const* hp, int npar, int64_t M,
                            double tau0, double dtau, const double* times, int64_t Tn, int order, int seg, double* A0,
                            double* A2, double* A4, uint8_t* nimg, uint8_t* flags) {
    if (!ctx) return fail(ctx, MLM_EINVAL, who);
    if (M < 0 || Tn < 0 || (order != 2 && order != 4) || seg < 0 || seg > 65536) return fail(ctx, MLM_EINVAL, who);
    if (M == 0 || Tn == 0) return 0;
    seg = models_segment(ctx, seg, M, Tn);             // from the WHOLE grid: the chunks below must not change a bit
    for (int i = 0; i < npar; ++i)
        if (!hp[i]) return fail(ctx, MLM_EINVAL, who);
    if (!(A0 || A2 || A4 || nimg || flags)) return fail(ctx, MLM_EINVAL, who);
    DeviceGuard g(ctx->device);
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    const int64_t n = M * Tn;
    const size_t npad = ((size_t)n + 255) & ~(size_t)255;
    const size_t mpad = ((size_t)M + 31) & ~(size_t)31, tpad = ((size_t)Tn + 31) & ~(size_t)31;
    int rc = ensure_dbuf(ctx, npad * 26 + (mpad * 8 + tpad) * 8);
    if (rc) return rc;
    Scratch sc;
    rc = carve(ctx, n, false, &sc);
    if (rc) return rc;
    double* par = (double*)((char*)ctx->dbuf + npad * 26);
    double* dtimes = par + 8 * mpad;
    cudaStream_t st0 = ctx->stream[0];
    for (int i = 0; i < npar; ++i) CK(cudaMemcpyAsync(par + i * mpad, hp[i], M * 8, cudaMemcpyHostToDevice, st0));
    if (times) CK(cudaMemcpyAsync(dtimes, times, Tn * 8, cudaMemcpyHostToDevice, st0));
    CK(cudaEventRecord(ctx->ev[0], st0));
    CK(cudaStreamWaitEvent(ctx->stream[1], ctx->ev[0], 0));
    int64_t models_per = MLM_CHUNK_POINTS / Tn;
    if (models_per < (M + 7) / 8) models_per = (M + 7) / 8;          // at most ~8 chunks
    if (models_per < 1) models_per = 1;
    int k = 0;
    for (int64_t m0 = 0; m0 < M; m0 += models_per, ++k) {
        const int64_t nm = (M - m0 < models_per) ? (M - m0) : models_per;
        const int64_t off = m0 * Tn, cnt = nm * Tn;
        cudaStream_t st = ctx->stream[k & 1];
        const Epochs ep = {tau0, dtau, times ? dtimes : nullptr, npar == 8 ? par + 6 * mpad + m0 : nullptr,
                           npar == 8 ? par + 7 * mpad + m0 : nullptr};
        rc = launch_models(ctx, who, par + m0, par + mpad + m0, par + 2 * mpad + m0, par + 3 * mpad + m0,
                           par + 4 * mpad + m0, par + 5 * mpad + m0, nm, ep, Tn, order, seg, A0 ? sc.A0 + off : nullptr,
                           A2 ? sc.A2 + off : nullptr, A4 ? sc.A4 + off : nullptr, nimg ? sc.nimg + off : nullptr,
                           flags ? sc.flags + off : nullptr, st);
        if (rc) return rc;
        rc = d2h_chunk(ctx, sc, off, cnt, A0, A2, A4, nimg, flags, st);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(ctx->stream[0]));
    CK(cudaStreamSynchronize(ctx->stream[1]));
    return 0;
}

int mlm_models_host(mlm_context* ctx, const double* s, const double* q, const double* rho, const double* Gamma,
                    const double* u0, const double* alpha, int64_t M, double tau0, double dtau, int64_t Tn, int order,
                    int seg, double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags) {
    const double* hp[6] = {s, q, rho, Gamma, u0, alpha};
    return models_host_impl(ctx, "mlm_models_host: bad argument", hp, 6, M, tau0, dtau, nullptr, Tn, order, seg, A0, A2, A4,
                            nimg, flags);
}

int mlm_models_times_host(mlm_context* ctx, const double* s, const double* q, const double* rho, const double* Gamma,
                          const double* u0, const double* alpha, const double* t0, const double* tE, int64_t M,
                          const double* times, int64_t Tn, int order, int seg, double* A0, double* A2, double* A4,
                          uint8_t* nimg, uint8_t* flags) {
    if (Tn > 0 && !times) return fail(ctx, MLM_EINVAL, "mlm_models_times_host: NULL times");
    const double* hp[8] = {s, q, rho, Gamma, u0, alpha, t0, tE};
    return models_host_impl(ctx, "mlm_models_times_host: bad argument", hp, 8, M, 0.0, 0.0, times, Tn, order, seg, A0, A2,
                            A4, nimg, flags);
}

int mlm_from_wk_host(mlm_context* ctx, const double* Wk, int64_t n, int order, double rho, double Gamma, double* out) {
    if (!ctx) return fail(ctx, MLM_EINVAL, "mlm_from_wk_host: NULL context");
    if (n < 0 || (order != 2 && order != 4) || !out) return fail(ctx, MLM_EINVAL, "mlm_from_wk_host: bad argument");
    if (n > 0 && !Wk) return fail(ctx, MLM_EINVAL, "mlm_from_wk_host: NULL Wk");
    DeviceGuard g(ctx->device);
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    const int nrows = order == 4 ? 7 : 5;              // rows 0..4 or 0..6 (rows 0-1 are skipped below)
    const size_t wbytes = (size_t)(nrows - 2) * n * 16;
    int rc = ensure_dbuf(ctx, (size_t)nrows * n * 16 + 64);
    if (rc) return rc;
    cudaStream_t st = ctx->stream[0];
    double* dW = (double*)ctx->dbuf;
    double* dout = (double*)((char*)ctx->dbuf + (size_t)nrows * n * 16);
    // rows 0-1 are uninitialised in the reference (np.empty, multipoles.py:186): never copied or read
    if (n > 0) CK(cudaMemcpyAsync(dW + 4 * n, Wk + 4 * n, wbytes, cudaMemcpyHostToDevice, st));
    rc = mlm_from_wk(ctx, dW, n, order, rho, Gamma, dout, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, dout, (order == 4 ? 3 : 2) * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

int mlm_solve_host(mlm_context* ctx, double s, double q, const double* zeta, int64_t n, double* roots) {
    if (!ctx) return fail(ctx, MLM_EINVAL, "mlm_solve_host: NULL context");
    if (n < 0 || bad_lens(s, q)) return fail(ctx, MLM_EINVAL, "mlm_solve_host: bad argument");
    if (n == 0) return 0;
    if (!zeta || !roots) return fail(ctx, MLM_EINVAL, "mlm_solve_host: NULL buffer");
    DeviceGuard g(ctx->device);
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    int rc = ensure_dbuf(ctx, (size_t)n * 96);
    if (rc) return rc;
    cudaStream_t st = ctx->stream[0];
    double* dz = (double*)ctx->dbuf;
    double* dr = dz + 2 * n;
    CK(cudaMemcpyAsync(dz, zeta, n * 16, cudaMemcpyHostToDevice, st));
    rc = mlm_solve(ctx, s, q, dz, n, dr, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(roots, dr, n * 80, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

int mlm_akl_host(mlm_context* ctx, const double* mu, const double* W2, const double* Q, int64_t n, double* out) {
    if (!ctx) return fail(ctx, MLM_EINVAL, "mlm_akl_host: NULL context");
    if (n < 0) return fail(ctx, MLM_EINVAL, "mlm_akl_host: bad argument");
    if (n == 0) return 0;
    if (!mu || !W2 || !Q || !out) return fail(ctx, MLM_EINVAL, "mlm_akl_host: NULL buffer");
    DeviceGuard g(ctx->device);
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    int rc = ensure_dbuf(ctx, (size_t)n * 64);
    if (rc) return rc;
    cudaStream_t st = ctx->stream[0];
    double* dmu = (double*)ctx->dbuf;
    double* dW = dmu + 2 * n;
    double* dQ = dW + 2 * n;
    double* dO = dQ + 2 * n;
    CK(cudaMemcpyAsync(dmu, mu, n * 16, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dW, W2, n * 16, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dQ, Q, n * 16, cudaMemcpyHostToDevice, st));
    rc = mlm_akl(ctx, dmu, dW, dQ, n, dO, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, dO, n * 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

// ---- measurement helpers ----------------------------------------------------------------------
int mlm_fp64_peak(mlm_context* ctx, double* tflops) {
    if (!ctx || !tflops) return fail(ctx, MLM_EINVAL, "mlm_fp64_peak: NULL argument");
    DeviceGuard g(ctx->device);
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    int rc = ensure_dbuf(ctx, 64);
    if (rc) return rc;
    cudaStream_t st = ctx->stream[0];
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int iters = 4096, blocks = ctx->sm_count * 8, threads = 256;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(e0, st));
        k_dfma_peak<<<blocks, threads, 0, st>>>((double*)ctx->dbuf, iters, 0.999999, 1e-7);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        ctx->launches++;
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double flop = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
        const double tf = flop / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CK(cudaGetLastError());
    *tflops = best;
    return 0;
}

/* experiment: instructions/s of the FP64 pipe under a realistic operand pattern, in units of the
 * nominal issue rate (fraction of 1 warp-instruction per 2 cycles per SM sub-partition is left to the caller):
 * returns executed FP64 thread-instructions per second */
int mlm_fp64_probe(mlm_context* ctx, int mode, int blocks_per_sm, double* inst_per_s) {
    if (!ctx || !inst_per_s || mode < 1 || mode > 2 || blocks_per_sm < 1) return fail(ctx, MLM_EINVAL, "mlm_fp64_probe: bad argument");
    DeviceGuard g(ctx->device);
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    int rc = ensure_dbuf(ctx, 256);
    if (rc) return rc;
    cudaStream_t st = ctx->stream[0];
    double h[24];
    for (int k = 0; k < 24; ++k) h[k] = (k < 8) ? 1.0 + 0.01 * k : (k < 16 ? 0.999999 - 1e-7 * k : 1e-7 * (k - 15));
    CK(cudaMemcpyAsync(ctx->dbuf, h, sizeof(h), cudaMemcpyHostToDevice, st));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int iters = 2048, blocks = ctx->sm_count * blocks_per_sm, threads = 128;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0, st));
        if (mode == 1) k_fp64_probe<1><<<blocks, threads, 0, st>>>((double*)ctx->dbuf + 32, iters, (const double*)ctx->dbuf);
        else k_fp64_probe<2><<<blocks, threads, 0, st>>>((double*)ctx->dbuf + 32, iters, (const double*)ctx->dbuf);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        ctx->launches++;
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double inst = 32.0 * (double)iters * (double)blocks * (double)threads;
        const double r = inst / (ms * 1e-3);
        if (rep > 0 && r > best) best = r;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CK(cudaGetLastError());
    *inst_per_s = best;
    return 0;
}

// Event counters of the -DMLM_STATS build: out[2 * e] = lanes, out[2 * e + 1] = warp-level occurrences of event e
// (order: mlm_stats_names()); the product build has no counters and answers MLM_ENOTSUP.
const char* mlm_stats_names(void) {
#define X(n) #n " "
    return MLM_CNT_NAMES(X);
#undef X
}

int mlm_stats_read(mlm_context* ctx, uint64_t* out, int n, int reset) {
    if (!ctx || !out || n < 2 * MLM_CNT_N) return fail(ctx, MLM_EINVAL, "mlm_stats_read: bad argument");
#ifdef MLM_STATS
    DeviceGuard g(ctx->device);
    CK(cudaDeviceSynchronize());
    unsigned long long h[2][MLM_CNT_N];
    CK(cudaMemcpyFromSymbol(h, g_mlm_cnt, sizeof(h)));
    for (int e = 0; e < MLM_CNT_N; ++e) { out[2 * e] = h[0][e]; out[2 * e + 1] = h[1][e]; }
    if (reset) {
        memset(h, 0, sizeof(h));
        CK(cudaMemcpyToSymbol(g_mlm_cnt, h, sizeof(h)));
    }
    return 0;
#else
    (void)reset;
    return fail(ctx, MLM_ENOTSUP, "mlm_stats_read: this is the product build (counters exist in libmlmultipoles_stats.so only)");
#endif
}

int mlm_store_probe(mlm_context* ctx, int mode, int64_t nx, int64_t nrows, int rows_per_block, double* A0, double* A2,
                    double* A4, uint8_t* nimg, uint8_t* flags, void* stream) {
    if (!ctx || mode < 0 || mode > 3 || nx <= 0 || nx % MLM_BLOCK || nrows <= 0 || rows_per_block <= 0 || !A0 || !A2 || !A4 ||
        !nimg || !flags)
        return fail(ctx, MLM_EINVAL, "mlm_store_probe: bad argument");
    DeviceGuard g(ctx->device);
    dim3 grid((unsigned)(nx / MLM_BLOCK), (unsigned)((nrows + rows_per_block - 1) / rows_per_block));
    k_store_probe<<<grid, MLM_BLOCK, 0, pick(ctx, stream)>>>(mode, nx, nrows, rows_per_block, A0, A2, A4, nimg, flags);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

int mlm_kernel_info(mlm_context* ctx, int32_t* out, int n) {
    if (!ctx || !out || n < 4) return fail(ctx, MLM_EINVAL, "mlm_kernel_info: bad argument");
    DeviceGuard g(ctx->device);
    cudaFuncAttributes a;
    CK(cudaFuncGetAttributes(&a, (const void*)k_points<4>));
    out[0] = a.numRegs;
    out[1] = (int32_t)a.localSizeBytes;
    CK(cudaFuncGetAttributes(&a, (const void*)k_map<4>));
    out[2] = a.numRegs;
    out[3] = (int32_t)a.localSizeBytes;
    return 0;
}

}  // extern "C"
