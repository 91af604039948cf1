/* mlmultipoles.h -- C ABI of libmlmultipoles.so (B200 / sm_100a, FP64).
 *
 * Drop-in boundary for the one hot path of ArnaudCassan/microlensing: the Cassan (2017)
 * binary-lens point-source (A0), quadrupole (A2) and hexadecapole (A4) magnifications of
 * microlensing/multipoles/multipoles.py.  The reference has no FFI; its boundary is the
 * Python function surface of that module.  Each entry point below names
 * the reference interface it replaces (file:line under /root/reference).  The Python side
 * (microlensing_b200/_capi.py, ctypes) is the only caller in this repo; INTEGRATION.md
 * shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; complex numbers are interleaved (re, im) doubles;
 *  - every function returns 0 on success, a negative MLM_E* code for argument errors, or a
 *    positive cudaError_t; mlm_last_error() gives the message; nothing throws;
 *  - the caller owns every buffer.  Un-suffixed functions take DEVICE pointers and enqueue
 *    on `stream` (a cudaStream_t passed as void*, NULL = the CUDA default stream) without
 *    synchronising; *_host functions take HOST pointers, copy in/out themselves and return
 *    when the results are in the host buffers;
 *  - frames: `zeta` arguments are in the PRIMARY-LENS frame (primary mass 1/(1+q) at 0,
 *    secondary q/(1+q) at -s), exactly what _solvelenseq receives (multipoles.py:156);
 *    map/light-curve generators take CENTRE-OF-MASS coordinates and apply
 *    zeta = zeta_cm - s q/(1+q) themselves (multipoles.py:177-178);
 *  - order = 2 computes A0, A2 (quadrupole, multipoles.py:12-65; A4 output is set to 0),
 *    order = 4 computes A0, A2, A4 (hexadecapole, multipoles.py:67-148);
 *  - any of the per-position outputs A0 / A2 / A4 / nimg / flags may be NULL (not stored, not
 *    copied back); at least one must be given;
 *  - threads: a context may be shared by host threads.  Entry points that use the context's own device
 *    scratch (every *_host function, mlm_from_wk) serialise on a mutex inside the context; the other
 *    device-pointer entry points only enqueue on the caller's stream and keep no state.
 * There is no CPU fallback anywhere in the library: without a CUDA device mlm_create fails.
 */
#ifndef MLMULTIPOLES_H
#define MLMULTIPOLES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mlm_context mlm_context;

/* argument-error codes (negative; CUDA errors are returned as positive cudaError_t) */
#define MLM_OK 0
#define MLM_EINVAL (-1)   /* bad argument (NULL pointer, order not 2/4, negative size, s<=0, q<=0) */
#define MLM_ENOMEM (-2)   /* host allocation failed */
#define MLM_ENOTSUP (-3)  /* not available in this build of the library */

/* per-position flag bits written to `flags` (the reference flags nothing) */
#define MLM_FLAG_NOCONV 1    /* a root-finder iteration cap was hit */
#define MLM_FLAG_EXTRA 2     /* nimg includes image(s) that the reference's plain test |z - conj(W1) - zeta| < 1e-6
                              * (multipoles.py:183) rejects on this solver's polynomial roots: kept because Newton on the
                              * lens equation converges from the root (root noise of the quintic at q <~ 1e-3), or found
                              * directly from the lens equation when an even count showed that an image next to a lens was
                              * lost.  Bits 6-7 hold how many: nimg - ((flags >> MLM_FLAG_EXTRA_SHIFT) & 3) is the count of the
                              * plain test.  Carried along while the images are tracked from position to position. */
#define MLM_FLAG_EXTRA_SHIFT 6
#define MLM_FLAG_NEARCRIT 4  /* min over images of |1-|W2|^2| < 1e-3 (near a critical curve) */
#define MLM_FLAG_AMBIG 8     /* a root's lens-equation residual within a decade of 1e-6 (multipoles.py:183) */
#define MLM_FLAG_SERIES 16   /* |A4-A2| > |A2-A0|: multipole series not converging */
#define MLM_FLAG_DEGEN 32    /* leading polynomial coefficient exactly 0 (zeta = 0 or zeta = -s) */

/* ---- context --------------------------------------------------------------------------- */
int mlm_create(int device, mlm_context** out);
int mlm_destroy(mlm_context* ctx);
const char* mlm_last_error(mlm_context* ctx);          /* ctx may be NULL: last global error */
const char* mlm_version(void);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int mlm_launch_count(mlm_context* ctx, int64_t* out);
int mlm_synchronize(mlm_context* ctx);
/* pinned host memory for *_host outputs (so the D2H copies overlap with the kernels) */
int mlm_alloc_host(mlm_context* ctx, int64_t bytes, void** out);
int mlm_free_host(mlm_context* ctx, void* p);

/* Device buffers that the other processes of the box can map into their address space
 * (cudaMalloc + CUDA IPC).  Multi-GPU maps use them for the final gather: rank 0
 * allocates the full-map outputs and exports the handle, every other rank opens it and passes
 * pointers INTO IT to mlm_map, whose kernel then stores its rows straight into rank 0's HBM over
 * NVLink (peer stores from the kernel epilogue; no separate collective, no staging copy). */
#define MLM_IPC_HANDLE_BYTES 64
int mlm_device_alloc(mlm_context* ctx, int64_t bytes, void** out);
int mlm_device_free(mlm_context* ctx, void* p);
int mlm_ipc_export(mlm_context* ctx, void* p, unsigned char* handle /* [MLM_IPC_HANDLE_BYTES] */);
int mlm_ipc_open(mlm_context* ctx, const unsigned char* handle, void** out);
int mlm_ipc_close(mlm_context* ctx, void* p);
/* cudaMemcpyAsync(cudaMemcpyDefault) on `stream`: bulk copy of a contiguous result block into a
 * peer-mapped buffer (model grids, whose kernel stores are strided and would cross NVLink 8 bytes at a time). */
int mlm_memcpy_async(mlm_context* ctx, void* dst, const void* src, int64_t bytes, void* stream);

/* 2-D variant (cudaMemcpy2DAsync): `height` rows of `width` bytes, pitches in bytes.  Pushes the strips a rank
 * owns under a cyclic / weighted deal (equally spaced row blocks of the full map) into the gather root's buffer
 * with one call per output array. */
int mlm_memcpy2d_async(mlm_context* ctx, void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width,
                       int64_t height, void* stream);
/* Page-lock an existing host range (cudaHostRegister, portable) / undo it.  The sharded host API maps ONE
 * POSIX shared-memory result map into every rank's process; each rank registers the rows it owns and its
 * D2H copies land directly in the map that rank 0 returns. */
int mlm_host_register(mlm_context* ctx, void* p, int64_t bytes);
int mlm_host_unregister(mlm_context* ctx, void* p);

/* Strip length the map entry points choose for strip = 0: a power of two in 16..64 from the size of the launch
 * (nx x nrows pixels on one GPU, about 12 waves of work items).  Callers that compute ONE map in several calls
 * (row blocks, several GPUs) evaluate it once for the rows per call and pass the value to every call. */
int mlm_auto_strip(mlm_context* ctx, int64_t nx, int64_t nrows);

/* ---- the fused path: replaces the per-position body of example(), multipoles.py:181-202
 *      (_solvelenseq :156-165, selection :182-183, Wk :185-201, hexadecapole :67-148) -------- */

/* N source positions, one (s,q,rho,Gamma).  zeta: n complex (primary frame), in any order.  Consecutive
 * entries that are close to each other (a light curve along any trajectory, a scan line) are exploited:
 * a thread walks along a run of consecutive entries carrying the images from one to the next and
 * solves cold only at the start of its run, after a jump of more than 0.02 Einstein radii, or where
 * tracking is rejected; an unordered list costs one cold solve per entry.  The result of an entry
 * agrees with its cold solve to rounding (<= 1e-11 away from caustics), the image count exactly. */
int mlm_points(mlm_context* ctx, double s, double q, double rho, double Gamma,
               const double* zeta, int64_t n, int order,
               double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags, void* stream);
/* The same for a list WITHOUT spatial order (random positions, a shuffled catalogue): every entry is solved on its
 * own, nothing is carried between entries, so the result of an entry does not depend on the rest of the list.  First
 * pass: the three images every binary-lens configuration has, by Newton on the lens equation from single-lens guesses,
 * the other two roots of the quintic by Vieta; entries that this does not solve (inside and next to resonant caustics)
 * are compacted into a work list and solved by the root finder (Laguerre + deflation + polish) in a second launch.
 * Image counts equal mlm_points' cold solve, values to rounding; MLM_FLAG_EXTRA / MLM_FLAG_AMBIG report polynomial
 * rounding noise and agree with the root-finder path in distribution, not position by position.  n < 2^31.
 * mlm_points_host checks a sample of the list for coherence and picks between the two. */
int mlm_points_unordered(mlm_context* ctx, double s, double q, double rho, double Gamma,
                         const double* zeta, int64_t n, int order,
                         double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags, void* stream);
int mlm_points_host(mlm_context* ctx, double s, double q, double rho, double Gamma,
                    const double* zeta, int64_t n, int order,
                    double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags);

/* Magnification map: pixel (row r, column c) has centre-of-mass source position
 *   x = x0 + (c + 0.5) dx,  y = y0 + (r + 0.5) dy          (generated in the kernel),
 * rows [row0, row0 + nrows) of an nx-wide map are written to out[(r - row0) * nx + c]
 * (row-shardable: "... to be computed for all source center positions", multipoles.py:184).
 * Rows are processed in strips of `strip` consecutive rows (aligned to absolute row numbers) with warm-started
 * root tracking inside a strip.  `strip` is a numerical parameter like a tile size: a shorter strip means more
 * independent work items (better tail behaviour of small launches, e.g. 1/8 of a map per GPU) and more cold root
 * solves (~ +2000/strip FP64 instructions per pixel); results for different strip lengths agree to rounding
 * (<= 4e-11), for ONE strip length the map is bit-identical under any strip-aligned sharding.  strip = 0 chooses
 * mlm_auto_strip(nx, nrows) (env MLM_MAP_STRIP, read at mlm_create, overrides that choice: experiments). */
int mlm_map(mlm_context* ctx, double s, double q, double rho, double Gamma,
            double x0, double y0, double dx, double dy, int64_t nx, int64_t row0, int64_t nrows, int order, int strip,
            double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags, void* stream);
/* Same, restricted to the strips this caller owns under a DEAL: strip k = rows [k L, (k+1) L), L = strip, belongs
 * to the call iff (k mod period) is one of sel[0..nsel) (ascending, nsel <= 32); rows of other strips are not
 * touched.  Sharding a map over GPUs: every rank calls it with row0 = 0, nrows = ny, output pointers to the FULL
 * map and its own positions of the round -- period = N, sel = {rank} deals the strips cyclically (every rank gets
 * the same mix of cheap and expensive, near-caustic, rows); unequal nsel gives a weighted deal (the gather root
 * takes more strips because its results do not cross NVLink).  The union over ranks is bit-identical to the
 * unsharded map.  mlm_map_strips is the cyclic special case (period = kstep, sel = {kphase}). */
int mlm_map_deal(mlm_context* ctx, double s, double q, double rho, double Gamma,
                 double x0, double y0, double dx, double dy, int64_t nx, int64_t row0, int64_t nrows,
                 int64_t period, int nsel, const int32_t* sel, int order, int strip,
                 double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags, void* stream);
int mlm_map_strips(mlm_context* ctx, double s, double q, double rho, double Gamma,
                   double x0, double y0, double dx, double dy, int64_t nx, int64_t row0, int64_t nrows,
                   int64_t kphase, int64_t kstep, int order, int strip,
                   double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags, void* stream);
int mlm_map_host(mlm_context* ctx, double s, double q, double rho, double Gamma,
                 double x0, double y0, double dx, double dy, int64_t nx, int64_t row0, int64_t nrows, int order, int strip,
                 double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags);

/* Batched model grid: M models x T epochs.  Model m has (s,q,rho,Gamma,u0,alpha)[m] (device
 * arrays of length M); epoch t has tau = tau0 + t dtau; centre-of-mass source position
 * zeta_cm = (tau + i u0) exp(i alpha).  Outputs are [m * T + t].  A thread walks a SEGMENT of `seg` consecutive
 * epochs of one model with warm-started root tracking (cold solve at the segment start); blocks of 64 threads serve
 * 1..16 models at a time.  `seg` is a numerical parameter like the strip length of a map: the bits of a light curve
 * depend on it (results for different seg agree to rounding), NOT on M or on how the grid is sharded over GPUs --
 * shards of one grid must be computed with ONE seg.  seg = 0 chooses mlm_auto_segment(M, T) of THIS call (long
 * segments as long as the grid fills the GPU: 128 epochs for 1e4 x 2000, 9 for 1 x 1e6; env MLM_SEGMENT overrides). */
int mlm_auto_segment(mlm_context* ctx, int64_t M, int64_t T);
int mlm_models(mlm_context* ctx, const double* s, const double* q, const double* rho, const double* Gamma,
               const double* u0, const double* alpha, int64_t M, double tau0, double dtau, int64_t T, int order, int seg,
               double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags, void* stream);
int mlm_models_host(mlm_context* ctx, const double* s, const double* q, const double* rho, const double* Gamma,
                    const double* u0, const double* alpha, int64_t M, double tau0, double dtau, int64_t T, int order, int seg,
                    double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags);

/* The same with OBSERVATION TIMES instead of a uniform grid ("in-kernel trajectory
 * zeta(t; t0, tE, u0, alpha)"): epoch t of model m is at tau = (times[t] - t0[m]) / tE[m]; t0, tE are device
 * arrays of length M (NULL: 0 and 1, i.e. times are already in Einstein times from closest approach).
 * Unevenly spaced epochs are tracked with the extrapolation scaled by the ratio of successive steps; a gap of
 * more than 0.02 Einstein times restarts cold. */
int mlm_models_times(mlm_context* ctx, const double* s, const double* q, const double* rho, const double* Gamma,
                     const double* u0, const double* alpha, const double* t0, const double* tE, int64_t M,
                     const double* times, int64_t T, int order, int seg,
                     double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags, void* stream);
int mlm_models_times_host(mlm_context* ctx, const double* s, const double* q, const double* rho, const double* Gamma,
                          const double* u0, const double* alpha, const double* t0, const double* tE, int64_t M,
                          const double* times, int64_t T, int order, int seg,
                          double* A0, double* A2, double* A4, uint8_t* nimg, uint8_t* flags);

/* Model grid with the light-curve fit fused in ("fused chi^2 reduction against observed
 * fluxes"): same models and epochs as mlm_models, but instead of M x T magnifications the kernel returns,
 * per model, the least-squares linear flux fit F_t = Fs A_t + Fb of the model magnifications A_t (A4 for
 * order 4, A2 for order 2) to the observed fluxes flux[T] with weights weight[T] = 1/sigma^2:
 *   stats[8 m + 0..7] = { chi2_min, Fs, Fb, sum w A, sum w A^2, sum w A f, #epochs with 5 images, #flagged epochs }
 * (fixed-order reduction over the threads of a model: deterministic; seg as in mlm_models). */
int mlm_models_chi2(mlm_context* ctx, const double* s, const double* q, const double* rho, const double* Gamma,
                    const double* u0, const double* alpha, int64_t M, double tau0, double dtau, int64_t T, int order, int seg,
                    const double* flux, const double* weight, double* stats, void* stream);
int mlm_models_chi2_host(mlm_context* ctx, const double* s, const double* q, const double* rho, const double* Gamma,
                         const double* u0, const double* alpha, int64_t M, double tau0, double dtau, int64_t T, int order, int seg,
                         const double* flux, const double* weight, double* stats);

/* ... and against a light curve observed at times[T] (t0, tE per model as in mlm_models_times). */
int mlm_models_chi2_times(mlm_context* ctx, const double* s, const double* q, const double* rho, const double* Gamma,
                          const double* u0, const double* alpha, const double* t0, const double* tE, int64_t M,
                          const double* times, int64_t T, int order, int seg,
                          const double* flux, const double* weight, double* stats, void* stream);
int mlm_models_chi2_times_host(mlm_context* ctx, const double* s, const double* q, const double* rho, const double* Gamma,
                               const double* u0, const double* alpha, const double* t0, const double* tE, int64_t M,
                               const double* times, int64_t T, int order, int seg,
                               const double* flux, const double* weight, double* stats);

/* ---- drop-ins for the individual reference functions ------------------------------------ */

/* quadrupole(Wk,rho,Gamma) multipoles.py:12 (order 2, out = [A0,A2]) and
 * hexadecapole(Wk,rho,Gamma) multipoles.py:67 (order 4, out = [A0,A2,A4]).
 * Wk: the reference's (7, n) complex128 C-ordered table, rows 2..(order==2 ? 4 : 6) are read,
 * rows 0-1 never (np.empty garbage in the reference, multipoles.py:186).  Sums over all n. */
int mlm_from_wk(mlm_context* ctx, const double* Wk, int64_t n, int order, double rho, double Gamma,
                double* out, void* stream);
int mlm_from_wk_host(mlm_context* ctx, const double* Wk, int64_t n, int order, double rho, double Gamma,
                     double* out);

/* _solvelenseq(s,q,zeta) multipoles.py:156: all roots of the lens polynomial, 5 complex per
 * position (NaN where np.roots would return fewer: zeta = 0, zeta = -s), order unspecified. */
int mlm_solve(mlm_context* ctx, double s, double q, const double* zeta, int64_t n, double* roots, void* stream);
int mlm_solve_host(mlm_context* ctx, double s, double q, const double* zeta, int64_t n, double* roots);

/* _akl(mu,W2,Qkl) multipoles.py:150: out = mu (conj(Q) + conj(W2) Q), elementwise; all four
 * arrays are n complex numbers (a real mu is passed with zero imaginary parts). */
int mlm_akl(mlm_context* ctx, const double* mu, const double* W2, const double* Q, int64_t n, double* out,
            void* stream);
int mlm_akl_host(mlm_context* ctx, const double* mu, const double* W2, const double* Q, int64_t n, double* out);

/* ---- measurement helpers ----------------------------------------------------------------- */
/* DFMA-chain microbenchmark: measured FP64 FMA throughput of this GPU in TFLOP/s
 * (2 flop per DFMA), the denominator of the FP64 roofline (MEASURED_PEAKS.json has none). */
int mlm_fp64_peak(mlm_context* ctx, double* tflops);
/* FP64-pipe probe with a realistic operand pattern (three distinct register pairs per instruction):
 * mode 1 = DFMA only, mode 2 = DFMA/DMUL/DADD 11:4:1; blocks_per_sm x 128 threads resident per SM.
 * Returns executed FP64 thread-instructions per second (nominal B200: 148 x 64 x 1.965e9 = 1.86e13). */
int mlm_fp64_probe(mlm_context* ctx, int mode, int blocks_per_sm, double* inst_per_s);
/* store-pattern probe: the store side of the map kernel without its arithmetic, into the given (possibly peer-mapped)
 * arrays of nx x nrows positions; mode 0 = the five SoA arrays like the map kernel, 1 = float64 arrays only,
 * 2 = uint8 arrays only, 3 = five arrays with the uint8 rows written 128 bytes at a time.  nx: multiple of 128. */
int mlm_store_probe(mlm_context* ctx, int mode, int64_t nx, int64_t nrows, int rows_per_block, double* A0, double* A2,
                    double* A4, uint8_t* nimg, uint8_t* flags, void* stream);
/* static description of the fused kernel (registers per thread, spill bytes) from
 * cudaFuncGetAttributes: [0]=regs points<4>, [1]=local bytes points<4>, [2]=regs map<4>, [3]=local bytes map<4> */
int mlm_kernel_info(mlm_context* ctx, int32_t* out, int n);
/* Event counters (measurement only).  libmlmultipoles_stats.so -- the same sources compiled with -DMLM_STATS, built
 * next to the product library, loaded by bench.py / tools only -- counts what the tracking kernels execute: source
 * positions, Laguerre / quintic-Newton / lens-equation-Newton evaluations, cold solves, classifications, images summed
 * (the MLM_COUNT hooks of csrc/mlm_core.cuh).  mlm_stats_names(): the event names, blank-separated, in order;
 * mlm_stats_read(): out[2e] = lanes, out[2e+1] = warp-level occurrences of event e since the last reset (synchronises
 * the device).  The product build counts nothing: mlm_stats_read returns MLM_ENOTSUP. */
const char* mlm_stats_names(void);
int mlm_stats_read(mlm_context* ctx, uint64_t* out, int n, int reset);

#ifdef __cplusplus
}
#endif
#endif /* MLMULTIPOLES_H */
