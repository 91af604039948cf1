// mlm_core.cuh -- per-source-position arithmetic of the binary-lens A0/A2/A4 path.
//
// Everything here runs in ONE thread (one source position; five roots of state in shared memory):
//   (a) lens polynomial from (s, q, zeta)        reference: multipoles.py:159-164
//   (b) all complex roots (Laguerre + deflation, Newton polish)   replaces np.roots, multipoles.py:165
//   (c) physical images by the residual test     reference: multipoles.py:182-183
//   (d) W_2..W_6 per image                        reference: multipoles.py:197-201
//   (e) mu0, mu2, mu4 per image and the sums      reference: multipoles.py:91-146
// Along a coherent sequence of positions (map column, light curve) (b)+(c) are replaced by TRACKING:
// images follow Newton on the lens equation, non-physical roots Newton on the quintic, and a cold
// (b)+(c) happens only at sequence starts and where tracking is rejected (advance_images).
//
// (e) is NOT a transcription of the reference's (xi, eta) recursion over 20 Taylor factors a_kl:
// mu2 and mu4 are the source-plane Laplacian and bi-Laplacian of mu0, evaluated in closed form with
// image-plane derivative operators (see image_terms_from_W and DESIGN.md, "operator form"), ~140
// FP64 instructions per image instead of ~1900 flop.
// tests/ verifies it against the oracle's literal restatement of the reference formulae.
//
// The functions are __host__ __device__ so that tests/ can compile them with g++ into a
// test-only checker of the kernel logic (tests/host_emu); the product only ever calls them
// from CUDA kernels (mlm_kernels.cu).
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MLM_HD __host__ __device__ __forceinline__
#define MLM_HDI __host__ __device__ inline
#else
#define MLM_HD inline __attribute__((always_inline))
#define MLM_HDI inline
#endif

namespace mlm {

// ---------------------------------------------------------------------------------------
// complex double in two registers
// ---------------------------------------------------------------------------------------
struct cd {
    double x, y;
};

MLM_HD cd mk(double x, double y) { cd r; r.x = x; r.y = y; return r; }
MLM_HD cd operator+(cd a, cd b) { return mk(a.x + b.x, a.y + b.y); }
MLM_HD cd operator-(cd a, cd b) { return mk(a.x - b.x, a.y - b.y); }
MLM_HD cd operator-(cd a) { return mk(-a.x, -a.y); }
MLM_HD cd conj(cd a) { return mk(a.x, -a.y); }
MLM_HD cd scale(double s, cd a) { return mk(s * a.x, s * a.y); }
MLM_HD double norm2(cd a) { return fma(a.x, a.x, a.y * a.y); }
// a*b
MLM_HD cd cmul(cd a, cd b) { return mk(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x)); }
// conj(a)*b
MLM_HD cd cmulc(cd a, cd b) { return mk(fma(a.x, b.x, a.y * b.y), fma(a.x, b.y, -(a.y * b.x))); }
// a*a
MLM_HD cd csq(cd a) { return mk(fma(a.x, a.x, -(a.y * a.y)), 2.0 * (a.x * a.y)); }
// a*b + c
MLM_HD cd cfma(cd a, cd b, cd c) {
    return mk(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
// conj(a)*b + c
MLM_HD cd cjfma(cd a, cd b, cd c) {
    return mk(fma(a.x, b.x, fma(a.y, b.y, c.x)), fma(a.x, b.y, fma(-a.y, b.x, c.y)));
}
// s*a + c (real s)
MLM_HD cd rfma(double s, cd a, cd c) { return mk(fma(s, a.x, c.x), fma(s, a.y, c.y)); }
// Re(a*b), Re(a*conj(b))
MLM_HD double re_mul(cd a, cd b) { return fma(a.x, b.x, -(a.y * b.y)); }
MLM_HD double re_mulc(cd a, cd b) { return fma(a.x, b.x, a.y * b.y); }

// ---------------------------------------------------------------------------------------
// fast FP64 reciprocal / reciprocal square root.
// On the device: MUFU.RCP64H / MUFU.RSQ64H seed (rcp.approx.ftz.f64, ~2^-20 relative) and
// two Newton steps on the FP64 pipe; no special-case branch (0 -> NaN/inf, which the callers
// rely on: a root at a lens position must yield a NaN residual exactly like 1/0 in NumPy).
// ---------------------------------------------------------------------------------------
MLM_HD double rcp(double a) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    return r;
#else
    return 1.0 / a;
#endif
}

// reciprocal good to ~1e-13 relative (one Newton step): for Newton CORRECTIONS, whose own relative
// accuracy need not exceed the size of the next correction
MLM_HD double rcp_step(double a) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    const double e = fma(-a, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / a;
#endif
}

MLM_HD double rsqrt_(double a) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    // Newton for 1/sqrt(a): r <- r + r*(1 - a r^2)/2
    double h = 0.5 * r;
    double e = fma(-a * r, r, 1.0);
    r = fma(h, e, r);
    h = 0.5 * r;
    e = fma(-a * r, r, 1.0);
    r = fma(h, e, r);
    return r;
#else
    return 1.0 / sqrt(a);
#endif
}

// 1/a (complex)
MLM_HD cd crcp(cd a) {
    double r = rcp(norm2(a));
    return mk(a.x * r, -(a.y * r));
}

// principal square root of a complex number (2 rsqrt, no division)
MLM_HD cd csqrt_(cd a) {
    double n2 = norm2(a);
    if (!(n2 > 0.0)) return mk(0.0, 0.0);            // 0 (or NaN -> 0, caller copes)
    double r = n2 * rsqrt_(n2);                       // |a|
    double h = 0.5 * (r + fabs(a.x));                 // >= |a|/2 > 0
    double rs = rsqrt_(h);
    double t = h * rs;                                // sqrt(h)
    double u = 0.5 * a.y * rs;                        // a.y / (2 sqrt(h))
    return (a.x >= 0.0) ? mk(t, u) : mk(fabs(u), (a.y >= 0.0) ? t : -t);
}

// ---------------------------------------------------------------------------------------
// single-precision complex helpers: only for the WATCH of the two non-physical roots on the tracking path
// (advance_images), a yes/no question with a threshold of 0.1 -- it runs on the FP32 pipe, which is otherwise
// idle in these kernels, instead of the FP64 pipe that bounds them.  MUFU seeds without refinement (1-2 ulp).
// ---------------------------------------------------------------------------------------
struct cf {
    float x, y;
};
MLM_HD cf mkf(float x, float y) { cf r; r.x = x; r.y = y; return r; }
MLM_HD cf tof(cd a) { return mkf((float)a.x, (float)a.y); }
MLM_HD float rcpf_(float a) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
#else
    return 1.0f / a;
#endif
}
MLM_HD float rsqrtf_(float a) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rsqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
#else
    return 1.0f / sqrtf(a);
#endif
}
MLM_HD float norm2f(cf a) { return fmaf(a.x, a.x, a.y * a.y); }
MLM_HD cf cmulf(cf a, cf b) { return mkf(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x)); }
MLM_HD cf crcpf(cf a) {
    const float r = rcpf_(norm2f(a));
    return mkf(a.x * r, -(a.y * r));
}
MLM_HD cf csqrtf_(cf a) {                             // principal square root, same scheme as csqrt_
    const float n2 = norm2f(a);
    if (!(n2 > 0.0f)) return mkf(0.0f, 0.0f);
    const float r = n2 * rsqrtf_(n2);
    const float h = 0.5f * (r + fabsf(a.x));
    const float rs = rsqrtf_(h);
    const float t = h * rs;
    const float u = 0.5f * a.y * rs;
    return (a.x >= 0.0f) ? mkf(t, u) : mkf(fabsf(u), (a.y >= 0.0f) ? t : -t);
}

// ---------------------------------------------------------------------------------------
// per-(s,q) table (kernel parameter -> constant bank; built on the host by mlm_make_table)
// ---------------------------------------------------------------------------------------
struct LensTable {
    double s, q;
    double m2;                                        // (1+q)^2
    double k40, k41, k42, k43;
    double k30, k31, k32, k33, k34, k35;
    double k20, k21, k22, k23, k24;
    double k10, k11, k12;
    double k00;
    double alpha[7];                                  // W_k = alpha_k z^-k + beta_k (z+s)^-k, k=1..6
    double beta[7];
    float s_f, alpha1_f, beta1_f, pad_f;              // s, alpha[1], beta[1] in single precision (watch_nonphysical)
};

MLM_HDI void make_table(double s, double q, LensTable* t) {
    const double m = 1.0 + q, s2 = s * s, s3 = s2 * s, m2 = m * m;
    t->s = s; t->q = q; t->m2 = m2;
    t->k40 = m * s * q; t->k41 = -m2 * s; t->k42 = 1.0 + 2.0 * s2; t->k43 = 2.0 * s;
    t->k30 = m * s2 * q; t->k31 = -s * m2; t->k32 = m * (2.0 * s + s3 * m); t->k33 = m2 * s2;
    t->k34 = -2.0 * m2 * (1.0 + s2); t->k35 = -2.0 * m2 * s;
    t->k20 = -m * s * q; t->k21 = -m * s2 * (q - 1.0); t->k22 = -m * (m + s2 * (2.0 + q));
    t->k23 = -m * (2.0 * s * (2.0 + q) + s3 * m); t->k24 = -m2 * s2;
    t->k10 = -s * m * (2.0 + s2); t->k11 = -2.0 * s2 * m; t->k12 = -s2 * q;
    t->k00 = -s2;
    double f = 1.0;
    t->alpha[0] = t->beta[0] = 0.0;
    for (int k = 1; k <= 6; ++k) {
        if (k > 1) f *= -(double)(k - 1);             // 1, -1, 2, -6, 24, -120
        t->alpha[k] = f / m;
        t->beta[k] = f * q / m;
    }
    t->s_f = (float)s; t->alpha1_f = (float)t->alpha[1]; t->beta1_f = (float)t->beta[1]; t->pad_f = 0.0f;
}

// (a) coefficients c[k] of z^k, k = 0..5        (multipoles.py:159-164, A.2 of SURVEY.md)
MLM_HD void lens_polynomial(const LensTable& T, cd zeta, cd c[6]) {
    const cd zb = conj(zeta);
    const double n = norm2(zeta);
    const cd szb = mk(T.s + zb.x, zb.y);
    c[5] = scale(T.m2, cmul(zb, szb));
    {   // c4 = k40 + k41 n + m2 zb (k42 - n + k43 zb)
        cd in = mk(fma(T.k43, zb.x, T.k42 - n), T.k43 * zb.y);
        cd t = scale(T.m2, cmul(zb, in));
        c[4] = mk(t.x + fma(T.k41, n, T.k40), t.y);
    }
    {   // c3 = k30 + k31 zeta + zb (k32 + k33 zb) + n (k34 + k35 zb)
        cd in = mk(fma(T.k33, zb.x, T.k32), T.k33 * zb.y);
        cd t = cmul(zb, in);
        t = rfma(T.k31, zeta, t);
        t = rfma(n, mk(fma(T.k35, zb.x, T.k34), T.k35 * zb.y), t);
        c[3] = mk(t.x + T.k30, t.y);
    }
    {   // c2 = k20 + k21 zb + k22 zeta + n (k23 + k24 zb)
        cd t = scale(T.k21, zb);
        t = rfma(T.k22, zeta, t);
        t = rfma(n, mk(fma(T.k24, zb.x, T.k23), T.k24 * zb.y), t);
        c[2] = mk(t.x + T.k20, t.y);
    }
    c[1] = mk(fma(T.k10, zeta.x, fma(T.k11, n, T.k12)), T.k10 * zeta.y);
    c[0] = scale(T.k00, zeta);
}

// c0, c4, c5 only (same arithmetic as lens_polynomial): what the tracking path needs for the sum and the product
// of the five roots
MLM_HD void lens_polynomial_c045(const LensTable& T, cd zeta, cd& c0, cd& c4, cd& c5) {
    const cd zb = conj(zeta);
    const double n = norm2(zeta);
    const cd szb = mk(T.s + zb.x, zb.y);
    c5 = scale(T.m2, cmul(zb, szb));
    const cd in = mk(fma(T.k43, zb.x, T.k42 - n), T.k43 * zb.y);
    const cd t = scale(T.m2, cmul(zb, in));
    c4 = mk(t.x + fma(T.k41, n, T.k40), t.y);
    c0 = scale(T.k00, zeta);
}

// ---------------------------------------------------------------------------------------
// (b) root finder
// ---------------------------------------------------------------------------------------
#ifndef MLM_COUNT
#define MLM_COUNT(name)                              /* test-only instrumentation hook */
#endif
// explicit reconvergence points of loops with lane-dependent trip counts (device only; no-ops in the host build)
#if defined(__CUDA_ARCH__)
#define MLM_ACTIVEMASK() __activemask()
#define MLM_SYNCWARP(m) __syncwarp(m)
#else
#define MLM_ACTIVEMASK() 0u
#define MLM_SYNCWARP(m) ((void)(m))
#endif
#define MLM_EPS 1.1102230246251565e-16               /* 2^-53 */
#define MLM_LAG_MAXIT 64
// Newton on the quintic stops once a step is below 1e-6 d, d = distance of the root to the nearer
// lens: by quadratic convergence the root is then good to ~K (1e-6 d)^2 with K = |p''/2p'| ~ 1/d
// (the roots cluster around the lenses on that scale; relative to |z| alone the rule fails next to
// a low-mass secondary, where d ~ sqrt(q) and the residual test amplifies root errors by q/d^2).
// That is enough for the 1e-6 residual test; the physical images are subsequently polished on
// the lens equation itself to rounding level.
#define MLM_NEWTON_TOL2 1e-12
#ifndef MLM_FLAG_NOCONV          /* same values as include/mlmultipoles.h */
#define MLM_FLAG_NOCONV 1        /* an iteration cap was hit */
#define MLM_FLAG_EXTRA 2         /* nimg counts image(s) that the plain residual test (multipoles.py:183) rejects */
#define MLM_FLAG_EXTRA_SHIFT 6   /* ... how many: (flags >> 6) & 3, so that nimg - that = count of the plain test */
#define MLM_FLAG_NEARCRIT 4      /* min_i |1-|W2_i|^2| < 1e-3 over the physical images */
#define MLM_FLAG_AMBIG 8         /* a root residual within a decade of the 1e-6 threshold */
#define MLM_FLAG_SERIES 16       /* |A4-A2| > |A2-A0| : multipole series not converging */
#define MLM_FLAG_DEGEN 32        /* degenerate polynomial (leading coefficient exactly 0) */
#endif

MLM_HD double mag1(cd a) { return fabs(a.x) + fabs(a.y); }   // |a| <= mag1 <= sqrt2 |a|

// Stopping rule of the Newton iterations on the quintic (see MLM_NEWTON_TOL2): the step is below
// 1e-6 of the distance to the nearer lens, or the iteration has reached the
// rounding noise of the polynomial evaluation (a step that is already below 1e-7 |z| and no longer
// shrinks: clustered roots next to a secondary of q <~ 1e-5 carry ~1e-9 of noise).
MLM_HD bool newton_stop(double dz2, double prev4, cd z, double s) {   // prev4 = 0.25 * previous dz2
    const double xs = z.x + s;
    const double yy = z.y * z.y;
    const double z2 = fma(z.x, z.x, yy);
    const double d2 = fmin(z2, fma(xs, xs, yy));
    // 1e-28 z2: a step of ~50 ulp of |z| is the end of the road whatever d is (a root that sits 5e-8
    // from the secondary when the source is almost behind it)
    return (dz2 <= fmax(MLM_NEWTON_TOL2 * d2, 1e-28 * z2)) || (dz2 >= prev4 && dz2 <= 1e-14 * z2);
}

// p, p' at x for the full quintic
MLM_HD void horner2(const cd c[6], cd x, cd& p, cd& dp) {
    p = c[5];
    dp = mk(0.0, 0.0);
#pragma unroll
    for (int k = 4; k >= 0; --k) {
        dp = cfma(dp, x, p);
        p = cfma(p, x, c[k]);
    }
}

// All roots of c[5] z^5 + ... + c[0].  Returns the degree actually solved (5 unless leading
// coefficients are exactly zero, as np.roots does); slots >= degree are set to NaN.
// Laguerre iterations from x=0 on the successively deflated polynomial (kept zero-padded in
// d[0..5] so one code path serves every degree), closed form for the last two roots, then
// Newton polishing of every root on the undeflated polynomial.
MLM_HD int solve_roots(const cd c[6], double s, cd z[5], unsigned& flags) {
    cd d[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) d[k] = c[k];
    int deg = 5;
    if (d[5].x == 0.0 && d[5].y == 0.0) {
        flags |= MLM_FLAG_DEGEN;
        deg = 4;
        if (d[4].x == 0.0 && d[4].y == 0.0) {
            deg = 3;
            if (d[3].x == 0.0 && d[3].y == 0.0) deg = 2;   // cannot happen for s>0; keeps the loop bounded
        }
    }
    const int deg0 = deg;
    const double qnan = nan("");
#pragma unroll
    for (int k = 0; k < 5; ++k) z[k] = mk(qnan, qnan);

    cd x = mk(0.0, 0.0);
    int it = 0;
    while (deg > 2) {
        // p (b), p' (dd), p''/2 (f) of the current deflated polynomial at x
        cd b = d[5], dd = mk(0.0, 0.0), f = mk(0.0, 0.0);
        const double ax = mag1(x);
        double E = mag1(d[5]);
#pragma unroll
        for (int k = 4; k >= 0; --k) {
            f = cfma(f, x, dd);
            dd = cfma(dd, x, b);
            b = cfma(b, x, d[k]);
            E = fma(E, ax, mag1(d[k]));
        }
        MLM_COUNT(laguerre_evals);
        const double b2 = norm2(b);
        const double tol = 16.0 * MLM_EPS * E;
        bool done = (b2 <= tol * tol);
        if (!done) {
            const double m = (double)deg;
            const cd ib = crcp(b);
            const cd g = cmul(dd, ib);
            const cd g2 = csq(g);
            const cd fb = cmul(f, ib);
            const cd h = mk(fma(-2.0, fb.x, g2.x), fma(-2.0, fb.y, g2.y));
            // sq = sqrt((m-1) (m h - g2))
            const cd rad = scale(m - 1.0, mk(fma(m, h.x, -g2.x), fma(m, h.y, -g2.y)));
            const cd sq = csqrt_(rad);
            cd gp = g + sq;
            const cd gm = g - sq;
            if (norm2(gp) < norm2(gm)) gp = gm;
            const double gp2 = norm2(gp);
            cd dx;
            if (gp2 > 0.0) {
                const double r = m * rcp(gp2);
                dx = mk(gp.x * r, -(gp.y * r));
            } else {
                // p' = p'' = 0: leave the stationary point with a kick of size 1+|x| whose
                // direction changes from one iteration to the next (unit vectors 3-4-5, 5-12-13)
                const double rr = 1.0 + ax;
                const double ux = (it & 1) ? 0.6 : -0.38461538461538464;
                const double uy = (it & 1) ? 0.8 : 0.9230769230769231;
                dx = (it & 2) ? mk(-rr * uy, rr * ux) : mk(rr * ux, rr * uy);
            }
            ++it;
            // break rare limit cycles with a fractional step every 10th iteration
            if (it % 10 == 0) dx = scale(0.37 + 0.061 * (double)(it / 10), dx);
            x = x - dx;
            // cubic convergence: once the step is below 1e-7 |x| the new x is at rounding level
            const double x2 = norm2(x);
            done = (norm2(dx) <= 1e-14 * x2) || (it >= MLM_LAG_MAXIT);
            if (it >= MLM_LAG_MAXIT) flags |= MLM_FLAG_NOCONV;
        }
        if (done) {
#pragma unroll
            for (int k = 2; k < 5; ++k)
                if (k == deg - 1) z[k] = x;
            // forward deflation by (z - x), zero-padded: d <- quotient
            cd q4 = d[5];
            cd q3 = cfma(q4, x, d[4]);
            cd q2 = cfma(q3, x, d[3]);
            cd q1 = cfma(q2, x, d[2]);
            cd q0 = cfma(q1, x, d[1]);
            d[5] = mk(0.0, 0.0); d[4] = q4; d[3] = q3; d[2] = q2; d[1] = q1; d[0] = q0;
            --deg;
            x = mk(0.0, 0.0);
            it = 0;
        }
    }
    // last two roots: d2 z^2 + d1 z + d0 (stable quadratic formula)
    {
        const cd disc = cfma(scale(-4.0, d[2]), d[0], csq(d[1]));
        cd sq = csqrt_(disc);
        if (re_mulc(d[1], sq) < 0.0) sq = -sq;       // Re(conj(d1) sq) >= 0
        const cd qq = scale(-0.5, d[1] + sq);
        z[1] = cmul(qq, crcp(d[2]));
        z[0] = cmul(d[0], crcp(qq));
    }
    // Newton polish on the full polynomial
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        if (k < deg0) {
            cd zk = z[k];
            double prev = 1e300;
#pragma unroll 1
            for (int j = 0; j < 8; ++j) {
                cd p, dp;
                horner2(c, zk, p, dp);
                MLM_COUNT(newton_evals);
                const cd dz = cmul(p, crcp(dp));
                const double dz2 = norm2(dz);
                if (!(dz2 < 1e300)) break;            // p'=0 or overflow: keep the deflated root
                zk = zk - dz;
                if (newton_stop(dz2, prev, zk, s)) break;
                prev = 0.25 * dz2;
                if (j == 7) flags |= MLM_FLAG_NOCONV;
            }
            z[k] = zk;
        }
    }
    return deg0;
}

// One Newton step on the quintic (used for the non-physical roots of the tracking path).
MLM_HD cd newton_step(const cd c[6], cd zk, double& dz2) {
    cd p, dp;
    horner2(c, zk, p, dp);
    MLM_COUNT(newton_evals);
    // dz = p / p' (a NaN from p' = 0 fails every comparison of newton_stop: no convergence -> cold solve)
    const double idp2 = rcp_step(norm2(dp));
    const cd dz = scale(idp2, cmulc(dp, p));
    dz2 = norm2(dz);
    return zk - dz;
}

// ---------------------------------------------------------------------------------------
// (d)+(e) one image: W_2..W_6 from r1 = 1/z, r2 = 1/(z+s); mu0, mu2, mu4
// ---------------------------------------------------------------------------------------
struct ImageTerms {
    double mu0, mu2, mu4;
};

// ORDER = 2: mu0, mu2 (quadrupole, needs W2..W4);  ORDER = 4: also mu4 (needs W2..W6)
//
// Closed form ("operator form").  With w = W2 and the image-plane operators
//   D = d/dz + w d/dzbar,  Dbar = wbar d/dz + d/dzbar   (so that d/dzeta = mu0 D, d/dzetabar = mu0 Dbar)
// acting on functions of (w, W3, W4, ... and conjugates; D W_k = W_{k+1}, D conj(W_k) = w conj(W_{k+1})),
// the source-plane Laplacian L = d/dzeta d/dzetabar of mu0 = 1/(1-|w|^2) is
//   mu2 = L mu0     = mu0^4 (3 mu0 |A|^2 + B),              A = Dbar |w|^2,  B = D A (real)
//   mu4 = 2 L^2 mu0 = 2 mu0^6 lap2,  lap2 = Re[6 mu0^2 conj(A) H + 5 mu0 conj(A) E + mu0 D H + D E]
// with C = Dbar A, E = Dbar B, H = 15 mu0 |A|^2 A + 7 A B + 3 conj(A) C.  These are the same numbers as
// the reference's mu2, mu4 (multipoles.py:142-143), which it reaches through the 20 Taylor factors
// a_kl; tests/ checks this function against the oracle's literal restatement of that recursion.
template <int ORDER>
MLM_HD ImageTerms image_terms_from_W(cd W2, cd W3, cd W4, cd W5, cd W6) {
    const cd w = W2, w1 = W3, w2 = W4;
    const double n = norm2(w);
    const double mu0 = rcp(1.0 - n);                 // signed, multipoles.py:91
    const double p = norm2(w1);
    const cd wb = conj(w);
    const cd wb2 = csq(wb);
    const cd ww1b = cmul(w, conj(w1));
    const cd A = cfma(wb2, w1, ww1b);                 // wbar^2 w1 + w conj(w1)
    const double B = fma(2.0, re_mul(wb2, w2), p * fma(2.0, n, 1.0));
    const double a2 = norm2(A);
    const double mu02 = mu0 * mu0;
    const double mu04 = mu02 * mu02;
    ImageTerms out;
    out.mu0 = mu0;
    out.mu2 = mu04 * fma(3.0 * mu0, a2, B);
    out.mu4 = 0.0;
    if (ORDER >= 4) {
        const cd w3 = W5, w4 = W6;
        const cd wb3 = cmul(wb2, wb);
        // C = Dbar A = 3 |w1|^2 wbar + wbar^3 w2 + w conj(w2)
        cd C = cmul(w, conj(w2));
        C = cfma(wb3, w2, C);
        C = rfma(3.0 * p, wb, C);
        const cd U = cmul(conj(w1), w2);
        const cd X = cmul(wb, U);                     // wbar conj(w1) w2
        const cd Y = cmul(w1, conj(w2));
        cd Z = cmul(wb3, w3);                         // wbar^3 w3 + w^2 conj(w3)
        Z = cfma(conj(wb2), conj(w3), Z);
        // S = 12 E + 3 D C = (45+33n) X + (15+57n) Y + 24 p A + 9 p w conj(w1) + 15 Z
        cd S = scale(15.0, Z);
        S = rfma(fma(33.0, n, 45.0), X, S);
        S = rfma(fma(57.0, n, 15.0), Y, S);
        S = rfma(24.0 * p, A, S);
        S = rfma(9.0 * p, ww1b, S);
        // Re(D E)
        const double T = re_mul(wb, cmul(conj(w1), w3));          // Re(wbar conj(w1) w3)
        const cd Dp = cfma(w, Y, U);                              // D |w1|^2
        const cd XY = rfma(2.0, Y, X);                            // X + 2 Y
        double reDE = 2.0 * re_mulc(XY, A);                       // Re(conj(A) (2 X + 4 Y))
        reDE = fma(fma(2.0, n, 3.0), re_mul(ww1b, U), reDE);
        reDE = fma(fma(fma(2.0, n, 7.0), n, 1.0), norm2(w2), reDE);
        reDE = fma(fma(9.0, n, 6.0), T, reDE);
        reDE = fma(2.0, re_mul(A, Dp) + fma(p, B, re_mul(wb3, w4)), reDE);
        // lap2 = mu0 { mu0 [ 105 mu0 a2^2 + 72 a2 B + 33 Re(conj(A)^2 C) ] + 7 B^2 + 3 |C|^2 + Re(conj(A) S) } + Re(D E)
        const cd Ab2 = csq(conj(A));
        double in1 = (105.0 * mu0) * (a2 * a2);
        in1 = fma(72.0 * a2, B, in1);
        in1 = fma(33.0, re_mul(Ab2, C), in1);
        double in2 = re_mulc(S, A);
        in2 = fma(3.0, norm2(C), in2);
        in2 = fma(7.0 * B, B, in2);
        in2 = fma(mu0, in1, in2);
        const double lap2 = fma(mu0, in2, reDE);
        out.mu4 = 2.0 * (mu04 * mu02) * lap2;
    }
    return out;
}

// Source-size / limb-darkening factors (multipoles.py:145-146):
//   A2 term: mu0 + h2 mu2,  A4 term: ... + h4 mu4
struct SourceFactors {
    double h2, h4;
};
MLM_HD SourceFactors source_factors(double rho, double Gamma) {
    SourceFactors f;
    const double r2 = rho * rho;
    f.h2 = 0.5 * (1.0 - 0.2 * Gamma) * r2;
    f.h4 = (1.0 / 24.0) * (1.0 - (11.0 / 35.0) * Gamma) * (r2 * r2);
    return f;
}

struct PointResult {
    double A0, A2, A4;
    unsigned nimg;
    unsigned flags;
};

// =======================================================================================
// Image-tracking path (k_map, k_models, k_points).
//
// Along a coherent sequence of source positions the PHYSICAL images are carried from one position to
// the next by Newton on the lens equation itself (well conditioned, and its converged iterate IS
// the polished image: no separate residual test, no polish), the non-physical roots by Newton on
// the quintic (they are watched only for the moment they turn physical).  The roots are kept
// sorted, physical images first, so that "k < nimg" is the image test.  A position whose tracking
// is rejected (an iteration that does not contract, or five results that are not five distinct
// roots: Vieta) is re-solved cold and re-classified with the reference's residual test.
// =======================================================================================

// One Newton iteration on f(z) = z - zeta - conj(W1(z)):  dz = -(f + conj(W2) conj(f)) / (1 - |W2|^2).
// z2, zs2 = squared distances of z to the two lenses (their minimum is the scale on which the images live).
MLM_HD cd lens_newton(const LensTable& T, cd zeta, cd z, double& dz2, double& z2, double& zs2) {
    const cd zs = mk(z.x + T.s, z.y);
    z2 = norm2(z);
    zs2 = norm2(zs);
    // one reciprocal for both: 1/|z|^2 = |z+s|^2 / (|z|^2 |z+s|^2)
    const double i12 = rcp(z2 * zs2);
    const double i1 = i12 * zs2, i2 = i12 * z2;
    const cd r1 = mk(z.x * i1, -(z.y * i1)), r2 = mk(zs.x * i2, -(zs.y * i2));
    const cd W1 = rfma(T.alpha[1], r1, scale(T.beta[1], r2));
    const cd f = z - zeta - conj(W1);
    const cd W2 = rfma(T.alpha[2], csq(r1), scale(T.beta[2], csq(r2)));
    const double iJ = rcp_step(1.0 - norm2(W2));
    const cd dz = scale(-iJ, cjfma(W2, conj(f), f));
    dz2 = norm2(dz);
    return z + dz;
}

// Classification of a fresh set of polynomial roots (cold solve, or a non-physical root that came
// close to the lens equation): the reference's residual test |z - conj(W1) - zeta| < 1e-6
// (multipoles.py:182-183; NaN -> rejected, like NumPy), the accepted images polished by Newton on the
// lens equation, roots with a residual in (1e-6, 1e-1) rescued iff that Newton, started with a step
// < 1e-7, converges (root noise of the ill-conditioned quintic at small q, amplified by |W2| up to
// 1e7).  Afterwards the store holds the physical images first (original order kept within each group).
// Returns the number of physical images.
MLM_HD int classify_roots(const LensTable& T, cd zeta, cd* sz, int stride, unsigned& flags, int& extra) {
    unsigned phys_mask = 0;
    extra = 0;                                        // images counted beyond the plain |residual| < 1e-6 test
    cd zk5[5];
#pragma unroll 1
    for (int k = 0; k < 5; ++k) {
        cd zk = sz[k * stride];
        cd zs = mk(zk.x + T.s, zk.y);
        cd r1 = crcp(zk), r2 = crcp(zs);
        cd W1 = rfma(T.alpha[1], r1, scale(T.beta[1], r2));
        cd f = zk - zeta - conj(W1);
        const double f2 = norm2(f);
        if (f2 > 1e-14 && f2 < 1e-10) flags |= MLM_FLAG_AMBIG;
        if (!(f2 < 1e-2)) continue;
        bool phys = (f2 < 1e-12);                     // the reference's test, multipoles.py:183
        bool conv = false;
        double lim = phys ? 1e-10 : 1e-14;
#pragma unroll 1
        for (int it = 0; it < 3; ++it) {
            MLM_COUNT(classify_newton);
            const cd W2 = rfma(T.alpha[2], csq(r1), scale(T.beta[2], csq(r2)));
            const double iJ = rcp_step(1.0 - norm2(W2));
            const cd dz = scale(-iJ, cjfma(W2, conj(f), f));
            const double dz2 = norm2(dz);
            // (8 ulp of |z|)^2: steps at the rounding level of z count as converged, not as stalled
            const double ulp2 = 3.2e-30 * norm2(zk);
            if (!(dz2 < fmax(lim, ulp2))) break;
            zk = zk + dz;
            zs = mk(zk.x + T.s, zk.y);
            r1 = crcp(zk);
            r2 = crcp(zs);
            if (dz2 <= fmax(1e-20 * fmin(norm2(zk), norm2(zs)), ulp2)) { conv = true; break; }
            lim = 1e-4 * dz2;
            W1 = rfma(T.alpha[1], r1, scale(T.beta[1], r2));
            f = zk - zeta - conj(W1);
        }
        if (!phys && conv) {
            // rescued: it must be an image of its own, not a second, noisier root next to an image
            // that is already counted (q <~ 1e-6: two roots 1e-9 apart next to the planet)
            phys = true;
            const double dl2 = fmin(norm2(zk), norm2(zs));
#pragma unroll
            for (int j = 0; j < 5; ++j)
                if (j < k && ((phys_mask >> j) & 1u) && norm2(sz[j * stride] - zk) <= 1e-12 * dl2) phys = false;
        }
        if (phys) {
            phys_mask |= 1u << k;
            sz[k * stride] = zk;                      // the polished image
            if (!(f2 < 1e-12)) ++extra;               // rescued: the reference's test alone would have dropped it
        }
    }
    // A binary lens has 3 or 5 images.  An EVEN count means that an image sitting almost on top of a lens
    // (distance d << 1: a very distant source, or a secondary of q <~ 1e-6) was lost to the rounding noise
    // of the quintic, which cannot resolve z + s to better than ~1e-16/d relative.  Such an image is found
    // directly: its position follows from the lens equation with the near lens dominating,
    //   delta = m_i / conj(p_i - zeta - m_j / conj(p_i - p_j)),   z = p_i + delta,
    // and Newton on the lens equation from there converges in two or three steps.  It replaces the
    // non-physical root next to it.  (The reference loses these images too; their magnification is < 1e-12.)
    {
        int nphys = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) nphys += (phys_mask >> k) & 1u;
        if ((nphys & 1) == 0 && nphys < 5) {
#pragma unroll 1
            for (int lens = 0; lens < 2; ++lens) {
                const cd p = mk(lens ? -T.s : 0.0, 0.0);
                const double mi = lens ? T.beta[1] : T.alpha[1];
                const double other = lens ? T.alpha[1] / T.s : -T.beta[1] / T.s;   // -m_j / conj(p_i - p_j), real
                const cd den = mk(p.x - zeta.x + other, zeta.y);                  // conj(p_i - zeta - m_j/conj(p_i - p_j))
                cd zn = p + scale(mi, crcp(den));
                const double del2 = norm2(zn - p);
                if (!(del2 < 1e-4)) continue;            // the image is not close to this lens: nothing was lost here
                bool conv = false;
                double lim = 1e300;
#pragma unroll 1
                for (int it = 0; it < 8; ++it) {
                    double dz2, z2, zs2;
                    const cd z1 = lens_newton(T, zeta, zn, dz2, z2, zs2);
                    if (!(dz2 < lim)) break;
                    zn = z1;
                    if (dz2 <= fmax(1e-24 * fmin(z2, zs2), 3.2e-30 * z2)) { conv = true; break; }
                    lim = 0.25 * dz2;
                }
                if (!conv) continue;
                // already among the images?  else it takes the place of the nearest non-physical root
                const double dl2 = fmin(norm2(zn), norm2(mk(zn.x + T.s, zn.y)));
                bool dup = false;
                int knear = -1;
                double best = 1e300;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const double dist2 = norm2(sz[k * stride] - zn);
                    if ((phys_mask >> k) & 1u) {
                        if (dist2 <= 1e-12 * dl2) dup = true;
                    } else if (!(dist2 >= best)) {       // NaN slots (degenerate polynomial) lose against any number... and win if nothing else
                        best = dist2; knear = k;
                    }
                }
                if (dup || knear < 0) continue;
                sz[knear * stride] = zn;
                phys_mask |= 1u << knear;
                ++extra;                                 // restored: no polynomial root passed the plain test for it
            }
        }
    }
    // stable partition: physical images first
#pragma unroll
    for (int k = 0; k < 5; ++k) zk5[k] = sz[k * stride];
    int w = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k)
        if ((phys_mask >> k) & 1u) sz[(w++) * stride] = zk5[k];
    const int nimg = w;
#pragma unroll
    for (int k = 0; k < 5; ++k)
        if (!((phys_mask >> k) & 1u)) sz[(w++) * stride] = zk5[k];
    return nimg;
}

// Tracking state of a thread along its sequence (registers).
struct TrackState {
    int hist;      // usable earlier positions: 0 cold start, 1 previous roots, 2 linear, 3 quadratic extrapolation
    int nimg;      // physical images = the first nimg entries of the root store
    int extra;     // of which the last classification accepted beyond the reference's plain residual test
                   // (Newton rescue, odd-count restoration); carried along while the images are tracked
};

// One step of a coherent sequence: bring the root store (sz current, szp previous, szpp -- optional --
// the one before; [k * stride], shared memory in the kernels) to the source position zeta.
// `ratio` = (this source step) / (previous one) for unevenly spaced sequences.
//
// Three images (the common state): only the images are carried (Newton on the lens equation); the other two roots
// of the quintic follow from the images by Vieta -- sum of the five roots = -c4/c5, product = -c0/c5 -- as the roots
// of a quadratic, afresh at every position, and are only looked at: a lens-equation residual below 0.1 means that
// they are about to turn physical (caustic crossing), then they are polished on the quintic and the position is
// re-classified.  Only c0, c4, c5 of the polynomial are needed for that; the full polynomial is formed on the rare
// paths (polish, cold solve) and in the other states (five images: Vieta check of the five tracks; 1, 2 or 4
// images -- degenerate positions: non-physical roots carried by Newton on the quintic).
// start value of root k from its history (mode 0: the previous root, 1: linear extrapolation scaled by `ratio`,
// 2: linear 2 z1 - z2, 3: quadratic 3 z1 - 3 z2 + z3), shifting the history of that root by one position
MLM_HD cd predict_root(cd* sz, cd* szp, cd* szpp, int mode, double ratio) {
    const cd zc = *sz;
    cd pred = zc;
    if (mode) {
        const cd zo = *szp;
        if (mode == 3) {
            const cd zoo = *szpp;
            pred = mk(fma(3.0, zc.x - zo.x, zoo.x), fma(3.0, zc.y - zo.y, zoo.y));
        } else if (mode == 2) {
            pred = mk(fma(2.0, zc.x, -zo.x), fma(2.0, zc.y, -zo.y));
        } else {
            pred = mk(fma(ratio, zc.x - zo.x, zc.x), fma(ratio, zc.y - zo.y, zc.y));
        }
        if (szpp != nullptr) *szpp = zo;
    }
    *szp = zc;
    return pred;
}

// |z - zeta - conj(W1(z))|^2: the quantity the reference compares with 1e-6 (multipoles.py:182-183)
MLM_HD double lens_residual2(const LensTable& T, cd zeta, cd z) {
    const cd zs = mk(z.x + T.s, z.y);
    const cd W1 = rfma(T.alpha[1], crcp(z), scale(T.beta[1], crcp(zs)));
    return norm2(z - zeta - conj(W1));
}

// WATCH of the two roots of the quintic that are not images (three-image state): is the lens-equation residual of
// one of them below ~0.1, i.e. are they about to turn physical?  They follow from the three images by Vieta (sum S,
// product P -> a quadratic).  The answer decides nothing but whether the position is looked at in double precision,
// and the threshold (1.5e-2 on the squared residual, against 1e-2 of the double-precision test behind it) is twelve
// orders of magnitude above the single-precision rounding of the residual, so this runs in FP32 -- on the pipe that
// these FP64-bound kernels leave idle.  Overflow, underflow or NaN anywhere answer "near" (conservative).
MLM_HD bool nonphysical_pair_near(const LensTable& T, cd zeta, cd S, cd c0, cd ic5, cd z0, cd z1, cd z2) {
    const cf Sf = tof(S), ztf = tof(zeta);
    const cf Pf = cmulf(cmulf(tof(c0), tof(ic5)), crcpf(cmulf(cmulf(tof(z0), tof(z1)), tof(z2))));   // = -P
    cf sq = csqrtf_(mkf(fmaf(4.0f, Pf.x, fmaf(Sf.x, Sf.x, -(Sf.y * Sf.y))), fmaf(4.0f, Pf.y, 2.0f * (Sf.x * Sf.y))));
    if (fmaf(Sf.x, sq.x, Sf.y * sq.y) < 0.0f) sq = mkf(-sq.x, -sq.y);
    const cf r4 = mkf(0.5f * (Sf.x + sq.x), 0.5f * (Sf.y + sq.y));
    const cf r5 = cmulf(mkf(-Pf.x, -Pf.y), crcpf(r4));
    float res[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const cf z = j ? r5 : r4;
        const cf i1 = crcpf(z), i2 = crcpf(mkf(z.x + T.s_f, z.y));
        // f = z - zeta - conj(W1)
        const float fx = z.x - ztf.x - fmaf(T.alpha1_f, i1.x, T.beta1_f * i2.x);
        const float fy = z.y - ztf.y + fmaf(T.alpha1_f, i1.y, T.beta1_f * i2.y);
        res[j] = fmaf(fx, fx, fy * fy);
    }
    return !(res[0] >= 1.5e-2f && res[1] >= 1.5e-2f && res[0] < 1e30f && res[1] < 1e30f);
}

// Newton on the quintic from zk until newton_stop (at most 6 steps); false: not converged
MLM_HD bool quintic_newton(const cd c[6], double s, cd& zk) {
    double prev = 1e300;
#pragma unroll 1
    for (int j = 0; j < 6; ++j) {
        double dz2;
        zk = newton_step(c, zk, dz2);
        if (newton_stop(dz2, prev, zk, s)) return true;
        prev = 0.25 * dz2;
    }
    return false;
}

MLM_HD void advance_images(const LensTable& T, cd zeta, cd* sz, cd* szp, cd* szpp, int stride,
                           TrackState& st, unsigned& flags, double ratio = 1.0) {
    bool need_cold = true, need_classify = false;
    cd c0, c4, c5;
    lens_polynomial_c045(T, zeta, c0, c4, c5);
    if (st.hist >= 1 && !(c5.x == 0.0 && c5.y == 0.0)) {
        //@phase track_image_newton       (markers: regions of the per-phase instruction table, tools/ncu_phase_table.py)
        const int nimg = st.nimg;
        const int mode = (st.hist < 2) ? 0 : ((ratio != 1.0) ? 1 : ((szpp != nullptr && st.hist >= 3) ? 3 : 2));
        bool ok = true;
        cd sum = mk(0.0, 0.0);
        double mag = 1.0;
        // the images: Newton on the lens equation from the extrapolated position.  A step below 1e-7 d leaves an
        // error of ~ (1e-7 d)^2 / d = 1e-14 d: done; every further step must contract by >= 10.  The results go
        // straight into the store: a rejected position is solved cold, which rewrites all of it.
        // (One flat loop over (image, iteration) was tried for the light-curve kernels, where the lanes of a warp need
        // different numbers of iterations: 15.2 -> 16.1 active lanes only -- the lanes next to the peak of a light
        // curve need more iterations for ALL their images -- and the warp no longer reconverged before the sums.)
#pragma unroll 1
        for (int k = 0; k < nimg; ++k) {
            cd zk = predict_root(sz + k * stride, szp + k * stride, szpp ? szpp + k * stride : nullptr, mode, ratio);
            bool conv = false;
            double lim = 1e300;
#pragma unroll 1
            for (int it = 0; it < 4; ++it) {
                double dz2, z2, zs2;
                const cd zn = lens_newton(T, zeta, zk, dz2, z2, zs2);
                MLM_COUNT(lens_evals);
                if (!(dz2 < lim)) break;
                zk = zn;
                // dz2 <= max(1e-14 min(z2, zs2), 3.2e-30 z2), written as comparisons (no min/max selects)
                if (dz2 <= 1e-14 * z2 && (dz2 <= 1e-14 * zs2 || dz2 <= 3.2e-30 * z2)) { conv = true; break; }
                lim = 1e-2 * dz2;
            }
            ok = ok && conv;
            sz[k * stride] = zk;
            sum = sum + zk;
            mag += norm2(zk);
        }
        if (nimg == 3) {
            //@phase track_quadratic
            MLM_COUNT(quadratics);
            const cd z0 = sz[0], z1 = sz[stride], z2 = sz[2 * stride];
            // three DISTINCT images?  (two tracks that collapsed onto one image differ by rounding only)
            const double tol = 1e-14 * mag;
            ok = ok && norm2(z0 - z1) > tol && norm2(z0 - z2) > tol && norm2(z1 - z2) > tol;
            // r4 + r5 = S, r4 r5 = P  ->  z^2 - S z + P = 0.  S in double (a difference of sums of roots: it cancels
            // for a distant source); everything after it is the WATCH, in single precision
            const cd ic5 = crcp(c5);
            const cd S = -(cfma(c4, ic5, sum));
            //@phase track_nonphys_check
            MLM_COUNT(nonphys_checks);
            MLM_COUNT(nonphys_checks);
            if (nonphysical_pair_near(T, zeta, S, c0, ic5, z0, z1, z2)) {
                //@phase track_polish
                MLM_COUNT(watch_near);
                // the same question in double, then: they may have turned physical (caustic crossing): classify them
                // like freshly solved roots, i.e. at the accuracy of a polished polynomial root (the quadratic loses
                // digits where the two merge)
                const cd P = -(cmul(cmul(c0, ic5), crcp(cmul(cmul(z0, z1), z2))));
                cd sq = csqrt_(rfma(-4.0, P, csq(S)));
                if (re_mulc(S, sq) < 0.0) sq = -sq;
                cd r4 = scale(0.5, S + sq);               // stable form: the larger root first
                cd r5 = cmul(P, crcp(r4));
                if (lens_residual2(T, zeta, r4) < 1e-2 || lens_residual2(T, zeta, r5) < 1e-2) {
                    need_classify = true;
                    cd c[6];
                    lens_polynomial(T, zeta, c);
                    ok = quintic_newton(c, T.s, r4) && ok;
                    ok = quintic_newton(c, T.s, r5) && ok;
                }
                sz[3 * stride] = r4;
                sz[4 * stride] = r5;
            }
        } else {
            //@phase track_general
            if (nimg < 5) {
                cd c[6];
                lens_polynomial(T, zeta, c);
#pragma unroll 1
                for (int k = nimg; k < 5; ++k) {
                    // non-physical root: Newton on the quintic, then a look at its lens-equation residual
                    cd zk = predict_root(sz + k * stride, szp + k * stride, szpp ? szpp + k * stride : nullptr, mode, ratio);
                    ok = quintic_newton(c, T.s, zk) && ok;
                    MLM_COUNT(nonphys_checks);
                    if (lens_residual2(T, zeta, zk) < 1e-2) need_classify = true;
                    sz[k * stride] = zk;
                    sum = sum + zk;
                    mag += norm2(zk);
                }
            }
            //@phase track_vieta
            // five DISTINCT roots?  (Vieta: sum z_k = -c4/c5; two tracks on one root break it by their separation)
            const cd v = cfma(c4, crcp(c5), sum);
            ok = ok && norm2(v) <= 1e-16 * mag;
        }
        //@phase track_accept
        if (ok) {
            need_cold = false;
            st.hist = (szpp != nullptr && st.hist >= 2) ? 3 : 2;
        } else {
            MLM_COUNT(track_fallbacks);
        }
    }
    //@phase cold_glue
    if (need_cold) {
        MLM_COUNT(cold_solves);
        cd c[6], z[5];
        lens_polynomial(T, zeta, c);
        const int deg = solve_roots(c, T.s, z, flags);
#pragma unroll
        for (int k = 0; k < 5; ++k) sz[k * stride] = z[k];
        need_classify = true;
        st.hist = (deg == 5) ? 1 : 0;
    }
    if (need_classify) {
        // images that were being TRACKED are at rounding level and pass the plain test whatever the polynomial
        // roots would have done: a re-classification without a cold solve keeps the earlier tally
        MLM_COUNT(classify_calls);
        const int carried = need_cold ? 0 : st.extra;
        st.nimg = classify_roots(T, zeta, sz, stride, flags, st.extra);
        if (st.extra < carried) st.extra = carried < st.nimg ? carried : st.nimg;
        if (st.hist > 1) st.hist = 1;                 // the order of the store changed: history is void
    }
    if (st.extra) flags |= MLM_FLAG_EXTRA | ((unsigned)(st.extra < 3 ? st.extra : 3) << MLM_FLAG_EXTRA_SHIFT);
}

// A0, A2, A4 of a position from its images (the first nimg entries of the root store):
// W_k (multipoles.py:197-201), mu0/mu2/mu4 per image, sums of |...| (multipoles.py:144-146).
template <int ORDER>
MLM_HD void sum_images(const LensTable& T, const SourceFactors sf, const cd* sz, int stride, int nimg,
                       PointResult& out) {
    double A0 = 0.0, A2 = 0.0, A4 = 0.0, minJ = 1e300;
    unsigned flags = out.flags;
#pragma unroll 1
    //@phase image_wk
    for (int k = 0; k < nimg; ++k) {
        MLM_COUNT(images);
        const cd zk = sz[k * stride];
        const cd zs = mk(zk.x + T.s, zk.y);
        // 1/z and 1/(z+s) from one reciprocal
        const double n1 = norm2(zk), n2 = norm2(zs);
        const double i12 = rcp(n1 * n2);
        const double i1 = i12 * n2, i2 = i12 * n1;
        const cd r1 = mk(zk.x * i1, -(zk.y * i1)), r2 = mk(zs.x * i2, -(zs.y * i2));
        const cd a2 = csq(r1), b2 = csq(r2);
        const cd a3 = cmul(a2, r1), b3 = cmul(b2, r2);
        const cd a4 = csq(a2), b4 = csq(b2);
        const cd W2 = rfma(T.alpha[2], a2, scale(T.beta[2], b2));
        const cd W3 = rfma(T.alpha[3], a3, scale(T.beta[3], b3));
        const cd W4 = rfma(T.alpha[4], a4, scale(T.beta[4], b4));
        cd W5 = mk(0.0, 0.0), W6 = mk(0.0, 0.0);
        if (ORDER >= 4) {
            const cd a5 = cmul(a4, r1), b5 = cmul(b4, r2);
            const cd a6 = csq(a3), b6 = csq(b3);
            W5 = rfma(T.alpha[5], a5, scale(T.beta[5], b5));
            W6 = rfma(T.alpha[6], a6, scale(T.beta[6], b6));
        }
        //@phase image_sums
        const ImageTerms t = image_terms_from_W<ORDER>(W2, W3, W4, W5, W6);
        const double v2 = fma(sf.h2, t.mu2, t.mu0);
        A0 += fabs(t.mu0);
        A2 += fabs(v2);
        if (ORDER >= 4) A4 += fabs(fma(sf.h4, t.mu4, v2));
        minJ = fmin(minJ, fabs(1.0 - norm2(W2)));
    }
    if (minJ < 1e-3) flags |= MLM_FLAG_NEARCRIT;
    if (ORDER >= 4 && fabs(A4 - A2) > fabs(A2 - A0)) flags |= MLM_FLAG_SERIES;
    out.A0 = A0; out.A2 = A2; out.A4 = A4; out.nimg = (unsigned)nimg; out.flags = flags;
}

// ---------------------------------------------------------------------------------------
// Cold start WITHOUT the root finder (k_points_direct: lists of unrelated positions).
//
// A binary lens always has (at least) three images, and they sit where three single-lens problems put them: next to
// either lens the minor image of that lens alone (the other lens acting as a constant deflection), and outside the
// major image of the total mass at the centre of mass.  Newton on the LENS EQUATION from these three guesses
// (~3 evaluations of 55 FP64 instructions each) gives three images at rounding level -- if it converges on three
// DISTINCT points they are solutions of the lens equation, hence roots of the quintic, whatever the guesses were --
// and the other two roots of the quintic follow from Vieta (a quadratic), polished on the quintic.  The five roots
// then take the usual classification (classify_roots).  ~1 000 FP64 instructions instead of ~3 400 for Laguerre +
// deflation + polish.  Where the guesses fail (inside and next to resonant caustics the three-point picture is wrong;
// the source on a lens; degenerate polynomial) the function answers false and the caller solves the position with
// the root finder: the result never depends on a guess having been good.
//
// `extra`: number of the three images whose own POLYNOMIAL root would fail the plain |residual| < 1e-6 test of
// multipoles.py:183 (the count MLM_FLAG_EXTRA reports): an image next to a lens (|W2| > 1e3) is moved by one Newton
// step on the quintic -- onto the root as the rounding noise of the polynomial defines it, which is what a root
// finder returns -- and tested there (MLM_FLAG_AMBIG likewise).  The flags are then those of the root-finder path
// in distribution, not position by position: they report rounding noise.
// ---------------------------------------------------------------------------------------
MLM_HD cd point_lens_image(cd u, double m, double sign) {        // u (1 + sign sqrt(1 + 4 m / |u|^2)) / 2
    const double n = norm2(u);
    const double t = 1.0 + 4.0 * m * rcp(n);
    const double f = 0.5 * (1.0 + sign * (t * rsqrt_(t)));
    return scale(f, u);
}

// Returns MLM_GUESS_FAILED (the root finder has to solve the position), MLM_GUESS_THREE (exactly the three images: the
// Vieta pair is nowhere near satisfying the lens equation -- no threshold, no rounding noise can change that) or
// MLM_GUESS_CLASSIFY: the pair is within 0.1 of the lens equation (or not a number) and the five roots must go through
// classify_roots.  With allow_classify the pair is polished on the quintic for that; without it the function returns
// MLM_GUESS_CLASSIFY at once so that the caller can set such positions aside (k_points_direct: next to caustics they
// are ~10 % of a light curve, i.e. some lane of almost every warp -- classifying them on the spot leaves 17 of 32
// lanes active over the whole kernel).
#define MLM_GUESS_FAILED 0
#define MLM_GUESS_THREE 1
#define MLM_GUESS_CLASSIFY 2
MLM_HD int guess_solve(const LensTable& T, cd zeta, cd* sz, int stride, int& extra, unsigned& flags, bool allow_classify) {
    extra = 0;
    int status = MLM_GUESS_THREE;
    // the lanes that enter together leave the Newton loops together: no early return before the loops, and an explicit
    // reconvergence point after each (without it the lanes drift apart over the rest of the function -- measured on a
    // shuffled light curve: 16.9 of 32 lanes active over the whole kernel instead of 25.5)
    const unsigned entry_mask = MLM_ACTIVEMASK();
    cd c0, c4, c5;
    lens_polynomial_c045(T, zeta, c0, c4, c5);
    const double m1 = T.alpha[1], m2 = T.beta[1], is = rcp(T.s);
    cd sum = mk(0.0, 0.0);
    double mag = 1.0, dmin2 = 1e300;
    bool ok = !(c5.x == 0.0 && c5.y == 0.0);            // degenerate polynomial: the root finder's business
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
        cd zk;
        if (k == 0) zk = point_lens_image(mk(zeta.x + m2 * is, zeta.y), m1, -1.0);
        else if (k == 1) zk = mk(-T.s, 0.0) + point_lens_image(mk(zeta.x - m1 * is + T.s, zeta.y), m2, -1.0);
        else zk = mk(-T.s * m2, 0.0) + point_lens_image(mk(zeta.x + T.s * m2, zeta.y), 1.0, 1.0);
        bool conv = false;
        double lim = 1e300, z2 = 0.0, zs2 = 0.0;
#pragma unroll 1
        for (int it = 0; it < 10; ++it) {
            double dz2;
            const cd zn = lens_newton(T, zeta, zk, dz2, z2, zs2);
            MLM_COUNT(lens_evals);
            if (!(dz2 < lim)) break;
            zk = zn;
            if (dz2 <= 1e-14 * z2 && (dz2 <= 1e-14 * zs2 || dz2 <= 3.2e-30 * z2)) { conv = true; break; }
            lim = 0.25 * dz2;                            // every step at most half the one before
        }
        ok = ok && conv;
        sz[k * stride] = zk;
        sum = sum + zk;
        mag += norm2(zk);
        dmin2 = fmin(dmin2, fmin(z2, zs2));              // distances of the last evaluation: the step after it is < 1e-7 of them
        MLM_SYNCWARP(entry_mask);
    }
    if (!ok) return MLM_GUESS_FAILED;
    const cd z0 = sz[0], z1 = sz[stride], z2 = sz[2 * stride];
    const double tol = 1e-12 * mag;
    if (!(norm2(z0 - z1) > tol && norm2(z0 - z2) > tol && norm2(z1 - z2) > tol)) return MLM_GUESS_FAILED;
    // the other two roots: r4 + r5 = S, r4 r5 = P
    const cd ic5 = crcp(c5);
    const cd S = -(cfma(c4, ic5, sum));
    const cd P = -(cmul(cmul(c0, ic5), crcp(cmul(cmul(z0, z1), z2))));
    cd sq = csqrt_(rfma(-4.0, P, csq(S)));
    if (re_mulc(S, sq) < 0.0) sq = -sq;
    cd r4 = scale(0.5, S + sq);
    cd r5 = cmul(P, crcp(r4));
    const bool near_lens = dmin2 < 1e-3;                 // an image with |W2| = m / d^2 > 1e3 m
    const bool pair_near = !(lens_residual2(T, zeta, r4) >= 1e-2 && lens_residual2(T, zeta, r5) >= 1e-2);
    if (pair_near && !allow_classify) return MLM_GUESS_CLASSIFY;
    if (pair_near || near_lens) {
        cd c[6];
        lens_polynomial(T, zeta, c);
        if (pair_near) {
            status = MLM_GUESS_CLASSIFY;
            if (!(quintic_newton(c, T.s, r4) && quintic_newton(c, T.s, r5))) return MLM_GUESS_FAILED;
            // five distinct roots?  (Vieta on the polished pair)
            const cd v = cfma(c4, ic5, sum + r4 + r5);
            if (!(norm2(v) <= 1e-16 * (mag + norm2(r4) + norm2(r5)))) return MLM_GUESS_FAILED;
        }
        // images next to a lens: would their polynomial root pass the plain test?
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
            const cd zk = sz[k * stride];
            const double d2 = fmin(norm2(zk), norm2(mk(zk.x + T.s, zk.y)));
            if (d2 < 1e-3) {
                double dz2;
                const cd zq = newton_step(c, zk, dz2);
                const double f2 = lens_residual2(T, zeta, zq);
                if (f2 > 1e-14 && f2 < 1e-10) flags |= MLM_FLAG_AMBIG;      // like classify_roots on a polynomial root
                if (!(f2 < 1e-12)) ++extra;
            }
        }
    }
    sz[3 * stride] = r4;
    sz[4 * stride] = r5;
    return status;
}

// whole path for one source position from the three-image guesses.  Returns MLM_GUESS_THREE or MLM_GUESS_CLASSIFY when
// `out` holds the result, MLM_GUESS_FAILED when the position is not solved this way, and -- without allow_classify --
// MLM_GUESS_CLASSIFY with nothing valid in `out` for a position that needs the classification (see guess_solve).
template <int ORDER>
MLM_HD int point_magnification_direct(const LensTable& T, const SourceFactors sf, cd zeta, cd* sz, int stride,
                                      PointResult& out, bool allow_classify = true) {
    int extra = 0, nimg = 3;
    out.flags = 0;
    const int status = guess_solve(T, zeta, sz, stride, extra, out.flags, allow_classify);
    if (status == MLM_GUESS_FAILED || (status == MLM_GUESS_CLASSIFY && !allow_classify)) return status;
    if (status == MLM_GUESS_CLASSIFY) {
        int extra_c = 0;
        MLM_COUNT(classify_calls);
        nimg = classify_roots(T, zeta, sz, stride, out.flags, extra_c);
        if (nimg < 3) return MLM_GUESS_FAILED;           // cannot happen with three converged images; be safe
        extra += extra_c;
    }
    if (extra) out.flags |= MLM_FLAG_EXTRA | ((unsigned)(extra < 3 ? extra : 3) << MLM_FLAG_EXTRA_SHIFT);
    sum_images<ORDER>(T, sf, sz, stride, nimg, out);
    return status;
}

// whole path for one source position (cold start)
template <int ORDER>
MLM_HD PointResult point_magnification(const LensTable& T, const SourceFactors sf, cd zeta) {
    cd z[5];
    PointResult out;
    out.flags = 0;
    TrackState st;
    st.hist = 0; st.nimg = 0; st.extra = 0;
    advance_images(T, zeta, z, z, nullptr, 1, st, out.flags);
    sum_images<ORDER>(T, sf, z, 1, st.nimg, out);
    return out;
}

// ---------------------------------------------------------------------------------------
// DEAL of the strips of a map over several launches / GPUs (k_map, mlm_map_deal): the strips are dealt in
// rounds of `period`; a launch owns the positions sel[0..nsel) of every round.  Its strips, in ascending order,
// carry the ownership index t = round * nsel + j  <->  strip k = round * period + sel[j].
// ---------------------------------------------------------------------------------------
#define MLM_DEAL_MAXSEL 32
struct Deal {
    int64_t period;
    int64_t t0;                  // ownership index of the first owned strip inside the row range of a launch
    int nsel;
    int sel[MLM_DEAL_MAXSEL];    // ascending positions inside a round
};

MLM_HD int64_t deal_strip(const Deal& d, int64_t t) { return (t / d.nsel) * d.period + d.sel[t % d.nsel]; }

// smallest ownership index whose strip is >= k
MLM_HDI int64_t deal_owned_from(const Deal& d, int64_t k) {
    const int64_t round = k / d.period, pos = k % d.period;
    int j = 0;
    while (j < d.nsel && d.sel[j] < pos) ++j;
    return round * d.nsel + j;                         // j == nsel: the first strip of the next round
}

// ---------------------------------------------------------------------------------------
// Literal (xi, eta) evaluation for user-supplied W_k tables: drop-in for
// quadrupole(Wk,...) / hexadecapole(Wk,...), multipoles.py:12-65 / 67-148.  Unlike
// image_terms_from_W nothing is assumed about where the W_k come from, and the result
// keeps the reference's definitions of every intermediate a_{kl}.
// The Q_{kl} sums are generated from set-partition counts (see tools/gen_qkl.py).
// ---------------------------------------------------------------------------------------
MLM_HD cd akl(double mu, cd W2, cd Q) {              // multipoles.py:150-154
    return scale(mu, cjfma(W2, Q, conj(Q)));
}

}  // namespace mlm
