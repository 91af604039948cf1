// TEST-ONLY host build of the per-position device functions (mlm_core.cuh).
// Lets the CPU test-suite check the kernel LOGIC (root finder, selection, Wirtinger
// recursion) against the oracle without a GPU.  Never loaded by microlensing_b200.
#ifdef MLM_STATS
static long g_cnt_laguerre_evals = 0, g_cnt_newton_evals = 0, g_cnt_track_fallbacks = 0, g_cnt_lens_evals = 0;
static long g_cnt_cold_solves = 0, g_cnt_classify_calls = 0, g_cnt_classify_newton = 0, g_cnt_nonphys_checks = 0, g_cnt_images = 0, g_cnt_quadratics = 0, g_cnt_anchor_positions = 0, g_cnt_watch_near = 0;
#define MLM_COUNT(name) (++g_cnt_##name)
#endif
#include "../../microlensing_b200/csrc/mlm_core.cuh"
#include "../../microlensing_b200/csrc/mlm_qkl_generated.cuh"

using namespace mlm;

extern "C" {
#ifdef MLM_STATS
void emu_stats(long* out, int reset) { out[0] = g_cnt_laguerre_evals; out[1] = g_cnt_newton_evals; out[2] = g_cnt_lens_evals; out[3] = g_cnt_track_fallbacks; if (reset) g_cnt_laguerre_evals = g_cnt_newton_evals = g_cnt_lens_evals = g_cnt_track_fallbacks = 0; }
// all counters: laguerre, newton, lens, fallbacks, cold_solves, classify_calls, classify_newton, nonphys_checks, images, quadratics
void emu_stats_all(long* out, int reset) {
    long* p[10] = {&g_cnt_laguerre_evals, &g_cnt_newton_evals, &g_cnt_lens_evals, &g_cnt_track_fallbacks, &g_cnt_cold_solves,
                   &g_cnt_classify_calls, &g_cnt_classify_newton, &g_cnt_nonphys_checks, &g_cnt_images, &g_cnt_quadratics};
    for (int i = 0; i < 10; ++i) { out[i] = *p[i]; if (reset) *p[i] = 0; }
}
#endif


void emu_points(double s, double q, double rho, double Gamma, const double* zeta, long n, int order,
                double* A0, double* A2, double* A4, unsigned char* nimg, unsigned char* flags) {
    LensTable T;
    make_table(s, q, &T);
    SourceFactors sf = source_factors(rho, Gamma);
    for (long i = 0; i < n; ++i) {
        cd z = mk(zeta[2 * i], zeta[2 * i + 1]);
        PointResult r = (order >= 4) ? point_magnification<4>(T, sf, z) : point_magnification<2>(T, sf, z);
        A0[i] = r.A0; A2[i] = r.A2; A4[i] = r.A4;
        nimg[i] = (unsigned char)r.nimg; flags[i] = (unsigned char)r.flags;
    }
}

// the same through the three-image guesses (k_points_direct), falling back to the root finder like the kernels do;
// used[i] = MLM_GUESS_THREE / MLM_GUESS_CLASSIFY where the guesses solved the position, 0 where the root finder did
void emu_points_direct(double s, double q, double rho, double Gamma, const double* zeta, long n, int order,
                       double* A0, double* A2, double* A4, unsigned char* nimg, unsigned char* flags, unsigned char* used) {
    LensTable T;
    make_table(s, q, &T);
    SourceFactors sf = source_factors(rho, Gamma);
    for (long i = 0; i < n; ++i) {
        cd z = mk(zeta[2 * i], zeta[2 * i + 1]);
        cd store[5];
        PointResult r;
        // k_points_direct sets positions that need the classification aside and k_points_list<.., 1> solves them with
        // allow_classify: the same arithmetic as allowing it at once
        const int st = (order >= 4) ? point_magnification_direct<4>(T, sf, z, store, 1, r) : point_magnification_direct<2>(T, sf, z, store, 1, r);
        const bool ok = st != MLM_GUESS_FAILED;
        used[i] = (unsigned char)st;
        if (!ok) r = (order >= 4) ? point_magnification<4>(T, sf, z) : point_magnification<2>(T, sf, z);
        A0[i] = r.A0; A2[i] = r.A2; A4[i] = r.A4;
        nimg[i] = (unsigned char)r.nimg; flags[i] = (unsigned char)r.flags;
    }
}

void emu_roots(double s, double q, const double* zeta, long n, double* roots, unsigned char* flags) {
    LensTable T;
    make_table(s, q, &T);
    for (long i = 0; i < n; ++i) {
        cd c[6], z[5];
        unsigned fl = 0;
        lens_polynomial(T, mk(zeta[2 * i], zeta[2 * i + 1]), c);
        solve_roots(c, T.s, z, fl);
        for (int k = 0; k < 5; ++k) { roots[10 * i + 2 * k] = z[k].x; roots[10 * i + 2 * k + 1] = z[k].y; }
        flags[i] = (unsigned char)fl;
    }
}

void emu_poly(double s, double q, const double* zeta, long n, double* coefs) {
    LensTable T;
    make_table(s, q, &T);
    for (long i = 0; i < n; ++i) {
        cd c[6];
        lens_polynomial(T, mk(zeta[2 * i], zeta[2 * i + 1]), c);
        for (int k = 0; k < 6; ++k) { coefs[12 * i + 2 * k] = c[k].x; coefs[12 * i + 2 * k + 1] = c[k].y; }
    }
}

// per-image terms from W2..W6 (interleaved re,im; 5 complex per image)
void emu_image_terms(const double* W, long n, int literal, double* mu) {
    for (long i = 0; i < n; ++i) {
        const double* w = W + 10 * i;
        cd W2 = mk(w[0], w[1]), W3 = mk(w[2], w[3]), W4 = mk(w[4], w[5]), W5 = mk(w[6], w[7]), W6 = mk(w[8], w[9]);
        ImageTerms t = literal ? image_terms_literal<4>(W2, W3, W4, W5, W6) : image_terms_from_W<4>(W2, W3, W4, W5, W6);
        mu[3 * i] = t.mu0; mu[3 * i + 1] = t.mu2; mu[3 * i + 2] = t.mu4;
    }
}
}

// column strip of a map with warm-started tracking, same control flow as k_map
static int g_quadratic = 0;   // 1: three-level history (k_models), 0: two-level (k_map)
extern "C" void emu_set_quadratic(int on) { g_quadratic = on; }

extern "C" void emu_strip(double s, double q, double rho, double Gamma, const double* zeta, long n, int strip,
                          double* A0, double* A2, double* A4, unsigned char* nimg, unsigned char* flags,
                          long* ncold) {
    LensTable T;
    make_table(s, q, &T);
    SourceFactors sf = source_factors(rho, Gamma);
    cd z[5], zp[5], zpp[5];
    TrackState st;
    st.hist = 0; st.nimg = 0; st.extra = 0;
    long cold = 0;
    for (long i = 0; i < n; ++i) {
        if (i % strip == 0) st.hist = 0;
        cd zeta_i = mk(zeta[2 * i], zeta[2 * i + 1]);
        PointResult r;
        r.flags = 0;
        const int before = st.hist;
        advance_images(T, zeta_i, z, zp, g_quadratic ? zpp : nullptr, 1, st, r.flags);
        if (before == 0 || st.hist < 2) ++cold;
        sum_images<4>(T, sf, z, 1, st.nimg, r);
        A0[i] = r.A0; A2[i] = r.A2; A4[i] = r.A4;
        nimg[i] = (unsigned char)r.nimg; flags[i] = (unsigned char)r.flags;
    }
    *ncold = cold;
}

// strips [k_first, k_last] owned under a deal, enumerated exactly like mlm_map_deal + k_map do
extern "C" long emu_deal_strips(long period, int nsel, const int* sel, long k_first, long k_last, long* out, long cap) {
    Deal d;
    d.period = period; d.nsel = nsel; d.t0 = 0;
    for (int j = 0; j < MLM_DEAL_MAXSEL; ++j) d.sel[j] = j < nsel ? sel[j] : 0;
    const long t_first = deal_owned_from(d, k_first), t_end = deal_owned_from(d, k_last + 1);
    d.t0 = t_first;
    long n = 0;
    for (long ii = 0; ii < t_end - t_first && n < cap; ++ii) out[n++] = deal_strip(d, d.t0 + ii);
    return t_end - t_first;
}
