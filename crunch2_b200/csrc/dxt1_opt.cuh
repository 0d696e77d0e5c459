// dxt1_opt.cuh -- one-warp DXT1 colour endpoint optimiser for 4x4 blocks (sm_100a).
//
// Replaces crnlib::dxt1_endpoint_optimizer (reference crnlib/crn_dxt1.cpp) as it is driven by
// dxt_image::set_block_pixels (crn_dxt_image.cpp:1436-1490) with endpoint caching disabled.
//
// Execution model.  The reference walks a long candidate list serially; a candidate replaces the
// running best when its error is strictly smaller (crn_dxt1.cpp:1585, :1773).  Here ONE LANE OWNS ONE
// CANDIDATE: a lane evaluates its (low, high) pair against all unique colours of the block (held in
// shared memory, read as broadcasts), and the serial outcome is then recovered exactly:
//   * "static" runs -- candidate lists that do not read the live best (LBG pairs :1289-1303, the
//     least-squares lattice :616-642, the probe sweep :895-903, solid-colour tables :93-153,
//     combinatorial recovery :1974-1992) -- reduce to the lexicographic minimum of
//     (error, sequence number), which is what strict '<' acceptance in sequence order yields;
//   * "live" runs -- the 26+6 lattice neighbours of one endpoint paired with the *current* other
//     endpoint (:908-1015) and the per-component search (:415-486) -- are evaluated speculatively,
//     the first improving lane is committed, and the run is replayed from the next sequence number
//     with the updated state.
// The reference's lower-bound gate (:1316-1322) and m_solutions_tried set (:1323-1330) only skip
// candidates that cannot pass strict '<' (see oracle/port/dxt1_port.c header), so neither is needed
// for parity; the scalar set-up phases (perceptual weights, PCA, LBG, least squares) are computed
// redundantly by every lane with the reference's operand types and evaluation order (the build uses
// -fmad=false so no multiply-add is contracted).
#pragma once
#include "warp_util.cuh"
#include "omatch_tables.h"

namespace crn {

// probe tables (crn_dxt1.cpp:43-53)
CRN_DEVICE_TABLE uint8_t g_uber_probe[15] = { 0, 1, 2, 3, 5, 7, 9, 10, 13, 15, 19, 27, 43, 59, 91 };
CRN_DEVICE_TABLE uint8_t g_better_probe[10] = { 0, 1, 2, 3, 5, 9, 15, 19, 27, 43 };
CRN_DEVICE_TABLE uint8_t g_normal_probe[5] = { 0, 1, 3, 5, 7 };
CRN_DEVICE_TABLE uint8_t g_fast_probe[4] = { 0, 1, 2, 3 };

struct Dxt1Params {
    int quality;               // crn_dxt_quality 0 superfast .. 4 uber (below better: evaluate_solution_fast, fewer probes / passes)
    int perceptual;
    int pixels_have_alpha;
    int use_alpha_blocks;
    int force_alpha_blocks;
    int grayscale_sampling;
    unsigned alpha_threshold;
    // 1: clusters with more than kDxt1ParallelSumMinColours unique colours form their float sums (perceptual averages, mean, covariance, the
    // 4-means of try_median4) lane-parallel instead of in the reference's member order.  The optimiser's result then depends on rounding
    // noise of those sums -- tolerance class; set by the cluster kernels unless the context is in exact mode (crn_gpu_set_vq_mode).  The 4x4
    // block kernels never set it.
    int parallel_sums;
};
constexpr int kDxt1ParallelSumMinColours = 64;

struct Dxt1Best {              // warp-uniform
    unsigned long long err;
    unsigned lo, hi;
    int alpha_block, alt_round, enforce, enforced_sel;
};

// Block state that travels between the phase kernels (global memory, 352 bytes per block).  The
// optimiser is split into five kernels (set-up / LBG / sweep passes / post passes / finish) because one
// fused kernel is ~11 K SASS instructions: with every warp of an SM in a different phase the
// instruction caches thrash (ncu: 59 % of stall samples "no_instructions", profiles/r1c).  Each phase
// kernel's hot code fits the 32 KB L1.5 instruction cache; the state round trip costs 0.7 KB of HBM
// traffic per block and phase boundary, i.e. microseconds.
struct Dxt1BlockState {
    int4 cw[16];               // unique colour i: r, g, b, weight (first-appearance order)
    Dxt1Best best;             // running best solution (warp-uniform; written by lane 0)
    float mean[3], axis[3], low[3], high[3];   // m_mean_norm_color, m_principle_axis, projected endpoints
    int U, total_w, pixels_have_alpha, stage;  // stage: 0 optimise, 1 solved during set-up, 2 fully transparent
};

struct Dxt1Scratch : Dxt1BlockState {          // per-warp shared memory
    int4 ce[16];               // unique colour i in evaluation form: 2wr*r | 2wg*g << 16, 2wb*b, C2, weight (see eval_colour)
    uint16_t probe[2][32];     // sweep candidates for the low / high endpoint
    uint16_t packed[64];       // combinatorial-recovery endpoint list
    uint8_t sel[16];           // selectors of the current best per unique colour
};
constexpr int kDxt1StateVec4 = (int)(sizeof(Dxt1BlockState) / 16);

struct Dxt1Cfg {               // warp-uniform evaluation mode
    int U;
    bool do4, do3;             // which block types evaluate_solution_* considers (:1377-1387)
    int wr, wg, wb;            // channel weights of color_distance (crn_color.h:720-745)
    bool gray;
    bool hc;                   // m_evaluate_hc (:2085)
    bool fast;                 // quality < better: evaluate_solution_fast (:1594-1757) instead of _uber
    bool perc;                 // m_perceptual (perceptual && !grayscale_sampling): scales the fast evaluator's axis by 8 / 24
    bool par;                  // Dxt1Params::parallel_sums applies to this cluster
};

__device__ __forceinline__ void unpack565(unsigned c, bool scaled, int& r, int& g, int& b)
{   // crn_dxt.cpp:167-182
    b = c & 31; g = (c >> 5) & 63; r = (c >> 11) & 31;
    if (scaled) { b = (b << 3) | (b >> 2); g = (g << 2) | (g >> 4); r = (r << 3) | (r >> 2); }
}
__device__ __forceinline__ unsigned pack565_unscaled(int r, int g, int b)
{   // crn_dxt.cpp:142-161 with scaled == false (inputs already in range)
    return (unsigned)(min(b, 31) | (min(g, 63) << 5) | (min(r, 31) << 11));
}
__device__ __forceinline__ unsigned pack565_scaled(int r, int g, int b)
{
    unsigned rr = ((unsigned)r * 31u + 127u) / 255u, gg = ((unsigned)g * 63u + 127u) / 255u, bb = ((unsigned)b * 31u + 127u) / 255u;
    return min(bb, 31u) | (min(gg, 63u) << 5) | (min(rr, 31u) << 11);
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// static_cast<int>(double) as x86-64 cvttsd2si does it (out of range -> INT_MIN); CUDA would saturate.
__device__ __forceinline__ int d2i_x86(double x) { return (x > -2147483649.0 && x < 2147483648.0) ? (int)x : (int)0x80000000; }

__device__ __forceinline__ unsigned dxt1_dist(const Dxt1Cfg cfg, int r, int g, int b, int pr, int pg, int pb)
{
    if (cfg.gray) {   // crn_dxt1.cpp:1348-1363 with color::RGB_to_Y (crn_color.h:788-796)
        int y0 = (r * 19595 + g * 38470 + b * 7471 + 32768) >> 16;
        int y1 = (pr * 19595 + pg * 38470 + pb * 7471 + 32768) >> 16;
        int yd = y0 - y1;
        return (unsigned)(yd * yd);
    }
    int dr = r - pr, dg = g - pg, db = b - pb;
    return (unsigned)(cfg.wr * dr * dr + cfg.wg * dg * dg + cfg.wb * db * db);
}

// Evaluation form of the colour distance.  color_distance (crn_color.h:720-745) is
//   d(c,p) = wr*(cr-pr)^2 + wg*(cg-pg)^2 + wb*(cb-pb)^2 = C2(c) + P2(p) - (2wr*cr*pr + 2wg*cg*pg + 2wb*cb*pb)
// (grayscale sampling, crn_dxt1.cpp:1348-1363: d = (Y(c)-Y(p))^2, the same shape in one dimension).
// C2 does not depend on the palette entry, so min_k d(c,p_k) = C2 + min_k (P2_k - dot_k): three integer
// multiply-adds per palette entry instead of eight, exactly the same integers.
__device__ __forceinline__ int rgb_to_y(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 32768) >> 16; }
__device__ __forceinline__ int4 eval_colour(const Dxt1Cfg cfg, int r, int g, int b, int weight)
{   // (2wr*r | 2wg*g << 16, 2wb*b, C2, weight): everything the evaluator needs about one unique colour in ONE 16-byte load
    // (2wg*g <= 12750, so the two halves of .x never touch)
    if (cfg.gray) { const int y = rgb_to_y(r, g, b); return make_int4(2 * y, 0, y * y, weight); }
    return make_int4((2 * cfg.wr * r) | ((2 * cfg.wg * g) << 16), 2 * cfg.wb * b, cfg.wr * r * r + cfg.wg * g * g + cfg.wb * b * b, weight);
}
__device__ __forceinline__ int4 eval_palette(const Dxt1Cfg cfg, int r, int g, int b)
{   // (pr, pg, pb, P2)
    if (cfg.gray) { const int y = rgb_to_y(r, g, b); return make_int4(y, 0, 0, y * y); }
    return make_int4(r, g, b, cfg.wr * r * r + cfg.wg * g * g + cfg.wb * b * b);
}
__device__ __forceinline__ int eval_dprime(const int cx, const int cy, const int cz, const int4 p) { return p.w - cx * p.x - cy * p.y - cz * p.z; }

// `bound` = error of the current best: a candidate whose partial sum has reached it can only be rejected (every
// acceptance test is a strict '<' against the best), so the lane stops there -- the reference's own early out
// (crn_dxt1.cpp:1436, :1475, :1512, :1544 against the trial solution, :1773, :1812 against the best).  Lanes leave independently; the warp moves on when the last one is done.
template <bool DO4, bool DO3, typename SC>
__device__ __forceinline__ void dxt1_eval_loop(const SC* sc, int U, const int4 p0, const int4 p1, const int4 p2, const int4 p3,
                                               const int4 pm, unsigned long long bound, unsigned long long& e4, unsigned long long& e3)
{
    e4 = 0; e3 = 0;
    for (int i = 0; i < U;) {
        const int stop = min(U, i + 8);
#pragma unroll 2
        for (; i < stop; i++) {
            const int4 c = sc->ce[i];
            const unsigned w = (unsigned)c.w;
            const int cx = c.x & 0xffff, cy = c.x >> 16, cz = c.y;
            const int d01 = min(eval_dprime(cx, cy, cz, p0), eval_dprime(cx, cy, cz, p1));
            if (DO4) {
                const int d = min(d01, min(eval_dprime(cx, cy, cz, p2), eval_dprime(cx, cy, cz, p3)));
                e4 += (unsigned long long)(unsigned)(d + c.z) * w;
            }
            if (DO3) {
                const int d = min(d01, eval_dprime(cx, cy, cz, pm));
                e3 += (unsigned long long)(unsigned)(d + c.z) * w;
            }
        }
        if ((DO4 && DO3) ? (e4 >= bound && e3 >= bound) : (DO4 ? e4 >= bound : e3 >= bound)) break;
    }
}

// evaluate_solution_fast (crn_dxt1.cpp:1594-1757), lane-private: the selector of a colour comes from the position of its projection on the
// endpoint axis between the palette entries' projections (NOT from the nearest entry), the error from color_distance to that entry.
// which entry a colour takes: 4-colour block 0 / 2 / 3 / 1 along the axis, 3-colour block 0 / 2 / 1
struct Dxt1FastAxis { int dirr, dirg, dirb, c0Point, halfPoint, c3Point, c02Point, c21Point; int c[4][3], m[3]; };
__device__ __forceinline__ Dxt1FastAxis dxt1_fast_axis(const Dxt1Cfg cfg, unsigned lo, unsigned hi, int alt)
{
    Dxt1FastAxis A;
    unpack565(lo, true, A.c[0][0], A.c[0][1], A.c[0][2]);
    unpack565(hi, true, A.c[1][0], A.c[1][1], A.c[1][2]);
    int vr = A.c[1][0] - A.c[0][0], vg = A.c[1][1] - A.c[0][1], vb = A.c[1][2] - A.c[0][2];
    if (cfg.perc) { vr *= 8; vg *= 24; }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        A.c[2][k] = (A.c[0][k] * 2 + A.c[1][k] + alt) / 3; A.c[3][k] = (A.c[1][k] * 2 + A.c[0][k] + alt) / 3;
        A.m[k] = (A.c[0][k] + A.c[1][k] + alt) >> 1;
    }
    const int s0 = A.c[0][0] * vr + A.c[0][1] * vg + A.c[0][2] * vb, s1 = A.c[1][0] * vr + A.c[1][1] * vg + A.c[1][2] * vb;
    const int s2 = A.c[2][0] * vr + A.c[2][1] * vg + A.c[2][2] * vb, s3 = A.c[3][0] * vr + A.c[3][1] * vg + A.c[3][2] * vb;
    const int sm = A.m[0] * vr + A.m[1] * vg + A.m[2] * vb;
    A.dirr = vr * 2; A.dirg = vg * 2; A.dirb = vb * 2;
    A.c0Point = s1 + s3; A.halfPoint = s3 + s2; A.c3Point = s2 + s0;
    A.c02Point = s0 + sm; A.c21Point = sm + s1;
    return A;
}
__device__ __forceinline__ unsigned dxt1_fast_sel4(const Dxt1FastAxis& A, int r, int g, int b)
{
    const int dot = r * A.dirr + g * A.dirg + b * A.dirb;
    return dot >= A.halfPoint ? (dot < A.c0Point ? 3u : 1u) : (dot < A.c3Point ? 0u : 2u);
}
__device__ __forceinline__ unsigned dxt1_fast_sel3(const Dxt1FastAxis& A, int r, int g, int b)
{
    const int dot = r * A.dirr + g * A.dirg + b * A.dirb;
    return dot < A.c02Point ? 0u : (dot < A.c21Point ? 2u : 1u);
}
template <typename SC>
__device__ __noinline__ void dxt1_eval_fast(SC* sc, const Dxt1Cfg cfg, unsigned lo, unsigned hi, int alt, unsigned long long& err, int& alpha)
{
    const Dxt1FastAxis A = dxt1_fast_axis(cfg, lo, hi, alt);
    unsigned long long best = sc->best.err;              // m_trial_solution.m_error starts at the best's (:1598)
    alpha = 0;
    if (cfg.do4) {
        unsigned long long te = 0;
        for (int i = cfg.U - 1; i >= 0; i--) {
            const int4 c = sc->cw[i];
            const unsigned bi = dxt1_fast_sel4(A, c.x, c.y, c.z);
            te += (unsigned long long)dxt1_dist(cfg, c.x, c.y, c.z, A.c[bi][0], A.c[bi][1], A.c[bi][2]) * (unsigned)c.w;
            if (te >= best) break;
        }
        if (te < best) { best = te; alpha = 0; }
    }
    if (cfg.do3) {
        unsigned long long te = 0;
        for (int i = cfg.U - 1; i >= 0; i--) {
            const int4 c = sc->cw[i];
            const unsigned bi = dxt1_fast_sel3(A, c.x, c.y, c.z);
            const int* p = bi == 2 ? A.m : A.c[bi];
            te += (unsigned long long)dxt1_dist(cfg, c.x, c.y, c.z, p[0], p[1], p[2]) * (unsigned)c.w;
            if (te >= best) break;
        }
        if (te < best) { best = te; alpha = 1; }
    }
    err = best;                                          // == the running best's error when neither block type improved: never accepted
}

template <typename SC> __device__ __forceinline__ void dxt1_count_eval(SC*, int) {}      // work counters: only the cluster scratch has them
// CTA-cooperative evaluation (cluster_kernels.cuh: several warps split the colours of one large cluster): only the cluster scratch can ask for it
template <typename SC> __device__ __forceinline__ bool dxt1_is_coop(const SC*) { return false; }
template <typename SC> __device__ __forceinline__ void dxt1_eval_coop(SC*, const Dxt1Cfg, unsigned, unsigned, int, unsigned long long&, int&, bool) {}
template <typename SC> __device__ __forceinline__ long long dxt1_prof_begin(SC*) { return 0; }                  // profiling build of the cluster kernels only
template <typename SC> __device__ __forceinline__ void dxt1_prof_end(SC*, long long) {}

// Lane-private evaluation of one candidate: evaluate_solution_uber / _hc_* without the bookkeeping
// (crn_dxt1.cpp:1370-1561, :1759-1835).  err = min over allowed block types, alpha = 3-colour won.
template <typename SC>
__device__ __noinline__ void dxt1_eval(SC* sc, const Dxt1Cfg cfg, unsigned lo, unsigned hi, int alt,
                                       unsigned long long& err, int& alpha, bool valid = true)
{
    // `valid` = this lane holds a candidate.  Callers pass it instead of branching around the call: with a CTA per cluster the evaluation
    // contains block-wide barriers, so every lane of every warp has to come through here the same number of times.
    if (!cfg.fast && dxt1_is_coop(sc)) { dxt1_eval_coop(sc, cfg, lo, hi, alt, err, alpha, valid); return; }     // (cluster scratch: always; see cluster_kernels.cuh)
    if (!valid) { err = ~0ull; alpha = 0; return; }
    dxt1_count_eval(sc, cfg.U);
    if (cfg.fast) { dxt1_eval_fast(sc, cfg, lo, hi, alt, err, alpha); return; }
    const long long prof_t0 = dxt1_prof_begin(sc);
    int r0, g0, b0, r1, g1, b1;
    unpack565(lo, true, r0, g0, b0);
    unpack565(hi, true, r1, g1, b1);
    const int4 p0 = eval_palette(cfg, r0, g0, b0), p1 = eval_palette(cfg, r1, g1, b1);
    const int4 p2 = eval_palette(cfg, (r0 * 2 + r1 + alt) / 3, (g0 * 2 + g1 + alt) / 3, (b0 * 2 + b1 + alt) / 3);
    const int4 p3 = eval_palette(cfg, (r1 * 2 + r0 + alt) / 3, (g1 * 2 + g0 + alt) / 3, (b1 * 2 + b0 + alt) / 3);
    const int4 pm = eval_palette(cfg, (r0 + r1 + alt) >> 1, (g0 + g1 + alt) >> 1, (b0 + b1 + alt) >> 1);
    unsigned long long e4, e3;
    const unsigned long long bound = sc->best.err;
    if (cfg.do4 && cfg.do3) {
        dxt1_eval_loop<true, true>(sc, cfg.U, p0, p1, p2, p3, pm, bound, e4, e3);
        alpha = e3 < e4; err = alpha ? e3 : e4;
    } else if (cfg.do4) {
        dxt1_eval_loop<true, false>(sc, cfg.U, p0, p1, p2, p3, pm, bound, e4, e3);
        alpha = 0; err = e4;
    } else {
        dxt1_eval_loop<false, true>(sc, cfg.U, p0, p1, p2, p3, pm, bound, e4, e3);
        alpha = 1; err = e3;
    }
    dxt1_prof_end(sc, prof_t0);
}

// Commit candidate (lo, hi, alt) with error e / block type alpha as the new best, applying the
// degenerate-endpoint fix-up of crn_dxt1.cpp:1563-1583 / :1781-1794.
template <typename SC>
__device__ __forceinline__ void dxt1_accept(SC* sc, unsigned lo, unsigned hi, int alt, unsigned long long e, int alpha)
{
    __syncwarp();                       // every lane has finished reading the previous best
    if (lane_id() == 0) {
        Dxt1Best nb;
        nb.lo = lo; nb.hi = hi; nb.err = e; nb.alpha_block = alpha; nb.alt_round = alt; nb.enforced_sel = 0;
        nb.enforce = !alpha && lo == hi;
        if (nb.enforce) {
            if ((nb.lo & 31u) != 31u) { nb.lo++; nb.enforced_sel = 1; }
            else { nb.hi--; nb.enforced_sel = 0; }
        }
        sc->best = nb;
    }
    __syncwarp();
}

// Static batch: every lane may hold one candidate (valid) whose sequence order is the lane index.
// Returns true if the best improved.
template <typename SC>
__device__ __noinline__ bool dxt1_commit_static(SC* sc, const Dxt1Cfg cfg,
                                                   bool valid, unsigned lo, unsigned hi, int alt)
{
    unsigned long long e = ~0ull; int alpha = 0;
    dxt1_eval(sc, cfg, lo, hi, alt, e, alpha, valid);
    unsigned long long key = e; unsigned idx = lane_id();
    warp_argmin_u64(key, idx);
    if (key >= sc->best.err) return false;
    const unsigned wlo = __shfl_sync(CRN_FULL_MASK, lo, idx), whi = __shfl_sync(CRN_FULL_MASK, hi, idx);
    const int walpha = __shfl_sync(CRN_FULL_MASK, alpha, idx);
    dxt1_accept(sc, wlo, whi, alt, key, walpha);
    return true;
}

__device__ __forceinline__ void canon(unsigned& lo, unsigned& hi)
{   // dxt1_solution_coordinates::canonicalize (crn_dxt1.h:78-85)
    if (lo < hi) { unsigned t = lo; lo = hi; hi = t; }
}

// Selectors of the current best for every unique colour -> sc->sel (first minimum in palette order;
// crn_dxt1.cpp:1407-1441 and :1845-1869 agree on ties).
template <typename SC>
__device__ __noinline__ void dxt1_best_selectors(SC* sc, const Dxt1Cfg cfg)
{
    for (int ci = (int)lane_id(); ci < cfg.U; ci += 32) {
        unsigned s;
        if (sc->best.enforce) s = (unsigned)sc->best.enforced_sel;
        else if (cfg.fast) {           // the selectors evaluate_solution_fast recorded for the winner (:1656, :1694)
            const Dxt1FastAxis A = dxt1_fast_axis(cfg, sc->best.lo, sc->best.hi, sc->best.alt_round);
            const int4 c = sc->cw[ci];
            s = sc->best.alpha_block ? dxt1_fast_sel3(A, c.x, c.y, c.z) : dxt1_fast_sel4(A, c.x, c.y, c.z);
        } else {
            int r0, g0, b0, r1, g1, b1;
            unpack565(sc->best.lo, true, r0, g0, b0);
            unpack565(sc->best.hi, true, r1, g1, b1);
            const int alt = sc->best.alt_round;
            const int4 c = sc->cw[ci];
            unsigned be = dxt1_dist(cfg, c.x, c.y, c.z, r0, g0, b0);
            s = 0;
            unsigned e = dxt1_dist(cfg, c.x, c.y, c.z, r1, g1, b1);
            if (e < be) { be = e; s = 1; }
            if (sc->best.alpha_block) {
                e = dxt1_dist(cfg, c.x, c.y, c.z, (r0 + r1 + alt) >> 1, (g0 + g1 + alt) >> 1, (b0 + b1 + alt) >> 1);
                if (e < be) { be = e; s = 2; }
            } else {
                e = dxt1_dist(cfg, c.x, c.y, c.z, (r0 * 2 + r1 + alt) / 3, (g0 * 2 + g1 + alt) / 3, (b0 * 2 + b1 + alt) / 3);
                if (e < be) { be = e; s = 2; }
                e = dxt1_dist(cfg, c.x, c.y, c.z, (r1 * 2 + r0 + alt) / 3, (g1 * 2 + g0 + alt) / 3, (b1 * 2 + b0 + alt) / 3);
                if (e < be) { be = e; s = 3; }
            }
        }
        sc->sel[ci] = (uint8_t)s;
    }
    __syncwarp();
}

// refine_solution (crn_dxt1.cpp:525-698), levels 0 and 1.
template <typename SC>
__device__ __noinline__ bool dxt1_refine(SC* sc, const Dxt1Cfg cfg, int level)
{
    dxt1_best_selectors(sc, cfg);
    double akku_0 = 0, akku_1 = 0, akku_2 = 0;
    double At1_r = 0, At1_g = 0, At1_b = 0, At2_r = 0, At2_g = 0, At2_b = 0;
    // Every term is an integer (small table value x colour component x pixel count) and every partial sum stays far below 2^53, so these
    // double sums are exact in any order: the colours are split over the lanes and the nine totals meet in a butterfly.
    for (int i = (int)lane_id(); i < cfg.U; i += 32) {
        const int4 c = sc->cw[i];
        const double weight = (double)(unsigned)c.w;
        const double r = c.x * weight, g = c.y * weight, b = c.z * weight;
        const int step = sc->sel[i] ^ 1;
        // w1Tab {3,0,2,1}; prods_0 {0,0,2,2}; prods_1 {0,9,1,4}; prods_2 {9,0,4,1}
        const int w1 = (0x1203 >> (4 * step)) & 15;
        const int p0 = (0x2200 >> (4 * step)) & 15, p1 = (0x4190 >> (4 * step)) & 15, p2 = (0x1409 >> (4 * step)) & 15;
        akku_0 += p0 * weight; akku_1 += p1 * weight; akku_2 += p2 * weight;
        At1_r += w1 * r; At1_g += w1 * g; At1_b += w1 * b;
        At2_r += r; At2_g += g; At2_b += b;
    }
    akku_0 = warp_sum_f64(akku_0); akku_1 = warp_sum_f64(akku_1); akku_2 = warp_sum_f64(akku_2);
    At1_r = warp_sum_f64(At1_r); At1_g = warp_sum_f64(At1_g); At1_b = warp_sum_f64(At1_b);
    At2_r = warp_sum_f64(At2_r); At2_g = warp_sum_f64(At2_g); At2_b = warp_sum_f64(At2_b);
    At2_r = 3 * At2_r - At1_r; At2_g = 3 * At2_g - At1_g; At2_b = 3 * At2_b - At1_b;
    const double xx = akku_2, yy = akku_1, xy = akku_0;
    const double t = xx * yy - xy * xy;
    if (!yy || !xx || (fabs(t) < (double).0000125f)) return false;
    const double frb = (double)(3.0f * 31.0f / 255.0f) / t;
    const double fg = frb * (double)(63.0f / 31.0f);
    int e0[3], e1[3];
    e0[0] = clampi(d2i_x86((At1_r * yy - At2_r * xy) * frb + (double)0.5f), 0, 31);
    e0[1] = clampi(d2i_x86((At1_g * yy - At2_g * xy) * fg + (double)0.5f), 0, 63);
    e0[2] = clampi(d2i_x86((At1_b * yy - At2_b * xy) * frb + (double)0.5f), 0, 31);
    e1[0] = clampi(d2i_x86((At2_r * xx - At1_r * xy) * frb + (double)0.5f), 0, 31);
    e1[1] = clampi(d2i_x86((At2_g * xx - At1_g * xy) * fg + (double)0.5f), 0, 63);
    e1[2] = clampi(d2i_x86((At2_b * xx - At1_b * xy) * frb + (double)0.5f), 0, 31);
    bool improved = false;
    if (level == 0) {
        unsigned mx = (unsigned)((e0[0] << 11) | (e0[1] << 5) | e0[2]);
        unsigned mn = (unsigned)((e1[0] << 11) | (e1[1] << 5) | e1[2]);
        canon(mn, mx);
        improved |= dxt1_commit_static(sc, cfg, lane_id() == 0, mn, mx, 0);
    } else {
        // 2 x 27 lattice neighbours, sequence = i*27 + (rr+1)*9 + (gr+1)*3 + (br+1)
#pragma unroll 1
        for (int base = 0; base < 64; base += 32) {
            const int k = base + (int)lane_id();
            const bool valid = k < 54;
            const int i = k >= 27, n = k - 27 * i;
            const int rr = n / 9 - 1, gr = (n / 3) % 3 - 1, br = n % 3 - 1;
            int c0[3] = { e0[0], e0[1], e0[2] }, c1[3] = { e1[0], e1[1], e1[2] };
            if (i) { c1[0] = clampi(c1[0] + rr, 0, 31); c1[1] = clampi(c1[1] + gr, 0, 63); c1[2] = clampi(c1[2] + br, 0, 31); }
            else { c0[0] = clampi(c0[0] + rr, 0, 31); c0[1] = clampi(c0[1] + gr, 0, 63); c0[2] = clampi(c0[2] + br, 0, 31); }
            unsigned lo = pack565_unscaled(c0[0], c0[1], c0[2]), hi = pack565_unscaled(c1[0], c1[1], c1[2]);
            canon(lo, hi);
            improved |= dxt1_commit_static(sc, cfg, valid, lo, hi, 0);
        }
    }
    return improved;
}

struct FastRandom { unsigned jsr, jcong; };   // crn_rand.cpp:310-405
__device__ __forceinline__ unsigned fr_u32(FastRandom& r)
{
    r.jsr ^= (r.jsr << 17); r.jsr ^= (r.jsr >> 13); r.jsr ^= (r.jsr << 5);
    r.jcong = 69069u * r.jcong + 1234567u;
    return r.jsr ^ r.jcong;
}
__device__ __forceinline__ float fr_frand(FastRandom& r, float l, float h)
{
    const double cNorm = 1.0 / 4294967296.0;
    float v = (float)((double)l + (double)(h - l) * ((double)fr_u32(r) * cNorm));
    return v < l ? l : (v > h ? h : v);
}

struct V3 { float x, y, z; };
__device__ __forceinline__ float v3_sqdist(const V3& a, const V3& b)
{
    float d2 = 0, d;
    d = a.x - b.x; d2 += d * d;
    d = a.y - b.y; d2 += d * d;
    d = a.z - b.z; d2 += d * d;
    return d2;
}
template <typename SC>
__device__ __forceinline__ V3 norm_color(SC* sc, int i, const V3& mean)
{   // m_norm_unique_colors[i] (crn_dxt1.cpp:168, :186)
    const int4 c = sc->cw[i];
    V3 v;
    v.x = (float)c.x * 1.0f / 255.0f - mean.x;
    v.y = (float)c.y * 1.0f / 255.0f - mean.y;
    v.z = (float)c.z * 1.0f / 255.0f - mean.z;
    return v;
}

// try_median4 (crn_dxt1.cpp:1181-1308)
template <typename SC>
__device__ __noinline__ bool dxt1_median4(SC* sc, const Dxt1Cfg cfg, int quality,
                                             const V3& mean, const V3& low_color, const V3& high_color)
{
    V3 means[4];
    const int U = cfg.U;
    if (U <= 4) {
#pragma unroll
        for (int i = 0; i < 4; i++) means[i] = norm_color(sc, min(U - 1, i), mean);
    } else {
        means[0].x = low_color.x - mean.x; means[0].y = low_color.y - mean.y; means[0].z = low_color.z - mean.z;
        means[3].x = high_color.x - mean.x; means[3].y = high_color.y - mean.y; means[3].z = high_color.z - mean.z;
        const float t1 = 1.0f / 3.0f, t2 = 2.0f / 3.0f;
        means[1].x = means[0].x + (means[3].x - means[0].x) * t1; means[1].y = means[0].y + (means[3].y - means[0].y) * t1; means[1].z = means[0].z + (means[3].z - means[0].z) * t1;
        means[2].x = means[0].x + (means[3].x - means[0].x) * t2; means[2].y = means[0].y + (means[3].y - means[0].y) * t2; means[2].z = means[0].z + (means[3].z - means[0].z) * t2;
        FastRandom rm = { 0xABCD917Au, 0x17F3DEADu };
        unsigned reassign_rover = 0;
        float prev_total_dist = 1.0e+37f;
#pragma unroll 1
        for (int iter = 0; iter < 8; iter++) {
            V3 nm[4]; float nw[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { nm[j].x = nm[j].y = nm[j].z = 0.0f; nw[j] = 0.0f; }
            float total_dist = 0;
            // cfg.par: lane l takes colours l, l + 32, ... and the 17 sums meet in a butterfly (every lane gets the same totals)
            const int i0 = cfg.par ? (int)lane_id() : 0, istep = cfg.par ? 32 : 1;
#pragma unroll 1
            for (int i = i0; i < U; i += istep) {
                const V3 v = norm_color(sc, i, mean);
                float best_dist = v3_sqdist(means[0], v);
                int best_index = 0;
#pragma unroll
                for (int j = 1; j < 4; j++) {
                    float dist = v3_sqdist(means[j], v);
                    if (dist < best_dist) { best_dist = dist; best_index = j; }
                }
                total_dist += best_dist;
                const float fw = (float)(unsigned)sc->cw[i].w;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (j == best_index) { nm[j].x += v.x * fw; nm[j].y += v.y * fw; nm[j].z += v.z * fw; nw[j] += fw; }
            }
            if (cfg.par) {
                total_dist = warp_sum_f32(total_dist);
#pragma unroll
                for (int j = 0; j < 4; j++) { nm[j].x = warp_sum_f32(nm[j].x); nm[j].y = warp_sum_f32(nm[j].y); nm[j].z = warp_sum_f32(nm[j].z); nw[j] = warp_sum_f32(nw[j]); }
            }
            unsigned highest_index = 0; float highest_weight = 0; bool empty_cell = false;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (nw[j] > 0.0f) {
                    means[j].x = nm[j].x / nw[j]; means[j].y = nm[j].y / nw[j]; means[j].z = nm[j].z / nw[j];
                    if (nw[j] > highest_weight) { highest_weight = nw[j]; highest_index = j; }
                } else empty_cell = true;
            }
            if (!empty_cell) {
                if (fabsf(total_dist - prev_total_dist) < .00001f) break;
                prev_total_dist = total_dist;
            } else prev_total_dist = 1.0e+37f;
            if (empty_cell && iter != 7) {
                const unsigned ri = (highest_index + reassign_rover) & 3;
                reassign_rover++;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (nw[j] == 0.0f) {
                        // means[j] = means[ri] reads the CURRENT means[ri] (it may itself have just been re-seeded)
                        V3 src = means[0];
#pragma unroll
                        for (int q = 1; q < 4; q++) if ((unsigned)q == ri) src = means[q];
                        means[j].x = src.x + fr_frand(rm, -.00196f, .00196f);
                        means[j].y = src.y + fr_frand(rm, -.00196f, .00196f);
                        means[j].z = src.z + fr_frand(rm, -.00196f, .00196f);
                    }
            }
        }
    }
    bool improved = false;
    {
        // 6 pairs (i<j): 01 02 03 12 13 23
        const unsigned lane = lane_id();
        const int pi = lane < 3 ? 0 : (lane < 5 ? 1 : 2);
        const int pj = lane < 3 ? (int)lane + 1 : (lane < 5 ? (int)lane - 1 : 3);
        V3 a = means[0], b = means[1];
#pragma unroll
        for (int j = 0; j < 4; j++) { if (j == pi) a = means[j]; if (j == pj) b = means[j]; }
        const float v0x = a.x + mean.x, v0y = a.y + mean.y, v0z = a.z + mean.z;
        const float v1x = b.x + mean.x, v1y = b.y + mean.y, v1z = b.z + mean.z;
        const int a0 = clampi((int)floorf(.5f + v0x * 31.0f), 0, 255), a1 = clampi((int)floorf(.5f + v0y * 63.0f), 0, 255), a2 = clampi((int)floorf(.5f + v0z * 31.0f), 0, 255);
        const int c0 = clampi((int)floorf(.5f + v1x * 31.0f), 0, 255), c1 = clampi((int)floorf(.5f + v1y * 63.0f), 0, 255), c2 = clampi((int)floorf(.5f + v1z * 31.0f), 0, 255);
        unsigned lo = pack565_unscaled(a0, a1, a2), hi = pack565_unscaled(c0, c1, c2);
        canon(lo, hi);
        improved |= dxt1_commit_static(sc, cfg, lane < 6, lo, hi, 0);
    }
    improved |= dxt1_refine(sc, cfg, quality == 4 ? 1 : 0);
    return improved;
}

// One "live" run over <= 32 lattice candidates around a fixed base colour of endpoint `which`
// (0 = low, 1 = high); the other endpoint is read from the live best (crn_dxt1.cpp:908-1015).
// delta(idx, dr, dg, db) supplies candidate idx's offset.
template <typename SC, typename DeltaFn>
__device__ __noinline__ void dxt1_live_neighbours(SC* sc, const Dxt1Cfg cfg, int which,
                                                     int ncand, DeltaFn delta)
{
    int cr, cg, cb;
    unpack565(which ? sc->best.hi : sc->best.lo, false, cr, cg, cb);
    int pos = 0;
    while (pos < ncand) {
        const int idx = pos + (int)lane_id();
        int dr = 0, dg = 0, db = 0;
        bool valid = idx < ncand;
        if (valid) delta(idx, dr, dg, db);
        const int r = cr + dr, g = cg + dg, b = cb + db;
        valid = valid && r >= 0 && r <= 31 && g >= 0 && g <= 63 && b >= 0 && b <= 31;
        unsigned lo, hi;
        const unsigned p = pack565_unscaled(max(r, 0), max(g, 0), max(b, 0));
        if (which) { lo = sc->best.lo; hi = p; } else { lo = p; hi = sc->best.hi; }
        canon(lo, hi);
        unsigned long long e = ~0ull; int alpha = 0;
        dxt1_eval(sc, cfg, lo, hi, 0, e, alpha, valid);
        const unsigned m = __ballot_sync(CRN_FULL_MASK, valid && e < sc->best.err);
        if (!m) { pos += 32; continue; }
        const int t = __ffs((int)m) - 1;
        const unsigned long long we = __shfl_sync(CRN_FULL_MASK, e, t);
        const unsigned wlo = __shfl_sync(CRN_FULL_MASK, lo, t), whi = __shfl_sync(CRN_FULL_MASK, hi, t);
        const int wa = __shfl_sync(CRN_FULL_MASK, alpha, t);
        dxt1_accept(sc, wlo, whi, 0, we, wa);
        pos = pos + t + 1;
    }
}

// try_average_block_as_solid (crn_dxt1.cpp:93-153)
template <typename SC>
__device__ __noinline__ bool dxt1_try_solid(SC* sc, const Dxt1Cfg cfg, const Dxt1Params& prm)
{
    unsigned long long tot_r = 0, tot_g = 0, tot_b = 0;
    unsigned total_weight = 0;
    for (int i = 0; i < cfg.U; i++) {
        const int4 c = sc->cw[i];
        total_weight += (unsigned)c.w;
        tot_r += (unsigned long long)c.x * (unsigned)c.w; tot_g += (unsigned long long)c.y * (unsigned)c.w; tot_b += (unsigned long long)c.z * (unsigned)c.w;
    }
    const unsigned half = total_weight >> 1;
    const int ar = (int)((tot_r + half) / total_weight), ag = (int)((tot_g + half) / total_weight), ab = (int)((tot_b + half) / total_weight);
    bool improved = false;
    // sequence: 0 ave/4, 1 ave/3, 2+2i colour i /4, 3+2i colour i /3
    const int ncand = prm.quality == 4 ? 2 + 2 * cfg.U : 2;
#pragma unroll 1
    for (int base = 0; base < ncand; base += 32) {
        const int k = base + (int)lane_id();
        bool valid = k < ncand;
        int r = ar, g = ag, b = ab;
        if (valid && k >= 2) {
            const int4 c = sc->cw[(k - 2) >> 1];
            r = c.x; g = c.y; b = c.z;
            if (r == ar && g == ag && b == ab) valid = false;   // :134-137
        }
        const bool three = k & 1;
        if (three && !prm.use_alpha_blocks) valid = false;
        unsigned lo, hi;
        if (three) {
            lo = ((unsigned)g_omatch5_3[2 * r] << 11) | ((unsigned)g_omatch6_3[2 * g] << 5) | g_omatch5_3[2 * b];
            hi = ((unsigned)g_omatch5_3[2 * r + 1] << 11) | ((unsigned)g_omatch6_3[2 * g + 1] << 5) | g_omatch5_3[2 * b + 1];
        } else {
            lo = ((unsigned)g_omatch5[2 * r] << 11) | ((unsigned)g_omatch6[2 * g] << 5) | g_omatch5[2 * b];
            hi = ((unsigned)g_omatch5[2 * r + 1] << 11) | ((unsigned)g_omatch6[2 * g + 1] << 5) | g_omatch5[2 * b + 1];
        }
        improved |= dxt1_commit_static(sc, cfg, valid, lo, hi, 0);
    }
    return improved;
}

// Closed form of compute_endpoint_component_errors (crn_dxt1.cpp:369-413): error[s][x] is the
// polynomial W[s]*p*p - WP2[s]*p + WPP[s] in wrapping 64-bit arithmetic (the reference's incremental
// d/dd recurrence for s >= 2 evaluates the same polynomial), so no table is materialised.
struct CompMoments { unsigned long long W[4], WP2[4], WPP[4], brem[4]; };
__device__ __forceinline__ unsigned long long comp_err(const CompMoments& m, int s, unsigned p)
{
    return m.W[s] * p * p - m.WP2[s] * p + m.WPP[s];
}
__device__ __forceinline__ unsigned expand_comp(int comp, unsigned c) { return comp == 1 ? ((c << 2) | (c >> 4)) : ((c << 3) | (c >> 2)); }
template <typename SC>
__device__ __noinline__ void comp_moments(SC* sc, const Dxt1Cfg cfg, int comp, CompMoments& m)
{
#pragma unroll
    for (int s = 0; s < 4; s++) m.W[s] = m.WP2[s] = m.WPP[s] = 0;
    const unsigned lane = lane_id();
    // wrapping integer sums: any order gives the same words, so the colours are split over the lanes
    for (int i = (int)lane; i < cfg.U; i += 32) {
        const int4 c = sc->cw[i];
        const unsigned long long p = (unsigned)(comp == 0 ? c.x : (comp == 1 ? c.y : c.z)), w = (unsigned)c.w;
        const int s = sc->sel[i];
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k == s) { m.W[k] += w; m.WP2[k] += w * p * 2; m.WPP[k] += w * p * p; }
    }
#pragma unroll
    for (int s = 0; s < 4; s++) { m.W[s] = warp_sum_u64(m.W[s]); m.WP2[s] = warp_sum_u64(m.WP2[s]); m.WPP[s] = warp_sum_u64(m.WPP[s]); }
    const unsigned limit = comp == 1 ? 64 : 32;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        unsigned long long mn = ~0ull;
        if (s < 2) {
            for (unsigned c = lane; c < limit; c += 32) mn = min(mn, comp_err(m, s, expand_comp(comp, c)));
        } else {
            for (unsigned p = lane; p < 256; p += 32) mn = min(mn, comp_err(m, s, p));
        }
        m.brem[s] = warp_min_u64(mn);
    }
    m.brem[2] += m.brem[3]; m.brem[1] += m.brem[2]; m.brem[0] += m.brem[1];
}

// optimize_endpoint_comps (crn_dxt1.cpp:415-486)
template <typename SC>
__device__ __noinline__ void dxt1_optimize_comps(SC* sc, const Dxt1Cfg cfg)
{
    dxt1_best_selectors(sc, cfg);
    if (sc->best.alpha_block || !sc->best.err) return;
    int sl[3], sh[3];
    unpack565(sc->best.lo, true, sl[0], sl[1], sl[2]);
    unpack565(sc->best.hi, true, sh[0], sh[1], sh[2]);
    const unsigned lane = lane_id();
#pragma unroll 1
    for (int comp = 0; comp < 3; comp++) {
        unsigned p0 = (unsigned)(comp == 0 ? sl[0] : (comp == 1 ? sl[1] : sl[2]));
        unsigned p1 = (unsigned)(comp == 0 ? sh[0] : (comp == 1 ? sh[1] : sh[2]));
        int low[3], high[3];
        unpack565(sc->best.lo, false, low[0], low[1], low[2]);
        unpack565(sc->best.hi, false, high[0], high[1], high[2]);
        CompMoments m;
        comp_moments(sc, cfg, comp, m);
        const unsigned lowc = (unsigned)(comp == 0 ? low[0] : (comp == 1 ? low[1] : low[2]));
        const unsigned highc = (unsigned)(comp == 0 ? high[0] : (comp == 1 ? high[1] : high[2]));
        unsigned long long best_error = comp_err(m, 0, expand_comp(comp, lowc)) + comp_err(m, 1, expand_comp(comp, highc)) +
                                        comp_err(m, 2, (p0 * 2 + p1) / 3) + comp_err(m, 3, (p0 + p1 * 2) / 3);
        if (m.brem[0] >= best_error) continue;
        const unsigned limit = comp == 1 ? 64 : 32;
#pragma unroll 1
        for (unsigned c0 = 0; c0 < limit; c0++) {
            unsigned long long e0 = comp_err(m, 0, expand_comp(comp, c0));
            if (e0 + m.brem[1] >= best_error) continue;
            if (comp == 0) low[0] = (int)c0; else if (comp == 1) low[1] = (int)c0; else low[2] = (int)c0;
            const unsigned packed_low = pack565_unscaled(low[0], low[1], low[2]);
            p0 = expand_comp(comp, c0);
            unsigned c1_start = 0;
            bool leave_c1 = false;
            while (c1_start < limit && !leave_c1) {
                const unsigned c1 = c1_start + lane;
                bool pass = c1 < limit;
                unsigned long long e = e0 + comp_err(m, 1, expand_comp(comp, c1 & 63));
                pass = pass && (e + m.brem[2] < best_error);
                const unsigned q1 = expand_comp(comp, c1 & 63);
                e += comp_err(m, 2, (p0 * 2 + q1) / 3);
                pass = pass && (e + m.brem[3] < best_error);
                e += comp_err(m, 3, (p0 + q1 * 2) / 3);
                pass = pass && (e < best_error);
                const unsigned surv = __ballot_sync(CRN_FULL_MASK, pass);
                if (!surv) { c1_start += 32; continue; }
                const unsigned wc1 = c1_start + (unsigned)(__ffs((int)surv) - 1);
                if (comp == 0) high[0] = (int)wc1; else if (comp == 1) high[1] = (int)wc1; else high[2] = (int)wc1;
                p1 = expand_comp(comp, wc1);
                c1_start = wc1 + 1;
                // single-candidate evaluation (every lane computes the same thing)
                unsigned long long ce; int ca;
                const unsigned chi = pack565_unscaled(high[0], high[1], high[2]);
                dxt1_eval(sc, cfg, packed_low, chi, 0, ce, ca);
                if (ce >= sc->best.err) continue;
                dxt1_accept(sc, packed_low, chi, 0, ce, ca);
                if (!sc->best.err) return;
                dxt1_best_selectors(sc, cfg);
                comp_moments(sc, cfg, comp, m);
                best_error = comp_err(m, 0, expand_comp(comp, c0)) + comp_err(m, 1, expand_comp(comp, wc1)) +
                             comp_err(m, 2, (p0 * 2 + p1) / 3) + comp_err(m, 3, (p0 + p1 * 2) / 3);
                e0 = comp_err(m, 0, expand_comp(comp, c0));
                if (e0 + m.brem[1] >= best_error) leave_c1 = true;
            }
        }
    }
}

__device__ __forceinline__ unsigned lerp_color_packed(const int4& a, const int4& b, float f, int rounding)
{   // lerp_color (crn_dxt1.cpp:1871-1882) followed by pack_color(.., scaled = true) (:1964)
    const float r = rounding ? 1.0f : 0.0f;
    const float ar = (float)a.x, ag = (float)a.y, ab = (float)a.z, br = (float)b.x, bg = (float)b.y, bb = (float)b.z;
    const int cr = clampi((int)(r + (ar + (br - ar) * f)), 0, 255);
    const int cg = clampi((int)(r + (ag + (bg - ag) * f)), 0, 255);
    const int cb = clampi((int)(r + (ab + (bb - ab) * f)), 0, 255);
    return pack565_scaled(cr, cg, cb);
}

// try_combinatorial_encoding (crn_dxt1.cpp:1886-1997)
template <typename SC>
__device__ __noinline__ void dxt1_combinatorial(SC* sc, const Dxt1Cfg cfg)
{
    const int U = cfg.U;
    if (U < 2 || U > 4) return;
    // build the de-duplicated packed list serially (tiny), identically on every lane; lane 0 stores it
    unsigned np = 0;
    const unsigned lane = lane_id();
    auto push = [&](unsigned pc) {
        for (unsigned j = 0; j < np; j++) if (sc->packed[j] == pc) return;
        if (lane == 0) sc->packed[np] = (uint16_t)pc;
        np++;
        __syncwarp();
    };
    for (int i = 0; i < U; i++) { const int4 c = sc->cw[i]; push(pack565_scaled(c.x, c.y, c.z)); }
    if (U == 2) {
        const float f2[10] = { 2.0f, 3.0f, .5f, 1.5f, -1.0f, 2.0f, -.5f, .5f, -2.0f, -1.0f };
        for (int k = 0; k < 2; k++)
            for (int q = 0; q < 2; q++)
                for (int mth = 0; mth < 10; mth++) push(lerp_color_packed(sc->cw[q], sc->cw[q ^ 1], f2[mth], k));
    } else if (U == 3) {
        const float f3[4] = { 1.5f, 2.0f / 3.0f, 1.0f / 3.0f, -.5f };
        for (int i = 0; i <= 2; i++)
            for (int j = 0; j <= 2; j++) {
                if (i == j) continue;
                for (int mth = 0; mth < 4; mth++) push(lerp_color_packed(sc->cw[i], sc->cw[j], f3[mth], 1));
            }
    }
    __syncwarp();
    const unsigned npairs = np * (np - 1) / 2;
#pragma unroll 1
    for (int alt = 0; alt < 2; alt++) {
        if (!sc->best.err) break;
        // per-lane running first-minimum over its pairs (sequence number k = i-major pair index)
        unsigned long long my_e = ~0ull; unsigned my_k = 0xffffffffu, my_lo = 0, my_hi = 0; int my_a = 0;
        unsigned i = 0, j = 1;
        for (unsigned s = 0; s < lane; s++) { if (++j >= np) { i++; j = i + 1; } }
        for (unsigned k0 = 0; k0 < npairs; k0 += 32) {              // uniform trip count: see dxt1_eval
            const unsigned k = k0 + lane;
            const bool valid = k < npairs;
            unsigned long long e; int a;
            const unsigned lo = valid ? sc->packed[i] : 0u, hi = valid ? sc->packed[j] : 0u;
            dxt1_eval(sc, cfg, lo, hi, alt, e, a, valid);
            if (valid && e < my_e) { my_e = e; my_k = k; my_lo = lo; my_hi = hi; my_a = a; }
            if (valid) for (int s = 0; s < 32; s++) { if (++j >= np) { i++; j = i + 1; if (i + 1 >= np) break; } }
        }
        unsigned long long key = my_e; unsigned idx = my_k;
        warp_argmin_u64(key, idx);
        // alt == 1: only a zero-error candidate is accepted (best error forced to 1, :1981-1996)
        const unsigned long long bar = alt ? 1ull : sc->best.err;
        if (key < bar) {
            const unsigned src = idx & 31u;
            dxt1_accept(sc, __shfl_sync(CRN_FULL_MASK, my_lo, src), __shfl_sync(CRN_FULL_MASK, my_hi, src), alt, key,
                        __shfl_sync(CRN_FULL_MASK, my_a, src));
        }
    }
}

// x^2.75 for x in [0,1], correctly rounded to float via exactly-rounded double operations.  The
// reference calls glibc powf (crn_dxt1.cpp:1097), which is itself not correctly rounded for ~2.4e-4
// of the float inputs in [0,1] (1-ulp differences; measured exhaustively) and whose
// result depends on the host's libm variant; a 1-ulp change of `p` moves the perceptual weights by
// one ulp and virtually never alters a candidate endpoint.
__device__ __forceinline__ float pow275(float x)
{
    const double d = (double)x;
    const double s = sqrt(d);
    const double q = sqrt(s);
    return (float)(d * d * s * q);
}

__device__ __forceinline__ Dxt1Cfg dxt1_make_cfg(const Dxt1Params& prm, int pixels_have_alpha, int U)
{
    Dxt1Cfg cfg;
    cfg.U = U;
    cfg.hc = prm.quality == 4 && !pixels_have_alpha && !prm.force_alpha_blocks && !prm.use_alpha_blocks && !prm.grayscale_sampling;
    const bool perceptual = prm.perceptual && !prm.grayscale_sampling;
    cfg.gray = !perceptual && prm.grayscale_sampling;
    cfg.wr = perceptual ? 8 : 1; cfg.wg = perceptual ? 25 : 1; cfg.wb = 1;
    cfg.fast = prm.quality < 3; cfg.perc = perceptual;
    cfg.par = prm.parallel_sums && U > kDxt1ParallelSumMinColours;
    if (pixels_have_alpha || prm.force_alpha_blocks) { cfg.do4 = false; cfg.do3 = true; }
    else if (!prm.use_alpha_blocks) { cfg.do4 = true; cfg.do3 = false; }
    else { cfg.do4 = true; cfg.do3 = true; }
    return cfg;
}

// (re)build the evaluation form of the unique colours after a state load
template <typename SC>
__device__ __forceinline__ void dxt1_build_eval_colours(SC* sc, const Dxt1Cfg cfg)
{
    for (int ci = (int)lane_id(); ci < cfg.U; ci += 32) { const int4 c = sc->cw[ci]; sc->ce[ci] = eval_colour(cfg, c.x, c.y, c.z, c.w); }
    __syncwarp();
}

// Phase 0, common part -- compute_internal after the unique colours exist (sc->cw / sc->ce filled,
// first-appearance order) up to the call of optimize_endpoints (crn_dxt1.cpp:2131-2232, :1069-1178).
template <typename SC>
__device__ __forceinline__ void dxt1_setup_common(SC* sc, const Dxt1Params& prm, int pixels_have_alpha, const int U, const unsigned total_w,
                                                  const bool has_transparent)
{
    const unsigned lane = lane_id();
    const Dxt1Cfg cfg = dxt1_make_cfg(prm, pixels_have_alpha, U);
    const bool perceptual = prm.perceptual && !prm.grayscale_sampling;
    const int i0 = cfg.par ? (int)lane : 0, istep = cfg.par ? 32 : 1;       // cfg.par: each O(U) float sum below is split over the lanes
    if (lane == 0) {
        sc->best.lo = sc->best.hi = 0; sc->best.err = ~0ull; sc->best.alpha_block = 0; sc->best.alt_round = 0; sc->best.enforce = 0; sc->best.enforced_sel = 0;
    }
    __syncwarp();

    if (lane == 0) { sc->U = U; sc->total_w = (int)total_w; sc->pixels_have_alpha = pixels_have_alpha; sc->stage = U == 0 ? 2 : 0; }
    __syncwarp();
    if (U == 0) return;   // :2205-2211
    if (U == 1 && !has_transparent) {   // :2212-2227
        const int4 c = sc->cw[0];
        const unsigned lo4 = ((unsigned)g_omatch5[2 * c.x] << 11) | ((unsigned)g_omatch6[2 * c.y] << 5) | g_omatch5[2 * c.z];
        const unsigned hi4 = ((unsigned)g_omatch5[2 * c.x + 1] << 11) | ((unsigned)g_omatch6[2 * c.y + 1] << 5) | g_omatch5[2 * c.z + 1];
        const unsigned lo3 = ((unsigned)g_omatch5_3[2 * c.x] << 11) | ((unsigned)g_omatch6_3[2 * c.y] << 5) | g_omatch5_3[2 * c.z];
        const unsigned hi3 = ((unsigned)g_omatch5_3[2 * c.x + 1] << 11) | ((unsigned)g_omatch6_3[2 * c.y + 1] << 5) | g_omatch5_3[2 * c.z + 1];
        dxt1_commit_static(sc, cfg, lane < (prm.use_alpha_blocks ? 2u : 1u), lane ? lo3 : lo4, lane ? hi3 : hi4, 0);
        if (lane == 0) sc->stage = 1;
        __syncwarp();
    } else {
        // ---- handle_multicolor_block (:1069-1178)
        int num_passes = 1;
        float pwx = 1.0f, pwy = 1.0f, pwz = 1.0f;
        if (perceptual) {
            float ave_redness = 0, ave_blueness = 0, ave_l = 0;
            for (int i = i0; i < U; i += istep) {
                const int4 c = sc->cw[i];
                const int l = (c.x + c.y + c.z + 1) / 3;
                const float fl = (float)l;
                const float scale = (float)(unsigned)c.w / (1.0f > fl ? 1.0f : fl);
                ave_redness += scale * (float)c.x;
                ave_blueness += scale * (float)c.z;
                ave_l += fl;
            }
            if (cfg.par) { ave_redness = warp_sum_f32(ave_redness); ave_blueness = warp_sum_f32(ave_blueness); ave_l = warp_sum_f32(ave_l); }
            const float ftw = (float)total_w;
            ave_redness /= ftw; ave_blueness /= ftw; ave_l /= ftw;
            ave_l = ave_l * 16.0f / 255.0f;
            ave_l = 1.0f < ave_l ? 1.0f : ave_l;
            const float mx = ave_redness > ave_blueness ? ave_redness : ave_blueness;
            float sat = mx * 1.0f / 3.0f;
            sat = sat < 0.0f ? 0.0f : (sat > 1.0f ? 1.0f : sat);
            const float p = ave_l * pow275(sat);
            if (!(p >= 1.0f)) {
                num_passes = 2;
                pwx = .212f + (pwx - .212f) * p; pwy = .72f + (pwy - .72f) * p; pwz = .072f + (pwz - .072f) * p;
            }
        }
        V3 mean, axis;
        mean.x = mean.y = mean.z = 0.0f; axis.x = axis.y = axis.z = 0.0f;
#pragma unroll 1
        for (int pass_index = 0; pass_index < num_passes; pass_index++) {
            // compute_vectors (:155-189)
            V3 meanw;
            mean.x = mean.y = mean.z = 0.0f; meanw.x = meanw.y = meanw.z = 0.0f;
            for (int i = i0; i < U; i += istep) {
                const int4 c = sc->cw[i];
                const float fw = (float)(unsigned)c.w;
                const float nx = (float)c.x * 1.0f / 255.0f, ny = (float)c.y * 1.0f / 255.0f, nz = (float)c.z * 1.0f / 255.0f;
                const float wx = pwx * nx, wy = pwy * ny, wz = pwz * nz;
                mean.x += nx * fw; mean.y += ny * fw; mean.z += nz * fw;
                meanw.x += wx * fw; meanw.y += wy * fw; meanw.z += wz * fw;
            }
            if (cfg.par) {
                mean.x = warp_sum_f32(mean.x); mean.y = warp_sum_f32(mean.y); mean.z = warp_sum_f32(mean.z);
                meanw.x = warp_sum_f32(meanw.x); meanw.y = warp_sum_f32(meanw.y); meanw.z = warp_sum_f32(meanw.z);
            }
            {
                const float inv = 1.0f / (float)total_w;
                mean.x *= inv; mean.y *= inv; mean.z *= inv;
                meanw.x *= inv; meanw.y *= inv; meanw.z *= inv;
            }
            // compute_pca on the weighted vectors (:192-256)
            double cov0 = 0, cov1 = 0, cov2 = 0, cov3 = 0, cov4 = 0, cov5 = 0;
            for (int i = i0; i < U; i += istep) {
                const int4 c = sc->cw[i];
                const float nx = (float)c.x * 1.0f / 255.0f, ny = (float)c.y * 1.0f / 255.0f, nz = (float)c.z * 1.0f / 255.0f;
                const float r = pwx * nx - meanw.x, g = pwy * ny - meanw.y, b = pwz * nz - meanw.z;
                const float rr = r * r, rg = r * g, rb = r * b, gg = g * g, gb = g * b, bb = b * b;
                if (c.w > 1) {
                    const double weight = (double)(unsigned)c.w;
                    cov0 += (double)rr * weight; cov1 += (double)rg * weight; cov2 += (double)rb * weight;
                    cov3 += (double)gg * weight; cov4 += (double)gb * weight; cov5 += (double)bb * weight;
                } else {
                    cov0 += (double)rr; cov1 += (double)rg; cov2 += (double)rb; cov3 += (double)gg; cov4 += (double)gb; cov5 += (double)bb;
                }
            }
            if (cfg.par) {
                cov0 = warp_sum_f64(cov0); cov1 = warp_sum_f64(cov1); cov2 = warp_sum_f64(cov2);
                cov3 = warp_sum_f64(cov3); cov4 = warp_sum_f64(cov4); cov5 = warp_sum_f64(cov5);
            }
            double vfr = (double).9f, vfg = 1.0, vfb = (double).7f;
#pragma unroll 1
            for (int iter = 0; iter < 8; iter++) {
                double r = vfr * cov0 + vfg * cov1 + vfb * cov2;
                double g = vfr * cov1 + vfg * cov3 + vfb * cov4;
                double b = vfr * cov2 + vfg * cov4 + vfb * cov5;
                double m = fabs(r) > fabs(g) ? fabs(r) : fabs(g);
                m = m > fabs(b) ? m : fabs(b);
                if (m > 1e-10) { m = 1.0 / m; r *= m; g *= m; b *= m; }
                const double delta = (vfr - r) * (vfr - r) + (vfg - g) * (vfg - g) + (vfb - b) * (vfb - b);
                vfr = r; vfg = g; vfb = b;
                if (iter > 2 && delta < 1e-8) break;
            }
            double len = vfr * vfr + vfg * vfg + vfb * vfb;
            if (len < 1e-10) { axis.x = .2837149f; axis.y = 0.9540631f; axis.z = 0.096277453f; }
            else {
                len = 1.0 / sqrt(len);
                axis.x = (float)(vfr * len); axis.y = (float)(vfg * len); axis.z = (float)(vfb * len);
            }
            axis.x /= pwx; axis.y /= pwy; axis.z /= pwz;
            {   // vec::normalize (crn_vec.h:674-689)
                double n = (double)(axis.x * axis.x);
                n += (double)(axis.y * axis.y);
                n += (double)(axis.z * axis.z);
                if (n != 0) { const float s = (float)(1.0 / sqrt(n)); axis.x *= s; axis.y *= s; axis.z *= s; }
            }
            if (num_passes > 1) {
                if (fabsf(axis.x) >= .795f) { pwx = .424f; pwy = .6f; pwz = .072f; }
                else if (fabsf(axis.z) >= .795f) { pwx = .212f; pwy = .6f; pwz = .212f; }
                else break;
            }
        }
        float l = 1e+9f, h = -1e+9f;
        for (int i = i0; i < U; i += istep) {
            const V3 v = norm_color(sc, i, mean);
            float d = v.x * axis.x;
            d += v.y * axis.y;
            d += v.z * axis.z;
            l = l < d ? l : d;
            h = h > d ? h : d;
        }
        if (cfg.par) {
#pragma unroll
            for (int ofs = 16; ofs > 0; ofs >>= 1) {
                const float ol = __shfl_xor_sync(CRN_FULL_MASK, l, ofs), oh = __shfl_xor_sync(CRN_FULL_MASK, h, ofs);
                l = fminf(l, ol); h = fmaxf(h, oh);                  // (commutative, so every lane ends with the same pair)
            }
        }
        V3 low_color, high_color;
        low_color.x = mean.x + axis.x * l; low_color.y = mean.y + axis.y * l; low_color.z = mean.z + axis.z * l;
        high_color.x = mean.x + axis.x * h; high_color.y = mean.y + axis.y * h; high_color.z = mean.z + axis.z * h;
        // ray/AABB clamp into the unit cube (crn_intersect.h:44-132), sign = +axis for low, -axis for high
#pragma unroll
        for (int which = 0; which < 2; which++) {
            V3& pt = which ? high_color : low_color;
            const float sgn = which ? -1.0f : 1.0f;
            const float o[3] = { pt.x, pt.y, pt.z };
            const float dir[3] = { sgn * axis.x, sgn * axis.y, sgn * axis.z };
            bool inside = true;
            int quad[3]; float plane[3] = { 0, 0, 0 };
#pragma unroll
            for (int i = 0; i < 3; i++) {
                if (o[i] < 0.0f) { quad[i] = 1; plane[i] = 0.0f; inside = false; }
                else if (o[i] > 1.0f) { quad[i] = 0; plane[i] = 1.0f; inside = false; }
                else quad[i] = 2;
            }
            if (!inside) {
                float max_t[3];
#pragma unroll
                for (int i = 0; i < 3; i++)
                    max_t[i] = (quad[i] != 2 && dir[i] != 0.0f) ? (plane[i] - o[i]) / dir[i] : -1.0f;
                int wp = 0;
                if (max_t[wp] < max_t[1]) wp = 1;
                if ((wp == 0 ? max_t[0] : max_t[1]) < max_t[2]) wp = 2;
                const float mt = wp == 0 ? max_t[0] : (wp == 1 ? max_t[1] : max_t[2]);
                if (!(mt < 0.0f)) {
                    float coord[3];
                    bool ok = true;
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        if (i != wp) {
                            coord[i] = o[i] + mt * dir[i];
                            if (coord[i] < 0.0f || coord[i] > 1.0f) ok = false;
                        } else coord[i] = plane[i];
                    }
                    if (ok) { pt.x = coord[0]; pt.y = coord[1]; pt.z = coord[2]; }
                }
            }
        }

        if (lane == 0) {
            sc->mean[0] = mean.x; sc->mean[1] = mean.y; sc->mean[2] = mean.z;
            sc->axis[0] = axis.x; sc->axis[1] = axis.y; sc->axis[2] = axis.z;
            sc->low[0] = low_color.x; sc->low[1] = low_color.y; sc->low[2] = low_color.z;
            sc->high[0] = high_color.x; sc->high[1] = high_color.y; sc->high[2] = high_color.z;
        }
        __syncwarp();
    }
}

// Phase 0 for a 4x4 block: lanes 0..15 hold pixel 4y+x as RGBA8 (r in the low byte).  Unique colours in
// first-appearance order (crn_dxt1.cpp:2113-2131) come from one __match_any_sync.
template <typename SC>
__device__ __forceinline__ void dxt1_phase_setup(SC* sc, uint32_t px, const Dxt1Params& prm, int pixels_have_alpha)
{
    const unsigned lane = lane_id();
    const Dxt1Cfg cfg = dxt1_make_cfg(prm, pixels_have_alpha, 0);
    const bool opaque = lane < 16 && (!pixels_have_alpha || (px >> 24) >= prm.alpha_threshold);
    const unsigned vmask = __ballot_sync(CRN_FULL_MASK, opaque);
    const unsigned key = px | 0xFF000000u;
    unsigned peers = 0;
    if (opaque) peers = __match_any_sync(vmask, key);
    const bool leader = opaque && (unsigned)(__ffs((int)peers) - 1) == lane;
    const unsigned leaders = __ballot_sync(CRN_FULL_MASK, leader);
    const int U = __popc(leaders);
    const unsigned my_u = __popc(leaders & lanemask_lt());
    if (leader) {
        const int r = (int)(px & 0xff), g = (int)((px >> 8) & 0xff), b = (int)((px >> 16) & 0xff);
        sc->cw[my_u] = make_int4(r, g, b, __popc(peers));
        sc->ce[my_u] = eval_colour(cfg, r, g, b, __popc(peers));
    }
    const unsigned total_w = (unsigned)__popc(vmask);
    __syncwarp();
    dxt1_setup_common(sc, prm, pixels_have_alpha, U, total_w, total_w != 16);
}

// Phase 1 -- try_median4 (+ its least-squares refinement), first step of optimize_endpoints (:765-771).
template <typename SC>
__device__ __forceinline__ void dxt1_phase_median4(SC* sc, const Dxt1Params& prm)
{
    if (sc->stage != 0 || prm.quality < 3) return;      // try_median4 only from better up (:765-771)
    const Dxt1Cfg cfg = dxt1_make_cfg(prm, sc->pixels_have_alpha, sc->U);
    V3 mean, low_color, high_color;
    mean.x = sc->mean[0]; mean.y = sc->mean[1]; mean.z = sc->mean[2];
    low_color.x = sc->low[0]; low_color.y = sc->low[1]; low_color.z = sc->low[2];
    high_color.x = sc->high[0]; high_color.y = sc->high[1]; high_color.z = sc->high[2];
    dxt1_median4(sc, cfg, prm.quality, mean, low_color, high_color);
}

// Phase 2 -- the probe-sweep / lattice-neighbour / refine passes of optimize_endpoints (:773-1028).
template <typename SC>
__device__ __forceinline__ void dxt1_phase_passes(SC* sc, const Dxt1Params& prm)
{
    if (sc->stage != 0) return;
    const unsigned lane = lane_id();
    const Dxt1Cfg cfg = dxt1_make_cfg(prm, sc->pixels_have_alpha, sc->U);
    V3 axis, low_color, high_color;
    axis.x = sc->axis[0]; axis.y = sc->axis[1]; axis.z = sc->axis[2];
    low_color.x = sc->low[0]; low_color.y = sc->low[1]; low_color.z = sc->low[2];
    high_color.x = sc->high[0]; high_color.y = sc->high[1]; high_color.z = sc->high[2];
    {
        // ---- optimize_endpoints (:703-1067)
        const int quality = prm.quality;
        int num_passes_o, probe_range;
        float dist_per_trial = .015625f;
        // probe tables (:43-53) packed as bytes
        const uint8_t* probe_tab;
        if (quality >= 4) { probe_tab = g_uber_probe; probe_range = 15; num_passes_o = 4; }
        else if (quality == 3) { probe_tab = g_better_probe; probe_range = 10; num_passes_o = 2; }
        else if (quality == 2) { probe_tab = g_normal_probe; probe_range = 5; num_passes_o = 2; dist_per_trial = .027063293f; }
        else { probe_tab = g_fast_probe; probe_range = 4; num_passes_o = quality == 1 ? 2 : 1; dist_per_trial = .027063293f; }

        float sx = axis.x * dist_per_trial, sy = axis.y * dist_per_trial, sz = axis.z * dist_per_trial;
        sx *= 31.0f; sy *= 63.0f; sz *= 31.0f;
        float lcx = low_color.x * 31.0f, lcy = low_color.y * 63.0f, lcz = low_color.z * 31.0f;
        float hcx = high_color.x * 31.0f, hcy = high_color.y * 63.0f, hcz = high_color.z * 31.0f;
        lcx = lcx < 0.0f ? 0.0f : (lcx > 31.0f ? 31.0f : lcx); lcy = lcy < 0.0f ? 0.0f : (lcy > 63.0f ? 63.0f : lcy); lcz = lcz < 0.0f ? 0.0f : (lcz > 31.0f ? 31.0f : lcz);
        hcx = hcx < 0.0f ? 0.0f : (hcx > 31.0f ? 31.0f : hcx); hcy = hcy < 0.0f ? 0.0f : (hcy > 63.0f ? 63.0f : hcy); hcz = hcz < 0.0f ? 0.0f : (hcz > 31.0f ? 31.0f : hcz);

#pragma unroll 1
        for (int pass = 0; pass < num_passes_o; pass++) {
            if (pass) {
                int r, g, b;
                unpack565(sc->best.lo, false, r, g, b); lcx = (float)r; lcy = (float)g; lcz = (float)b;
                unpack565(sc->best.hi, false, r, g, b); hcx = (float)r; hcy = (float)g; hcz = (float)b;
            }
            const unsigned long long prev_best_error = sc->best.err;
            if (!prev_best_error) break;
            // probe sweeps (:840-892): sequence index t: 0 -> (i=0,s=1); 2i-1 -> (i,s=0); 2i -> (i,s=1)
            int n_probe[2];
#pragma unroll
            for (int which = 0; which < 2; which++) {
                const float ix = (which ? hcx : lcx) + .5f, iy = (which ? hcy : lcy) + .5f, iz = (which ? hcz : lcz) + .5f;
                const int t = (int)lane;
                const int nseq = 2 * probe_range - 1;
                const int i = (t + 1) >> 1, s = (t == 0) ? 1 : ((t & 1) ? 0 : 1);
                int packed = -1, prev = -1;
                if (t < nseq) {
                    const int x = probe_tab[min(i, probe_range - 1)];
                    const float fx = (float)x;
                    const float ax = s ? sx : -sx, ay = s ? sy : -sy, az = s ? sz : -sz;
                    const float px_ = ix + ax * fx, py_ = iy + ay * fx, pz_ = iz + az * fx;
                    packed = clampi((int)floorf(pz_), 0, 31) | (clampi((int)floorf(py_), 0, 63) << 5) | (clampi((int)floorf(px_), 0, 31) << 11);
                    const bool has_prev = s ? (i >= 1) : (i >= 2);
                    if (has_prev) {
                        const int xp = probe_tab[i - 1];
                        const float fp = (float)xp;
                        const float qx = ix + ax * fp, qy = iy + ay * fp, qz = iz + az * fp;
                        prev = clampi((int)floorf(qz), 0, 31) | (clampi((int)floorf(qy), 0, 63) << 5) | (clampi((int)floorf(qx), 0, 31) << 11);
                    }
                }
                const bool keep = t < nseq && packed != prev;
                const unsigned km = __ballot_sync(CRN_FULL_MASK, keep);
                if (keep) sc->probe[which][__popc(km & lanemask_lt())] = (uint16_t)packed;
                n_probe[which] = __popc(km);
            }
            __syncwarp();
            // all pairs (:895-903): static run, per-lane running first-minimum then one reduction
            {
                const int nl = n_probe[0], nh = n_probe[1], total = nl * nh;
                unsigned long long my_e = ~0ull; unsigned my_k = 0xffffffffu, my_lo = 0, my_hi = 0; int my_a = 0;
                for (int k0 = 0; k0 < total; k0 += 32) {                // uniform trip count: see dxt1_eval
                    const int k = k0 + (int)lane;
                    const bool valid = k < total;
                    unsigned lo = valid ? sc->probe[0][k / nh] : 0u, hi = valid ? sc->probe[1][k % nh] : 0u;
                    canon(lo, hi);
                    unsigned long long e; int a;
                    dxt1_eval(sc, cfg, lo, hi, 0, e, a, valid);
                    if (valid && e < my_e) { my_e = e; my_k = (unsigned)k; my_lo = lo; my_hi = hi; my_a = a; }
                }
                unsigned long long keyv = my_e; unsigned idx = my_k;
                warp_argmin_u64(keyv, idx);
                if (keyv < sc->best.err) {
                    const unsigned src = idx & 31u;
                    dxt1_accept(sc, __shfl_sync(CRN_FULL_MASK, my_lo, src), __shfl_sync(CRN_FULL_MASK, my_hi, src), 0, keyv,
                                __shfl_sync(CRN_FULL_MASK, my_a, src));
                }
            }
            // lattice neighbours (:905-1016), from normal quality up
#pragma unroll 1
            for (int which = 0; which < 2 && quality >= 2; which++) {
                dxt1_live_neighbours(sc, cfg, which, 26, [](int idx, int& dr, int& dg, int& db) {
                    const int n = idx < 13 ? idx : idx + 1;   // g_adjacency (:489-522): x fastest, centre skipped
                    dr = n % 3 - 1; dg = (n / 3) % 3 - 1; db = n / 9 - 1;
                });
                if (quality == 4)
                    dxt1_live_neighbours(sc, cfg, which, 6, [](int idx, int& dr, int& dg, int& db) {
                        const int a = idx >> 1, s = (idx & 1) ? 2 : -2;
                        dr = a == 0 ? s : 0; dg = a == 1 ? s : 0; db = a == 2 ? s : 0;
                    });
            }
            if (!sc->best.err || (pass && sc->best.err == prev_best_error)) break;
            if (quality >= 4) dxt1_refine(sc, cfg, 1);
        }
    }
}

// Phase 3 -- solid-colour and per-component post passes (:1030-1046).
template <typename SC>
__device__ __forceinline__ void dxt1_phase_post(SC* sc, const Dxt1Params& prm)
{
    if (sc->stage != 0) return;
    const Dxt1Cfg cfg = dxt1_make_cfg(prm, sc->pixels_have_alpha, sc->U);
    const int quality = prm.quality;
    const int U = sc->U;
    if (quality >= 2 && sc->best.err && !sc->pixels_have_alpha) {      // :1030
        bool choose_solid_block = false;
        dxt1_best_selectors(sc, cfg);
        bool all_equal = true;
        {
            const uint8_t s0 = sc->sel[0];
            for (int i = 1 + (int)lane_id(); i < U; i += 32) all_equal = all_equal && sc->sel[i] == s0;
            all_equal = __all_sync(CRN_FULL_MASK, all_equal);
        }
        if (all_equal) choose_solid_block = dxt1_try_solid(sc, cfg, prm);
        if (!choose_solid_block && quality == 4) dxt1_optimize_comps(sc, cfg);
    }
}

// Phase 4 -- combinatorial recovery (:1048-1056) and return_solution (:263-365).  Returns the packed 8-byte
// DXT1 element (low565, high565, 16 x 2-bit selectors; crn_dxt.h:109-172) on every lane.
template <typename SC>
__device__ __forceinline__ unsigned long long dxt1_phase_finish(SC* sc, uint32_t px, const Dxt1Params& prm)
{
    const unsigned lane = lane_id();
    if (sc->stage == 2) return 0xFFFFFFFF00000000ull;
    const Dxt1Cfg cfg = dxt1_make_cfg(prm, sc->pixels_have_alpha, sc->U);
    if (sc->stage == 0 && prm.quality == 4 && sc->best.err) dxt1_combinatorial(sc, cfg);
    // pixel -> unique colour index, as in phase 0
    const bool opaque = lane < 16 && (!sc->pixels_have_alpha || (px >> 24) >= prm.alpha_threshold);
    const unsigned vmask = __ballot_sync(CRN_FULL_MASK, opaque);
    unsigned peers = 0;
    if (opaque) peers = __match_any_sync(vmask, px | 0xFF000000u);
    const bool leader = opaque && (unsigned)(__ffs((int)peers) - 1) == lane;
    const unsigned leaders = __ballot_sync(CRN_FULL_MASK, leader);
    const unsigned my_u = __popc(leaders & lanemask_lt());
    const unsigned uidx = __shfl_sync(CRN_FULL_MASK, my_u, opaque ? __ffs((int)peers) - 1 : 0);
    // ---- return_solution (:263-365)
    dxt1_best_selectors(sc, cfg);
    const bool invert = sc->best.alpha_block ? (sc->best.lo > sc->best.hi) : (sc->best.lo < sc->best.hi);
    const unsigned out_lo = invert ? sc->best.hi : sc->best.lo, out_hi = invert ? sc->best.lo : sc->best.hi;
    unsigned s = 3;
    if (opaque) {
        s = sc->sel[uidx];
        if (invert) s = sc->best.alpha_block ? (s < 2 ? s ^ 1 : s) : (s ^ 1);   // g_invTableAlpha {1,0,2,3} / g_invTableColor {1,0,3,2}
    }
    unsigned bits = lane < 16 ? (s << (2 * lane)) : 0u;
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) bits |= __shfl_xor_sync(CRN_FULL_MASK, bits, ofs);
    __syncwarp();
    return (unsigned long long)out_lo | ((unsigned long long)out_hi << 16) | ((unsigned long long)bits << 32);
}

}  // namespace crn
