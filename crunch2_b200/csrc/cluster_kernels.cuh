// cluster_kernels.cuh -- N-pixel ("cluster") form of the endpoint optimisers and the selector re-vote
// (SURVEY 8(a) rows a7, a9, a15/a20 core, a21): clusters given as CSR lists of member blocks over a
// [n_blocks][16] RGBA8 block array.  Colour: one warp per small cluster, one 8-warp CTA per large one
// (dxt1_optimize_clusters_cta_kernel); alpha: one CTA per cluster, candidates scored from prefix sums.
//
// Replaces the bodies of qdxt1::pack_endpoints_task / qdxt5::pack_endpoints_task (reference
// crnlib/crn_qdxt1.cpp:471-699, crnlib/crn_qdxt5.cpp:452-576: concatenate the member blocks' pixels,
// run dxt1_/dxt5_endpoint_optimizer over all of them, write the shared endpoints and the per-pixel
// selectors into every member block) and qdxt1::optimize_selectors_task (crn_qdxt1.cpp:714-865).
// The optimiser code is the same as for 4x4 blocks (dxt1_opt.cuh / dxt5a_opt.cuh, templated on the scratch
// type); what changes is where the unique colours live (global workspace instead of shared memory) and
// how they are found: a hash table built with atomics, then ordered by first appearance so that every
// order-sensitive float sum runs over the colours in exactly the reference's order.
#pragma once
#include "dxt1_opt.cuh"
#include "dxt5a_opt.cuh"

namespace crn {

// CTA-per-cluster mode (large clusters): the owning warp runs the optimiser as always; each time it evaluates a batch of 32 candidates it
// publishes them here and the CTA's other warps each take a slice of the unique colours (dxt1_eval_coop below).
#ifndef CRN_COOP_WARPS
#define CRN_COOP_WARPS 8
#endif
#ifndef CRN_COOP_OCC
#define CRN_COOP_OCC 3
#endif
constexpr int kClusterCoopWarps = CRN_COOP_WARPS;   // warps of a cooperative CTA
constexpr int kClusterCoopChunk = 32;           // unique colours per warp per slice
#ifndef CRN_COOP_SLICES
#define CRN_COOP_SLICES 2
#endif
constexpr int kClusterCoopSlices = CRN_COOP_SLICES;   // slices per warp per round; the early-out test runs once a round (every 8 x 32 x this colours)
constexpr uint32_t kClusterCoopMinBlocks = 128; // clusters with at least this many member blocks are optimised by a whole CTA
struct Dxt1CoopShared {
    unsigned lo[32], hi[32];                    // the batch: one candidate per lane of the owning warp
    unsigned long long bound;                   // error of the best so far (the early-out bound)
    const int4* ce;                             // evaluation colours of the cluster
    Dxt1Cfg cfg;
    unsigned valid_mask;
    int alt, cmd;                               // cmd: 1 = evaluate this batch, 0 = no more batches for the helper warps
    unsigned long long part[2][kClusterCoopWarps][2][32];   // per round (double buffered), warp, block type, candidate: partial error sums
    int4 cbuf[kClusterCoopWarps][kClusterCoopChunk];        // the colours of each warp's current slice
};

struct Dxt1ClusterScratch {            // per-warp shared memory; colour arrays live in the global workspace
    int4* cw; int4* ce; uint8_t* sel;
    Dxt1CoopShared* coop;              // non-null: this warp owns a cluster in CTA-per-cluster mode
    int4 cbuf[32];                     // one-warp evaluation: the 32 evaluation colours being scored (coalesced load -> broadcast reads)
    int4 pal[32][5];                   // one-warp evaluation, second phase: every candidate's palette (p0 p1 p2 p3 pm)
    Dxt1Best best;
    float mean[3], axis[3], low[3], high[3];
    int U, total_w, pixels_have_alpha, stage;
    uint16_t probe[2][32];
    uint16_t packed[64];
    uint32_t hist[64], cursor[64];     // eval-colour ordering: counts / write cursors per magnitude class
    uint32_t n_eval[32];               // per lane: candidates evaluated (dxt1_eval calls) and unique colours they range over (sum of U):
    unsigned long long n_cu[32];       // the algorithmic work of SURVEY 8(d), U * (11 P + 1) integer ops per evaluation
#ifdef CRN_B200_PHASE_CLOCKS
    unsigned long long phase_clk[12];
#endif
};
// found by dxt1_eval through argument-dependent lookup; the 4x4-block scratch type has no counters (generic no-op in dxt1_opt.cuh)
__device__ __forceinline__ void dxt1_count_eval(Dxt1ClusterScratch* sc, int U) { sc->n_eval[lane_id()]++; sc->n_cu[lane_id()] += (unsigned)U; }

__device__ __forceinline__ bool dxt1_is_coop(const Dxt1ClusterScratch*) { return true; }     // every evaluation of a cluster goes through dxt1_eval_coop below
#ifdef CRN_B200_PHASE_CLOCKS
// one-warp evaluation: every lane adds its own cycles ([10]) and 1 ([11]); [10] / 32 is then a lower bound of the warp's time in dxt1_eval
__device__ __forceinline__ long long dxt1_prof_begin(Dxt1ClusterScratch*) { return clock64(); }
__device__ __forceinline__ void dxt1_prof_end(Dxt1ClusterScratch* sc, long long t0)
{
    atomicAdd(&sc->phase_clk[10], (unsigned long long)(clock64() - t0));
    atomicAdd(&sc->phase_clk[11], 1ull);
}
#endif

// One batch of <= 32 candidates against the U evaluation colours, by all warps of the CTA: warp w sums colours [base + w * chunk, + chunk) of every
// round for each candidate, the partial sums meet in shared memory, and every warp adds them up in the same (warp) order -- integer sums, so the
// totals are the ones dxt1_eval_loop produces, and every warp sees the same totals and leaves the loop in the same round.  A candidate whose
// totals have reached the bound is dropped from the following rounds (same rule as dxt1_eval_loop, tested every 256 colours instead of every 8).
template <bool DO4, bool DO3>
__device__ __forceinline__ void dxt1_coop_rounds(Dxt1CoopShared* cs, unsigned w, int U, const int4* __restrict__ ce, const int4 p0, const int4 p1, const int4 p2,
                                                 const int4 p3, const int4 pm, bool valid, unsigned long long bound, unsigned long long& e4, unsigned long long& e3)
{
    const unsigned lane = lane_id();
    e4 = 0; e3 = 0;
    bool active = valid;
    int buf = 0;
    static_assert(kClusterCoopChunk == 32, "one colour per lane per slice");
    // Slice k of warp w = colours [(k W + w) 32, + 32); a round is kClusterCoopSlices slices per warp, then one exchange of partial sums.
    // The next slice is fetched (one coalesced 512-byte load per warp) while the current one is being scored out of shared memory.
    int4* cbuf = cs->cbuf[w];
    int k = 0;
    int4 nxt = make_int4(0, 0, 0, 0);
    if ((int)(w * 32 + lane) < U) nxt = ce[w * 32 + lane];
    do {                                                     // at least one round, so that the batch is not republished while a warp still reads it
        unsigned long long s4 = 0, s3 = 0;
#pragma unroll 1
        for (int sl = 0; sl < kClusterCoopSlices; sl++, k++) {
            const int i0 = (k * kClusterCoopWarps + (int)w) * 32;
            if (sl && i0 - (int)w * 32 >= U) { k += kClusterCoopSlices - sl; break; }    // (warp-uniform) nothing left for any warp in this round
            __syncwarp();
            cbuf[lane] = nxt;
            __syncwarp();
            {
                const int ni = ((k + 1) * kClusterCoopWarps + (int)w) * 32 + (int)lane;
                if (ni < U) nxt = ce[ni];
            }
            if (active) {
                const int cnt = min(U, i0 + 32) - i0;
#pragma unroll 4
                for (int j = 0; j < cnt; j++) {
                    const int4 c = cbuf[j];
                    const unsigned wt = (unsigned)c.w;
                    const int cx = c.x & 0xffff, cy = c.x >> 16, cz = c.y;
                    const int d01 = min(eval_dprime(cx, cy, cz, p0), eval_dprime(cx, cy, cz, p1));
                    if (DO4) {
                        const int d = min(d01, min(eval_dprime(cx, cy, cz, p2), eval_dprime(cx, cy, cz, p3)));
                        s4 += (unsigned long long)(unsigned)(d + c.z) * wt;
                    }
                    if (DO3) {
                        const int d = min(d01, eval_dprime(cx, cy, cz, pm));
                        s3 += (unsigned long long)(unsigned)(d + c.z) * wt;
                    }
                }
            }
        }
        __syncwarp();
        if (DO4) cs->part[buf][w][0][lane] = s4;
        if (DO3) cs->part[buf][w][1][lane] = s3;
        __syncthreads();
#pragma unroll
        for (int ww = 0; ww < kClusterCoopWarps; ww++) {
            if (DO4) e4 += cs->part[buf][ww][0][lane];
            if (DO3) e3 += cs->part[buf][ww][1][lane];
        }
        buf ^= 1;
        if (active && ((DO4 && DO3) ? (e4 >= bound && e3 >= bound) : (DO4 ? e4 >= bound : e3 >= bound))) active = false;
    } while (__any_sync(CRN_FULL_MASK, active) && k * kClusterCoopWarps * 32 < U);
}
// what every warp of the CTA does with a published batch; returns this lane's candidate's (err, alpha) as dxt1_eval does
__device__ __noinline__ void dxt1_coop_batch(Dxt1CoopShared* cs, unsigned w, unsigned long long& err, int& alpha)
{
    const unsigned lane = lane_id();
    const Dxt1Cfg cfg = cs->cfg;
    const unsigned lo = cs->lo[lane], hi = cs->hi[lane];
    const int alt = cs->alt;
    const bool valid = (cs->valid_mask >> lane) & 1u;
    const unsigned long long bound = cs->bound;
    const int4* ce = cs->ce;
    int r0, g0, b0, r1, g1, b1;
    unpack565(lo, true, r0, g0, b0);
    unpack565(hi, true, r1, g1, b1);
    const int4 p0 = eval_palette(cfg, r0, g0, b0), p1 = eval_palette(cfg, r1, g1, b1);
    const int4 p2 = eval_palette(cfg, (r0 * 2 + r1 + alt) / 3, (g0 * 2 + g1 + alt) / 3, (b0 * 2 + b1 + alt) / 3);
    const int4 p3 = eval_palette(cfg, (r1 * 2 + r0 + alt) / 3, (g1 * 2 + g0 + alt) / 3, (b1 * 2 + b0 + alt) / 3);
    const int4 pm = eval_palette(cfg, (r0 + r1 + alt) >> 1, (g0 + g1 + alt) >> 1, (b0 + b1 + alt) >> 1);
    unsigned long long e4, e3;
    if (cfg.do4 && cfg.do3) {
        dxt1_coop_rounds<true, true>(cs, w, cfg.U, ce, p0, p1, p2, p3, pm, valid, bound, e4, e3);
        alpha = e3 < e4; err = alpha ? e3 : e4;
    } else if (cfg.do4) {
        dxt1_coop_rounds<true, false>(cs, w, cfg.U, ce, p0, p1, p2, p3, pm, valid, bound, e4, e3);
        alpha = 0; err = e4;
    } else {
        dxt1_coop_rounds<false, true>(cs, w, cfg.U, ce, p0, p1, p2, p3, pm, valid, bound, e4, e3);
        alpha = 1; err = e3;
    }
    if (!valid) { err = ~0ull; alpha = 0; }
}

// dxt1_eval of the owning warp in CTA-per-cluster mode (found through argument-dependent lookup): publish the batch, meet the helper warps
// at the barrier, take part as warp 0.  All 32 lanes come through here together (dxt1_eval's `valid` argument).
// One-warp evaluation of a cluster's candidates (one per lane), in two phases.
//   1. candidate per lane over the first 32 evaluation colours (the heaviest ones, cluster_order_eval_colours), read through sc->cbuf: most
//      candidates of a batch pass the bound here and are dropped (same rule as dxt1_eval_loop, tested every 8 colours);
//   2. the survivors one after the other, COLOUR per lane: every lane scores its own colours of the remaining U - 32 against the survivor's
//      palette (broadcast from sc->pal) and the partial sums meet in a butterfly -- every 8 x 32 colours for the early-out, and at the end.
// A lane-private loop keeps the whole warp waiting for its slowest lane (13.7 of 32 lanes active on configs[1], profiles/r2n); with the roles
// swapped the cost of a batch is proportional to the number of survivors.  The sums are integers: totals equal dxt1_eval_loop's.
template <bool DO4, bool DO3>
__device__ __forceinline__ void dxt1_eval_loop_staged(Dxt1ClusterScratch* sc, int U, const int4 p0, const int4 p1, const int4 p2, const int4 p3, const int4 pm,
                                                      bool valid, unsigned long long bound, unsigned long long& e4, unsigned long long& e3)
{
    const unsigned lane = lane_id();
    const int4* __restrict__ ce = sc->ce;
    e4 = 0; e3 = 0;
    bool active = valid;
    sc->cbuf[lane] = (int)lane < U ? ce[lane] : make_int4(0, 0, 0, 0);
    sc->pal[lane][0] = p0; sc->pal[lane][1] = p1;
    if (DO4) { sc->pal[lane][2] = p2; sc->pal[lane][3] = p3; }
    if (DO3) sc->pal[lane][4] = pm;
    __syncwarp();
    if (active) {
        const int cnt = min(32, U);
        for (int j = 0; j < cnt;) {
            const int stop = min(cnt, j + 8);
#pragma unroll 2
            for (; j < stop; j++) {
                const int4 c = sc->cbuf[j];
                const unsigned wt = (unsigned)c.w;
                const int cx = c.x & 0xffff, cy = c.x >> 16, cz = c.y;
                const int d01 = min(eval_dprime(cx, cy, cz, p0), eval_dprime(cx, cy, cz, p1));
                if (DO4) {
                    const int d = min(d01, min(eval_dprime(cx, cy, cz, p2), eval_dprime(cx, cy, cz, p3)));
                    e4 += (unsigned long long)(unsigned)(d + c.z) * wt;
                }
                if (DO3) {
                    const int d = min(d01, eval_dprime(cx, cy, cz, pm));
                    e3 += (unsigned long long)(unsigned)(d + c.z) * wt;
                }
            }
            if ((DO4 && DO3) ? (e4 >= bound && e3 >= bound) : (DO4 ? e4 >= bound : e3 >= bound)) { active = false; break; }
        }
    }
    unsigned mask = __ballot_sync(CRN_FULL_MASK, active);
    if (U <= 32) mask = 0;
#ifdef CRN_B200_PHASE_CLOCKS
    { const unsigned vm = __ballot_sync(CRN_FULL_MASK, valid); if (lane == 0) { sc->phase_clk[8] += (unsigned)__popc(vm); } }
    if (lane == 0) { sc->phase_clk[10] += (unsigned)__popc(mask); sc->phase_clk[11] += 1; }
#endif
    while (mask) {
        const int s = __ffs((int)mask) - 1;
        mask &= mask - 1;
        const int4 q0 = sc->pal[s][0], q1 = sc->pal[s][1];
        int4 q2 = q0, q3 = q0, qm = q0;
        if (DO4) { q2 = sc->pal[s][2]; q3 = sc->pal[s][3]; }
        if (DO3) qm = sc->pal[s][4];
        const unsigned long long h4 = __shfl_sync(CRN_FULL_MASK, e4, s), h3 = __shfl_sync(CRN_FULL_MASK, e3, s);      // the survivor's sums so far
        unsigned long long t4 = 0, t3 = 0;                           // this lane's share of the rest; totals at the checks
        unsigned long long a4 = 0, a3 = 0;
        for (int base = 32; base < U; base += 32 * 8) {
            // the eight loads of a group go out together (clamped index + zero weight past the end: no predicate between them)
            int4 cbatch[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int i = base + 32 * k + (int)lane;
                cbatch[k] = ce[min(i, U - 1)];
                if (i >= U) cbatch[k].w = 0;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                {
                    const int4 c = cbatch[k];
                    const unsigned wt = (unsigned)c.w;
                    const int cx = c.x & 0xffff, cy = c.x >> 16, cz = c.y;
                    const int d01 = min(eval_dprime(cx, cy, cz, q0), eval_dprime(cx, cy, cz, q1));
                    if (DO4) {
                        const int d = min(d01, min(eval_dprime(cx, cy, cz, q2), eval_dprime(cx, cy, cz, q3)));
                        a4 += (unsigned long long)(unsigned)(d + c.z) * wt;
                    }
                    if (DO3) {
                        const int d = min(d01, eval_dprime(cx, cy, cz, qm));
                        a3 += (unsigned long long)(unsigned)(d + c.z) * wt;
                    }
                }
            }
            if (base + 32 * 8 < U) {                                 // early out: the totals so far already reach the bound
                if (DO4) t4 = h4 + warp_sum_u64(a4);
                if (DO3) t3 = h3 + warp_sum_u64(a3);
                if ((DO4 && DO3) ? (t4 >= bound && t3 >= bound) : (DO4 ? t4 >= bound : t3 >= bound)) break;
            }
        }
        if (DO4) t4 = warp_sum_u64(a4);
        if (DO3) t3 = warp_sum_u64(a3);
        if ((int)lane == s) { e4 += t4; e3 += t3; }
    }
    __syncwarp();
}

__device__ __noinline__ void dxt1_eval_warp(Dxt1ClusterScratch* sc, const Dxt1Cfg cfg, unsigned lo, unsigned hi, int alt, unsigned long long& err, int& alpha, bool valid)
{
    __syncwarp();
    if (valid) dxt1_count_eval(sc, cfg.U);
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
        dxt1_eval_loop_staged<true, true>(sc, cfg.U, p0, p1, p2, p3, pm, valid, bound, e4, e3);
        alpha = e3 < e4; err = alpha ? e3 : e4;
    } else if (cfg.do4) {
        dxt1_eval_loop_staged<true, false>(sc, cfg.U, p0, p1, p2, p3, pm, valid, bound, e4, e3);
        alpha = 0; err = e4;
    } else {
        dxt1_eval_loop_staged<false, true>(sc, cfg.U, p0, p1, p2, p3, pm, valid, bound, e4, e3);
        alpha = 1; err = e3;
    }
    if (!valid) { err = ~0ull; alpha = 0; }
}

__device__ __forceinline__ void dxt1_eval_coop(Dxt1ClusterScratch* sc, const Dxt1Cfg cfg, unsigned lo, unsigned hi, int alt,
                                               unsigned long long& err, int& alpha, bool valid)
{
    Dxt1CoopShared* cs = sc->coop;
    if (!cs) { dxt1_eval_warp(sc, cfg, lo, hi, alt, err, alpha, valid); return; }     // one warp per cluster
    const unsigned lane = lane_id();
    if (valid) dxt1_count_eval(sc, cfg.U);
    cs->lo[lane] = lo; cs->hi[lane] = hi;
    const unsigned m = __ballot_sync(CRN_FULL_MASK, valid);
    if (lane == 0) { cs->bound = sc->best.err; cs->ce = sc->ce; cs->cfg = cfg; cs->valid_mask = m; cs->alt = alt; cs->cmd = 1; }
#ifdef CRN_B200_PHASE_CLOCKS
    const long long t0 = clock64();
#endif
    __syncthreads();
    dxt1_coop_batch(cs, 0, err, alpha);
#ifdef CRN_B200_PHASE_CLOCKS
    if (lane == 0) { sc->phase_clk[8] += (unsigned long long)(clock64() - t0); sc->phase_clk[9]++; }
#endif
}

struct ClusterHashEntry { uint32_t key, first_inv, count, uidx; };   // first_inv = ~(index of first appearance)

struct ClusterWorkspace {              // global scratch, all sized by the number of member pixels P
    ClusterHashEntry* hash;            // 2 entries per pixel
    uint32_t* mark;                    // 1 per pixel
    int4* cw; int4* ce;                // 1 per pixel each
    int4* ce2;                         // 1 per pixel: the evaluation colours again, heaviest contributors first
    uint8_t* sel;                      // 1 per pixel
};
// + first-appearance flags and their scan (4 + 4 per pixel)
constexpr size_t kClusterWorkspaceBytesPerPixel = 2 * sizeof(ClusterHashEntry) + 4 + 16 + 16 + 16 + 1 + 8;

constexpr int kClusterWarpsPerCta = 4;

__device__ __forceinline__ uint32_t cluster_pixel(const uint32_t* __restrict__ blocks, const uint32_t* __restrict__ members, uint32_t i)
{
    return blocks[(size_t)members[i >> 4] * 16 + (i & 15)];
}

// cluster that owns member position m: last c with offsets[c] <= m
__device__ __forceinline__ uint32_t cluster_of_member(const uint32_t* __restrict__ offsets, uint32_t n_clusters, uint32_t m)
{
    uint32_t lo = 0, hi = n_clusters;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (offsets[mid] <= m) lo = mid; else hi = mid; }
    return lo;
}

// The unique-colour extraction (crn_dxt1.cpp:2113-2131) is data parallel over ALL member pixels of ALL clusters:
//   insert  (thread per pixel)      hash table of the pixel's cluster: key, earliest index, count
//   mark    (thread per table slot) mark[first index] = slot + 1
//   scan    (vq_scan_* kernels)     rank of every first appearance
//   compact (thread per pixel)      unique colour u of the cluster = colour whose first appearance has rank u
// so that the per-cluster warp below only runs the optimiser over the U unique colours.
__global__ void __launch_bounds__(256)
cluster_hash_insert_kernel(const uint32_t* __restrict__ blocks, const uint32_t* __restrict__ cluster_offsets, const uint32_t* __restrict__ cluster_blocks,
                           uint32_t n_clusters, uint32_t total_pixels, int dxt1a, uint32_t alpha_threshold, ClusterWorkspace ws, uint32_t* __restrict__ transparent)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_pixels) return;
    const uint32_t m = g >> 4, c = cluster_of_member(cluster_offsets, n_clusters, m);
    const uint32_t P = cluster_offsets[c] * 16, N = (cluster_offsets[c + 1] - cluster_offsets[c]) * 16, i = g - P;
    const uint32_t px = blocks[(size_t)cluster_blocks[m] * 16 + (g & 15)];
    if (dxt1a && (px >> 24) < alpha_threshold) { atomicAdd(&transparent[c], 1u); return; }     // crn_qdxt1.cpp:625-656
    ClusterHashEntry* tab = ws.hash + 2 * (size_t)P;
    const uint32_t cap = 2 * N, key = px | 0xFF000000u;
    uint32_t h = (key * 2654435761u) % cap;
    for (;;) {
        const uint32_t old = atomicCAS(&tab[h].key, 0u, key);
        if (old == 0u || old == key) break;
        h = h + 1 == cap ? 0 : h + 1;
    }
    atomicMax(&tab[h].first_inv, ~i);
    atomicAdd(&tab[h].count, 1u);
}

__global__ void __launch_bounds__(256)
cluster_mark_kernel(const uint32_t* __restrict__ cluster_offsets, uint32_t n_clusters, uint32_t total_pixels, ClusterWorkspace ws)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= 2 * total_pixels) return;
    const ClusterHashEntry e = ws.hash[s];
    if (!e.key) return;
    const uint32_t c = cluster_of_member(cluster_offsets, n_clusters, s >> 5), P = cluster_offsets[c] * 16;
    ws.mark[P + ~e.first_inv] = s - 2 * P + 1;
}

__global__ void __launch_bounds__(256)
cluster_first_flags_kernel(const uint32_t* __restrict__ mark, uint32_t* __restrict__ flags, uint32_t total_pixels)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g <= total_pixels) flags[g] = (g < total_pixels && mark[g]) ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
cluster_compact_kernel(const uint32_t* __restrict__ cluster_offsets, uint32_t n_clusters, uint32_t total_pixels, ClusterWorkspace ws, const uint32_t* __restrict__ rank)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_pixels) return;
    const uint32_t v = ws.mark[g];
    if (!v) return;
    const uint32_t c = cluster_of_member(cluster_offsets, n_clusters, g >> 4), P = cluster_offsets[c] * 16;
    const uint32_t u = rank[g] - rank[P];
    ClusterHashEntry& e = ws.hash[2 * (size_t)P + v - 1];
    e.uidx = u;
    ws.cw[P + u] = make_int4((int)(e.key & 0xff), (int)((e.key >> 8) & 0xff), (int)((e.key >> 16) & 0xff), (int)e.count);
}

// Evaluation order of the unique colours.  A candidate's error is an integer sum over the colours, so the order does not
// change any result; what it changes is how soon a poor candidate's partial sum passes the best error and the lane stops
// (dxt1_eval_loop).  The reference sorts m_evaluated_colors by weighted projection for the same reason
// (crn_dxt1.cpp:798-816).  Here: one counting pass over 64 magnitude classes of weight * squared distance to the cluster's
// mean colour, heaviest class first -- O(U), against the ~10^4 evaluations that follow.
__device__ __forceinline__ void cluster_order_eval_colours(Dxt1ClusterScratch* sc, const Dxt1Cfg cfg, int4* __restrict__ dst)
{
    const unsigned lane = lane_id();
    const int U = cfg.U;
    if (U <= 64) return;
    long long sw = 0, sr = 0, sg = 0, sb = 0;
    for (int i = (int)lane; i < U; i += 32) { const int4 c = sc->cw[i]; sw += c.w; sr += (long long)c.x * c.w; sg += (long long)c.y * c.w; sb += (long long)c.z * c.w; }
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) {
        sw += __shfl_xor_sync(CRN_FULL_MASK, sw, ofs); sr += __shfl_xor_sync(CRN_FULL_MASK, sr, ofs);
        sg += __shfl_xor_sync(CRN_FULL_MASK, sg, ofs); sb += __shfl_xor_sync(CRN_FULL_MASK, sb, ofs);
    }
    const int mr = (int)(sr / sw), mg = (int)(sg / sw), mb = (int)(sb / sw);
    const int wr = cfg.gray ? 1 : cfg.wr, wg = cfg.gray ? 1 : cfg.wg, wb = cfg.gray ? 1 : cfg.wb;
    for (unsigned k = lane; k < 64; k += 32) { sc->hist[k] = 0; sc->cursor[k] = 0; }
    __syncwarp();
    auto cls_of = [&](const int4 c) -> unsigned {
        const int dr = c.x - mr, dg = c.y - mg, db = c.z - mb;
        const unsigned long long key = (unsigned long long)(unsigned)(wr * dr * dr + wg * dg * dg + wb * db * db) * (unsigned)c.w;
        return key ? 63u - (unsigned)__clzll((long long)key) : 0u;
    };
    for (int i = (int)lane; i < U; i += 32) atomicAdd(&sc->hist[cls_of(sc->cw[i])], 1u);
    __syncwarp();
    if (lane == 0) { unsigned run = 0; for (int k = 63; k >= 0; k--) { const unsigned h = sc->hist[k]; sc->hist[k] = run; run += h; } }   // hist -> start of class, heaviest first
    __syncwarp();
    for (int i = (int)lane; i < U; i += 32) {
        const unsigned k = cls_of(sc->cw[i]);
        dst[sc->hist[k] + atomicAdd(&sc->cursor[k], 1u)] = sc->ce[i];
    }
    __syncwarp();
    if (lane == 0) sc->ce = dst;
    __syncwarp();
}

struct ClusterResult { uint32_t endpoints; uint32_t flags; };      // flags: bit 0 invert, bit 1 alpha_block, bits 2-3 stage

struct ClusterOptArgs {
    const uint32_t* cluster_offsets; Dxt1Params prm; int dxt1a; ClusterWorkspace ws; const uint32_t* rank; const uint32_t* transparent;
    ClusterResult* results; uint32_t* out_endpoints; unsigned long long* out_error; uint32_t* out_flags;
};

// Profiling build only (-DCRN_B200_PHASE_CLOCKS, tools/prof_cluster_phases.sh): SM cycles of the owning warp per optimiser phase, summed over
// the clusters into the 64-bit words at next_cluster + 20 + 2 * phase.
#ifdef CRN_B200_PHASE_CLOCKS
#define CRN_PHASE_BEGIN() long long ph_t0_ = clock64()
#define CRN_PHASE_END(k) do { __syncwarp(); const long long t_ = clock64(); if (lane_id() == 0) sc->phase_clk[k] += (unsigned long long)(t_ - ph_t0_); ph_t0_ = t_; } while (0)
#else
#define CRN_PHASE_BEGIN() do { } while (0)
#define CRN_PHASE_END(k) do { } while (0)
#endif

// one cluster, by the calling warp (with the CTA's other warps behind dxt1_eval when sc->coop is set)
__device__ __forceinline__ void dxt1_optimize_one_cluster(Dxt1ClusterScratch* sc, const ClusterOptArgs& a, uint32_t c)
{
    const unsigned lane = lane_id();
    const Dxt1Params prm = a.prm;
    const uint32_t b0 = a.cluster_offsets[c], nb = a.cluster_offsets[c + 1] - b0;
    const uint32_t N = nb * 16, P = b0 * 16;
    if (!nb) return;
    const uint32_t ntransp = a.dxt1a ? a.transparent[c] : 0u;
    const int pha = ntransp != 0;                       // pixels_have_alpha as qdxt1 computes it
    const unsigned opaque_cnt = N - ntransp;
    const int U = (int)(a.rank[P + N] - a.rank[P]);
    if (lane == 0) { sc->cw = a.ws.cw + P; sc->ce = a.ws.ce + P; sc->sel = a.ws.sel + P; }
    __syncwarp();
    // ---- the optimiser proper: same phases as the 4x4 block kernels, fused
    CRN_PHASE_BEGIN();
    dxt1_build_eval_colours(sc, dxt1_make_cfg(prm, pha, U));
    cluster_order_eval_colours(sc, dxt1_make_cfg(prm, pha, U), a.ws.ce2 + P);
    CRN_PHASE_END(0);
    dxt1_setup_common(sc, prm, pha, U, opaque_cnt, opaque_cnt != N);
    CRN_PHASE_END(1);
    dxt1_phase_median4(sc, prm);
    CRN_PHASE_END(2);
    dxt1_phase_passes(sc, prm);
    CRN_PHASE_END(3);
    dxt1_phase_post(sc, prm);
    CRN_PHASE_END(4);
    // ---- finish: combinatorial recovery + return_solution (crn_dxt1.cpp:1048-1056, :263-365)
    unsigned out_lo = 0, out_hi = 0;
    bool invert = false;
    int alpha_block = 1;
    const int stage = sc->stage;
    if (stage != 2) {
        const Dxt1Cfg cfg = dxt1_make_cfg(prm, pha, U);
        if (stage == 0 && prm.quality == 4 && sc->best.err) dxt1_combinatorial(sc, cfg);
        CRN_PHASE_END(5);
        dxt1_best_selectors(sc, cfg);
        CRN_PHASE_END(6);
        alpha_block = sc->best.alpha_block;
        invert = alpha_block ? (sc->best.lo > sc->best.hi) : (sc->best.lo < sc->best.hi);
        out_lo = invert ? sc->best.hi : sc->best.lo; out_hi = invert ? sc->best.lo : sc->best.hi;
    }
    if (lane == 0) {
        a.results[c].endpoints = out_lo | (out_hi << 16);
        a.results[c].flags = (invert ? 1u : 0u) | (alpha_block ? 2u : 0u) | ((unsigned)stage << 2);
        if (a.out_endpoints) a.out_endpoints[c] = out_lo | (out_hi << 16);
        if (a.out_error) a.out_error[c] = stage == 2 ? 0ull : sc->best.err;
        // dxt_hc wants results::m_reordered (bit 0) and m_alternate_rounding (bit 4) (crn_dxt1.cpp:279-282)
        if (a.out_flags) a.out_flags[c] = (invert ? 1u : 0u) | (alpha_block ? 2u : 0u) | ((unsigned)stage << 2) | ((stage != 2 && sc->best.alt_round) ? 16u : 0u);
    }
    __syncwarp();
}

// work counters of a warp -> the two 64-bit words behind the work-stealing counter (next_cluster + 16 / + 18)
__device__ __forceinline__ void cluster_flush_counters(Dxt1ClusterScratch* sc, unsigned int* next_cluster)
{
    const unsigned lane = lane_id();
    unsigned long long ne = sc->n_eval[lane], nc = sc->n_cu[lane];
    ne = warp_sum_u64(ne); nc = warp_sum_u64(nc);
    if (lane == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(next_cluster + 16), ne);
        atomicAdd(reinterpret_cast<unsigned long long*>(next_cluster + 18), nc);
#ifdef CRN_B200_PHASE_CLOCKS
        for (int k = 0; k < 12; k++) atomicAdd(reinterpret_cast<unsigned long long*>(next_cluster + 20 + 2 * k), sc->phase_clk[k]);
#endif
    }
}

// the work-stealing loop of one warp over order[first .. n_clusters) (order == nullptr: the clusters in index order)
__device__ __forceinline__ void dxt1_cluster_warp_loop(Dxt1ClusterScratch* sc, const ClusterOptArgs& a, uint32_t first, uint32_t n_clusters,
                                                       unsigned int* next_cluster, const uint32_t* __restrict__ order)
{
    const unsigned lane = lane_id();
    for (;;) {
        uint32_t c = 0;
        if (lane == 0) c = first + atomicAdd(next_cluster, 1u);
        c = __shfl_sync(CRN_FULL_MASK, c, 0);
        if (c >= n_clusters) break;
        if (order) c = order[c];                             // largest clusters first: the work-stealing tail is one small cluster, not one huge one
        dxt1_optimize_one_cluster(sc, a, c);
    }
}

__global__ void __launch_bounds__(kClusterWarpsPerCta * 32, 5)
dxt1_optimize_clusters_kernel(const uint32_t* __restrict__ cluster_offsets, uint32_t n_clusters, Dxt1Params prm, int dxt1a,
                              ClusterWorkspace ws, const uint32_t* __restrict__ rank, const uint32_t* __restrict__ transparent,
                              unsigned int* __restrict__ next_cluster, ClusterResult* __restrict__ results,
                              uint32_t* __restrict__ out_endpoints, unsigned long long* __restrict__ out_error, uint32_t* __restrict__ out_flags,
                              const uint32_t* __restrict__ order)
{
    __shared__ Dxt1ClusterScratch scratch[kClusterWarpsPerCta];
    const unsigned warp = threadIdx.x >> 5, lane = lane_id();
    Dxt1ClusterScratch* sc = &scratch[warp];
    sc->n_eval[lane] = 0; sc->n_cu[lane] = 0;
#ifdef CRN_B200_PHASE_CLOCKS
    if (lane < 12) sc->phase_clk[lane] = 0;
#endif
    if (lane == 0) sc->coop = nullptr;
    __syncwarp();
    const ClusterOptArgs a = { cluster_offsets, prm, dxt1a, ws, rank, transparent, results, out_endpoints, out_error, out_flags };
    dxt1_cluster_warp_loop(sc, a, 0, n_clusters, next_cluster, order);
    cluster_flush_counters(sc, next_cluster);
}

// Large clusters first, a CTA each; then the same CTAs turn into kClusterCoopWarps independent warps for the small ones.
// `order` lists the clusters by descending size, its first n_big entries are the large ones.  One warp per cluster leaves a dxt_hc pass at low
// quality (a few hundred clusters of thousands of blocks) waiting for the single warp that owns the largest cluster; here that cluster's
// colour loop -- >90 % of the optimiser's instructions at that size -- runs on eight warps.  Work-stealing counters: next_cluster[0] for the
// large clusters (one fetch per CTA), next_cluster[1] for the rest.
__global__ void __launch_bounds__(kClusterCoopWarps * 32, CRN_COOP_OCC)
dxt1_optimize_clusters_cta_kernel(const uint32_t* __restrict__ cluster_offsets, uint32_t n_clusters, uint32_t n_big, Dxt1Params prm, int dxt1a,
                                  ClusterWorkspace ws, const uint32_t* __restrict__ rank, const uint32_t* __restrict__ transparent,
                                  unsigned int* __restrict__ next_cluster, ClusterResult* __restrict__ results,
                                  uint32_t* __restrict__ out_endpoints, unsigned long long* __restrict__ out_error, uint32_t* __restrict__ out_flags,
                                  const uint32_t* __restrict__ order)
{
    __shared__ Dxt1ClusterScratch scratch[kClusterCoopWarps];
    __shared__ Dxt1CoopShared coop;
    const unsigned warp = threadIdx.x >> 5, lane = lane_id();
    Dxt1ClusterScratch* sc = &scratch[warp];
    sc->n_eval[lane] = 0; sc->n_cu[lane] = 0;
#ifdef CRN_B200_PHASE_CLOCKS
    if (lane < 12) sc->phase_clk[lane] = 0;
#endif
    if (lane == 0) sc->coop = nullptr;
    __syncwarp();
    const ClusterOptArgs a = { cluster_offsets, prm, dxt1a, ws, rank, transparent, results, out_endpoints, out_error, out_flags };
    if (warp == 0) {
        if (lane == 0) sc->coop = &coop;
        __syncwarp();
        for (;;) {
            uint32_t c = 0;
            if (lane == 0) c = atomicAdd(next_cluster, 1u);
            c = __shfl_sync(CRN_FULL_MASK, c, 0);
            if (c >= n_big) break;
            dxt1_optimize_one_cluster(sc, a, order[c]);
        }
        if (lane == 0) { coop.cmd = 0; sc->coop = nullptr; }
        __syncwarp();
        __syncthreads();                                     // releases the helpers
    } else {
        for (;;) {
            __syncthreads();                                 // the owning warp has published a batch, or is done
            if (!coop.cmd) break;
            unsigned long long e; int al;
            dxt1_coop_batch(&coop, warp, e, al);
        }
    }
    dxt1_cluster_warp_loop(sc, a, n_big, n_clusters, next_cluster + 1, order);
    cluster_flush_counters(sc, next_cluster);
}

// selectors of every member pixel from its unique colour's selector; 16 lanes per member block
__global__ void __launch_bounds__(256)
cluster_write_kernel(const uint32_t* __restrict__ blocks, const uint32_t* __restrict__ cluster_offsets, const uint32_t* __restrict__ cluster_blocks,
                     uint32_t n_clusters, uint32_t total_pixels, int dxt1a, uint32_t alpha_threshold, ClusterWorkspace ws, const uint32_t* __restrict__ transparent,
                     const ClusterResult* __restrict__ results, uint8_t* __restrict__ out, uint32_t out_stride, uint32_t out_ofs)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = lane_id();
    unsigned s = 3;
    uint32_t endpoints = 0, block = 0;
    if (g < total_pixels) {
        const uint32_t m = g >> 4, c = cluster_of_member(cluster_offsets, n_clusters, m);
        const uint32_t P = cluster_offsets[c] * 16, N = (cluster_offsets[c + 1] - cluster_offsets[c]) * 16;
        const ClusterResult r = results[c];
        const int stage = (int)(r.flags >> 2), invert = r.flags & 1, alpha_block = (r.flags >> 1) & 1;
        const int pha = dxt1a && transparent[c] != 0;
        endpoints = r.endpoints; block = cluster_blocks[m];
        if (stage != 2) {
            const uint32_t px = blocks[(size_t)block * 16 + (g & 15)];
            if (!(pha && (px >> 24) < alpha_threshold)) {
                const ClusterHashEntry* tab = ws.hash + 2 * (size_t)P;
                const uint32_t cap = 2 * N, key = px | 0xFF000000u;
                uint32_t h = (key * 2654435761u) % cap;
                while (tab[h].key != key) h = h + 1 == cap ? 0 : h + 1;
                s = ws.sel[P + tab[h].uidx];
                if (invert) s = alpha_block ? (s < 2 ? s ^ 1 : s) : (s ^ 1);
            }
        }
    }
    unsigned bits = s << (2 * (lane & 15));
#pragma unroll
    for (int ofs = 8; ofs > 0; ofs >>= 1) bits |= __shfl_xor_sync(CRN_FULL_MASK, bits, ofs);
    if ((lane & 15) == 0 && g < total_pixels)
        *reinterpret_cast<unsigned long long*>(out + (size_t)block * out_stride + out_ofs) = (unsigned long long)endpoints | ((unsigned long long)bits << 32);
}

// ---- DXT5A clusters (qdxt5::pack_endpoints_task, crn_qdxt5.cpp:452-576 -> dxt5_endpoint_optimizer) -----
struct Dxt5aClusterScratch : Dxt5aScratch {
    uint32_t first[256];               // ~index of first appearance per 8-bit value, 0 = absent
    uint32_t count[256];
    uint8_t uidx_of_value[256];
};

// Candidate error from prefix sums over the 8-bit value axis.  A DXT5A palette is a set of points on a line, so the value that takes palette
// entry e_m is decided by the midpoints to its sorted neighbours, and the error over such an interval is S2 - 2 e S1 + e^2 S0 with
// S0 / S1 / S2 = sums of w, w v, w v^2 over the interval -- the same integer the reference reaches by trying all 8 entries for every unique
// value (crn_dxt5a.cpp:198-262; a tie between two entries contributes the same d^2 either way), in O(16) instead of O(16 U) per candidate.
// Not valid where the reference's 32-bit product d*d*weight wraps: the kernel checks the largest weight and keeps dxt5a_eval<true> there.
struct Dxt5aPrefix {
    unsigned long long s1[257], s2[257];   // [x] = sum over values v < x
    uint32_t s0[257];
};
__device__ __forceinline__ unsigned long long dxt5a_interval_err(const Dxt5aPrefix* pf, unsigned e, int t_prev, int t)
{   // values in (t_prev, t] against palette entry e
    const unsigned long long c = pf->s0[t + 1] - pf->s0[t_prev + 1], a = pf->s1[t + 1] - pf->s1[t_prev + 1], b = pf->s2[t + 1] - pf->s2[t_prev + 1];
    return b - 2ull * e * a + (unsigned long long)(e * e) * c;
}
__device__ __forceinline__ void dxt5a_eval_prefix(const Dxt5aPrefix* pf, unsigned l, unsigned h, bool both, unsigned long long& err, unsigned& type)
{
    const unsigned lo = min(l, h), hi = max(l, h);       // both palettes are symmetric in l <-> h as sets (crn_dxt.cpp:404-430)
    unsigned p[8];
    dxt5a_values8(lo, hi, p);                            // ascending: p0, p2 .. p7, p1
    unsigned long long e8 = 0;
    {
        int tp = -1;
        unsigned cur = p[0];
#pragma unroll
        for (int m = 2; m <= 8; m++) {
            const unsigned nxt = m == 8 ? p[1] : p[m];
            const int t = (int)((cur + nxt) >> 1);
            e8 += dxt5a_interval_err(pf, cur, tp, t);
            tp = t; cur = nxt;
        }
        e8 += dxt5a_interval_err(pf, cur, tp, 255);
    }
    err = e8; type = 0;
    if (both) {
        dxt5a_values6(lo, hi, p);                        // ascending: 0, p0, p2 .. p5, p1, 255
        unsigned long long e6 = 0;
        int tp = -1;
        unsigned cur = 0;
#pragma unroll
        for (int m = 0; m < 7; m++) {
            const unsigned nxt = m == 0 ? p[0] : (m <= 4 ? p[m + 1] : (m == 5 ? p[1] : 255u));
            const int t = (int)((cur + nxt) >> 1);
            e6 += dxt5a_interval_err(pf, cur, tp, t);
            tp = t; cur = nxt;
        }
        e6 += dxt5a_interval_err(pf, cur, tp, 255);
        if (e6 < e8) { err = e6; type = 1; }
    }
}

constexpr int kAlphaClusterWarps = 8;
struct Dxt5aCtaScratch {
    Dxt5aClusterScratch c;
    Dxt5aPrefix pf;
    unsigned long long red_err[kAlphaClusterWarps];
    uint32_t red_k[kAlphaClusterWarps], red_l[kAlphaClusterWarps], red_h[kAlphaClusterWarps], red_t[kAlphaClusterWarps];
    uint32_t cluster, max_weight, n_unique;
};

// pair number k (i-major over i < j < U) -> (i, j)
__device__ __forceinline__ void dxt5a_pair_of(unsigned k, int U, int& i, int& j)
{
    // row i starts at off(i) = i (2U - i - 1) / 2; first guess from the quadratic, then corrected
    const float fu = (float)(2 * U - 1);
    int r = (int)((fu - sqrtf(fmaxf(fu * fu - 8.0f * (float)k, 0.0f))) * 0.5f);
    r = max(0, min(r, U - 2));
    while (r > 0 && (unsigned)(r * (2 * U - r - 1) / 2) > k) r--;
    while ((unsigned)((r + 1) * (2 * U - r - 2) / 2) <= k) r++;
    i = r; j = r + 1 + (int)(k - (unsigned)(r * (2 * U - r - 1) / 2));
}

// One CTA per cluster (qdxt5::pack_endpoints_task, crn_qdxt5.cpp:452-576 -> dxt5_endpoint_optimizer::compute): the O(N) passes (histogram of the
// member pixels, selector write-back) and the all-pairs phase of the search run on all 256 threads; candidates are scored from prefix sums.
// Results equal the one-warp kernel's (and the reference's): the all-pairs winner is the first minimum in pair order, reduced over
// (error, pair number); the probe window around the live best stays on one warp, 32 speculative candidates at a time.
__global__ void __launch_bounds__(kAlphaClusterWarps * 32)
dxt5_optimize_clusters_kernel(const uint32_t* __restrict__ blocks, const uint32_t* __restrict__ cluster_offsets,
                              const uint32_t* __restrict__ cluster_blocks, uint32_t n_clusters, uint32_t comp, int quality, int both_types,
                              unsigned int* __restrict__ next_cluster, uint8_t* __restrict__ out, uint32_t out_stride, uint32_t out_ofs,
                              uint32_t* __restrict__ out_endpoints, unsigned long long* __restrict__ out_error, uint32_t* __restrict__ out_flags,
                              const uint32_t* __restrict__ order)
{
    __shared__ Dxt5aCtaScratch S;
    Dxt5aClusterScratch* sc = &S.c;
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = lane_id();
    constexpr unsigned T = kAlphaClusterWarps * 32;
    const bool both = both_types != 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) S.cluster = atomicAdd(next_cluster, 1u);
        __syncthreads();
        uint32_t c = S.cluster;
        if (c >= n_clusters) break;
        if (order) c = order[c];
        const uint32_t b0 = cluster_offsets[c], nb = cluster_offsets[c + 1] - b0;
        const uint32_t* members = cluster_blocks + b0;
        const uint32_t N = nb * 16;
        if (!nb) continue;
        sc->first[tid] = 0; sc->count[tid] = 0;
        if (tid == 0) { S.max_weight = 0; S.n_unique = 0; }
        __syncthreads();
        for (uint32_t i = tid; i < N; i += T) {
            const uint32_t a = (cluster_pixel(blocks, members, i) >> (8 * comp)) & 0xffu;
            atomicMax(&sc->first[a], ~i);
            atomicAdd(&sc->count[a], 1u);
        }
        __syncthreads();
        // unique values ordered by first appearance (crn_dxt5a.cpp:58-75): rank = number of present values seen earlier
        {
            const uint32_t f = sc->first[tid];
            if (f) {
                int rank = 0;
                for (uint32_t o = 0; o < 256; o++) rank += sc->first[o] > f;     // larger ~index == earlier
                sc->val[rank] = (uint8_t)tid; sc->wgt[rank] = sc->count[tid]; sc->uidx_of_value[tid] = (uint8_t)rank;
                atomicMax(&S.max_weight, sc->count[tid]);
                atomicAdd(&S.n_unique, 1u);
            }
        }
        __syncthreads();
        const int U = (int)S.n_unique;
        // prefix sums over the value axis (warp 0: 8 values per lane, then a scan over the lanes)
        if (warp == 0) {
            unsigned long long a0 = 0, a1 = 0, a2 = 0;
            uint32_t c0[8]; unsigned long long c1[8], c2[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const unsigned v = lane * 8 + q;
                const unsigned long long w = sc->count[v];
                c0[q] = (uint32_t)a0; c1[q] = a1; c2[q] = a2;
                a0 += w; a1 += w * v; a2 += w * v * v;
            }
            unsigned long long x0 = a0, x1 = a1, x2 = a2;
#pragma unroll
            for (int ofs = 1; ofs < 32; ofs <<= 1) {
                const unsigned long long y0 = __shfl_up_sync(CRN_FULL_MASK, x0, ofs), y1 = __shfl_up_sync(CRN_FULL_MASK, x1, ofs), y2 = __shfl_up_sync(CRN_FULL_MASK, x2, ofs);
                if (lane >= (unsigned)ofs) { x0 += y0; x1 += y1; x2 += y2; }
            }
            const unsigned long long base0 = x0 - a0, base1 = x1 - a1, base2 = x2 - a2;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const unsigned v = lane * 8 + q;
                S.pf.s0[v] = (uint32_t)base0 + c0[q]; S.pf.s1[v] = base1 + c1[q]; S.pf.s2[v] = base2 + c2[q];
            }
            if (lane == 31) { S.pf.s0[256] = (uint32_t)x0; S.pf.s1[256] = x1; S.pf.s2[256] = x2; }
        }
        __syncthreads();
        // d*d*weight stays below 2^31 for every value: the prefix-sum evaluator equals the reference's sum (255^2 * 33025 < 2^31)
        const bool exact_prefix = S.max_weight <= 33025u;
        unsigned first = 0, second = 0;
        unsigned long long err = 0;
        unsigned reordered = 0;                              // results::m_reordered (crn_dxt5a.cpp:150-182)
        if (U == 1) {
            first = second = sc->val[0];
            if (tid == 0) sc->sel[0] = 0;
        } else {
            // ---- phase 1: every pair (i<j) of unique values, i-major (crn_dxt5a.cpp:93-103), over all threads
            unsigned long long my_err = ~0ull;
            unsigned my_k = 0xffffffffu, my_type = 0, my_l = 0, my_h = 0;
            const unsigned npairs = (unsigned)(U * (U - 1) / 2);
            for (unsigned k = tid; k < npairs; k += T) {
                int i, j;
                dxt5a_pair_of(k, U, i, j);
                unsigned long long e; unsigned ty;
                const unsigned l = sc->val[i], h = sc->val[j];
                if (exact_prefix) dxt5a_eval_prefix(&S.pf, l, h, both, e, ty);
                else dxt5a_eval<true>(sc, U, l, h, both, e, ty);
                if (e < my_err) { my_err = e; my_k = k; my_type = ty; my_l = l; my_h = h; }
            }
            {
                unsigned long long key = my_err; unsigned idx = my_k;
                warp_argmin_u64(key, idx);                   // (error, pair number): the first minimum in sequence order
                const unsigned owner = __ballot_sync(CRN_FULL_MASK, my_err == key && my_k == idx);
                const int src = owner ? __ffs((int)owner) - 1 : 0;
                const unsigned wl = __shfl_sync(CRN_FULL_MASK, my_l, src), wh = __shfl_sync(CRN_FULL_MASK, my_h, src), wt = __shfl_sync(CRN_FULL_MASK, my_type, src);
                if (lane == 0) { S.red_err[warp] = key; S.red_k[warp] = idx; S.red_l[warp] = wl; S.red_h[warp] = wh; S.red_t[warp] = wt; }
            }
            __syncthreads();
            if (warp == 0) {
                Dxt5aBest best;
                int bw = 0;
                for (int w = 1; w < kAlphaClusterWarps; w++)
                    if (S.red_err[w] < S.red_err[bw] || (S.red_err[w] == S.red_err[bw] && S.red_k[w] < S.red_k[bw])) bw = w;
                best.error = S.red_err[bw]; best.first = S.red_l[bw]; best.second = S.red_h[bw]; best.block_type = S.red_t[bw];
                // ---- phase 2: probe window around the live best (crn_dxt5a.cpp:105-149), one warp
                if (quality >= 3 && best.error) {
                    const int P = (quality == 4) ? 16 : 8;
                    const int W = 2 * P + 1;
                    int k0 = 0, row_ld = -1000, row_l = 0;
                    while (k0 < W * W && best.error) {
                        const int ld0 = k0 / W - P;
                        if (ld0 != row_ld) { row_ld = ld0; row_l = (int)best.first + ld0; }
                        if (row_l > 255) break;                                   // :119-122
                        const int k = k0 + (int)lane;
                        const int ld = k / W - P, hd = k % W - P;
                        const int l = (ld == row_ld) ? row_l : (int)best.first + ld;
                        const int h = (int)best.second + hd;
                        const bool valid = (k < W * W) && l >= 0 && l <= 255 && h >= 0 && h <= 255;
                        unsigned long long e = ~0ull; unsigned ty = 0;
                        if (valid) {
                            if (exact_prefix) dxt5a_eval_prefix(&S.pf, (unsigned)l, (unsigned)h, both, e, ty);
                            else dxt5a_eval<true>(sc, U, (unsigned)l, (unsigned)h, both, e, ty);
                        }
                        const unsigned m = __ballot_sync(CRN_FULL_MASK, valid && e < best.error);
                        if (!m) { k0 += 32; continue; }
                        const int t = __ffs((int)m) - 1;
                        best.error = __shfl_sync(CRN_FULL_MASK, e, t);
                        best.block_type = __shfl_sync(CRN_FULL_MASK, ty, t);
                        const int wl = __shfl_sync(CRN_FULL_MASK, l, t), wh = __shfl_sync(CRN_FULL_MASK, h, t);
                        const int wld = __shfl_sync(CRN_FULL_MASK, ld, t);
                        best.first = (unsigned)wl; best.second = (unsigned)wh;
                        row_ld = wld; row_l = wl;                                  // the rest of this row keeps its l
                        k0 = k0 + t + 1;
                    }
                }
                err = best.error;
                dxt5a_finish(sc, U, best, first, second);
                reordered = (best.first != best.second && first != best.first) ? 1u : 0u;
                if (lane == 0) { S.red_l[0] = first; S.red_h[0] = second; }
            }
        }
        __syncthreads();
        if (U != 1) { first = S.red_l[0]; second = S.red_h[0]; }
        if (tid == 0) {
            if (out_endpoints) out_endpoints[c] = first | (second << 8);
            if (out_error) out_error[c] = err;
            if (out_flags) out_flags[c] = reordered;
        }
        for (uint32_t base = 0; base < N; base += T) {
            const uint32_t i = base + tid;
            unsigned long long bits = 0;
            if (i < N) {
                const uint32_t a = (cluster_pixel(blocks, members, i) >> (8 * comp)) & 0xffu;
                bits = (unsigned long long)sc->sel[sc->uidx_of_value[a]] << (3 * (lane & 15));
            }
#pragma unroll
            for (int ofs = 8; ofs > 0; ofs >>= 1) bits |= __shfl_xor_sync(CRN_FULL_MASK, bits, ofs);
            if ((lane & 15) == 0 && i < N)
                *reinterpret_cast<unsigned long long*>(out + (size_t)members[i >> 4] * out_stride + out_ofs) =
                    (unsigned long long)first | ((unsigned long long)second << 8) | (bits << 16);
        }
    }
}

// ---- selector re-vote (qdxt1::optimize_selectors_task, crn_qdxt1.cpp:714-865; qdxt5:: crn_qdxt5.cpp:578-687) --
// One warp per selector cluster.  For each of the two block categories (colour: 4-colour / opaque 3-colour
// blocks; alpha: 8-value / 6-value blocks) and each of the 16 pixel positions, the selector minimising the
// error summed over the category's member blocks (each with ITS OWN endpoints) replaces the selectors of
// every member.  Lane = pixel position (x16) x block parity (x2); sums are integers, so the order of
// accumulation does not matter; ties keep the lowest selector like the reference's strict '<'.
template <bool IS_ALPHA>
__global__ void __launch_bounds__(kClusterWarpsPerCta * 32)
optimize_selectors_kernel(const uint32_t* __restrict__ blocks, const uint32_t* __restrict__ cluster_offsets,
                          const uint32_t* __restrict__ cluster_blocks, uint32_t n_clusters, uint32_t comp, int perceptual,
                          uint32_t alpha_threshold, uint8_t* __restrict__ elems, uint32_t stride, uint32_t ofs)
{
    constexpr int NS = IS_ALPHA ? 8 : 4;
    const unsigned lane = lane_id(), pos = lane & 15, half = lane >> 4;
    const uint32_t warps = gridDim.x * kClusterWarpsPerCta;
    for (uint32_t c = blockIdx.x * kClusterWarpsPerCta + (threadIdx.x >> 5); c < n_clusters; c += warps) {
        const uint32_t b0 = cluster_offsets[c], nb = cluster_offsets[c + 1] - b0;
        if (nb <= 1) continue;
        const uint32_t* members = cluster_blocks + b0;
        unsigned long long err[2][NS];
#pragma unroll
        for (int k = 0; k < NS; k++) { err[0][k] = 0; err[1][k] = 0; }
        unsigned cnt0 = 0, cnt1 = 0;
        for (uint32_t j = half; j < nb; j += 2) {
            const uint32_t b = members[j];
            const unsigned long long e = *reinterpret_cast<const unsigned long long*>(elems + (size_t)b * stride + ofs);
            const uint32_t px = blocks[(size_t)b * 16 + pos];
            if (IS_ALPHA) {
                const unsigned l = (unsigned)(e & 0xff), h = (unsigned)((e >> 8) & 0xff);
                unsigned p[8];
                const int cat = l <= h;                         // is_alpha6_block (crn_dxt.h:306-309)
                if (cat) dxt5a_values6(l, h, p); else dxt5a_values8(l, h, p);
                const int v = (int)((px >> (8 * comp)) & 0xffu);
                if (cat) cnt1++; else cnt0++;
#pragma unroll
                for (int k = 0; k < 8; k++) { const int d = v - (int)p[k]; if (cat) err[1][k] += (unsigned)(d * d); else err[0][k] += (unsigned)(d * d); }
            } else {
                const unsigned lo = (unsigned)(e & 0xffff), hi = (unsigned)((e >> 16) & 0xffff);
                int cat = lo <= hi;                             // is_alpha_block
                if (cat && alpha_threshold > 0) {               // 3-colour blocks with transparent pixels are left alone (:772-792)
                    const unsigned any = __ballot_sync(0xffffu << (16 * half), (px >> 24) < alpha_threshold);
                    if (any) cat = 2;
                }
                if (cat == 2) continue;
                int r0, g0, b0c, r1, g1, b1;
                unpack565(lo, true, r0, g0, b0c);
                unpack565(hi, true, r1, g1, b1);
                int pr[4], pg[4], pb[4];
                pr[0] = r0; pg[0] = g0; pb[0] = b0c; pr[1] = r1; pg[1] = g1; pb[1] = b1;
                if (!cat) {                                     // dxt1_block::get_block_colors4 (crn_dxt.cpp:246-260)
                    pr[2] = (r0 * 2 + r1) / 3; pg[2] = (g0 * 2 + g1) / 3; pb[2] = (b0c * 2 + b1) / 3;
                    pr[3] = (r1 * 2 + r0) / 3; pg[3] = (g1 * 2 + g0) / 3; pb[3] = (b1 * 2 + b0c) / 3;
                } else {                                        // get_block_colors3 (:234-244)
                    pr[2] = (r0 + r1) >> 1; pg[2] = (g0 + g1) >> 1; pb[2] = (b0c + b1) >> 1;
                    pr[3] = 0; pg[3] = 0; pb[3] = 0;
                }
                const int r = (int)(px & 0xff), g = (int)((px >> 8) & 0xff), bl = (int)((px >> 16) & 0xff);
                if (cat) cnt1++; else cnt0++;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int dr = r - pr[k], dg = g - pg[k], db = bl - pb[k];
                    const unsigned d = perceptual ? (unsigned)(8 * dr * dr + 25 * dg * dg + db * db) : (unsigned)(dr * dr + dg * dg + db * db);
                    if (cat) err[1][k] += d; else err[0][k] += d;
                }
            }
        }
        __syncwarp();
        cnt0 += __shfl_xor_sync(CRN_FULL_MASK, cnt0, 16);
        cnt1 += __shfl_xor_sync(CRN_FULL_MASK, cnt1, 16);
        unsigned long long bits[2] = { 0, 0 };
#pragma unroll
        for (int cat = 0; cat < 2; cat++) {
            unsigned best_s = 0;
            unsigned long long best_e = 0xFFFFFFFFFFull;
            const int max_s = IS_ALPHA ? 8 : (cat ? 3 : 4);
#pragma unroll
            for (int k = 0; k < NS; k++) {
                const unsigned long long t = err[cat][k] + __shfl_xor_sync(CRN_FULL_MASK, err[cat][k], 16);
                if (k < max_s && t < best_e) { best_e = t; best_s = (unsigned)k; }
            }
            unsigned long long b = half == 0 ? (unsigned long long)best_s << ((IS_ALPHA ? 3 : 2) * pos) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) b |= __shfl_xor_sync(CRN_FULL_MASK, b, o);
            bits[cat] = b;
        }
        // write back: every member of a category with more than one block gets the voted selectors
        for (uint32_t j = lane; j < nb; j += 32) {
            const uint32_t b = members[j];
            unsigned long long* pe = reinterpret_cast<unsigned long long*>(elems + (size_t)b * stride + ofs);
            const unsigned long long e = *pe;
            if (IS_ALPHA) {
                const int cat = (unsigned)(e & 0xff) <= (unsigned)((e >> 8) & 0xff);
                if ((cat ? cnt1 : cnt0) > 1) *pe = (e & 0xffffull) | (bits[cat] << 16);
            } else {
                const unsigned lo = (unsigned)(e & 0xffff), hi = (unsigned)((e >> 16) & 0xffff);
                int cat = lo <= hi;
                if (cat && alpha_threshold > 0) {
                    bool any = false;
                    for (int k = 0; k < 16; k++) any = any || (blocks[(size_t)b * 16 + k] >> 24) < alpha_threshold;
                    if (any) cat = 2;
                }
                if (cat != 2 && (cat ? cnt1 : cnt0) > 1) *pe = (e & 0xffffffffull) | (bits[cat] << 32);
            }
        }
        __syncwarp();
    }
}

}  // namespace crn
