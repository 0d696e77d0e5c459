// refiner_kernels.cuh -- dxt_hc building blocks for sm_100a.
//
// refine_endpoints_kernel replaces crnlib::dxt_endpoint_refiner::refine (crnlib/crn_dxt_endpoint_refiner.cpp:36-301),
// which dxt_hc::determine_color/alpha_endpoint_codebook_task calls once per endpoint cluster after the optimiser
// (crnlib/crn_dxt_hc.cpp:739-753, :1102-1129): with the cluster's selectors fixed, a least-squares start in double,
// then an integer search over the closed-form error  sum_s hist[s] v_s^2 - D2[s] v_s + DD[s].
// One warp per cluster:
//   * the eleven least-squares sums are DOUBLE accumulations in pixel order, so their value depends on that order:
//     lanes 0..10 own one sum each and walk the pixels in order (exact, eleven chains side by side);
//   * histogram moments: integer, lanes stride over the pixels, shared-memory atomics;
//   * candidates: one per lane.  The reference accepts a candidate when its error is strictly below the running best
//     and walks them in a fixed order (DXT5A: the start, then the window row-major; DXT1: ascending packed value after
//     sort + dedup), so the outcome is the lexicographic minimum of (error, position) -- a warp reduction.
//
// nearest_codebook_kernel replaces dxt_hc::determine_color_endpoint_clusters_task (crnlib/crn_dxt_hc.cpp:836-886) and
// determine_alpha_endpoint_clusters_task (:1132-1163): first codebook entry at minimum float squared distance.  The
// reference's early-outs against the tree-search leaf only skip entries that cannot win, so this is the plain first
// arg-min with the reference's float operation order (-fmad=false build).  Codebook tiles in shared memory, one
// thread per vector.
#pragma once
#include "warp_util.cuh"

namespace crn {

constexpr int kRefineWarpsPerCta = 4;

struct RefineSmem {
    unsigned long long hist[8];
    unsigned long long D2[8][3], DD[8][3];
};

__device__ __forceinline__ void refine_colors4(unsigned c0, unsigned c1, unsigned (&out)[4][3])
{   // dxt1_block::get_block_colors4 (crn_dxt.cpp:247-260) on unpack_color(scaled)
    unsigned a[3] = { (c0 >> 11) & 31u, (c0 >> 5) & 63u, c0 & 31u }, b[3] = { (c1 >> 11) & 31u, (c1 >> 5) & 63u, c1 & 31u };
    a[0] = a[0] << 3 | a[0] >> 2; a[1] = a[1] << 2 | a[1] >> 4; a[2] = a[2] << 3 | a[2] >> 2;
    b[0] = b[0] << 3 | b[0] >> 2; b[1] = b[1] << 2 | b[1] >> 4; b[2] = b[2] << 3 | b[2] >> 2;
#pragma unroll
    for (int c = 0; c < 3; c++) { out[0][c] = a[c]; out[1][c] = b[c]; out[2][c] = (a[c] * 2 + b[c]) / 3; out[3][c] = (b[c] * 2 + a[c]) / 3; }
}

__device__ __forceinline__ float refine_clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
__device__ __forceinline__ int refine_clampi(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

// lexicographic warp minimum of (err, key); every lane returns the winner
__device__ __forceinline__ void refine_warp_min(unsigned long long& err, unsigned& key, unsigned& payload)
{
#pragma unroll
    for (int ofs = 16; ofs; ofs >>= 1) {
        const unsigned long long e2 = __shfl_xor_sync(CRN_FULL_MASK, err, ofs);
        const unsigned k2 = __shfl_xor_sync(CRN_FULL_MASK, key, ofs), p2 = __shfl_xor_sync(CRN_FULL_MASK, payload, ofs);
        if (e2 < err || (e2 == err && k2 < key)) { err = e2; key = k2; payload = p2; }
    }
}

// pixels: RGBA8 of all clusters back to back; selectors: one byte per pixel; offsets: n_clusters + 1 (CSR).
// out_endpoints[c] = low | high << 16, out_error[c], out_ok[c] = error < error_to_beat[c] (error_to_beat may be null).
__global__ void __launch_bounds__(kRefineWarpsPerCta * 32)
refine_endpoints_kernel(const uint32_t* __restrict__ pixels, const uint8_t* __restrict__ selectors, const uint32_t* __restrict__ offsets,
                        uint32_t n_clusters, int dxt1_selectors, int perceptual, uint32_t comp,
                        const unsigned long long* __restrict__ error_to_beat,
                        uint32_t* __restrict__ out_endpoints, unsigned long long* __restrict__ out_error, uint8_t* __restrict__ out_ok, int parallel_sums)
{
    __shared__ RefineSmem smem[kRefineWarpsPerCta];
    const unsigned warp = threadIdx.x >> 5, lane = lane_id();
    RefineSmem& sm = smem[warp];
    const uint32_t warps = gridDim.x * kRefineWarpsPerCta;
    for (uint32_t cl = blockIdx.x * kRefineWarpsPerCta + warp; cl < n_clusters; cl += warps) {
        const uint32_t p0 = offsets[cl], n = offsets[cl + 1] - p0;
        if (!n) {                                           // refine() returns false without touching the results
            if (lane == 0) { out_endpoints[cl] = 0; out_error[cl] = ~0ull; out_ok[cl] = 0; }
            continue;
        }
        // ---- least squares (:51-128): lane j owns sum j -- 0 alpha^2, 1 beta^2, 2 alpha beta, 3-5 alpha x, 6-8 beta x
        // parallel_sums (the dxt_hc pipeline outside exact mode): every lane sums its own pixels of all nine sums and the totals meet in a
        // butterfly -- the sums then depend on rounding noise, tolerance class like the quantiser that formed the cluster.  A warp spends
        // n / 32 steps there instead of n (13 ms of a configs[2] pass were eleven ordered chains over ~14 K pixels per cluster).
        double acc = 0.0;
        double a2, b2, ab, axs[3], bxs[3];
        if (parallel_sums) {
            double t0 = 0, t1 = 0, t2 = 0, tx[3] = { 0, 0, 0 }, ty[3] = { 0, 0, 0 };
            for (uint32_t i = lane; i < n; i += 32) {
                const uint32_t px = pixels[p0 + i];
                const unsigned c = selectors[p0 + i];
                double k;
                if (dxt1_selectors) { const unsigned lin = (0x2130u >> (4 * c)) & 3u; k = (float)lin * 1.0f / 3.0f; }
                else { const unsigned lin = (0x65432170u >> (4 * c)) & 7u; k = (float)lin * 1.0f / 7.0f; }
                const double alpha = 1.0f - k, beta = k;
                t0 += alpha * alpha; t1 += beta * beta; t2 += alpha * beta;
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const unsigned ch = dxt1_selectors ? (unsigned)j : comp;
                    const float xf = dxt1_selectors ? (float)((px >> (8 * ch)) & 0xffu) * 1.0f / 255.0f : (float)((px >> (8 * ch)) & 0xffu) / 255.0f;
                    tx[j] += alpha * (double)xf; ty[j] += beta * (double)xf;
                }
            }
            a2 = warp_sum_f64(t0); b2 = warp_sum_f64(t1); ab = warp_sum_f64(t2);
#pragma unroll
            for (int j = 0; j < 3; j++) { axs[j] = warp_sum_f64(tx[j]); bxs[j] = warp_sum_f64(ty[j]); }
        } else {
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t px = pixels[p0 + i];
            const unsigned c = selectors[p0 + i];
            double k;
            if (dxt1_selectors) { const unsigned lin = (0x2130u >> (4 * c)) & 3u; k = (float)lin * 1.0f / 3.0f; }            // g_dxt1_to_linear {0,3,1,2}
            else { const unsigned lin = (0x65432170u >> (4 * c)) & 7u; k = (float)lin * 1.0f / 7.0f; }                        // g_dxt5_to_linear {0,7,1,2,3,4,5,6}
            const double alpha = 1.0f - k, beta = k;
            const unsigned ch = dxt1_selectors ? (lane >= 3 ? (lane - 3) % 3 : 0u) : comp;
            const float xf = dxt1_selectors ? (float)((px >> (8 * ch)) & 0xffu) * 1.0f / 255.0f : (float)((px >> (8 * ch)) & 0xffu) / 255.0f;
            const double x = xf;
            double term;
            if (lane == 0) term = alpha * alpha;
            else if (lane == 1) term = beta * beta;
            else if (lane == 2) term = alpha * beta;
            else if (lane < 6) term = alpha * x;
            else term = beta * x;
            acc += term;
        }
        a2 = __shfl_sync(CRN_FULL_MASK, acc, 0); b2 = __shfl_sync(CRN_FULL_MASK, acc, 1); ab = __shfl_sync(CRN_FULL_MASK, acc, 2);
#pragma unroll
        for (int j = 0; j < 3; j++) { axs[j] = __shfl_sync(CRN_FULL_MASK, acc, 3 + j); bxs[j] = __shfl_sync(CRN_FULL_MASK, acc, 6 + j); }
        }
        float l[3], h[3];
        {
            const uint32_t px0 = pixels[p0];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const double ax = axs[j], bx = bxs[j];
                const unsigned ch = dxt1_selectors ? (unsigned)j : comp;
                const double first = dxt1_selectors ? (double)((float)((px0 >> (8 * ch)) & 0xffu) * 1.0f / 255.0f) : (double)((float)((px0 >> (8 * ch)) & 0xffu) / 255.0f);
                double a, b;
                if (b2 == 0.0f) { a = ax / a2; b = 0.0; }
                else if (a2 == 0.0f) { a = 0.0; b = bx / b2; }
                else {
                    const double factor = a2 * b2 - ab * ab;
                    if (factor != 0.0f) { a = (ax * b2 - bx * ab) / factor; b = (bx * a2 - ax * ab) / factor; }
                    else { a = first; b = first; }
                }
                l[j] = refine_clamp01((float)a); h[j] = refine_clamp01((float)b);
            }
        }
        // ---- moments per selector value
        if (lane < 8) { sm.hist[lane] = 0; for (int c = 0; c < 3; c++) { sm.D2[lane][c] = 0; sm.DD[lane][c] = 0; } }
        __syncwarp();
        for (uint32_t i = lane; i < n; i += 32) {
            const uint32_t px = pixels[p0 + i];
            const unsigned s = selectors[p0 + i];
            atomicAdd(&sm.hist[s], 1ull);
            if (dxt1_selectors) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const unsigned long long v = (px >> (8 * c)) & 0xffu;
                    atomicAdd(&sm.D2[s][c], v * 2); atomicAdd(&sm.DD[s][c], v * v);
                }
            } else {
                const unsigned long long v = (px >> (8 * comp)) & 0xffu;
                atomicAdd(&sm.D2[s][0], v * 2); atomicAdd(&sm.DD[s][0], v * v);
            }
        }
        __syncwarp();
        unsigned long long best_err = ~0ull;
        unsigned best_lh = 0;
        if (!dxt1_selectors) {
            // ---- optimize_dxt5 (:146-201): the start, then the +-11 window row-major, first minimum wins
            const unsigned L0 = (unsigned)refine_clampi((int)(l[0] * 256.0f), 255), H0 = (unsigned)refine_clampi((int)(h[0] * 256.0f), 255);
            const unsigned minL = L0 <= 11 ? 0 : L0 - 11, maxL = L0 >= 244 ? 255 : L0 + 11;
            const unsigned minH = H0 <= 11 ? 0 : H0 - 11, maxH = H0 >= 244 ? 255 : H0 + 11;
            const unsigned wH = maxH - minH + 1, total = (maxL - minL + 1) * wH;
            unsigned long long e_best = ~0ull;
            unsigned k_best = 0xffffffffu, p_best = 0;
            for (unsigned q = lane; q < total + 1; q += 32) {           // q == 0 is the start, q - 1 indexes the window
                unsigned L, H;
                bool valid = true;
                if (q == 0) { L = L0; H = H0; }
                else {
                    L = minL + (q - 1) / wH; H = minH + (q - 1) % wH;
                    valid = (maxH < L || L <= H || H < minL) && (L != L0 || H != H0) && (L != H0 || H != L0);
                }
                if (!valid) continue;
                const unsigned sol = L == H ? (H ? ((H - 1) << 8 | L) : 1u) : (L > H ? (H << 8 | L) : (L << 8 | H));
                const unsigned lo = sol & 0xffu, hi = sol >> 8;
                const unsigned v[8] = { lo, hi, (lo * 6 + hi) / 7, (lo * 5 + hi * 2) / 7, (lo * 4 + hi * 3) / 7, (lo * 3 + hi * 4) / 7, (lo * 2 + hi * 5) / 7, (lo + hi * 6) / 7 };
                unsigned long long e = 0;
#pragma unroll
                for (int s = 0; s < 8; s++) e += sm.hist[s] * v[s] * v[s] - sm.D2[s][0] * v[s] + sm.DD[s][0];
                if (e < e_best) { e_best = e; k_best = q; p_best = lo | hi << 16; }      // q ascending per lane: first minimum kept
            }
            refine_warp_min(e_best, k_best, p_best);
            best_err = e_best; best_lh = p_best;
        } else {
            // ---- optimize_dxt1 (:203-301): up to eight rounds over the lattice neighbours of the current pair
            unsigned L0 = (unsigned)(refine_clampi((int)(l[0] * 32.0f), 31) << 11 | refine_clampi((int)(l[1] * 64.0f), 63) << 5 | refine_clampi((int)(l[2] * 32.0f), 31));
            unsigned H0 = (unsigned)(refine_clampi((int)(h[0] * 32.0f), 31) << 11 | refine_clampi((int)(h[1] * 64.0f), 63) << 5 | refine_clampi((int)(h[2] * 32.0f), 31));
            const bool preserveL = sm.hist[0] + sm.hist[2] > sm.hist[1] + sm.hist[3];
            bool improved = true;
            for (int it = 8; improved && it; it--) {
                improved = false;
                unsigned long long e_best = ~0ull;
                unsigned k_best = 0xffffffffu, p_best = 0;
                // 54 slots: 27 neighbours of L0 (against H0), 27 of H0 (against L0); slot = (db * 3 + dg) * 3 + dr
                for (unsigned slot = lane; slot < 54; slot += 32) {
                    const bool second = slot >= 27;
                    const unsigned sl = second ? slot - 27 : slot;
                    const unsigned C = second ? H0 : L0, O = second ? L0 : H0;
                    const unsigned b0 = C & 31u, g0 = (C >> 5) & 63u, r0 = (C >> 11) & 31u;
                    const unsigned b = (b0 ? b0 - 1 : b0) + sl / 9, g = (g0 ? g0 - 1 : g0) + (sl / 3) % 3, r = (r0 ? r0 - 1 : r0) + sl % 3;
                    if (b > b0 + 1 || b > 31 || g > g0 + 1 || g > 63 || r > r0 + 1 || r > 31) continue;
                    const unsigned X = r << 11 | g << 5 | b;
                    if (X == C) continue;
                    const unsigned packed = X > O ? (X | O << 16) : (O | X << 16);
                    unsigned L = packed & 0xffffu, H = packed >> 16;
                    if (L == H) {
                        L = (L + (!preserveL ? ((~L & 0x1Fu) ? 0x1u : (~L & 0xF800u) ? 0x800u : (~L & 0x7E0u) ? 0x20u : 0u) : (!L ? 0x1u : 0u))) & 0xffffu;
                        H = (H - (preserveL ? ((H & 0x1Fu) ? 0x1u : (H & 0xF800u) ? 0x800u : (H & 0x7E0u) ? 0x20u : 0u) : (H == 0xFFFFu ? 0x1u : 0u))) & 0xffffu;
                    }
                    unsigned bc[4][3];
                    refine_colors4(L, H, bc);
                    unsigned long long e = 0;
#pragma unroll
                    for (int s = 0; s < 4; s++) {
                        unsigned long long d[3];
#pragma unroll
                        for (int c = 0; c < 3; c++) d[c] = sm.hist[s] * bc[s][c] * bc[s][c] - sm.D2[s][c] * bc[s][c] + sm.DD[s][c];
                        e += perceptual ? d[0] * 8 + d[1] * 25 + d[2] : d[0] + d[1] + d[2];
                    }
                    if (e < e_best || (e == e_best && packed < k_best)) { e_best = e; k_best = packed; p_best = L | H << 16; }
                }
                refine_warp_min(e_best, k_best, p_best);
                if (e_best < best_err) {
                    best_err = e_best; best_lh = p_best;
                    L0 = p_best & 0xffffu; H0 = p_best >> 16;
                    improved = best_err != 0;
                }
            }
        }
        if (lane == 0) {
            out_endpoints[cl] = best_lh;
            out_error[cl] = best_err;
            out_ok[cl] = best_err < (error_to_beat ? error_to_beat[cl] : ~0ull) ? 1 : 0;
        }
        __syncwarp();
    }
}

// ---- nearest codebook entry ---------------------------------------------------------------------------------------
constexpr int kNearestThreads = 256;
constexpr int kNearestTile = 1024;               // codebook entries per shared-memory tile

template <int D>
__global__ void __launch_bounds__(kNearestThreads)
nearest_codebook_kernel(const float* __restrict__ vecs, uint32_t n, const float* __restrict__ codebook, uint32_t k, uint32_t* __restrict__ out)
{
    __shared__ float tile[kNearestTile * D];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float v[D];
#pragma unroll
    for (int d = 0; d < D; d++) v[d] = i < n ? vecs[(size_t)i * D + d] : 0.0f;
    float best = 1.0e+37f;                       // math::cNearlyInfinite
    uint32_t bi = 0;
    for (uint32_t base = 0; base < k; base += kNearestTile) {
        const uint32_t cnt = k - base < (uint32_t)kNearestTile ? k - base : (uint32_t)kNearestTile;
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < cnt * D; t += blockDim.x) tile[t] = codebook[(size_t)base * D + t];
        __syncthreads();
        if (best != 0.0f) {                      // the reference stops at the first exact match
            for (uint32_t j = 0; j < cnt; j++) {
                float dist = 0.0f;
#pragma unroll
                for (int d = 0; d < D; d++) { const float e = tile[j * D + d] - v[d]; dist += e * e; }
                if (dist < best) { best = dist; bi = base + j; if (best == 0.0f) break; }
            }
        }
    }
    if (i < n) out[i] = bi;
}

}  // namespace crn

// ---- selector codebook assignment + re-vote -----------------------------------------------------
// assign_selectors_kernel replaces dxt_hc::create_color_selector_codebook_task (KIND 0, crnlib/crn_dxt_hc.cpp:1306-1360)
// and create_alpha_selector_codebook_task (KIND 1, :1516-1586): for every block, the codebook entry (16 selectors of 2 or
// 3 bits) with the smallest summed error against the block's palette values -- an exhaustive blocks x codebook search.
// As in the reference the 16-term sum is folded into table lookups (colour: E2[16][4] -> E4[8][16] -> E8[4][256], four
// lookups per entry; alpha: E3[16][8] -> E6[8][64], eight lookups), here with the tables of one block in shared memory,
// one warp per block and one lane per codebook entry; the first minimum wins (strict < in entry order).  The block's
// per-pixel error table is then added into the winner's table (uint32 wrap-around sums, so the order does not matter).
// revote_selectors_kernel is the tail of create_color/alpha_selector_codebook (:1488-1503, :1702-1720): every entry's
// selectors are replaced by the per-pixel arg-min of its accumulated table, with the reference's pairwise tie order.
namespace crn {

constexpr int kAssignWarpsPerCta = 4;

template <int KIND> struct AssignSmem {
    static constexpr int V = KIND == 0 ? 4 : 8;
    uint32_t E1[16][V];                               // E2 (colour) / E3 (alpha)
    uint32_t Emid[KIND == 0 ? 8 * 16 : 1];            // E4 (colour only)
    uint32_t Ebig[KIND == 0 ? 4 * 256 : 8 * 64];      // E8 / E6
};

template <int KIND>
__device__ __forceinline__ void assign_fill_e1(AssignSmem<KIND>& sm, const uint32_t* __restrict__ block, const uint8_t* __restrict__ values, int perceptual, uint32_t comp)
{
    constexpr int V = AssignSmem<KIND>::V;
    const unsigned lane = lane_id();
    for (unsigned t = lane; t < 16u * V; t += 32) {
        const unsigned p = t / V, s = t % V;
        const uint32_t px = block[p];
        if (KIND == 0) {
            const uint32_t c = reinterpret_cast<const uint32_t*>(values)[s];
            const int dr = (int)(px & 0xffu) - (int)(c & 0xffu), dg = (int)((px >> 8) & 0xffu) - (int)((c >> 8) & 0xffu), db = (int)((px >> 16) & 0xffu) - (int)((c >> 16) & 0xffu);
            sm.E1[p][s] = perceptual ? (uint32_t)(8 * dr * dr) + (uint32_t)(25 * dg * dg) + (uint32_t)(db * db) : (uint32_t)(dr * dr + dg * dg + db * db);
        } else {
            const int d = (int)((px >> (8 * comp)) & 0xffu) - (int)values[s];
            sm.E1[p][s] = (uint32_t)(d * d);
        }
    }
}

// blocks: n x 16 RGBA8.  values: per block 4 RGBA8 colours (KIND 0, 16 bytes) or 8 alpha values (KIND 1, 8 bytes);
// values_accum (KIND 1, optional): the values whose error table is accumulated (the cluster's refined alpha values).
template <int KIND>
__global__ void __launch_bounds__(kAssignWarpsPerCta * 32)
assign_selectors_kernel(const uint32_t* __restrict__ blocks, uint32_t n_blocks, const uint8_t* __restrict__ values, const uint8_t* __restrict__ values_accum,
                        const unsigned long long* __restrict__ codebook, uint32_t K, int perceptual, uint32_t comp,
                        uint32_t* __restrict__ best_index, uint32_t* __restrict__ total_errors, uint8_t* __restrict__ used)
{
    constexpr int V = AssignSmem<KIND>::V, VB = KIND == 0 ? 16 : 8;
    __shared__ AssignSmem<KIND> smem[kAssignWarpsPerCta];
    const unsigned warp = threadIdx.x >> 5, lane = lane_id();
    AssignSmem<KIND>& sm = smem[warp];
    const uint32_t warps = gridDim.x * kAssignWarpsPerCta;
    for (uint32_t b = blockIdx.x * kAssignWarpsPerCta + warp; b < n_blocks; b += warps) {
        assign_fill_e1<KIND>(sm, blocks + (size_t)b * 16, values + (size_t)b * VB, perceptual, comp);
        __syncwarp();
        if (KIND == 0) {
            for (unsigned t = lane; t < 8u * 16u; t += 32) { const unsigned p = t >> 4, s = t & 15u; sm.Emid[t] = sm.E1[p << 1][s & 3u] + sm.E1[p << 1 | 1][s >> 2]; }
            __syncwarp();
            for (unsigned t = lane; t < 4u * 256u; t += 32) { const unsigned p = t >> 8, s = t & 255u; sm.Ebig[t] = sm.Emid[(p << 1) * 16 + (s & 15u)] + sm.Emid[(p << 1 | 1) * 16 + (s >> 4)]; }
        } else {
            for (unsigned t = lane; t < 8u * 64u; t += 32) { const unsigned p = t >> 6, s = t & 63u; sm.Ebig[t] = sm.E1[p << 1][s & 7u] + sm.E1[p << 1 | 1][s >> 3]; }
        }
        __syncwarp();
        uint32_t e_best = 0xffffffffu, i_best = 0xffffffffu;
        for (uint32_t s = lane; s < K; s += 32) {
            const unsigned long long sel = codebook[s];
            uint32_t e;
            if (KIND == 0) {
                const uint32_t q = (uint32_t)sel;
                e = sm.Ebig[q & 255u] + sm.Ebig[256 + ((q >> 8) & 255u)] + sm.Ebig[512 + ((q >> 16) & 255u)] + sm.Ebig[768 + (q >> 24)];
            } else {
                e = sm.Ebig[sel & 63u];
#pragma unroll
                for (int k = 1; k < 8; k++) e += sm.Ebig[64 * k + ((sel >> (6 * k)) & 63u)];
            }
            if (e < e_best) { e_best = e; i_best = s; }           // s ascending per lane: the first minimum is kept
        }
        // (error, index) lexicographic minimum; best_error starts at UINT32_MAX in the reference, so an entry whose error is
        // exactly UINT32_MAX never wins and index 0 stays
#pragma unroll
        for (int ofs = 16; ofs; ofs >>= 1) {
            const uint32_t e2 = __shfl_xor_sync(CRN_FULL_MASK, e_best, ofs), i2 = __shfl_xor_sync(CRN_FULL_MASK, i_best, ofs);
            if (e2 < e_best || (e2 == e_best && i2 < i_best)) { e_best = e2; i_best = i2; }
        }
        const uint32_t best = e_best == 0xffffffffu ? 0u : i_best;
        if (KIND == 1 && values_accum) {
            __syncwarp();
            assign_fill_e1<KIND>(sm, blocks + (size_t)b * 16, values_accum + (size_t)b * VB, perceptual, comp);
            __syncwarp();
        }
        uint32_t* tot = total_errors + (size_t)best * 16 * V;
        for (unsigned t = lane; t < 16u * V; t += 32) atomicAdd(&tot[t], sm.E1[t / V][t % V]);
        if (lane == 0) { used[best] = 1; best_index[b] = best; }
        __syncwarp();
    }
}

template <int KIND>
__global__ void __launch_bounds__(256)
revote_selectors_kernel(const uint32_t* __restrict__ total_errors, uint32_t K, unsigned long long* __restrict__ refined)
{
    constexpr int V = KIND == 0 ? 4 : 8;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const uint32_t* tab = total_errors + (size_t)i * 16 * V;
    unsigned long long out = 0;
    for (unsigned p = 0; p < 16; p++) {
        const uint32_t* e = tab + p * V;
        unsigned s;
        if (KIND == 0) {
            const unsigned s03 = e[3] < e[0] ? 3 : 0, s12 = e[2] < e[1] ? 2 : 1;
            s = e[s12] < e[s03] ? s12 : s03;
        } else {
            const unsigned s07 = e[7] < e[0] ? 7 : 0, s12 = e[2] < e[1] ? 2 : 1, s34 = e[4] < e[3] ? 4 : 3, s56 = e[6] < e[5] ? 6 : 5;
            const unsigned s02 = e[s12] < e[s07] ? s12 : s07, s36 = e[s56] < e[s34] ? s56 : s34;
            s = e[s36] < e[s02] ? s36 : s02;
        }
        out |= (unsigned long long)s << (p * (KIND == 0 ? 2 : 3));
    }
    refined[i] = out;
}

}  // namespace crn
