// qdxt_kernels.cuh -- tile analysis of the clustered-DDS quantisers for sm_100a.
//
// Replaces the single-threaded front half of qdxt1::init / qdxt5::init (reference
// crnlib/crn_qdxt1.cpp:103-438, crnlib/crn_qdxt5.cpp:103-418): per 8x8-pixel chunk, fit the nine tile
// layouts with dxt_fast (crnlib/crn_dxt_fast.cpp:725-764, :788-826), score the eight encodings, and turn
// every tile of the winning encoding into an endpoint training vector with find_representative_colors
// (:855-995); then count the distinct dxt_fast selector patterns that bound the selector codebook.
//
// One warp per chunk; lane L holds chunk pixels L and L+32 (index = x + 8y).  All pixel statistics are
// integer warp reductions (exact in any order), the few scalar double-precision steps (4 power iterations,
// the Cramer solve, log10/sqrt scoring) are computed redundantly by every lane.
#pragma once
#include "dxt1_opt.cuh"
#include "dxt5a_opt.cuh"

namespace crn {

struct TileSel {                 // membership of this lane's two pixels in the current tile
    bool m[2];
    int li[2];                   // index of the pixel inside the tile (x - xo) + (y - yo) * w
};

__device__ __forceinline__ TileSel tile_select(int xo, int yo, int w, int h)
{
    TileSel t;
    const int lane = (int)lane_id();
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const int p = lane + 32 * s, x = p & 7, y = p >> 3;
        t.m[s] = x >= xo && x < xo + w && y >= yo && y < yo + h;
        t.li[s] = (x - xo) + (y - yo) * w;
    }
    return t;
}

__device__ __forceinline__ int wsum(int v) { return __reduce_add_sync(CRN_FULL_MASK, v); }
__device__ __forceinline__ int wmin(int v) { return __reduce_min_sync(CRN_FULL_MASK, v); }
__device__ __forceinline__ int wmax(int v) { return __reduce_max_sync(CRN_FULL_MASK, v); }
__device__ __forceinline__ unsigned long long wmin64(unsigned long long v) { return warp_min_u64(v); }

__device__ __forceinline__ int mul8bit(int a, int b) { const int t = a * b + 128; return (t + (t >> 8)) >> 8; }
__device__ __forceinline__ unsigned pack_fast565(unsigned px) { return (unsigned)((mul8bit(px & 0xff, 31) << 11) + (mul8bit((px >> 8) & 0xff, 63) << 5) + mul8bit((px >> 16) & 0xff, 31)); }

__device__ __forceinline__ void fast_eval_colors(int (&c)[4][3], unsigned c0, unsigned c1)
{
    unpack565(c0, true, c[0][0], c[0][1], c[0][2]);
    unpack565(c1, true, c[1][0], c[1][1], c[1][2]);
#pragma unroll
    for (int k = 0; k < 3; k++) { c[2][k] = (c[0][k] * 2 + c[1][k]) / 3; c[3][k] = (c[1][k] * 2 + c[0][k]) / 3; }
}

// match_block_colors / determine_selectors (crn_dxt_fast.cpp:84-136, :339-352).  Returns false if all equal.
__device__ __forceinline__ bool fast_determine_selectors(const unsigned (&px)[2], const TileSel& t, unsigned min16, unsigned max16, unsigned (&sel)[2])
{
    if (max16 == min16) { sel[0] = sel[1] = 0; return false; }
    int c[4][3];
    fast_eval_colors(c, min16, max16);
    const int dr = c[0][0] - c[1][0], dg = c[0][1] - c[1][1], db = c[0][2] - c[1][2];
    int stops[4];
#pragma unroll
    for (int i = 0; i < 4; i++) stops[i] = c[i][0] * dr + c[i][1] * dg + c[i][2] * db;
    const int c0p = (stops[1] + stops[3]) >> 1, half = (stops[3] + stops[2]) >> 1, c3p = (stops[2] + stops[0]) >> 1;
    unsigned first = 0xffffffffu;
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const int dot = (int)(px[s] & 0xff) * dr + (int)((px[s] >> 8) & 0xff) * dg + (int)((px[s] >> 16) & 0xff) * db;
        sel[s] = dot < half ? (dot < c0p ? 1u : 3u) : (dot < c3p ? 2u : 0u);
        if (t.m[s] && t.li[s] == 0) first = sel[s];
    }
    first = __reduce_min_sync(CRN_FULL_MASK, first);
    const bool differs = (t.m[0] && sel[0] != first) || (t.m[1] && sel[1] != first);
    return __any_sync(CRN_FULL_MASK, differs) != 0;
}

// dxt_fast::compress_color_block(n, ..., refine = false) over the tile's pixels (crn_dxt_fast.cpp:725-764).
__device__ __forceinline__ void fast_color_fit(const unsigned (&px)[2], const TileSel& t, int n, unsigned& low16, unsigned& high16, unsigned (&sel)[2])
{
    int r[2], g[2], b[2];
#pragma unroll
    for (int s = 0; s < 2; s++) { r[s] = px[s] & 0xff; g[s] = (px[s] >> 8) & 0xff; b[s] = (px[s] >> 16) & 0xff; }
    // optimize_block_colors (:138-265)
    int ave[3], mn[3], mx[3];
    {
        const int sr = wsum((t.m[0] ? r[0] : 0) + (t.m[1] ? r[1] : 0)), sg = wsum((t.m[0] ? g[0] : 0) + (t.m[1] ? g[1] : 0)), sb = wsum((t.m[0] ? b[0] : 0) + (t.m[1] ? b[1] : 0));
        ave[0] = (sr + n / 2) / n; ave[1] = (sg + n / 2) / n; ave[2] = (sb + n / 2) / n;
        mn[0] = wmin(min(t.m[0] ? r[0] : 255, t.m[1] ? r[1] : 255)); mx[0] = wmax(max(t.m[0] ? r[0] : 0, t.m[1] ? r[1] : 0));
        mn[1] = wmin(min(t.m[0] ? g[0] : 255, t.m[1] ? g[1] : 255)); mx[1] = wmax(max(t.m[0] ? g[0] : 0, t.m[1] ? g[1] : 0));
        mn[2] = wmin(min(t.m[0] ? b[0] : 255, t.m[1] ? b[1] : 255)); mx[2] = wmax(max(t.m[0] ? b[0] : 0, t.m[1] ? b[1] : 0));
    }
    bool solid = mn[0] == mx[0] && mn[1] == mx[1] && mn[2] == mx[2];
    unsigned max16 = 0, min16 = 0;
    if (!solid) {
        int cv[6] = { 0, 0, 0, 0, 0, 0 };
#pragma unroll
        for (int s = 0; s < 2; s++)
            if (t.m[s]) {
                const int dr = r[s] - ave[0], dg = g[s] - ave[1], db = b[s] - ave[2];
                cv[0] += dr * dr; cv[1] += dr * dg; cv[2] += dr * db; cv[3] += dg * dg; cv[4] += dg * db; cv[5] += db * db;
            }
        double covf[6];
#pragma unroll
        for (int i = 0; i < 6; i++) covf[i] = (double)wsum(cv[i]) * (double)(1.0f / 255.0f);
        double vfr = mx[0] - mn[0], vfg = mx[1] - mn[1], vfb = mx[2] - mn[2];
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            const double rr = vfr * covf[0] + vfg * covf[1] + vfb * covf[2];
            const double gg = vfr * covf[1] + vfg * covf[3] + vfb * covf[4];
            const double bb = vfr * covf[2] + vfg * covf[4] + vfb * covf[5];
            vfr = rr; vfg = gg; vfb = bb;
        }
        double magn = fabs(vfr) > fabs(vfg) ? fabs(vfr) : fabs(vfg);
        magn = magn > fabs(vfb) ? magn : fabs(vfb);
        int v_r, v_g, v_b;
        if (magn < 4.0) { v_r = 148; v_g = 300; v_b = 58; }
        else { magn = 512.0 / magn; v_r = d2i_x86(vfr * magn); v_g = d2i_x86(vfg * magn); v_b = d2i_x86(vfb * magn); }
        // first minimum / first maximum of the projection in tile order
        unsigned long long kmin = ~0ull, kmax = ~0ull;
#pragma unroll
        for (int s = 0; s < 2; s++)
            if (t.m[s]) {
                const long long dot = (long long)r[s] * v_r + (long long)g[s] * v_g + (long long)b[s] * v_b;
                const unsigned long long tag = ((unsigned long long)t.li[s] << 6) | (unsigned long long)(lane_id() + 32 * s);
                kmin = min(kmin, ((unsigned long long)(dot + (1ll << 30)) << 12) | tag);
                kmax = min(kmax, ((unsigned long long)((1ll << 31) - (dot + (1ll << 30))) << 12) | tag);
            }
        kmin = wmin64(kmin); kmax = wmin64(kmax);
        const int pmin = (int)(kmin & 63), pmax = (int)(kmax & 63);
        const unsigned a0 = __shfl_sync(CRN_FULL_MASK, px[0], pmin & 31), a1 = __shfl_sync(CRN_FULL_MASK, px[1], pmin & 31);
        const unsigned b0 = __shfl_sync(CRN_FULL_MASK, px[0], pmax & 31), b1 = __shfl_sync(CRN_FULL_MASK, px[1], pmax & 31);
        min16 = pack_fast565(pmin < 32 ? a0 : a1);
        max16 = pack_fast565(pmax < 32 ? b0 : b1);
        // NB the reference passes (low16 = max colour, high16 = min colour) into determine_selectors(min16, max16):
        // palette entry 0 is the MAX colour (crn_dxt_fast.cpp:738 with the signatures of :138 and :339)
        if (!fast_determine_selectors(px, t, max16, min16, sel)) solid = true;
    }
    if (solid) {   // compress_solid_block (:713-723)
        sel[0] = sel[1] = 2;
        max16 = ((unsigned)g_omatch5[2 * ave[0]] << 11) | ((unsigned)g_omatch6[2 * ave[1]] << 5) | g_omatch5[2 * ave[2]];
        min16 = ((unsigned)g_omatch5[2 * ave[0] + 1] << 11) | ((unsigned)g_omatch6[2 * ave[1] + 1] << 5) | g_omatch5[2 * ave[2] + 1];
    } else {
        // refine_block (:270-336): integer sums, then Cramer's rule in double
        int ak[3] = { 0, 0, 0 }, a1s[3] = { 0, 0, 0 }, a2s[3] = { 0, 0, 0 };
#pragma unroll
        for (int s = 0; s < 2; s++)
            if (t.m[s]) {
                const int step = (int)sel[s];
                const int w1 = (0x1203 >> (4 * step)) & 15;
                ak[0] += (0x2200 >> (4 * step)) & 15; ak[1] += (0x4190 >> (4 * step)) & 15; ak[2] += (0x1409 >> (4 * step)) & 15;
                a1s[0] += w1 * r[s]; a1s[1] += w1 * g[s]; a1s[2] += w1 * b[s];
                a2s[0] += r[s]; a2s[1] += g[s]; a2s[2] += b[s];
            }
        const double xy = wsum(ak[0]), yy = wsum(ak[1]), xx = wsum(ak[2]);
        const double A1r = wsum(a1s[0]), A1g = wsum(a1s[1]), A1b = wsum(a1s[2]);
        const double A2r = 3.0 * wsum(a2s[0]) - A1r, A2g = 3.0 * wsum(a2s[1]) - A1g, A2b = 3.0 * wsum(a2s[2]) - A1b;
        const double tt = xx * yy - xy * xy;
        if (!(!yy || !xx || fabs(tt) < (double).0000125f)) {
            const double frb = (double)(3.0f * 31.0f / 255.0f) / tt, fg = frb * (double)(63.0f / 31.0f);
            const unsigned omin = min16, omax = max16;
            max16 = ((unsigned)clampi(d2i_x86((A1r * yy - A2r * xy) * frb + 0.5), 0, 31) << 11) |
                    ((unsigned)clampi(d2i_x86((A1g * yy - A2g * xy) * fg + 0.5), 0, 63) << 5) |
                    (unsigned)clampi(d2i_x86((A1b * yy - A2b * xy) * frb + 0.5), 0, 31);
            min16 = ((unsigned)clampi(d2i_x86((A2r * xx - A1r * xy) * frb + 0.5), 0, 31) << 11) |
                    ((unsigned)clampi(d2i_x86((A2g * xx - A1g * xy) * fg + 0.5), 0, 63) << 5) |
                    (unsigned)clampi(d2i_x86((A2b * xx - A1b * xy) * frb + 0.5), 0, 31);
            if (omin != min16 || omax != max16) fast_determine_selectors(px, t, max16, min16, sel);
        }
    }
    // the reference names the (max16, min16) pair (low16, high16) and keeps low16 >= high16 (:756-763)
    low16 = max16; high16 = min16;
    if (low16 < high16) { const unsigned x = low16; low16 = high16; high16 = x; sel[0] ^= 1; sel[1] ^= 1; }
}

// squared error of the tile against its dxt_fast encoding (crn_qdxt1.cpp:187-196)
__device__ __forceinline__ unsigned fast_color_error(const unsigned (&px)[2], const TileSel& t, unsigned lo, unsigned hi, const unsigned (&sel)[2])
{
    int c[4][3];
    fast_eval_colors(c, lo, hi);
    if (lo <= hi) {   // dxt1_block::get_block_colors picks the 3-colour palette (crn_dxt.cpp:234-244)
#pragma unroll
        for (int k = 0; k < 3; k++) { c[2][k] = (c[0][k] + c[1][k]) >> 1; c[3][k] = 0; }
    }
    int e = 0;
#pragma unroll
    for (int s = 0; s < 2; s++)
        if (t.m[s]) {
            int pr = c[0][0], pg = c[0][1], pb = c[0][2];
#pragma unroll
            for (int k = 1; k < 4; k++) if (sel[s] == (unsigned)k) { pr = c[k][0]; pg = c[k][1]; pb = c[k][2]; }
            const int dr = (int)(px[s] & 0xff) - pr, dg = (int)((px[s] >> 8) & 0xff) - pg, db = (int)((px[s] >> 16) & 0xff) - pb;
            e += dr * dr + dg * dg + db * db;
        }
    return (unsigned)wsum(e);
}

// dxt_fast::compress_alpha_block + the tile's squared error (crn_dxt_fast.cpp:788-826, crn_qdxt5.cpp:183-194)
__device__ __forceinline__ unsigned fast_alpha_fit(const int (&a)[2], const TileSel& t, unsigned& low8, unsigned& high8, unsigned (&sel)[2])
{
    const int mn = wmin(min(t.m[0] ? a[0] : 255, t.m[1] ? a[1] : 255)), mx = wmax(max(t.m[0] ? a[0] : 0, t.m[1] ? a[1] : 0));
    low8 = (unsigned)mx; high8 = (unsigned)mn;
    const int dist = mx - mn, bias = mn * 7 - (dist >> 1), dist4 = dist * 4, dist2 = dist * 2;
    unsigned v[8];
    if (low8 > high8) dxt5a_values8(low8, high8, v); else dxt5a_values6(low8, high8, v);
    int e = 0;
#pragma unroll
    for (int s = 0; s < 2; s++) {
        int x = a[s] * 7 - bias, ind, q;
        q = (dist4 - x) >> 31; ind = q & 4; x -= dist4 & q;
        q = (dist2 - x) >> 31; ind += q & 2; x -= dist2 & q;
        q = (dist - x) >> 31; ind += q & 1;
        ind = -ind & 7;
        ind ^= (2 > ind);
        sel[s] = (unsigned)ind;
        if (t.m[s]) {
            unsigned pv = v[0];
#pragma unroll
            for (int k = 1; k < 8; k++) if (ind == k) pv = v[k];
            const int d = a[s] - (int)pv;
            e += d * d;
        }
    }
    return (unsigned)wsum(e);
}

// dxt_fast::find_representative_colors over the tile (crn_dxt_fast.cpp:855-995); channels as ints.
__device__ __forceinline__ void find_rep_colors(const int (&r)[2], const int (&g)[2], const int (&b)[2], const TileSel& t, int n, int (&lo)[3], int (&hi)[3])
{
    int ave[3];
    ave[0] = (wsum((t.m[0] ? r[0] : 0) + (t.m[1] ? r[1] : 0)) + n / 2) / n;
    ave[1] = (wsum((t.m[0] ? g[0] : 0) + (t.m[1] ? g[1] : 0)) + n / 2) / n;
    ave[2] = (wsum((t.m[0] ? b[0] : 0) + (t.m[1] ? b[1] : 0)) + n / 2) / n;
    auto first_max = [&](int cr, int cg, int cb, int (&out)[3]) {
        unsigned long long k = ~0ull;
#pragma unroll
        for (int s = 0; s < 2; s++)
            if (t.m[s]) {
                const int dr = r[s] - cr, dg = g[s] - cg, db = b[s] - cb;
                const unsigned d = (unsigned)(dr * dr + dg * dg + db * db);
                k = min(k, ((unsigned long long)(0x7fffffffu - d) << 12) | ((unsigned long long)t.li[s] << 6) | (unsigned long long)(lane_id() + 32 * s));
            }
        k = wmin64(k);
        const int p = (int)(k & 63);
        const int r0 = __shfl_sync(CRN_FULL_MASK, r[0], p & 31), r1 = __shfl_sync(CRN_FULL_MASK, r[1], p & 31);
        const int g0 = __shfl_sync(CRN_FULL_MASK, g[0], p & 31), g1 = __shfl_sync(CRN_FULL_MASK, g[1], p & 31);
        const int b0 = __shfl_sync(CRN_FULL_MASK, b[0], p & 31), b1 = __shfl_sync(CRN_FULL_MASK, b[1], p & 31);
        out[0] = p < 32 ? r0 : r1; out[1] = p < 32 ? g0 : g1; out[2] = p < 32 ? b0 : b1;
    };
    int lc[3], hc[3];
    first_max(ave[0], ave[1], ave[2], lc);
    first_max(lc[0], lc[1], lc[2], hc);
#pragma unroll
    for (int k = 0; k < 3; k++) { lc[k] = (lc[k] + ave[k]) >> 1; hc[k] = (hc[k] + ave[k]) >> 1; }
#pragma unroll 1
    for (int it = 0; it < 4; it++) {
        if (lc[0] == hc[0] && lc[1] == hc[1] && lc[2] == hc[2]) break;
        int vr = hc[0] - lc[0], vg = hc[1] - lc[1], vb = hc[2] - lc[2];
        const int mid = vr * lc[0] + vg * lc[1] + vb * lc[2] + vr * hc[0] + vg * hc[1] + vb * hc[2];
        vr *= 2; vg *= 2; vb *= 2;
        int s0[3] = { 0, 0, 0 }, s1[3] = { 0, 0, 0 }, w0 = 0, w1 = 0;
#pragma unroll
        for (int s = 0; s < 2; s++)
            if (t.m[s]) {
                const int dot = r[s] * vr + g[s] * vg + b[s] * vb;
                if (dot > mid) { s1[0] += r[s]; s1[1] += g[s]; s1[2] += b[s]; w1++; }
                else { s0[0] += r[s]; s0[1] += g[s]; s0[2] += b[s]; w0++; }
            }
        w0 = wsum(w0); w1 = wsum(w1);
        if (!w0 || !w1) break;
        int n0[3], n1[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { n0[k] = (wsum(s0[k]) + w0 / 2) / w0; n1[k] = (wsum(s1[k]) + w1 / 2) / w1; }
        if (n0[0] == lc[0] && n0[1] == lc[1] && n0[2] == lc[2] && n1[0] == hc[0] && n1[1] == hc[1] && n1[2] == hc[2]) break;
#pragma unroll
        for (int k = 0; k < 3; k++) { lc[k] = n0[k]; hc[k] = n1[k]; }
    }
    const int e0 = lc[0] * lc[0] + lc[1] * lc[1] + lc[2] * lc[2], e1 = hc[0] * hc[0] + hc[1] * hc[1] + hc[2] * hc[2];
    const bool sw = e0 > e1;
#pragma unroll
    for (int k = 0; k < 3; k++) { lo[k] = sw ? hc[k] : lc[k]; hi[k] = sw ? lc[k] : hc[k]; }
}

struct QdxtMip { uint32_t first_block, block_width, block_height, first_chunk; };
constexpr int kQdxtMaxMips = 6 * 16;
struct QdxtMipTable { QdxtMip m[kQdxtMaxMips]; uint32_t num_mips, total_chunks; };

CRN_DEVICE_TABLE uint8_t g_layout[9][4] = { { 0, 0, 8, 8 }, { 0, 0, 8, 4 }, { 0, 4, 8, 4 }, { 0, 0, 4, 8 }, { 4, 0, 4, 8 },
                                            { 0, 0, 4, 4 }, { 4, 0, 4, 4 }, { 0, 4, 4, 4 }, { 4, 4, 4, 4 } };   // crn_dxt_hc_common.cpp:42-58
CRN_DEVICE_TABLE uint8_t g_enc_nt[8] = { 1, 2, 2, 3, 3, 3, 3, 4 };                                              // :30-39
CRN_DEVICE_TABLE uint8_t g_enc_tiles[8][4] = { { 0, 0, 0, 0 }, { 1, 2, 0, 0 }, { 3, 4, 0, 0 }, { 1, 7, 8, 0 }, { 2, 5, 6, 0 }, { 3, 6, 8, 0 }, { 4, 5, 7, 0 }, { 5, 6, 7, 8 } };

constexpr int kQdxtWarpsPerCta = 8;

// KIND 0: colour (6-byte vectors lo.rgb, hi.rgb), KIND 1: alpha component `comp` (2-byte vectors lo, hi).
template <int KIND>
__global__ void __launch_bounds__(kQdxtWarpsPerCta * 32)
qdxt_training_kernel(const uint32_t* __restrict__ blocks, QdxtMipTable mt, uint32_t comp,
                     uint8_t* __restrict__ out_vecs, uint32_t* __restrict__ out_weights, uint8_t* __restrict__ out_encoding,
                     unsigned long long* __restrict__ out_sel_keys, int flat)
{
    const unsigned lane = lane_id();
    const uint32_t warps = gridDim.x * kQdxtWarpsPerCta;
    for (uint32_t ch = blockIdx.x * kQdxtWarpsPerCta + (threadIdx.x >> 5); ch < mt.total_chunks; ch += warps) {
        uint32_t level = 0;
        while (level + 1 < mt.num_mips && ch >= mt.m[level + 1].first_chunk) level++;
        const QdxtMip mp = mt.m[level];
        const uint32_t ncx = (mp.block_width + 1) / 2;
        const uint32_t local = ch - mp.first_chunk, cx = local % ncx, cy = local / ncx;
        const uint32_t lw = mp.block_width * 4, lh = mp.block_height * 4;
        // chunk pixels with edge clamp (crn_qdxt1.cpp:134-156)
        unsigned px[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const uint32_t p = lane + 32 * s;
            const uint32_t x = min(cx * 8 + (p & 7), lw - 1), y = min(cy * 8 + (p >> 3), lh - 1);
            px[s] = blocks[(size_t)(mp.first_block + (y >> 2) * mp.block_width + (x >> 2)) * 16 + (y & 3) * 4 + (x & 3)];
        }
        int rr[2], gg[2], bb[2], aa[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            rr[s] = px[s] & 0xff; gg[s] = (px[s] >> 8) & 0xff; bb[s] = (px[s] >> 16) & 0xff; aa[s] = (px[s] >> (8 * comp)) & 0xff;
        }
        // the nine tile layouts
        unsigned lerr[9];
#pragma unroll 1
        for (int l = 0; l < 9; l++) {
            const int xo = g_layout[l][0], yo = g_layout[l][1], w = g_layout[l][2], h = g_layout[l][3];
            const TileSel t = tile_select(xo, yo, w, h);
            unsigned sel[2];
            if (KIND == 0) {
                unsigned lo, hi;
                fast_color_fit(px, t, w * h, lo, hi, sel);
                lerr[l] = fast_color_error(px, t, lo, hi, sel);
            } else {
                unsigned lo, hi;
                lerr[l] = fast_alpha_fit(aa, t, lo, hi, sel);
            }
            // layouts 5..8 are the chunk's four blocks: their dxt_fast selectors are what qdxt1/qdxt5::init hash to bound
            // the selector codebook (crn_qdxt1.cpp:415-438, crn_qdxt5.cpp:395-421)
            if (out_sel_keys && l >= 5) {
                unsigned long long key = 0;
#pragma unroll
                for (int s2 = 0; s2 < 2; s2++)
                    if (t.m[s2]) key |= (unsigned long long)sel[s2] << ((KIND ? 3 : 2) * t.li[s2]);
                const unsigned klo = __reduce_or_sync(CRN_FULL_MASK, (unsigned)key), khi = __reduce_or_sync(CRN_FULL_MASK, (unsigned)(key >> 32));
                const uint32_t bx = cx * 2 + (xo >> 2), by = cy * 2 + (yo >> 2);
                if (lane == 0 && bx < mp.block_width && by < mp.block_height)
                    out_sel_keys[mp.first_block + bx + by * mp.block_width] = ((unsigned long long)khi << 32) | klo;
            }
        }
        // the eight encodings (crn_qdxt1.cpp:216-254, crn_qdxt5.cpp:197-239)
        float derating = KIND ? 2.4f : 1.5f;
        if (level && derating > .25f) {
            // powf(3.1f | 3.0f, level) in the reference: small integer powers, restated as repeated float multiplies is NOT
            // bit-identical to powf in general, so use the double power and round once like a correctly rounded powf
            const double base = KIND ? (double)3.0f : (double)3.1f;
            double pw = 1.0;
            for (uint32_t i = 0; i < level; i++) pw *= base;
            const float d = derating / (float)pw;
            derating = d > .25f ? d : .25f;
        }
        double best = -1.0; int best_e = 0;
#pragma unroll 1
        for (int e = 0; e < 8; e++) {
            double total = 0;
            const int nt = g_enc_nt[e];
            for (int q = 0; q < nt; q++) {
                const int li = g_enc_tiles[e][q];
                unsigned v = lerr[0];
#pragma unroll
                for (int k = 1; k < 9; k++) if (li == k) v = lerr[k];
                total += (double)v;
            }
            const double ms = total * (KIND ? (double)(1.0f / 64.0f) : (double)(1.0f / (64.0f * 3.0f)));
            double psnr = (double)999999.0f;
            if (ms != 0.0) {
                psnr = log10((double)255.0f / sqrt(ms)) * (double)20.0f;
                psnr = psnr < 0.0 ? 0.0 : (psnr > 500.0 ? 500.0 : psnr);
            }
            const float der = 0.0f + (derating - 0.0f) * ((float)(nt - 1) / 3.0f);
            psnr = psnr - (double)der;
            if (psnr > best) { best = psnr; best_e = e; }
        }
        // m_hierarchical == false (crn_qdxt1.cpp:370-403, crn_qdxt5.cpp:346-382): every block stands alone = the four-4x4-tile encoding
        if (flat) best_e = 7;
        if (out_encoding && lane == 0) out_encoding[ch] = (uint8_t)best_e;
        // training vectors of the winning encoding's tiles
        const int nt = g_enc_nt[best_e];
#pragma unroll 1
        for (int q = 0; q < nt; q++) {
            const int l = g_enc_tiles[best_e][q];
            const int xo = g_layout[l][0], yo = g_layout[l][1], w = g_layout[l][2], h = g_layout[l][3];
            const TileSel t = tile_select(xo, yo, w, h);
            int lo[3], hi[3];
            if (KIND == 0) find_rep_colors(rr, gg, bb, t, w * h, lo, hi);
            else find_rep_colors(aa, aa, aa, t, w * h, lo, hi);
            unsigned dist, wgt;
            if (KIND == 0) { dist = (unsigned)((lo[0] - hi[0]) * (lo[0] - hi[0]) + (lo[1] - hi[1]) * (lo[1] - hi[1]) + (lo[2] - hi[2]) * (lo[2] - hi[2])); wgt = dist / 5000u; }
            else { dist = (unsigned)((lo[0] - hi[0]) * (lo[0] - hi[0])); wgt = dist / 8u; }
            wgt = wgt < 1 ? 1 : (wgt > 8 ? 8 : wgt);
            // every block the tile covers (:310-341)
            const int nbx = w >> 2, nby = h >> 2;
            if ((int)lane < nbx * nby) {
                const uint32_t bx = cx * 2 + (lane % nbx) + (xo >> 2), by = cy * 2 + (lane / nbx) + (yo >> 2);
                if (bx < mp.block_width && by < mp.block_height) {
                    const uint32_t bi = mp.first_block + bx + by * mp.block_width;
                    if (KIND == 0) {
                        uint8_t* o = out_vecs + (size_t)bi * 6;
                        o[0] = (uint8_t)lo[0]; o[1] = (uint8_t)lo[1]; o[2] = (uint8_t)lo[2]; o[3] = (uint8_t)hi[0]; o[4] = (uint8_t)hi[1]; o[5] = (uint8_t)hi[2];
                    } else { out_vecs[(size_t)bi * 2] = (uint8_t)lo[0]; out_vecs[(size_t)bi * 2 + 1] = (uint8_t)hi[0]; }
                    out_weights[bi] = wgt;
                }
            }
        }
    }
}

// number of distinct keys (qdxt1/qdxt5::init's selector_hash.size()): open-addressing insert, `table` holds
// `mask + 1` entries preset to ~0 (no key of this path has bits above 47)
__global__ void __launch_bounds__(256) count_distinct_kernel(const unsigned long long* __restrict__ keys, uint32_t n, unsigned long long* __restrict__ table,
                                                             uint32_t mask, unsigned* __restrict__ counter)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    int fresh = 0;
    if (i < n) {
        const unsigned long long key = keys[i];
        uint32_t h = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32) & mask;
        for (;;) {
            const unsigned long long old = atomicCAS(&table[h], ~0ull, key);
            if (old == ~0ull) { fresh = 1; break; }
            if (old == key) break;
            h = (h + 1) & mask;
        }
    }
    const int total = __reduce_add_sync(CRN_FULL_MASK, fresh);
    if (lane_id() == 0 && total) atomicAdd(counter, (unsigned)total);
}

// Selector training vectors from the packed elements (qdxt1::create_selector_clusters, crn_qdxt1.cpp:871-908;
// qdxt5::create_selector_clusters, crn_qdxt5.cpp:701-760).  category: colour 0; alpha 0 = 8-value block,
// 1 = 6-value block, 255 = skipped (a 6-value block that uses the absolute selectors 6/7).
CRN_DEVICE_TABLE uint8_t g_sel_dxt1_to_linear[4] = { 0, 3, 1, 2 };                        // crn_dxt.cpp:39
CRN_DEVICE_TABLE uint8_t g_sel_dxt5_to_linear[8] = { 0, 7, 1, 2, 3, 4, 5, 6 };            // crn_dxt.cpp:34
CRN_DEVICE_TABLE uint8_t g_sel_dxt5_alpha6_to_linear[8] = { 0, 5, 1, 2, 3, 4, 0, 0 };     // crn_dxt.cpp:36

template <int KIND>
__global__ void __launch_bounds__(256) selector_vectors_kernel(const uint8_t* __restrict__ elements, uint32_t stride, uint32_t offset, uint32_t n, int perceptual,
                                                               uint8_t* __restrict__ out_vecs, uint32_t* __restrict__ out_weights, uint8_t* __restrict__ out_category)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const uint2 e = *reinterpret_cast<const uint2*>(elements + (size_t)b * stride + offset);
    uint8_t v[16];
    uint32_t weight; uint8_t cat = 0;
    if (KIND == 0) {
        const unsigned lo = e.x & 0xffff, hi = e.x >> 16;
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = g_sel_dxt1_to_linear[(e.y >> (2 * i)) & 3];
        int c0[3], c1[3];
        unpack565(lo, true, c0[0], c0[1], c0[2]);
        unpack565(hi, true, c1[0], c1[1], c1[2]);
        const int dr = c0[0] - c1[0], dg = c0[1] - c1[1], db = c0[2] - c1[2];
        const unsigned dist = perceptual ? (unsigned)(8 * dr * dr + 25 * dg * dg + db * db) : (unsigned)(dr * dr + dg * dg + db * db);   // crn_color.h:724-745
        weight = dist / 2000u;
    } else {
        const int lo = e.x & 0xff, hi = (e.x >> 8) & 0xff;
        const unsigned long long bits = ((unsigned long long)e.y << 16) | (e.x >> 16);
        const bool six = lo <= hi;
        bool absolute = false;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const unsigned s = (unsigned)(bits >> (3 * i)) & 7;
            if (six) { if (s >= 6) absolute = true; v[i] = g_sel_dxt5_alpha6_to_linear[s]; }
            else v[i] = g_sel_dxt5_to_linear[s];
        }
        cat = absolute ? 255 : (six ? 1 : 0);
        weight = (unsigned)((lo - hi) * (lo - hi)) / 8u;
    }
    weight = weight < 1 ? 1 : (weight > 2048 ? 2048 : weight);
    uint4* o = reinterpret_cast<uint4*>(out_vecs + (size_t)b * 16);
    uint32_t w4[4];
#pragma unroll
    for (int k = 0; k < 4; k++) w4[k] = v[4 * k] | (v[4 * k + 1] << 8) | (v[4 * k + 2] << 16) | ((uint32_t)v[4 * k + 3] << 24);
    *o = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    out_weights[b] = weight;
    if (out_category) out_category[b] = cat;
}

// dxt_pixel_block layout of one image level with edge clamping (mipmapped_texture::qdxt_pack_init,
// crn_mipmapped_texture.cpp:2457-2471)
__global__ void __launch_bounds__(256) blockify_kernel(const uint8_t* __restrict__ rgba, uint32_t width, uint32_t height, uint32_t pitch, uint32_t* __restrict__ blocks)
{
    const uint32_t bw = (width + 3) >> 2, bh = (height + 3) >> 2;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bw * bh * 16) return;
    const uint32_t b = i >> 4, p = i & 15;
    const uint32_t x = min((b % bw) * 4 + (p & 3), width - 1), y = min((b / bw) * 4 + (p >> 2), height - 1);
    blocks[i] = *reinterpret_cast<const uint32_t*>(rgba + (size_t)y * pitch + (size_t)x * 4);
}


// same gather with the level padded to a multiple of `pad` pixels (8 for crn_comp::quantize_images, crn_comp.cpp:717-741)
__global__ void __launch_bounds__(256) blockify_padded_kernel(const uint8_t* __restrict__ rgba, uint32_t width, uint32_t height, uint32_t pitch, uint32_t pad,
                                                              uint32_t* __restrict__ blocks)
{
    const uint32_t bw = ((width + pad - 1) / pad * pad) >> 2, bh = ((height + pad - 1) / pad * pad) >> 2;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bw * bh * 16) return;
    const uint32_t b = i >> 4, p = i & 15;
    const uint32_t x = min((b % bw) * 4 + (p & 3), width - 1), y = min((b / bw) * 4 + (p >> 2), height - 1);
    blocks[i] = *reinterpret_cast<const uint32_t*>(rgba + (size_t)y * pitch + (size_t)x * 4);
}

// retrieve_clusters() on the device: the clusters are ranges [offsets[c], offsets[c + 1]) of the tree's final permutation; block perm[i] goes to the
// cluster whose range holds position i.  A stable sort of the block ids by that key then lists every cluster's members in ascending order, as
// the reference's scan over cluster indices does (crn_qdxt1.cpp:941-960).
__global__ void __launch_bounds__(256) range_cluster_of_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ offsets, uint32_t n_clusters, uint32_t n,
                                                               uint32_t* __restrict__ cluster_of, uint32_t* __restrict__ ids)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t lo = 0, hi = n_clusters;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (offsets[mid] <= i) lo = mid; else hi = mid; }
    cluster_of[perm[i]] = lo;
    ids[i] = i;
}

}  // namespace crn
