// hc_kernels.cuh -- device side of the dxt_hc pipeline for sm_100a.
//
// Replaces, in crnlib/crn_dxt_hc.cpp of the reference:
//   determine_tiles_task (:386-561) + palettize_color / palettize_alpha (:318-384) + split_vectors<>
//   (crn_tree_clusterizer.h:573-753)            -> hc_tiles_kernel, hc_palettize_kernel
//   tree_clusterizer<V>::split_node (crn_tree_clusterizer.h:265-570), V = vec2F / vec6F / vec16F
//                                                -> hc_tree_root_kernel, hc_tree_split_kernel
//   the per-block half of determine_color/alpha_endpoint_codebook_task (:692-737, :1035-1100)
//                                                -> hc_color_blocks_kernel, hc_alpha_blocks_kernel
// The per-cluster optimiser, the refiner, the nearest-codebook search and the selector search / re-vote are the
// kernels of cluster_kernels.cuh and refiner_kernels.cuh; hc_host.h strings them together.
#pragma once
#include "qdxt_kernels.cuh"
#include "cluster_kernels.cuh"

namespace crn {

constexpr int kHcMaxLevels = 16;                       // cCRNMaxLevels (inc/crnlib.h:46)
struct HcLevel { uint32_t first_block, num_blocks, block_width, first_chunk; float weight; };
struct HcTileParams {
    HcLevel levels[kHcMaxLevels];
    uint32_t num_levels, num_faces, total_chunks;
    float color_derating[kHcMaxLevels][8];             // m_color_derating (crn_dxt_hc.cpp:126-142)
    float alpha_derating[8];
    float color_alpha_ratio;                           // m_adaptive_tile_color_alpha_weighting_ratio
    int has_color, num_alpha;
    uint32_t alpha_comp[2];
};

// level weights by block index (dxt_hc::params::m_levels[].m_weight): blocks and tile slots of a level share its index range
struct HcLevelWeights { uint32_t first_block[kHcMaxLevels + 1]; float weight[kHcMaxLevels]; uint32_t num_levels; };
__device__ __forceinline__ float hc_level_weight(const HcLevelWeights& LW, uint32_t block)
{
    uint32_t l = 0;
    while (l + 1 < LW.num_levels && block >= LW.first_block[l + 1]) l++;
    return LW.weight[l];
}

// tile t of determine_tiles_task as a rectangle of the 8x8 chunk (x, y, w, h): 0-3 the four blocks in the order
// b, b+width, b+1, b+width+1; 4/5 left / right column; 6/7 top / bottom row; 8 the whole chunk (:389-390, :433-440).
// The reference's pixel order inside every tile is row-major over that rectangle.
CRN_DEVICE_TABLE uint8_t g_hc_rect[9][4] = { { 0, 0, 4, 4 }, { 0, 4, 4, 4 }, { 4, 0, 4, 4 }, { 4, 4, 4, 4 }, { 0, 0, 4, 8 }, { 4, 0, 4, 8 },
                                             { 0, 0, 8, 4 }, { 0, 4, 8, 4 }, { 0, 0, 8, 8 } };
CRN_DEVICE_TABLE uint8_t g_hc_enc_tiles[8][4] = { { 8, 0, 0, 0 }, { 6, 7, 0, 0 }, { 4, 5, 0, 0 }, { 6, 1, 3, 0 }, { 7, 0, 2, 0 }, { 4, 2, 3, 0 }, { 5, 0, 1, 0 }, { 0, 2, 1, 3 } };
CRN_DEVICE_TABLE uint8_t g_hc_enc_nt[8] = { 1, 2, 2, 3, 3, 3, 3, 4 };
// g_tile_map[encoding][by][bx] (crn_dxt_hc_common.cpp / crn_dxt_hc.cpp:36-45): tile index of each of the chunk's blocks
CRN_DEVICE_TABLE uint8_t g_hc_tile_map[8][2][2] = { { { 0, 0 }, { 0, 0 } }, { { 0, 0 }, { 1, 1 } }, { { 0, 1 }, { 0, 1 } }, { { 0, 0 }, { 1, 2 } },
                                                   { { 1, 2 }, { 0, 0 } }, { { 0, 1 }, { 0, 2 } }, { { 1, 0 }, { 2, 0 } }, { { 0, 1 }, { 2, 3 } } };

struct HcTileScratch : Dxt5aScratch { uint32_t count[256]; };
constexpr int kHcTileWarps = 4;

// squared error of the tile against the dxt_fast fit, always with the 4-colour palette (get_block_colors4, crn_dxt_hc.cpp:462-473)
__device__ __forceinline__ unsigned hc_color_error4(const unsigned (&px)[2], const TileSel& t, unsigned lo, unsigned hi, const unsigned (&sel)[2])
{
    int c[4][3];
    fast_eval_colors(c, lo, hi);
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

// dxt5_endpoint_optimizer at cCRNDXTQualityNormal, one block type, over the tile's values of one channel
// (crn_dxt_hc.cpp:475-491): only the error is used, so the order of the unique values is irrelevant.
__device__ __forceinline__ unsigned hc_alpha_tile_error(HcTileScratch* sc, const int (&a)[2], const TileSel& t)
{
    const unsigned lane = lane_id();
    for (unsigned v = lane; v < 256; v += 32) sc->count[v] = 0;
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 2; s++) if (t.m[s]) atomicAdd(&sc->count[a[s]], 1u);
    __syncwarp();
    unsigned mine = 0;
    for (unsigned k = 0; k < 8; k++) mine += sc->count[lane * 8 + k] != 0;
    unsigned incl = mine;
#pragma unroll
    for (int ofs = 1; ofs < 32; ofs <<= 1) { const unsigned o = __shfl_up_sync(CRN_FULL_MASK, incl, ofs); if ((int)lane >= ofs) incl += o; }
    const int U = (int)__shfl_sync(CRN_FULL_MASK, incl, 31);
    unsigned pos = incl - mine;
    for (unsigned k = 0; k < 8; k++) {
        const unsigned c = sc->count[lane * 8 + k];
        if (c) { sc->val[pos] = (uint8_t)(lane * 8 + k); sc->wgt[pos] = c; pos++; }
    }
    __syncwarp();
    unsigned err = 0;
    if (U > 1) err = (unsigned)dxt5a_search<false>(sc, U, 2, false).error;
    __syncwarp();
    return err;
}

// One warp per 8x8 chunk.  Outputs: per block its encoding and tile slot; per tile slot (4 per chunk, at the reference's
// boustrophedon tile_offset) the pixel count and where its pixels start in `vpix`, the tile-ordered copy of the pixels
// that the cluster optimisers read as virtual 16-pixel blocks.
__global__ void __launch_bounds__(kHcTileWarps * 32)
hc_tiles_kernel(const uint32_t* __restrict__ blocks, HcTileParams P, uint8_t* __restrict__ block_encoding, uint32_t* __restrict__ block_tile,
                uint8_t* __restrict__ tile_npix, uint8_t* __restrict__ tile_pixofs, uint32_t* __restrict__ vpix)
{
    __shared__ HcTileScratch scratch[kHcTileWarps];
    HcTileScratch* sc = &scratch[threadIdx.x >> 5];
    const unsigned lane = lane_id();
    const uint32_t warps = gridDim.x * kHcTileWarps;
    for (uint32_t ch = blockIdx.x * kHcTileWarps + (threadIdx.x >> 5); ch < P.total_chunks; ch += warps) {
        uint32_t level = 0;
        while (level + 1 < P.num_levels && ch >= P.levels[level + 1].first_chunk) level++;
        const HcLevel L = P.levels[level];
        const uint32_t width = L.block_width, ncx = width >> 1;
        const uint32_t local = ch - L.first_chunk, cx = local % ncx, cy = local / ncx;
        const uint32_t h = cy * 2, face_height = (L.num_blocks / width) / P.num_faces;
        const uint32_t b = L.first_block + h * width + cx * 2;
        // boustrophedon tile offset (:407-428)
        const bool reversed = ((h % face_height) & 2) != 0;
        const uint32_t row0 = L.first_block + h * width;
        const uint32_t tile_offset = reversed ? row0 + 2 * width - 4 - 4 * cx : row0 + 4 * cx;
        unsigned px[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const uint32_t p = lane + 32 * s, x = p & 7, y = p >> 3;
            px[s] = blocks[(size_t)(b + (y >> 2) * width + (x >> 2)) * 16 + (y & 3) * 4 + (x & 3)];
        }
        unsigned terr[3][9];
#pragma unroll 1
        for (int t = 0; t < 9; t++) {
            const int xo = g_hc_rect[t][0], yo = g_hc_rect[t][1], w = g_hc_rect[t][2], hh = g_hc_rect[t][3];
            const TileSel ts = tile_select(xo, yo, w, hh);
            unsigned e0 = 0, e1 = 0, e2 = 0;
            if (P.has_color) {
                unsigned lo, hi, sel[2];
                fast_color_fit(px, ts, w * hh, lo, hi, sel);
                e0 = hc_color_error4(px, ts, lo, hi, sel);
            }
            for (int a = 0; a < P.num_alpha; a++) {
                int av[2];
                av[0] = (px[0] >> (8 * P.alpha_comp[a])) & 0xff; av[1] = (px[1] >> (8 * P.alpha_comp[a])) & 0xff;
                const unsigned e = hc_alpha_tile_error(sc, av, ts);
                if (a == 0) e1 = e; else e2 = e;
            }
#pragma unroll
            for (int k = 0; k < 9; k++) if (k == t) { terr[0][k] = e0; terr[1][k] = e1; terr[2][k] = e2; }
        }
        // the eight encodings (:494-531)
        float best_quality = 0.0f; int best_e = 0;
#pragma unroll 1
        for (int e = 0; e < 8; e++) {
            unsigned tot[3] = { 0, 0, 0 };
            const int nt = g_hc_enc_nt[e];
            for (int q = 0; q < nt; q++) {
                const int t = g_hc_enc_tiles[e][q];
#pragma unroll
                for (int k = 0; k < 9; k++) if (k == t) { tot[0] += terr[0][k]; tot[1] += terr[1][k]; tot[2] += terr[2][k]; }
            }
            float quality = 0;
            if (P.has_color) {
                const double psnr = tot[0] ? log10((double)255.0f / sqrt((double)tot[0] / 192.0)) * (double)20.0f : (double)999999.0f;
                const double q = psnr - (double)P.color_derating[level][e];
                quality = (float)(q > 0.0 ? q : 0.0);
                if (P.num_alpha) quality *= P.color_alpha_ratio;
            }
            for (int a = 0; a < P.num_alpha; a++) {
                const unsigned te = tot[1 + a];
                const double psnr = te ? log10((double)255.0f / sqrt((double)te / 64.0)) * (double)20.0f : (double)999999.0f;
                const double q = psnr - (double)P.alpha_derating[e];
                quality += (float)(q > 0.0 ? q : 0.0);
            }
            if (quality > best_quality) { best_quality = quality; best_e = e; }
        }
        // tiles of the winning encoding: pixels in tile order, slot bookkeeping
        const int nt = g_hc_enc_nt[best_e];
        unsigned pixofs = 0;
#pragma unroll 1
        for (int q = 0; q < 4; q++) {
            if (q < nt) {
                const int t = g_hc_enc_tiles[best_e][q];
                const int xo = g_hc_rect[t][0], yo = g_hc_rect[t][1], w = g_hc_rect[t][2], hh = g_hc_rect[t][3];
                const TileSel ts = tile_select(xo, yo, w, hh);
#pragma unroll
                for (int s = 0; s < 2; s++)
                    if (ts.m[s]) vpix[(size_t)tile_offset * 16 + pixofs + ts.li[s]] = px[s];
                if (lane == 0) { tile_npix[tile_offset + q] = (uint8_t)(w * hh); tile_pixofs[tile_offset + q] = (uint8_t)pixofs; }
                pixofs += w * hh;
            } else if (lane == 0) { tile_npix[tile_offset + q] = 0; tile_pixofs[tile_offset + q] = 0; }
        }
        if (lane < 4) {
            const unsigned by = lane >> 1, bx = lane & 1;
            const uint32_t bi = b + by * width + bx;
            block_encoding[bi] = (uint8_t)best_e;
            block_tile[bi] = tile_offset | g_hc_tile_map[best_e][by][bx];
        }
        __syncwarp();
    }
}

// ---- split_vectors<V> (crn_tree_clusterizer.h:573-753), thread-serial with the reference's operand types and order ----
template <int D>
__device__ void hc_split_vectors(const float (*vec)[D], const unsigned* wts, unsigned size, float (&res0)[D], float (&res1)[D])
{
    // weightedVectors[i] = v * (float)weight and weightedDotProducts[i] = v.dot(v) * weight are recomputed where used: the same
    // operands give the same floats, and the per-thread local-memory footprint halves
    float centroid[D];
    for (int d = 0; d < D; d++) centroid[d] = 0.0f;
    unsigned long long total_weight = 0;
    double ttsum = 0.0;
    for (unsigned i = 0; i < size; i++) {
        const unsigned weight = wts[i];
        float dot = vec[i][0] * vec[i][0];
        for (int d = 1; d < D; d++) dot += vec[i][d] * vec[i][d];
        for (int d = 0; d < D; d++) centroid[d] += vec[i][d] * (float)weight;
        total_weight += weight;
        ttsum += (double)(dot * (float)weight);
    }
    float cdot = centroid[0] * centroid[0];
    for (int d = 1; d < D; d++) cdot += centroid[d] * centroid[d];
    const float variance = (float)(ttsum - (double)(cdot / (float)total_weight));
    const float inv_tw = 1.0f / (float)total_weight;
    for (int d = 0; d < D; d++) { centroid[d] *= inv_tw; res0[d] = res1[d] = centroid[d]; }
    if (variance <= 0.0f || size == 1) return;
    float furthest[D], opposite[D];
    double best = -1.0;
    for (int d = 0; d < D; d++) furthest[d] = opposite[d] = 0.0f;
    for (unsigned i = 0; i < size; i++) {
        float d2 = 0;
        for (int d = 0; d < D; d++) { const float x = vec[i][d] - centroid[d]; d2 += x * x; }
        if ((double)d2 > best) { best = d2; for (int d = 0; d < D; d++) furthest[d] = vec[i][d]; }
    }
    best = -1.0;
    for (unsigned i = 0; i < size; i++) {
        float d2 = 0;
        for (int d = 0; d < D; d++) { const float x = vec[i][d] - furthest[d]; d2 += x * x; }
        if ((double)d2 > best) { best = d2; for (int d = 0; d < D; d++) opposite[d] = vec[i][d]; }
    }
    float left[D], right[D];
    for (int d = 0; d < D; d++) { left[d] = (furthest[d] + centroid[d]) * .5f; right[d] = (opposite[d] + centroid[d]) * .5f; }
    if (size > 2) {
        float covar[D][D];
        for (int x = 0; x < D; x++) for (int y = 0; y < D; y++) covar[x][y] = 0.0f;
        for (unsigned i = 0; i < size; i++) {
            float v[D], w[D];
            for (int d = 0; d < D; d++) { v[d] = vec[i][d] - centroid[d]; w[d] = v[d] * (float)wts[i]; }
            for (int x = 0; x < D; x++) for (int y = x; y < D; y++) covar[x][y] = covar[x][y] + v[x] * w[y];
        }
        const float divider = (float)total_weight;
        for (int x = 0; x < D; x++) for (int y = x; y < D; y++) { covar[x][y] /= divider; covar[y][x] = covar[x][y]; }
        float axis[D];
        for (int d = 0; d < D; d++) axis[d] = 1.0f;
        for (int iter = 0; iter < 10; iter++) {
            float x[D];
            double max_sum = 0;
            for (int i = 0; i < D; i++) {
                double sum = 0;
                for (int j = 0; j < D; j++) sum += (double)(axis[j] * covar[i][j]);
                x[i] = (float)sum;
                max_sum = i ? (max_sum > sum ? max_sum : sum) : sum;
            }
            if (max_sum != 0.0) { const float sc = (float)(1.0 / max_sum); for (int i = 0; i < D; i++) x[i] *= sc; }
            for (int i = 0; i < D; i++) axis[i] = x[i];
        }
        {
            double n = (double)(axis[0] * axis[0]);
            for (int i = 1; i < D; i++) n += (double)(axis[i] * axis[i]);
            if (n != 0) { const float sc = (float)(1.0 / sqrt(n)); for (int i = 0; i < D; i++) axis[i] *= sc; }
        }
        float nl[D], nr[D];
        for (int d = 0; d < D; d++) nl[d] = nr[d] = 0.0f;
        double lw = 0.0, rw = 0.0;
        for (unsigned i = 0; i < size; i++) {
            float t = (vec[i][0] - centroid[0]) * axis[0];
            for (int d = 1; d < D; d++) t += (vec[i][d] - centroid[d]) * axis[d];
            if ((double)t < 0.0) { for (int d = 0; d < D; d++) nl[d] += vec[i][d] * (float)wts[i]; lw += (double)(float)wts[i]; }
            else { for (int d = 0; d < D; d++) nr[d] += vec[i][d] * (float)wts[i]; rw += (double)(float)wts[i]; }
        }
        if (lw > 0.0 && rw > 0.0) {
            const float sl = (float)(1.0 / lw), sr = (float)(1.0 / rw);
            for (int d = 0; d < D; d++) { left[d] = nl[d] * sl; right[d] = nr[d] * sr; }
        }
    }
    float prev_total_variance = 1e+10f;
    for (unsigned loops = 0; loops < 1024; loops++) {
        float nl[D], nr[D];
        for (int d = 0; d < D; d++) nl[d] = nr[d] = 0.0f;
        double lt = 0.0, rt = 0.0;
        unsigned long long lw = 0, rw = 0;
        for (unsigned i = 0; i < size; i++) {
            float dl = 0, dr = 0;
            for (int d = 0; d < D; d++) { const float x = left[d] - vec[i][d]; dl += x * x; }
            for (int d = 0; d < D; d++) { const float x = right[d] - vec[i][d]; dr += x * x; }
            float dot = vec[i][0] * vec[i][0];
            for (int d = 1; d < D; d++) dot += vec[i][d] * vec[i][d];
            const double wdp = (double)(dot * (float)wts[i]);
            if ((double)dl < (double)dr) { for (int d = 0; d < D; d++) nl[d] += vec[i][d] * (float)wts[i]; lt += wdp; lw += wts[i]; }
            else { for (int d = 0; d < D; d++) nr[d] += vec[i][d] * (float)wts[i]; rt += wdp; rw += wts[i]; }
        }
        if (!lw || !rw) return;
        float ldot = nl[0] * nl[0], rdot = nr[0] * nr[0];
        for (int d = 1; d < D; d++) { ldot += nl[d] * nl[d]; rdot += nr[d] * nr[d]; }
        const float lvar = (float)(lt - (double)(ldot / (float)lw)), rvar = (float)(rt - (double)(rdot / (float)rw));
        const float sl = 1.0f / (float)lw, sr = 1.0f / (float)rw;
        for (int d = 0; d < D; d++) { left[d] = nl[d] * sl; right[d] = nr[d] * sr; }
        const float total_variance = lvar + rvar;
        if (total_variance < .00001f) break;
        if (((prev_total_variance - total_variance) / total_variance) < .00001f) break;
        prev_total_variance = total_variance;
    }
    for (int d = 0; d < D; d++) { res0[d] = left[d]; res1[d] = right[d]; }
}

// palettize_color / palettize_alpha (crn_dxt_hc.cpp:318-384): one thread per (tile slot, component).  Component 0 is the
// colour vector when the format has one, the rest are the alpha channels.  color_vec: 6 floats per slot, alpha_vec:
// [num_alpha][n_slots][2] floats.
__global__ void __launch_bounds__(128)
hc_palettize_kernel(const uint32_t* __restrict__ vpix, const uint8_t* __restrict__ tile_npix, const uint8_t* __restrict__ tile_pixofs, uint32_t n_slots,
                    int has_color, int num_alpha, uint32_t comp0, uint32_t comp1, int perceptual, float* __restrict__ color_vec, float* __restrict__ alpha_vec)
{
    const int ncomp = (has_color ? 1 : 0) + num_alpha;
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_slots * (uint32_t)ncomp) return;
    const uint32_t slot = g / ncomp; const int c = (int)(g % ncomp);
    const unsigned n = tile_npix[slot];
    if (!n) return;
    const uint32_t* pix = vpix + (size_t)(slot & ~3u) * 16 + tile_pixofs[slot];
    unsigned key[64];
    const bool is_color = has_color && c == 0;
    const unsigned comp = is_color ? 0 : ((c - (has_color ? 1 : 0)) == 0 ? comp0 : comp1);
    for (unsigned i = 0; i < n; i++) {
        const unsigned p = pix[i];
        key[i] = is_color ? (((p & 0xff) << 16) | (p & 0xff00) | ((p >> 16) & 0xff)) : ((p >> (8 * comp)) & 0xff);
    }
    for (unsigned i = 1; i < n; i++) {      // insertion sort, ascending
        const unsigned k = key[i]; int j = (int)i - 1;
        while (j >= 0 && key[j] > k) { key[j + 1] = key[j]; j--; }
        key[j + 1] = k;
    }
    unsigned wts[64];
    unsigned size = 0;
    if (is_color) {
        float vec[64][3];
        for (unsigned i = 0; i < n; i++) {
            if (!i || key[i] != key[i - 1]) {
                const float r = (float)(key[i] >> 16) * 1.0f / 255.0f, gch = (float)((key[i] >> 8) & 0xff) * 1.0f / 255.0f, bch = (float)(key[i] & 0xff) * 1.0f / 255.0f;
                vec[size][0] = perceptual ? r * 0.5f : r; vec[size][1] = gch; vec[size][2] = perceptual ? bch * 0.25f : bch;
                wts[size] = 1; size++;
            } else wts[size - 1]++;
        }
        float r0[3], r1[3];
        hc_split_vectors<3>(vec, wts, size, r0, r1);
        const float l0 = (float)sqrt((double)(r0[0] * r0[0] + r0[1] * r0[1] + r0[2] * r0[2])), l1 = (float)sqrt((double)(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]));
        float* o = color_vec + (size_t)slot * 6;
        if (l0 > l1) { for (int d = 0; d < 3; d++) { o[d] = r1[d]; o[3 + d] = r0[d]; } }
        else { for (int d = 0; d < 3; d++) { o[d] = r0[d]; o[3 + d] = r1[d]; } }
    } else {
        float vec[64][1];
        for (unsigned i = 0; i < n; i++) {
            if (!i || key[i] != key[i - 1]) { vec[size][0] = (float)key[i] * 1.0f / 255.0f; wts[size] = 1; size++; }
            else wts[size - 1]++;
        }
        float r0[1], r1[1];
        hc_split_vectors<1>(vec, wts, size, r0, r1);
        const int a = c - (has_color ? 1 : 0);
        float* o = alpha_vec + ((size_t)a * n_slots + slot) * 2;
        if (r0[0] > r1[0]) { o[0] = r1[0]; o[1] = r0[0]; } else { o[0] = r0[0]; o[1] = r1[0]; }
    }
}

// ---- tree_clusterizer<V> (crn_tree_clusterizer.h:89-570): frontier-batched split_node, one CTA per node ------------
// Float sums run over the members in a fixed parallel order (per-thread strided partial sums in double, fixed reduction
// tree), so results are deterministic but not the reference's member-order float sums: this quantiser is
// tolerance-class, like the reference's own (its output depends on the helper-thread count).
template <int D> struct HcTreeSlot {
    uint32_t begin, end;                 // in: member range of the node in `perm`
    float centroid[D];                   // in
    unsigned long long total_weight;     // in
    int state;                           // out: 1 split, 2 unsplittable
    uint32_t n_left;
    float lc[D], rc[D];
    unsigned long long lw, rw;
    float lvar, rvar;
};

// Reductions over the threads that work on one node: one CTA of T threads (G == 1) or a thread-block cluster of G CTAs
// (G > 1, nvcc build only) whose per-CTA partials are exchanged through distributed shared memory.  The summation order is
// fixed (lanes by shuffle tree, warps in order, CTAs in rank order), so results are deterministic and identical in every CTA.
template <int T, int K> struct HcRed {
    double part_d[T / 32][K];
    unsigned long long part_u[T / 32][4];
    double cta_d[2][K];                       // double-buffered by call parity: one cluster barrier per reduction
    unsigned long long cta_u[2][4];
    double tot_d[K];
    unsigned long long tot_u[4], pre_u[4];
};

#ifdef __CUDACC__
}  // namespace crn
#include <cooperative_groups.h>
namespace crn {
namespace cg = cooperative_groups;
#endif

template <int G> __device__ __forceinline__ unsigned hc_cta_rank()
{
#ifdef __CUDACC__
    if (G > 1) return cg::this_cluster().block_rank();
#endif
    return 0;
}
template <int G> __device__ __forceinline__ void hc_group_sync()
{
#ifdef __CUDACC__
    if (G > 1) { cg::this_cluster().sync(); return; }
#endif
    __syncthreads();
}

// sums nd doubles (v) and nu <= 4 uint64 (u) over the group; u_prefix receives the sum over the CTAs of lower rank.
// SUM_U false: the uint64 entries are reduced with max instead (u_prefix unused).
template <int T, int G, int K, bool SUM_U>
__device__ __forceinline__ void hc_group_reduce(HcRed<T, K>& R, unsigned& parity, double* v, int nd, unsigned long long* u, int nu, unsigned long long* u_prefix)
{
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k = 0; k < nd; k++) {
        double x = v[k];
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) x += __shfl_xor_sync(CRN_FULL_MASK, x, ofs);
        if (lane == 0) R.part_d[warp][k] = x;
    }
    for (int k = 0; k < nu; k++) {
        unsigned long long x = u[k];
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) { const unsigned long long o = __shfl_xor_sync(CRN_FULL_MASK, x, ofs); x = SUM_U ? x + o : (o > x ? o : x); }
        if (lane == 0) R.part_u[warp][k] = x;
    }
    __syncthreads();
    const unsigned pb = parity & 1; parity++;
    if ((int)tid < nd) { double s = 0; for (int w = 0; w < T / 32; w++) s += R.part_d[w][tid]; R.cta_d[pb][tid] = s; if (G == 1) R.tot_d[tid] = s; }
    if ((int)tid >= nd && (int)tid < nd + nu) {
        const int k = tid - nd; unsigned long long s = 0;
        for (int w = 0; w < T / 32; w++) s = SUM_U ? s + R.part_u[w][k] : (R.part_u[w][k] > s ? R.part_u[w][k] : s);
        R.cta_u[pb][k] = s; if (G == 1) { R.tot_u[k] = s; R.pre_u[k] = 0; }
    }
    if (T == 32) __syncwarp();
#ifdef __CUDACC__
    if (G > 1) {
        cg::cluster_group cl = cg::this_cluster();
        cl.sync();
        const unsigned me = cl.block_rank();
        if ((int)tid < nd) { double s = 0; for (unsigned r = 0; r < (unsigned)G; r++) s += cl.map_shared_rank(&R.cta_d[pb][0], r)[tid]; R.tot_d[tid] = s; }
        if ((int)tid >= nd && (int)tid < nd + nu) {
            const int k = tid - nd; unsigned long long s = 0, pre = 0;
            for (unsigned r = 0; r < (unsigned)G; r++) {
                const unsigned long long x = cl.map_shared_rank(&R.cta_u[pb][0], r)[k];
                if (r == me) pre = s;
                s = SUM_U ? s + x : (x > s ? x : s);
            }
            R.tot_u[k] = s; R.pre_u[k] = pre;
        }
    }
#endif
    __syncthreads();
    for (int k = 0; k < nd; k++) v[k] = R.tot_d[k];
    for (int k = 0; k < nu; k++) { u[k] = R.tot_u[k]; if (u_prefix) u_prefix[k] = R.pre_u[k]; }
    __syncthreads();                             // tot_* may be rewritten by the next call
}

template <int D> __device__ __forceinline__ void hc_load_vec(const float* __restrict__ vecs, uint32_t id, float (&v)[D])
{
    if (D == 16) {
        const float4* p = reinterpret_cast<const float4*>(vecs + (size_t)id * 16);
#pragma unroll
        for (int q = 0; q < 4; q++) { const float4 x = p[q]; v[4 * q] = x.x; v[(4 * q + 1) % D] = x.y; v[(4 * q + 2) % D] = x.z; v[(4 * q + 3) % D] = x.w; }
    } else {
        const float2* p = reinterpret_cast<const float2*>(vecs + (size_t)id * D);
#pragma unroll
        for (int q = 0; q < D / 2; q++) { const float2 x = p[q]; v[2 * q] = x.x; v[2 * q + 1] = x.y; }
    }
}
// Walks the members first, first + stride, ... below e, calling f(i, id, v, w) once per member in that order.  The row and weight of the NEXT member
// are loaded before f runs on the current one and the member index after that is already in flight, so the two dependent gathers (perm -> row) overlap
// the arithmetic instead of being paid in series every iteration: the root split ran at long_scoreboard 13.5 of 16 cycles per issue with one chain per
// thread (profiles/r2ad_hc_tree_ncu.txt).  The order of the calls, and so every per-thread sum, is unchanged.
template <int D, typename F>
__device__ __forceinline__ void hc_for_members(const float* __restrict__ vecs, const uint32_t* __restrict__ wts, const uint32_t* __restrict__ perm,
                                               uint32_t first, uint32_t e, uint32_t stride, F&& f)
{
    if (first >= e) return;
    uint32_t i = first, id = perm[i];
    uint32_t id_next = (i + stride < e) ? perm[i + stride] : 0;
    float v[D]; hc_load_vec<D>(vecs, id, v);
    uint32_t w = wts[id];
    for (;;) {
        const uint32_t in = i + stride;
        const bool more = in < e;
        float vn[D]; uint32_t wn = 0, id2 = 0;
#pragma unroll
        for (int d = 0; d < D; d++) vn[d] = 0.0f;
        if (more) {
            hc_load_vec<D>(vecs, id_next, vn); wn = wts[id_next];
            if (in + stride < e) id2 = perm[in + stride];
        }
        f(i, id, v, w);
        if (!more) break;
        i = in; id = id_next; id_next = id2; w = wn;
#pragma unroll
        for (int d = 0; d < D; d++) v[d] = vn[d];
    }
}
template <int D> __device__ __forceinline__ float hc_sqdist(const float (&a)[D], const float (&b)[D])
{
    float s = 0;
#pragma unroll
    for (int d = 0; d < D; d++) { const float x = a[d] - b[d]; s += x * x; }
    return s;
}

// root statistics (generate_codebook, :101-125): sum of weighted vectors, total weight, sum of weighted dot products
template <int D>
__global__ void __launch_bounds__(512)
hc_tree_root_kernel(const float* __restrict__ vecs, const uint32_t* __restrict__ wts, uint32_t n, uint32_t* __restrict__ perm, double* __restrict__ out)
{
    constexpr int T = 512;
    __shared__ HcRed<T, D + 1> red;
    unsigned parity = 0;
    double s[D + 1]; unsigned long long tw = 0;
#pragma unroll
    for (int d = 0; d <= D; d++) s[d] = 0;
    for (uint32_t i = threadIdx.x; i < n; i += T) {
        float v[D]; hc_load_vec<D>(vecs, i, v);
        const unsigned w = wts[i];
        float dot = v[0] * v[0];
#pragma unroll
        for (int d = 1; d < D; d++) dot += v[d] * v[d];
#pragma unroll
        for (int d = 0; d < D; d++) s[d] += (double)(v[d] * (float)w);
        s[D] += (double)(dot * (float)w); tw += w;
        perm[i] = i;
    }
    hc_group_reduce<T, 1, D + 1, true>(red, parity, s, D + 1, &tw, 1, nullptr);
    if (threadIdx.x == 0) { for (int d = 0; d <= D; d++) out[d] = s[d]; out[D + 1] = (double)tw; }
}

// One node per group (CTA, or cluster of G CTAs).  CTA r of the group owns the contiguous member segment r.
template <int D, int T, int G>
__global__ void __launch_bounds__(T)
hc_tree_split_kernel(const float* __restrict__ vecs, const uint32_t* __restrict__ wts, uint32_t* __restrict__ perm, uint32_t* __restrict__ perm_tmp,
                     HcTreeSlot<D>* __restrict__ slots, const uint32_t* __restrict__ slot_list, uint32_t nslots)
{
    constexpr int K = D == 16 ? 18 : D * (D + 1) / 2 + D + 2;       // widest batch: Lloyd sums (D + 1) / totals (D + 1) / D <= 6 covariance
    constexpr int COVN = D == 16 ? T * 16 : D * D;
    __shared__ HcRed<T, K> red;
    __shared__ float s_cov[COVN];
    __shared__ double s_cpart[D == 16 ? 256 : 1];
    __shared__ float s_axis[D];
    __shared__ uint32_t s_warp_cnt[T / 32 + 1];
    const unsigned tid = threadIdx.x;
    const unsigned rank = hc_cta_rank<G>();
    unsigned parity = 0;
    for (uint32_t si = blockIdx.x / G; si < nslots; si += gridDim.x / G) {
        HcTreeSlot<D>& S = slots[slot_list[si]];
        const uint32_t begin = S.begin, end = S.end;
        const uint32_t seg = (end - begin + G - 1) / G;
        const uint32_t sb = min(end, begin + rank * seg), se = min(end, sb + seg);          // this CTA's members
        float centroid[D];
#pragma unroll
        for (int d = 0; d < D; d++) centroid[d] = S.centroid[d];
        const unsigned long long total_weight = S.total_weight;
        // furthest from the centroid, then furthest from that one (:305-330): first maximum in member order
        float seed[2][D];
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            unsigned long long key = 0;
            hc_for_members<D>(vecs, wts, perm, sb + tid, se, T, [&](uint32_t i, uint32_t, const float (&v)[D], uint32_t) {
                const float d2 = pass ? hc_sqdist<D>(v, seed[0]) : hc_sqdist<D>(v, centroid);
                const unsigned long long k = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)(~i);
                key = k > key ? k : key;
            });
            hc_group_reduce<T, G, K, false>(red, parity, nullptr, 0, &key, 1, nullptr);
            const uint32_t pos = ~(unsigned)key;
            hc_load_vec<D>(vecs, perm[pos], seed[pass]);
        }
        float left[D], right[D];
#pragma unroll
        for (int d = 0; d < D; d++) { left[d] = (seed[0][d] + centroid[d]) * .5f; right[d] = (seed[1][d] + centroid[d]) * .5f; }
        // node totals (right sums = total - left sums)
        double tot[D + 1];
        {
#pragma unroll
            for (int d = 0; d <= D; d++) tot[d] = 0;
            hc_for_members<D>(vecs, wts, perm, sb + tid, se, T, [&](uint32_t, uint32_t, const float (&v)[D], uint32_t wi) {
                const float w = (float)wi;
                float dot = v[0] * v[0];
#pragma unroll
                for (int d = 1; d < D; d++) dot += v[d] * v[d];
#pragma unroll
                for (int d = 0; d < D; d++) tot[d] += (double)(v[d] * w);
                tot[D] += (double)(dot * w);
            });
            hc_group_reduce<T, G, K, true>(red, parity, tot, D + 1, nullptr, 0, nullptr);
        }
        if (begin + 2 < end) {
            // covariance (:335-357)
            if (D == 16) {
                constexpr int LANES = T / 16;
                const int x = tid & 15, ml = tid >> 4;
                float acc[16];
#pragma unroll
                for (int y = 0; y < 16; y++) acc[y] = 0.0f;
                hc_for_members<D>(vecs, wts, perm, sb + ml, se, LANES, [&](uint32_t, uint32_t, const float (&vr)[D], uint32_t wi) {
                    const float w = (float)wi;
                    float v[D], vx = 0;
#pragma unroll
                    for (int d = 0; d < D; d++) { v[d] = vr[d] - centroid[d]; if (d == x) vx = v[d]; }
#pragma unroll
                    for (int y = 0; y < 16; y++) acc[y % D] += vx * (v[y % D] * w);
                });
                __syncthreads();
#pragma unroll
                for (int y = 0; y < 16; y++) s_cov[tid * 16 + y] = acc[y];
                __syncthreads();
                for (int e = tid; e < 256; e += T) {
                    const int ex = e >> 4, ey = e & 15;
                    double s = 0;
                    for (int l = 0; l < LANES; l++) s += (double)s_cov[(l * 16 + ex) * 16 + ey];
                    s_cpart[e % (D == 16 ? 256 : 1)] = s;
                }
                hc_group_sync<G>();
                for (int e = tid; e < 256; e += T) {
                    double s = s_cpart[e % (D == 16 ? 256 : 1)];
#ifdef __CUDACC__
                    if (G > 1) { s = 0; for (unsigned r = 0; r < (unsigned)G; r++) s += cg::this_cluster().map_shared_rank(&s_cpart[0], r)[e % (D == 16 ? 256 : 1)]; }
#endif
                    s_cov[e] = (float)s / (float)total_weight;
                }
                hc_group_sync<G>();
            } else {
                constexpr int NC = D * (D + 1) / 2;
                float acc[NC];
#pragma unroll
                for (int k = 0; k < NC; k++) acc[k] = 0.0f;
                for (uint32_t i = sb + tid; i < se; i += T) {
                    const uint32_t id = perm[i];
                    float v[D]; hc_load_vec<D>(vecs, id, v);
                    const float w = (float)wts[id];
#pragma unroll
                    for (int d = 0; d < D; d++) v[d] -= centroid[d];
                    int k = 0;
#pragma unroll
                    for (int x = 0; x < D; x++)
#pragma unroll
                        for (int y = x; y < D; y++) acc[k++] += v[x] * (v[y] * w);
                }
                double dacc[NC];
#pragma unroll
                for (int k = 0; k < NC; k++) dacc[k] = (double)acc[k];
                hc_group_reduce<T, G, K, true>(red, parity, dacc, NC, nullptr, 0, nullptr);
                if (tid == 0) {
                    int k = 0;
                    for (int x = 0; x < D; x++) for (int y = x; y < D; y++) { const float c = (float)dacc[k++] / (float)total_weight; s_cov[(x * D + y) % COVN] = c; s_cov[(y * D + x) % COVN] = c; }
                }
                __syncthreads();
            }
            // 10 power iterations from (1, ..., 1) with max-normalisation, then normalise (:358-386)
            if (tid == 0) {
                float axis[D];
#pragma unroll
                for (int d = 0; d < D; d++) axis[d] = 1.0f;
                for (int iter = 0; iter < 10; iter++) {
                    float xv[D];
                    double max_sum = 0;
                    for (int i = 0; i < D; i++) {
                        double sum = 0;
                        for (int j = 0; j < D; j++) {
                            const float c = i <= j ? s_cov[i * D + j] : s_cov[j * D + i];      // the reference mirrors the upper triangle
                            sum += (double)(axis[j] * c);
                        }
                        xv[i] = (float)sum;
                        max_sum = i ? (max_sum > sum ? max_sum : sum) : sum;
                    }
                    if (max_sum != 0.0) { const float sc = (float)(1.0 / max_sum); for (int i = 0; i < D; i++) xv[i] *= sc; }
                    for (int i = 0; i < D; i++) axis[i] = xv[i];
                }
                double n = (double)(axis[0] * axis[0]);
                for (int i = 1; i < D; i++) n += (double)(axis[i] * axis[i]);
                if (n != 0) { const float sc = (float)(1.0 / sqrt(n)); for (int i = 0; i < D; i++) axis[i] *= sc; }
                for (int i = 0; i < D; i++) s_axis[i] = axis[i];
            }
            __syncthreads();
            float axis[D];
#pragma unroll
            for (int d = 0; d < D; d++) axis[d] = s_axis[d];
            // split by the sign of the projection (:387-412)
            double sl[D + 2];
#pragma unroll
            for (int d = 0; d < D + 2; d++) sl[d] = 0;
            hc_for_members<D>(vecs, wts, perm, sb + tid, se, T, [&](uint32_t, uint32_t, const float (&v)[D], uint32_t wi) {
                const float w = (float)wi;
                float t = (v[0] - centroid[0]) * axis[0];
#pragma unroll
                for (int d = 1; d < D; d++) t += (v[d] - centroid[d]) * axis[d];
                sl[D + 1] += (double)w;
                if (t < 0.0f) {
#pragma unroll
                    for (int d = 0; d < D; d++) sl[d] += (double)(v[d] * w);
                    sl[D] += (double)w;
                }
            });
            hc_group_reduce<T, G, K, true>(red, parity, sl, D + 2, nullptr, 0, nullptr);
            const double lw = sl[D], rw = sl[D + 1] - sl[D];
            if (lw > 0.0 && rw > 0.0) {
                const float fl = (float)(1.0 / lw), fr = (float)(1.0 / rw);
#pragma unroll
                for (int d = 0; d < D; d++) { left[d] = (float)sl[d] * fl; right[d] = (float)(tot[d] - sl[d]) * fr; }
            }
        }
        // Lloyd iterations until the variance stops improving (:413-519)
        float prev_total_variance = 1e+10f, lvar = 0, rvar = 0;
        unsigned long long lw = 0, rw = 0;
        uint32_t n_left = 0, left_before = 0;
        bool unsplittable = false;
        float used_left[D], used_right[D];
#pragma unroll 1
        for (unsigned loops = 0; loops < 1024; loops++) {
            double sl[D + 1]; unsigned long long uu[2] = { 0, 0 }, upre[2];
#pragma unroll
            for (int d = 0; d <= D; d++) sl[d] = 0;
#pragma unroll
            for (int d = 0; d < D; d++) { used_left[d] = left[d]; used_right[d] = right[d]; }
            hc_for_members<D>(vecs, wts, perm, sb + tid, se, T, [&](uint32_t, uint32_t, const float (&v)[D], uint32_t wi) {
                if (hc_sqdist<D>(left, v) < hc_sqdist<D>(right, v)) {
                    const float w = (float)wi;
                    float dot = v[0] * v[0];
#pragma unroll
                    for (int d = 1; d < D; d++) dot += v[d] * v[d];
#pragma unroll
                    for (int d = 0; d < D; d++) sl[d] += (double)(v[d] * w);
                    sl[D] += (double)(dot * w); uu[0] += wi; uu[1]++;
                }
            });
            hc_group_reduce<T, G, K, true>(red, parity, sl, D + 1, uu, 2, upre);
            lw = uu[0]; n_left = (uint32_t)uu[1]; left_before = (uint32_t)upre[1];
            rw = total_weight - lw;
            if (!lw || !rw) { unsplittable = true; break; }
            float nl[D], nr[D];
#pragma unroll
            for (int d = 0; d < D; d++) { nl[d] = (float)sl[d]; nr[d] = (float)(tot[d] - sl[d]); }
            float ldot = nl[0] * nl[0], rdot = nr[0] * nr[0];
#pragma unroll
            for (int d = 1; d < D; d++) { ldot += nl[d] * nl[d]; rdot += nr[d] * nr[d]; }
            lvar = (float)(sl[D] - (double)(ldot / (float)lw)); rvar = (float)((tot[D] - sl[D]) - (double)(rdot / (float)rw));
            const float fl = 1.0f / (float)lw, fr = 1.0f / (float)rw;
#pragma unroll
            for (int d = 0; d < D; d++) { left[d] = nl[d] * fl; right[d] = nr[d] * fr; }
            const float total_variance = lvar + rvar;
            if (total_variance < .00001f) break;
            if (((prev_total_variance - total_variance) / total_variance) < .00001f) break;
            prev_total_variance = total_variance;
        }
        if (!unsplittable) {
            // stable partition by the last assignment (:521-535): lefts of lower-ranked CTAs come first
            uint32_t base_l = begin + left_before, base_r = begin + n_left + ((sb - begin) - left_before);
            for (uint32_t i0 = sb; i0 < se; i0 += T) {
                const uint32_t i = i0 + tid;
                uint32_t id = 0; bool valid = i < se, is_left = false;
                if (valid) {
                    id = perm[i];
                    float v[D]; hc_load_vec<D>(vecs, id, v);
                    is_left = hc_sqdist<D>(used_left, v) < hc_sqdist<D>(used_right, v);
                }
                const unsigned bl = __ballot_sync(CRN_FULL_MASK, valid && is_left), br = __ballot_sync(CRN_FULL_MASK, valid && !is_left);
                uint32_t pre_l = 0, pre_r = 0, tot_l = __popc(bl), tot_r = __popc(br);
                if (T > 32) {
                    __syncthreads();
                    if ((tid & 31) == 0) { s_warp_cnt[tid >> 5] = (uint32_t)__popc(bl) | ((uint32_t)__popc(br) << 16); }
                    __syncthreads();
                    tot_l = tot_r = 0;
                    for (unsigned w = 0; w < T / 32; w++) {
                        const uint32_t c = s_warp_cnt[w];
                        if (w < (tid >> 5)) { pre_l += c & 0xffff; pre_r += c >> 16; }
                        tot_l += c & 0xffff; tot_r += c >> 16;
                    }
                }
                if (valid) {
                    const unsigned lt_mask = (1u << (tid & 31)) - 1u;
                    if (is_left) perm_tmp[base_l + pre_l + __popc(bl & lt_mask)] = id;
                    else perm_tmp[base_r + pre_r + __popc(br & lt_mask)] = id;
                }
                base_l += tot_l; base_r += tot_r;
            }
            hc_group_sync<G>();
            for (uint32_t i = sb + tid; i < se; i += T) perm[i] = perm_tmp[i];
        }
        if (tid == 0 && rank == 0) {
            S.state = unsplittable ? 2 : 1;
            S.n_left = n_left;
#pragma unroll
            for (int d = 0; d < D; d++) { S.lc[d] = left[d]; S.rc[d] = right[d]; }
            S.lw = lw; S.rw = rw; S.lvar = lvar; S.rvar = rvar;
        }
        hc_group_sync<G>();
    }
}

// ---- per-block selectors against the cluster palette (determine_*_endpoint_codebook_task, block loops) -------------
// colour (crn_dxt_hc.cpp:692-737): cluster_endpoints = low | high << 16, cluster_flags bit 0 = m_reordered, bit 4 =
// m_alternate_rounding.  Writes m_block_selectors[cColor][b] = selector << 32 | weight and the four RGBA8 colour values
// of the block's cluster in linear order (color_cluster::color_values, without the alternate rounding) for the selector search.
__global__ void __launch_bounds__(256)
hc_color_blocks_kernel(const uint32_t* __restrict__ blocks, uint32_t n_blocks, const uint32_t* __restrict__ block_cluster, const uint32_t* __restrict__ cluster_endpoints,
                       const uint32_t* __restrict__ cluster_flags, HcLevelWeights LW, const uint8_t* __restrict__ block_encoding, int perceptual,
                       unsigned long long* __restrict__ block_selectors, uint32_t* __restrict__ block_values)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint32_t c = block_cluster[b], ep = cluster_endpoints[c], fl = cluster_flags[c];
    const bool reordered = fl & 1, alt = (fl >> 4) & 1;
    int cv[4][3], cs[4][3];
    unpack565(ep & 0xffff, true, cv[0][0], cv[0][1], cv[0][2]);
    unpack565(ep >> 16, true, cv[3][0], cv[3][1], cv[3][2]);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        cv[1][k] = (cv[0][k] * 2 + cv[3][k]) / 3; cv[2][k] = (cv[3][k] * 2 + cv[0][k]) / 3;
        cs[0][k] = cv[0][k]; cs[3][k] = cv[3][k];
        cs[1][k] = alt ? ((cv[0][k] << 1) + cv[3][k] + 1) / 3 : cv[1][k];
        cs[2][k] = alt ? ((cv[3][k] << 1) + cv[0][k] + 1) / 3 : cv[2][k];
    }
    const int wr = perceptual ? 8 : 1, wg = perceptual ? 25 : 1, wb = 1;
    const int d0r = cs[0][0] - cs[3][0], d0g = cs[0][1] - cs[3][1], d0b = cs[0][2] - cs[3][2];
    const unsigned endpoint_weight = (unsigned)(wr * d0r * d0r + wg * d0g * d0g + wb * d0b * d0b) / 2000u;
    const float ew = 1.15f + (1.0f - 1.15f) * ((float)block_encoding[b] / 7.0f);                // math::lerp(1.15f, 1.0f, i / 7.0f)
    float wf = (float)endpoint_weight * hc_level_weight(LW, b);                                // uint * float (m_block_weights[b])
    // math::clamp<uint>(float, 1, 2048): the float converts to uint first (x86 cvttss2si semantics for in-range values)
    unsigned wu = (unsigned)wf; wu = wu < 1 ? 1 : (wu > 2048 ? 2048 : wu);
    const unsigned weight = (unsigned)((float)wu * ew);
    unsigned selector = 0;
    const uint4* pb = reinterpret_cast<const uint4*>(blocks + (size_t)b * 16);
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
        const uint4 four = pb[q];
        const unsigned pxs[4] = { four.x, four.y, four.z, four.w };
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int r = pxs[j] & 0xff, g = (pxs[j] >> 8) & 0xff, bl = (pxs[j] >> 16) & 0xff;
            unsigned best = 0xffffffffu, sb = 0;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int lin = t == 0 ? 0 : (t == 1 ? 3 : (t == 2 ? 1 : 2));       // g_dxt1_to_linear
                const int s = reordered ? 3 - lin : lin;
                const int dr = r - cs[s][0], dg = g - cs[s][1], db = bl - cs[s][2];
                const unsigned e = (unsigned)(wr * dr * dr + wg * dg * dg + wb * db * db);
                if (e < best) { best = e; sb = (unsigned)s; }
            }
            selector = (selector << 2) | sb;
        }
    }
    block_selectors[b] = ((unsigned long long)selector << 32) | weight;
#pragma unroll
    for (int s = 0; s < 4; s++) block_values[(size_t)b * 4 + s] = (unsigned)cv[s][0] | ((unsigned)cv[s][1] << 8) | ((unsigned)cv[s][2] << 16) | 0xff000000u;
}

// alpha (crn_dxt_hc.cpp:1035-1100): one thread per (component a, block).  cluster_endpoints = first | second << 8 as the
// optimiser returned them, cluster_flags bit 0 = m_reordered.  block_selectors[a * n + b] = selector << 16 | weight,
// block_values[(a * n + b) * 8 ..] the eight alpha values in linear order.
__global__ void __launch_bounds__(256)
hc_alpha_blocks_kernel(const uint32_t* __restrict__ blocks, uint32_t n_blocks, int num_alpha, uint32_t comp0, uint32_t comp1, const uint32_t* __restrict__ block_cluster,
                       const uint32_t* __restrict__ cluster_endpoints, const uint32_t* __restrict__ cluster_flags, const uint8_t* __restrict__ block_encoding,
                       unsigned long long* __restrict__ block_selectors, uint8_t* __restrict__ block_values)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_blocks * (uint32_t)num_alpha) return;
    const uint32_t a = g / n_blocks, b = g % n_blocks, comp = a ? comp1 : comp0;
    const uint32_t c = block_cluster[g], ep = cluster_endpoints[c];
    const bool reordered = cluster_flags[c] & 1;
    const unsigned first = ep & 0xff, second = (ep >> 8) & 0xff;
    unsigned bv[8], av[8];
    if (first > second) dxt5a_values8(first, second, bv); else dxt5a_values6(first, second, bv);
    const int from_linear[8] = { 0, 2, 3, 4, 5, 6, 7, 1 }, to_linear[8] = { 0, 7, 1, 2, 3, 4, 5, 6 };
#pragma unroll
    for (int i = 0; i < 8; i++) av[i] = bv[from_linear[i]];
    const int delta = (int)first - (int)second;
    unsigned endpoint_weight = (unsigned)(delta * delta) >> 3;
    endpoint_weight = endpoint_weight < 1 ? 1 : (endpoint_weight > 2048 ? 2048 : endpoint_weight);
    const unsigned weight = (unsigned)((float)endpoint_weight * (1.15f + (1.0f - 1.15f) * ((float)block_encoding[b] / 7.0f)));
    unsigned long long selector = 0;
    for (int p = 0; p < 16; p++) {
        const int v = (blocks[(size_t)b * 16 + p] >> (8 * comp)) & 0xff;
        unsigned best = 0xffffffffu, sb = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const int s = reordered ? 7 - to_linear[t] : to_linear[t];
            const int d = v - (int)av[s];
            const unsigned e = (unsigned)(d >= 0 ? d : -d);
            if (e < best) { best = e; sb = (unsigned)s; }
        }
        selector = (selector << 3) | sb;
    }
    block_selectors[g] = (selector << 16) | weight;
#pragma unroll
    for (int i = 0; i < 8; i++) block_values[(size_t)g * 8 + i] = (uint8_t)av[i];
}

// refined alpha values of every block's cluster (crn_dxt_hc.cpp:1112-1128): ok ? values of the refined endpoints : the optimiser's
__global__ void __launch_bounds__(256)
hc_alpha_refined_values_kernel(uint32_t n_total, const uint32_t* __restrict__ block_cluster, const uint32_t* __restrict__ refined_endpoints, const uint8_t* __restrict__ refined_ok,
                               const uint8_t* __restrict__ block_values, uint8_t* __restrict__ block_values_accum)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_total) return;
    const uint32_t c = block_cluster[g];
    if (refined_ok[c]) {
        const unsigned first = refined_endpoints[c] & 0xffff, second = refined_endpoints[c] >> 16;
        unsigned bv[8];
        if (first > second) dxt5a_values8(first, second, bv); else dxt5a_values6(first, second, bv);
        const int from_linear[8] = { 0, 2, 3, 4, 5, 6, 7, 1 };
#pragma unroll
        for (int i = 0; i < 8; i++) block_values_accum[(size_t)g * 8 + i] = (uint8_t)bv[from_linear[i]];
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) block_values_accum[(size_t)g * 8 + i] = block_values[(size_t)g * 8 + i];
    }
}

// cluster pixel lists for the refiner: pixel g of the concatenated member (virtual) blocks and its selector out of the
// element the cluster optimiser wrote for that virtual block.  kind 0: DXT1 element (2-bit selectors at byte 4), kind 1:
// DXT5A element (3-bit selectors from bit 16).
__global__ void __launch_bounds__(256)
hc_gather_cluster_pixels_kernel(const uint32_t* __restrict__ vblocks, const uint32_t* __restrict__ members, uint32_t total_pixels, int kind,
                                const unsigned long long* __restrict__ elements, uint32_t* __restrict__ out_pixels, uint8_t* __restrict__ out_selectors)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_pixels) return;
    const uint32_t vb = members[g >> 4], p = g & 15;
    out_pixels[g] = vblocks[(size_t)vb * 16 + p];
    const unsigned long long e = elements[vb];
    out_selectors[g] = kind ? (uint8_t)((e >> (16 + 3 * p)) & 7) : (uint8_t)((e >> (32 + 2 * p)) & 3);
}

// grey copies of one channel: out[b][p] = value * 0x01010101 (color_quad_u8(value), crn_dxt_hc.cpp:1257)
__global__ void __launch_bounds__(256)
hc_grey_kernel(const uint32_t* __restrict__ pixels, size_t n_pixels, uint32_t comp, uint32_t* __restrict__ out)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_pixels) out[g] = ((pixels[g] >> (8 * comp)) & 0xff) * 0x01010101u;
}

// ---- selector training set on the device (create_*_selector_codebook, crn_dxt_hc.cpp:1379-1444, :1588-1660) ---------
// keys = m_block_selectors sorted ascending (selector << wshift | weight).  head[i] = 1 where a new selector starts.
__global__ void __launch_bounds__(256)
hc_sel_heads_kernel(const unsigned long long* __restrict__ keys, uint32_t n, int wshift, uint32_t* __restrict__ head)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    head[i] = (i < n && (i == 0 || (keys[i] >> wshift) != (keys[i - 1] >> wshift))) ? 1u : 0u;
}
// one thread per run: the 16-D vector of the selector (pixel p of the packed selector is vector component p) and the
// saturating sum of the run's weights
__global__ void __launch_bounds__(256)
hc_sel_vectors_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ head, const uint32_t* __restrict__ rank, uint32_t n, int kind,
                      float* __restrict__ vecs, uint32_t* __restrict__ wts)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !head[i]) return;
    const int bits = kind ? 3 : 2, wshift = kind ? 16 : 32;
    const unsigned long long wmask = kind ? 0xffffull : 0xffffffffull;
    unsigned long long sel = keys[i] >> wshift;
    unsigned long long w = keys[i] & wmask;
    for (uint32_t j = i + 1; j < n && !head[j]; j++) { w += keys[j] & wmask; if (w > 0xffffffffull) w = 0xffffffffull; }
    const uint32_t o = rank[i];
    wts[o] = (uint32_t)w;
    float v[16];
#pragma unroll
    for (int p = 0; p < 16; p++, sel >>= bits) {
        const unsigned sv = (unsigned)(sel & ((1u << bits) - 1));
        v[15 - p] = kind ? ((float)sv + 0.5f) * 0.125f : ((float)sv + 0.5f) * 0.25f;
    }
    float4* dst = reinterpret_cast<float4*>(vecs + (size_t)o * 16);
#pragma unroll
    for (int q = 0; q < 4; q++) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// ---- endpoint training set on the device (determine_color/alpha_endpoints, crn_dxt_hc.cpp:888-968, :1165-1244) --------

// compact (component, tile) list: entry i = a * num_tiles + t takes the vector of tile slot used[t] of plane a and the
// weight the reference gives it: (uint)(pixels * level weight) for colour, pixels for alpha
template <int D>
__global__ void __launch_bounds__(256)
hc_compact_tiles_kernel(const float* __restrict__ src, const uint32_t* __restrict__ used, const uint8_t* __restrict__ tile_npix, uint32_t n_slots, uint32_t num_tiles,
                        uint32_t total, int kind, HcLevelWeights LW, float* __restrict__ out_vecs, uint32_t* __restrict__ out_wts)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint32_t a = i / num_tiles, slot = used[i % num_tiles];
    const float* v = src + ((size_t)a * n_slots + slot) * D;
#pragma unroll
    for (int d = 0; d < D; d++) out_vecs[(size_t)i * D + d] = v[d];
    const unsigned np = tile_npix[slot];
    if (kind) out_wts[i] = np;
    else out_wts[i] = (uint32_t)((float)np * hc_level_weight(LW, slot));
}
__global__ void __launch_bounds__(256) hc_iota_kernel(uint32_t* __restrict__ p, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}
// sort key of one LSD pass: component `comp` of the vectors in the current order (non-negative floats order like their bits)
__global__ void __launch_bounds__(256)
hc_gather_key_kernel(const float* __restrict__ vecs, const uint32_t* __restrict__ perm, int D, int comp, uint32_t n, uint32_t* __restrict__ keys)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = __float_as_uint(vecs[(size_t)perm[i] * D + comp]);
}
template <int D>
__global__ void __launch_bounds__(256)
hc_vec_heads_kernel(const float* __restrict__ vecs, const uint32_t* __restrict__ perm, uint32_t n, uint32_t* __restrict__ head)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    unsigned h = 0;
    if (i < n) {
        h = i == 0;
        if (i) {
            const float* a = vecs + (size_t)perm[i] * D; const float* b = vecs + (size_t)perm[i - 1] * D;
#pragma unroll
            for (int d = 0; d < D; d++) h |= a[d] != b[d];
        }
    }
    head[i] = h;
}
template <int D>
__global__ void __launch_bounds__(256)
hc_vec_unique_kernel(const float* __restrict__ vecs, const uint32_t* __restrict__ wts, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ head,
                     const uint32_t* __restrict__ rank, uint32_t n, float* __restrict__ out_vecs, uint32_t* __restrict__ out_wts)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !head[i]) return;
    unsigned long long w = wts[perm[i]];
    for (uint32_t j = i + 1; j < n && !head[j]; j++) { w += wts[perm[j]]; if (w > 0xffffffffull) w = 0xffffffffull; }
    const uint32_t o = rank[i];
    out_wts[o] = (uint32_t)w;
    const float* v = vecs + (size_t)perm[i] * D;
#pragma unroll
    for (int d = 0; d < D; d++) out_vecs[(size_t)o * D + d] = v[d];
}

}  // namespace crn
