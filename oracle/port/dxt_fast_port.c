/* oracle/port/dxt_fast_port.c -- TEST INFRASTRUCTURE ONLY (see oracle_port.h).
 *
 * Restatement of the parts of crnlib::dxt_fast the clustered-DDS quantisers use
 * (reference crnlib/crn_dxt_fast.cpp): compress_color_block without the optional probe refinement
 * (:725-764 with refine == false, i.e. optimize_block_colors :138-265, determine_selectors /
 * match_block_colors :84-136 / :339-352, refine_block :270-336, compress_solid_block :713-723),
 * compress_alpha_block (:788-826), find_representative_colors (:855-995); and of the per-chunk tile
 * analysis of qdxt1::init / qdxt5::init that turns blocks into endpoint training vectors
 * (crnlib/crn_qdxt1.cpp:103-361, crnlib/crn_qdxt5.cpp:103-330; layouts and encodings from
 * crnlib/crn_dxt_hc_common.cpp:28-58).
 */
#include "oracle_port.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

extern const uint8_t* op_omatch_table(int which);

static int mul8(int a, int b) { int t = a * b + 128; return (t + (t >> 8)) >> 8; }
static int e5(int v) { return (v << 3) | (v >> 2); }
static int e6(int v) { return (v << 2) | (v >> 4); }
static unsigned pack_fast(const uint8_t* c) { return (unsigned)((mul8(c[0], 31) << 11) + (mul8(c[1], 63) << 5) + mul8(c[2], 31)); }
static int d2i(double x) { return (x > -2147483649.0 && x < 2147483648.0) ? (int)x : (int)0x80000000; }
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void eval_colors(int (*c)[3], unsigned c0, unsigned c1)
{
    c[0][0] = e5((c0 >> 11) & 31); c[0][1] = e6((c0 >> 5) & 63); c[0][2] = e5(c0 & 31);
    c[1][0] = e5((c1 >> 11) & 31); c[1][1] = e6((c1 >> 5) & 63); c[1][2] = e5(c1 & 31);
    for (int k = 0; k < 3; k++) { c[2][k] = (c[0][k] * 2 + c[1][k]) / 3; c[3][k] = (c[1][k] * 2 + c[0][k]) / 3; }
}

static int match_colors(uint32_t n, const uint8_t* px, int (*c)[3], uint8_t* sel)
{
    int dr = c[0][0] - c[1][0], dg = c[0][1] - c[1][1], db = c[0][2] - c[1][2];
    int stops[4];
    for (int i = 0; i < 4; i++) stops[i] = c[i][0] * dr + c[i][1] * dg + c[i][2] * db;
    int c0p = (stops[1] + stops[3]) >> 1, half = (stops[3] + stops[2]) >> 1, c3p = (stops[2] + stops[0]) >> 1;
    int status = 0;
    for (uint32_t i = 0; i < n; i++) {
        int dot = px[4 * i] * dr + px[4 * i + 1] * dg + px[4 * i + 2] * db;
        uint8_t s = dot < half ? (dot < c0p ? 1 : 3) : (dot < c3p ? 2 : 0);
        sel[i] = s;
        if (s != sel[0]) status = 1;
    }
    return status;
}

static int determine_selectors(uint32_t n, const uint8_t* px, unsigned min16, unsigned max16, uint8_t* sel)
{
    if (max16 != min16) { int c[4][3]; eval_colors(c, min16, max16); return match_colors(n, px, c, sel); }
    memset(sel, 0, n);
    return 0;
}

static int optimize_block_colors(uint32_t n, const uint8_t* px, unsigned* max16, unsigned* min16, unsigned* ave)
{
    int mn[3], mx[3];
    for (int ch = 0; ch < 3; ch++) {
        int64_t mu = px[ch]; int lo = px[ch], hi = px[ch];
        for (uint32_t i = 1; i < n; i++) { int v = px[4 * i + ch]; mu += v; if (v < lo) lo = v; if (v > hi) hi = v; }
        ave[ch] = (unsigned)((mu + (n / 2)) / n); mn[ch] = lo; mx[ch] = hi;
    }
    if (mn[0] == mx[0] && mn[1] == mx[1] && mn[2] == mx[2]) return 0;
    double cov[6] = { 0, 0, 0, 0, 0, 0 };
    for (uint32_t i = 0; i < n; i++) {
        double r = (int)px[4 * i] - (int)ave[0], g = (int)px[4 * i + 1] - (int)ave[1], b = (int)px[4 * i + 2] - (int)ave[2];
        cov[0] += r * r; cov[1] += r * g; cov[2] += r * b; cov[3] += g * g; cov[4] += g * b; cov[5] += b * b;
    }
    double covf[6], vfr = mx[0] - mn[0], vfg = mx[1] - mn[1], vfb = mx[2] - mn[2];
    for (int i = 0; i < 6; i++) covf[i] = cov[i] * (1.0f / 255.0f);
    for (int it = 0; it < 4; it++) {
        double r = vfr * covf[0] + vfg * covf[1] + vfb * covf[2];
        double g = vfr * covf[1] + vfg * covf[3] + vfb * covf[4];
        double b = vfr * covf[2] + vfg * covf[4] + vfb * covf[5];
        vfr = r; vfg = g; vfb = b;
    }
    double magn = fabs(vfr) > fabs(vfg) ? fabs(vfr) : fabs(vfg);
    magn = magn > fabs(vfb) ? magn : fabs(vfb);
    int v_r, v_g, v_b;
    if (magn < 4.0f) { v_r = 148; v_g = 300; v_b = 58; }
    else { magn = 512.0f / magn; vfr *= magn; vfg *= magn; vfb *= magn; v_r = d2i(vfr); v_g = d2i(vfg); v_b = d2i(vfb); }
    int mind = px[0] * v_r + px[1] * v_g + px[2] * v_b, maxd = mind;
    const uint8_t* minp = px; const uint8_t* maxp = px;
    for (uint32_t i = 1; i < n; i++) {
        int dot = px[4 * i] * v_r + px[4 * i + 1] * v_g + px[4 * i + 2] * v_b;
        if (dot < mind) { mind = dot; minp = px + 4 * i; }
        if (dot > maxd) { maxd = dot; maxp = px + 4 * i; }
    }
    *max16 = pack_fast(maxp); *min16 = pack_fast(minp);
    return 1;
}

static int refine_block(uint32_t n, const uint8_t* px, unsigned* max16, unsigned* min16, const uint8_t* sel)
{
    static const int w1Tab[4] = { 3, 0, 2, 1 }, p0[4] = { 0, 0, 2, 2 }, p1[4] = { 0, 9, 1, 4 }, p2[4] = { 9, 0, 4, 1 };
    double ak0 = 0, ak1 = 0, ak2 = 0, A1r = 0, A1g = 0, A1b = 0, A2r = 0, A2g = 0, A2b = 0;
    for (uint32_t i = 0; i < n; i++) {
        double r = px[4 * i], g = px[4 * i + 1], b = px[4 * i + 2];
        int step = sel[i], w1 = w1Tab[step];
        ak0 += p0[step]; ak1 += p1[step]; ak2 += p2[step];
        A1r += w1 * r; A1g += w1 * g; A1b += w1 * b; A2r += r; A2g += g; A2b += b;
    }
    A2r = 3 * A2r - A1r; A2g = 3 * A2g - A1g; A2b = 3 * A2b - A1b;
    double xx = ak2, yy = ak1, xy = ak0, t = xx * yy - xy * xy;
    if (!yy || !xx || fabs(t) < .0000125f) return 0;
    double frb = (3.0f * 31.0f / 255.0f) / t, fg = frb * (63.0f / 31.0f);
    unsigned oldMin = *min16, oldMax = *max16;
    *max16 = (unsigned)clampi(d2i((A1r * yy - A2r * xy) * frb + 0.5f), 0, 31) << 11;
    *max16 |= (unsigned)clampi(d2i((A1g * yy - A2g * xy) * fg + 0.5f), 0, 63) << 5;
    *max16 |= (unsigned)clampi(d2i((A1b * yy - A2b * xy) * frb + 0.5f), 0, 31);
    *min16 = (unsigned)clampi(d2i((A2r * xx - A1r * xy) * frb + 0.5f), 0, 31) << 11;
    *min16 |= (unsigned)clampi(d2i((A2g * xx - A1g * xy) * fg + 0.5f), 0, 63) << 5;
    *min16 |= (unsigned)clampi(d2i((A2b * xx - A1b * xy) * frb + 0.5f), 0, 31);
    return oldMin != *min16 || oldMax != *max16;
}

static void solid_block(uint32_t n, const unsigned* ave, unsigned* lo, unsigned* hi, uint8_t* sel)
{
    const uint8_t* o5 = op_omatch_table(0); const uint8_t* o6 = op_omatch_table(1);
    memset(sel, 2, n);
    *lo = (unsigned)((o5[2 * ave[0]] << 11) | (o6[2 * ave[1]] << 5) | o5[2 * ave[2]]);
    *hi = (unsigned)((o5[2 * ave[0] + 1] << 11) | (o6[2 * ave[1] + 1] << 5) | o5[2 * ave[2] + 1]);
}

/* dxt_fast::compress_color_block(n, ..., refine = false), crn_dxt_fast.cpp:725-764 */
void op_fast_color_block(uint32_t n, const uint8_t* px, uint32_t* low16, uint32_t* high16, uint8_t* sel)
{
    unsigned ave[3], lo = 0, hi = 0;
    if (!optimize_block_colors(n, px, &lo, &hi, ave)) solid_block(n, ave, &lo, &hi, sel);
    else if (!determine_selectors(n, px, lo, hi, sel)) solid_block(n, ave, &lo, &hi, sel);
    else if (refine_block(n, px, &lo, &hi, sel)) determine_selectors(n, px, lo, hi, sel);
    /* NB: the reference calls the (max16, min16) pair (low16, high16): low16 receives max16 */
    if (lo < hi) { unsigned t = lo; lo = hi; hi = t; for (uint32_t i = 0; i < n; i++) sel[i] ^= 1; }
    *low16 = lo; *high16 = hi;
}

/* dxt_fast::compress_alpha_block, crn_dxt_fast.cpp:788-826 */
void op_fast_alpha_block(uint32_t n, const uint8_t* px, uint32_t comp, uint32_t* low8, uint32_t* high8, uint8_t* sel)
{
    int mn = px[comp], mx = px[comp];
    for (uint32_t i = 1; i < n; i++) { int v = px[4 * i + comp]; if (v < mn) mn = v; if (v > mx) mx = v; }
    *low8 = (uint32_t)mx; *high8 = (uint32_t)mn;
    int dist = mx - mn, bias = mn * 7 - (dist >> 1), dist4 = dist * 4, dist2 = dist * 2;
    for (uint32_t i = 0; i < n; i++) {
        int a = px[4 * i + comp] * 7 - bias, ind, t;
        t = (dist4 - a) >> 31; ind = t & 4; a -= dist4 & t;
        t = (dist2 - a) >> 31; ind += t & 2; a -= dist2 & t;
        t = (dist - a) >> 31; ind += t & 1;
        ind = -ind & 7;
        ind ^= (2 > ind);
        sel[i] = (uint8_t)ind;
    }
}

/* dxt_fast::find_representative_colors, crn_dxt_fast.cpp:855-995 */
void op_find_representative_colors(uint32_t n, const uint8_t* px, uint8_t* lo, uint8_t* hi)
{
    uint64_t ave64[3] = { 0, 0, 0 };
    for (uint32_t i = 0; i < n; i++) for (int k = 0; k < 3; k++) ave64[k] += px[4 * i + k];
    unsigned ave[3];
    for (int k = 0; k < 3; k++) ave[k] = (unsigned)((ave64[k] + (n / 2)) / n);
    int fd = -1; uint32_t fi = 0;
    for (uint32_t i = 0; i < n; i++) {
        int r = px[4 * i] - (int)ave[0], g = px[4 * i + 1] - (int)ave[1], b = px[4 * i + 2] - (int)ave[2];
        int d = r * r + g * g + b * b;
        if (d > fd) { fd = d; fi = i; }
    }
    uint8_t lc[4], hc[4];
    memcpy(lc, px + 4 * fi, 4);
    int od = -1; uint32_t oi = 0;
    for (uint32_t i = 0; i < n; i++) {
        int r = px[4 * i] - lc[0], g = px[4 * i + 1] - lc[1], b = px[4 * i + 2] - lc[2];
        int d = r * r + g * g + b * b;
        if (d > od) { od = d; oi = i; }
    }
    memcpy(hc, px + 4 * oi, 4);
    for (int k = 0; k < 3; k++) { lc[k] = (uint8_t)((lc[k] + ave[k]) >> 1); hc[k] = (uint8_t)((hc[k] + ave[k]) >> 1); }
    for (int it = 0; it < 4; it++) {
        if (lc[0] == hc[0] && lc[1] == hc[1] && lc[2] == hc[2]) break;
        uint64_t nc[2][3] = { { 0, 0, 0 }, { 0, 0, 0 } }; unsigned w[2] = { 0, 0 };
        int vr = hc[0] - lc[0], vg = hc[1] - lc[1], vb = hc[2] - lc[2];
        int lod = vr * lc[0] + vg * lc[1] + vb * lc[2], hid = vr * hc[0] + vg * hc[1] + vb * hc[2], mid = lod + hid;
        vr *= 2; vg *= 2; vb *= 2;
        for (uint32_t i = 0; i < n; i++) {
            int dot = px[4 * i] * vr + px[4 * i + 1] * vg + px[4 * i + 2] * vb;
            unsigned m = dot > mid;
            nc[m][0] += px[4 * i]; nc[m][1] += px[4 * i + 1]; nc[m][2] += px[4 * i + 2]; w[m]++;
        }
        if (!w[0] || !w[1]) break;
        uint8_t n8[2][3];
        for (int j = 0; j < 2; j++) for (int k = 0; k < 3; k++) n8[j][k] = (uint8_t)((nc[j][k] + (w[j] / 2)) / w[j]);
        if (!memcmp(n8[0], lc, 3) && !memcmp(n8[1], hc, 3)) break;
        memcpy(lc, n8[0], 3); memcpy(hc, n8[1], 3);
    }
    unsigned en0 = 0, en1 = 0;
    for (int k = 0; k < 3; k++) { en0 += lc[k] * lc[k]; en1 += hc[k] * hc[k]; }
    if (en0 > en1) { uint8_t t[4]; memcpy(t, lc, 4); memcpy(lc, hc, 4); memcpy(hc, t, 4); }
    memcpy(lo, lc, 3); memcpy(hi, hc, 3);
}

/* chunk tile layouts and encodings, crn_dxt_hc_common.cpp:28-58: {x, y, w, h, layout index} */
static const uint8_t LAYOUT[9][4] = { { 0, 0, 8, 8 }, { 0, 0, 8, 4 }, { 0, 4, 8, 4 }, { 0, 0, 4, 8 }, { 4, 0, 4, 8 },
                                      { 0, 0, 4, 4 }, { 4, 0, 4, 4 }, { 0, 4, 4, 4 }, { 4, 4, 4, 4 } };
static const uint8_t ENC_NT[8] = { 1, 2, 2, 3, 3, 3, 3, 4 };
static const uint8_t ENC_TILES[8][4] = { { 0 }, { 1, 2 }, { 3, 4 }, { 1, 7, 8 }, { 2, 5, 6 }, { 3, 6, 8 }, { 4, 5, 7 }, { 5, 6, 7, 8 } };

/* Endpoint training vectors of qdxt1::init (kind 0: 6 bytes lo.rgb hi.rgb) / qdxt5::init (kind 1: 2 bytes
 * lo, hi of component `comp`) + weights, for blocks laid out as mipmapped_texture::qdxt_pack_init does
 * (mips[i] = first_block, block_width, block_height).  hierarchical == 0 is the per-block branch. */
void op_qdxt_training(int kind, uint32_t comp, const uint8_t* blocks, uint32_t n_blocks, const uint32_t* mips, uint32_t num_mips,
                      int hierarchical, uint8_t* out_vecs, uint32_t* out_weights, uint8_t* out_encoding)
{
    const int D = kind ? 2 : 6;
    if (!hierarchical || !num_mips) {
        for (uint32_t b = 0; b < n_blocks; b++) {
            uint8_t px[64], l[3], h[3];
            memcpy(px, blocks + 64 * (size_t)b, 64);
            if (kind) for (int i = 0; i < 16; i++) { uint8_t a = px[4 * i + comp]; px[4 * i] = px[4 * i + 1] = px[4 * i + 2] = a; }
            op_find_representative_colors(16, px, l, h);
            uint32_t dist, w;
            if (kind) { int d = (int)l[0] - (int)h[0]; dist = (uint32_t)(d * d); w = dist / 8; out_vecs[2 * b] = l[0]; out_vecs[2 * b + 1] = h[0]; }
            else { dist = 0; for (int k = 0; k < 3; k++) { int d = (int)l[k] - (int)h[k]; dist += (uint32_t)(d * d); } w = dist / 5000; memcpy(out_vecs + 6 * b, l, 3); memcpy(out_vecs + 6 * b + 3, h, 3); }
            out_weights[b] = w < 1 ? 1 : (w > 8 ? 8 : w);
        }
        return;
    }
    uint32_t chunk_counter = 0;
    for (uint32_t level = 0; level < num_mips; level++) {
        const uint32_t first = mips[3 * level], bw = mips[3 * level + 1], bh = mips[3 * level + 2];
        const uint32_t ncx = (bw + 1) / 2, ncy = (bh + 1) / 2, lw = bw * 4, lh = bh * 4;
        float derating = kind ? 2.4f : 1.5f;
        if (level && derating > .25f) { float d = derating / powf(kind ? 3.0f : 3.1f, (float)level); derating = d > .25f ? d : .25f; }
        for (uint32_t cy = 0; cy < ncy; cy++)
            for (uint32_t cx = 0; cx < ncx; cx++, chunk_counter++) {
                uint8_t chunk[64][4];
                for (uint32_t y = 0; y < 8; y++) {
                    uint32_t py = cy * 8 + y < lh - 1 ? cy * 8 + y : lh - 1;
                    for (uint32_t x = 0; x < 8; x++) {
                        uint32_t pxx = cx * 8 + x < lw - 1 ? cx * 8 + x : lw - 1;
                        const uint8_t* src = blocks + 64 * (size_t)(first + (py >> 2) * bw + (pxx >> 2)) + 4 * ((py & 3) * 4 + (pxx & 3));
                        memcpy(chunk[x + y * 8], src, 4);
                    }
                }
                uint64_t lerr[9];
                for (int l = 0; l < 9; l++) {
                    const uint32_t xo = LAYOUT[l][0], yo = LAYOUT[l][1], w = LAYOUT[l][2], h = LAYOUT[l][3], n = w * h;
                    uint8_t lp[64][4], sel[64];
                    for (uint32_t y = 0; y < h; y++) for (uint32_t x = 0; x < w; x++) memcpy(lp[x + y * w], chunk[(xo + x) + (yo + y) * 8], 4);
                    uint64_t err = 0;
                    if (!kind) {
                        uint32_t lo, hi; int c[4][3];
                        op_fast_color_block(n, &lp[0][0], &lo, &hi, sel);
                        /* dxt1_block::get_block_colors: 4-colour iff lo > hi, else 3-colour with transparent black */
                        eval_colors(c, lo, hi);
                        if (lo <= hi) for (int k = 0; k < 3; k++) { c[2][k] = (c[0][k] + c[1][k]) >> 1; c[3][k] = 0; }
                        for (uint32_t i = 0; i < n; i++) for (int k = 0; k < 3; k++) { int d = (int)lp[i][k] - c[sel[i]][k]; err += (uint64_t)(d * d); }
                    } else {
                        uint32_t lo, hi, v[8];
                        op_fast_alpha_block(n, &lp[0][0], comp, &lo, &hi, sel);
                        if (lo > hi) { v[0] = lo; v[1] = hi; for (int k = 1; k < 7; k++) v[k + 1] = (lo * (7 - k) + hi * k) / 7; }
                        else { v[0] = lo; v[1] = hi; for (int k = 1; k < 5; k++) v[k + 1] = (lo * (5 - k) + hi * k) / 5; v[6] = 0; v[7] = 255; }
                        for (uint32_t i = 0; i < n; i++) { int d = (int)lp[i][comp] - (int)v[sel[i]]; err += (uint64_t)(d * d); }
                    }
                    lerr[l] = err;
                }
                double best = -1.0f; uint32_t best_e = 0;
                for (uint32_t e = 0; e < 8; e++) {
                    double total = 0;
                    for (uint32_t t = 0; t < ENC_NT[e]; t++) total += (double)lerr[ENC_TILES[e][t]];
                    double ms = total * (kind ? (1.0f / 64.0f) : (1.0f / (64.0f * 3.0f)));
                    double rms = sqrt(ms), psnr = 999999.0f;
                    if (ms) { psnr = log10(255.0f / rms) * 20.0f; psnr = psnr < 0.0f ? 0.0f : (psnr > 500.0f ? 500.0f : psnr); }
                    float der = 0.0f + (derating - 0.0f) * ((ENC_NT[e] - 1) / 3.0f);
                    psnr = psnr - der;
                    if (psnr > best) { best = psnr; best_e = e; }
                }
                if (out_encoding) out_encoding[chunk_counter] = (uint8_t)best_e;
                for (uint32_t t = 0; t < ENC_NT[best_e]; t++) {
                    const int l = ENC_TILES[best_e][t];
                    const uint32_t xo = LAYOUT[l][0], yo = LAYOUT[l][1], w = LAYOUT[l][2], h = LAYOUT[l][3];
                    uint8_t tp[64][4], lo3[3], hi3[3];
                    for (uint32_t y = 0; y < h; y++) for (uint32_t x = 0; x < w; x++) {
                        const uint8_t* s = chunk[(xo + x) + (yo + y) * 8];
                        if (kind) { tp[x + y * w][0] = tp[x + y * w][1] = tp[x + y * w][2] = s[comp]; tp[x + y * w][3] = 255; }
                        else memcpy(tp[x + y * w], s, 4);
                    }
                    op_find_representative_colors(w * h, &tp[0][0], lo3, hi3);
                    uint32_t dist = 0, wgt;
                    if (kind) { int d = (int)lo3[0] - (int)hi3[0]; dist = (uint32_t)(d * d); wgt = dist / 8; }
                    else { for (int k = 0; k < 3; k++) { int d = (int)lo3[k] - (int)hi3[k]; dist += (uint32_t)(d * d); } wgt = dist / 5000; }
                    wgt = wgt < 1 ? 1 : (wgt > 8 ? 8 : wgt);
                    for (uint32_t y = 0; y < (h >> 2); y++) {
                        uint32_t by = cy * 2 + y + (yo >> 2);
                        if (by >= bh) continue;
                        for (uint32_t x = 0; x < (w >> 2); x++) {
                            uint32_t bx = cx * 2 + x + (xo >> 2);
                            if (bx >= bw) break;
                            uint32_t bi = first + bx + by * bw;
                            if (kind) { out_vecs[D * bi] = lo3[0]; out_vecs[D * bi + 1] = hi3[0]; }
                            else { memcpy(out_vecs + D * bi, lo3, 3); memcpy(out_vecs + D * bi + 3, hi3, 3); }
                            out_weights[bi] = wgt;
                        }
                    }
                }
            }
    }
    (void)n_blocks;
}
