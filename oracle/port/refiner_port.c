/* oracle/port/refiner_port.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of crnlib::dxt_endpoint_refiner (crnlib/crn_dxt_endpoint_refiner.cpp:36-301): with the
 * selectors of a pixel cluster fixed, a least-squares start (double, Squish's solver, :51-128) followed by an integer
 * search over the closed-form error  sum_s hist[s] v_s^2 - D2[s] v_s + DD[s]  (DXT5A: window around the start,
 * :146-201; DXT1: up to eight rounds over the 26+26 lattice neighbours of the current pair, :203-301).
 * Also the nearest-codebook search of dxt_hc (crn_dxt_hc.cpp:836-886, :1132-1163).
 * Pinned against the unmodified reference by tests/test_refiner_cpu.py (ref_refine in oracle/ref_shim.cpp). */
#include "oracle_port.h"
#include <string.h>

static const uint8_t k_dxt1_to_linear[4] = { 0, 3, 1, 2 };               /* crn_dxt.cpp:39 */
static const uint8_t k_dxt5_to_linear[8] = { 0, 7, 1, 2, 3, 4, 5, 6 };   /* crn_dxt.cpp:34 */

static float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

typedef struct { uint32_t low, high; uint64_t err; } refine_best;

static void refine_dxt5(const uint8_t* px, uint32_t n, const uint8_t* sel, uint32_t comp, float l0, float h0, refine_best* r)
{   /* :146-201 */
    const uint8_t L0 = (uint8_t)clampi((int)(l0 * 256.0f), 0, 255), H0 = (uint8_t)clampi((int)(h0 * 256.0f), 0, 255);
    uint64_t hist[8] = { 0 }, D2[8] = { 0 }, DD[8] = { 0 };
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t a = px[4 * i + comp], s = sel[i];
        hist[s]++; D2[s] += a * 2; DD[s] += a * a;
    }
    uint16_t sols[530];
    uint32_t cnt = 0;
    sols[cnt++] = (uint16_t)(L0 == H0 ? (H0 ? ((H0 - 1) << 8 | L0) : 1) : (L0 > H0 ? (H0 << 8 | L0) : (L0 << 8 | H0)));
    const uint8_t minL = L0 <= 11 ? 0 : L0 - 11, maxL = L0 >= 244 ? 255 : L0 + 11;
    const uint8_t minH = H0 <= 11 ? 0 : H0 - 11, maxH = H0 >= 244 ? 255 : H0 + 11;
    for (uint32_t L = minL; L <= maxL; L++)
        for (uint32_t H = minH; H <= maxH; H++)
            if ((maxH < L || L <= H || H < minL) && (L != L0 || H != H0) && (L != H0 || H != L0))
                sols[cnt++] = (uint16_t)(L == H ? (H ? ((H - 1) << 8 | L) : 1) : (L > H ? (H << 8 | L) : (L << 8 | H)));
    for (uint32_t i = 0; i < cnt; i++) {
        const uint32_t l = sols[i] & 0xFF, h = sols[i] >> 8;
        const uint32_t v[8] = { l, h, (l * 6 + h) / 7, (l * 5 + h * 2) / 7, (l * 4 + h * 3) / 7, (l * 3 + h * 4) / 7, (l * 2 + h * 5) / 7, (l + h * 6) / 7 };
        uint64_t e = 0;
        for (uint32_t s = 0; s < 8; s++) e += hist[s] * v[s] * v[s] - D2[s] * v[s] + DD[s];
        if (e < r->err) { r->low = l; r->high = h; r->err = e; if (!e) return; }
    }
}

static void colors4(uint32_t c0, uint32_t c1, uint32_t out[4][3])
{   /* dxt1_block::get_block_colors4, crn_dxt.cpp:247-260, unpack_color(scaled) */
    uint32_t a[3] = { (c0 >> 11) & 31, (c0 >> 5) & 63, c0 & 31 }, b[3] = { (c1 >> 11) & 31, (c1 >> 5) & 63, c1 & 31 };
    a[0] = a[0] << 3 | a[0] >> 2; a[1] = a[1] << 2 | a[1] >> 4; a[2] = a[2] << 3 | a[2] >> 2;
    b[0] = b[0] << 3 | b[0] >> 2; b[1] = b[1] << 2 | b[1] >> 4; b[2] = b[2] << 3 | b[2] >> 2;
    for (int c = 0; c < 3; c++) { out[0][c] = a[c]; out[1][c] = b[c]; out[2][c] = (a[c] * 2 + b[c]) / 3; out[3][c] = (b[c] * 2 + a[c]) / 3; }
}

static int cmp_u32(const void* a, const void* b) { const uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b; return x < y ? -1 : x > y; }
#include <stdlib.h>

static void refine_dxt1(const uint8_t* px, uint32_t n, const uint8_t* sel, int perceptual, const float* l, const float* h, refine_best* r)
{   /* :203-301 */
    uint32_t L0 = (uint32_t)(clampi((int)(l[0] * 32.0f), 0, 31) << 11 | clampi((int)(l[1] * 64.0f), 0, 63) << 5 | clampi((int)(l[2] * 32.0f), 0, 31));
    uint32_t H0 = (uint32_t)(clampi((int)(h[0] * 32.0f), 0, 31) << 11 | clampi((int)(h[1] * 64.0f), 0, 63) << 5 | clampi((int)(h[2] * 32.0f), 0, 31));
    uint64_t hist[4] = { 0 }, D2[4][3], DD[4][3];
    memset(D2, 0, sizeof(D2)); memset(DD, 0, sizeof(DD));
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t s = sel[i];
        hist[s]++;
        for (int c = 0; c < 3; c++) { const uint32_t p = px[4 * i + c]; D2[s][c] += p * 2; DD[s][c] += p * p; }
    }
    const int preserveL = hist[0] + hist[2] > hist[1] + hist[3];
    int improved = 1;
    uint32_t sols[54];
    for (uint32_t it = 8; improved && it; it--) {
        improved = 0;
        uint32_t cnt = 0;
        for (int pass = 0; pass < 2; pass++) {                       /* neighbours of L0 against H0, then of H0 against L0 */
            const uint32_t C = pass ? H0 : L0, O = pass ? L0 : H0;
            const uint32_t b0 = C & 31, g0 = C >> 5 & 63, r0 = C >> 11 & 31;
            for (uint32_t b = b0 ? b0 - 1 : b0; b <= b0 + 1 && b <= 31; b++)
                for (uint32_t g = g0 ? g0 - 1 : g0; g <= g0 + 1 && g <= 63; g++)
                    for (uint32_t rr = r0 ? r0 - 1 : r0; rr <= r0 + 1 && rr <= 31; rr++) {
                        const uint32_t X = rr << 11 | g << 5 | b;
                        if (X != C) sols[cnt++] = X > O ? X | O << 16 : O | X << 16;
                    }
        }
        qsort(sols, cnt, sizeof(uint32_t), cmp_u32);
        for (uint32_t i = 0; i < cnt; i++) {
            if (i && sols[i] == sols[i - 1]) continue;
            uint32_t L = sols[i] & 0xFFFF, H = sols[i] >> 16;
            if (L == H) {                                            /* uint16 arithmetic in the reference */
                L = (L + (!preserveL ? ((~L & 0x1F) ? 0x1 : (~L & 0xF800) ? 0x800 : (~L & 0x7E0) ? 0x20 : 0) : (!L ? 0x1 : 0))) & 0xFFFF;
                H = (H - (preserveL ? ((H & 0x1F) ? 0x1 : (H & 0xF800) ? 0x800 : (H & 0x7E0) ? 0x20 : 0) : (H == 0xFFFF ? 0x1 : 0))) & 0xFFFF;
            }
            uint32_t bc[4][3];
            colors4(L, H, bc);
            uint64_t e = 0;
            for (uint32_t s = 0; s < 4; s++) {
                uint64_t d[3];
                for (int c = 0; c < 3; c++) d[c] = hist[s] * bc[s][c] * bc[s][c] - D2[s][c] * bc[s][c] + DD[s][c];
                e += perceptual ? d[0] * 8 + d[1] * 25 + d[2] : d[0] + d[1] + d[2];
            }
            if (e < r->err) {
                r->low = L0 = L; r->high = H0 = H; r->err = e;
                if (!e) return;
                improved = 1;
            }
        }
    }
}

int op_refine(int dxt1_selectors, int perceptual, uint32_t comp, const uint8_t* px, uint32_t n, const uint8_t* sel,
              uint64_t error_to_beat, uint32_t* low, uint32_t* high, uint64_t* error)
{   /* dxt_endpoint_refiner::refine, :36-144 */
    if (!n) return 0;
    refine_best r = { 0, 0, UINT64_MAX };
    double a2 = 0.0f, b2 = 0.0f, ab = 0.0f, ax[3] = { 0, 0, 0 }, bx[3] = { 0, 0, 0 }, first[3] = { 0, 0, 0 };
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t c = sel[i];
        double k;
        if (dxt1_selectors) k = k_dxt1_to_linear[c] * 1.0f / 3.0f; else k = k_dxt5_to_linear[c] * 1.0f / 7.0f;
        const double alpha = 1.0f - k, beta = k;
        double x[3];
        if (dxt1_selectors) { x[0] = px[4 * i] * 1.0f / 255.0f; x[1] = px[4 * i + 1] * 1.0f / 255.0f; x[2] = px[4 * i + 2] * 1.0f / 255.0f; }
        else x[0] = x[1] = x[2] = px[4 * i + comp] / 255.0f;
        if (!i) { first[0] = x[0]; first[1] = x[1]; first[2] = x[2]; }
        a2 += alpha * alpha; b2 += beta * beta; ab += alpha * beta;
        for (int j = 0; j < 3; j++) { ax[j] += alpha * x[j]; bx[j] += beta * x[j]; }
    }
    double a[3], b[3];
    if (b2 == 0.0f) { for (int j = 0; j < 3; j++) { a[j] = ax[j] / a2; b[j] = 0; } }
    else if (a2 == 0.0f) { for (int j = 0; j < 3; j++) { a[j] = 0; b[j] = bx[j] / b2; } }
    else {
        const double factor = a2 * b2 - ab * ab;
        if (factor != 0.0f) for (int j = 0; j < 3; j++) { a[j] = (ax[j] * b2 - bx[j] * ab) / factor; b[j] = (bx[j] * a2 - ax[j] * ab) / factor; }
        else for (int j = 0; j < 3; j++) { a[j] = first[j]; b[j] = first[j]; }
    }
    float l[3], h[3];
    for (int j = 0; j < 3; j++) { l[j] = clampf((float)a[j], 0.0f, 1.0f); h[j] = clampf((float)b[j], 0.0f, 1.0f); }
    if (dxt1_selectors) refine_dxt1(px, n, sel, perceptual, l, h, &r); else refine_dxt5(px, n, sel, comp, l[0], h[0], &r);
    *low = r.low; *high = r.high; *error = r.err;
    return r.err < error_to_beat;
}

/* dxt_hc::determine_color_endpoint_clusters_task (crn_dxt_hc.cpp:836-886; dims 6) and
 * determine_alpha_endpoint_clusters_task (:1132-1163; dims 2): index of the first codebook entry at minimum float
 * squared distance, the sum taken in component order.  The reference skips entries whose partial sum already exceeds
 * the distance to the tree-search leaf; such an entry cannot be the minimum, so the result is the plain first arg-min.
 * Pinned against the reference's own task, run on a tree_clusterizer codebook (ref_hc_nearest_codebook in oracle/ref_shim.cpp,
 * tests/test_refiner_cpu.py). */
void op_nearest_codebook(uint32_t dims, const float* vecs, uint32_t n, const float* codebook, uint32_t k, uint32_t* out)
{
    for (uint32_t i = 0; i < n; i++) {
        const float* v = vecs + (size_t)i * dims;
        float best = 1.0e+37f;    /* math::cNearlyInfinite, crn_math.h:38 */
        uint32_t bi = 0;
        for (uint32_t j = 0; j < k; j++) {
            const float* c = codebook + (size_t)j * dims;
            float d = 0;
            for (uint32_t t = 0; t < dims; t++) { const float e = c[t] - v[t]; d += e * e; }
            if (d < best) { best = d; bi = j; if (best == 0.0f) break; }
        }
        out[i] = bi;
    }
}

/* dxt_hc::create_color_selector_codebook_task (kind 0, crn_dxt_hc.cpp:1306-1360) / create_alpha_selector_codebook_task
 * (kind 1, :1516-1586) over all blocks, then the re-vote tail of create_color/alpha_selector_codebook (:1488-1503,
 * :1702-1720).  values: per block 4 RGBA8 colours (kind 0) or 8 alpha values (kind 1); values_accum: optional (kind 1). */
void op_assign_selectors(int kind, int perceptual, uint32_t comp, const uint8_t* blocks, uint32_t n, const uint8_t* values, const uint8_t* values_accum,
                         const uint64_t* codebook, uint32_t K, uint32_t* best_index, uint64_t* refined, uint8_t* used)
{
    const uint32_t V = kind ? 8 : 4;
    uint32_t* tot = (uint32_t*)calloc((size_t)K * 16 * V, 4);
    memset(used, 0, K);
    for (uint32_t b = 0; b < n; b++) {
        uint32_t E[16][8];
        const uint8_t* px = blocks + (size_t)b * 64;
        for (int pass = 0; pass < 2; pass++) {
            const uint8_t* vals = (pass && kind && values_accum ? values_accum : values) + (size_t)b * (kind ? 8 : 16);
            for (uint32_t p = 0; p < 16; p++)
                for (uint32_t s = 0; s < V; s++) {
                    if (!kind) {
                        const int dr = (int)px[4 * p] - vals[4 * s], dg = (int)px[4 * p + 1] - vals[4 * s + 1], db = (int)px[4 * p + 2] - vals[4 * s + 2];
                        E[p][s] = perceptual ? (uint32_t)(8 * dr * dr) + (uint32_t)(25 * dg * dg) + (uint32_t)(db * db) : (uint32_t)(dr * dr + dg * dg + db * db);
                    } else { const int d = (int)px[4 * p + comp] - vals[s]; E[p][s] = (uint32_t)(d * d); }
                }
            if (pass) break;
            uint32_t best = 0, best_err = UINT32_MAX;
            for (uint32_t s = 0; s < K; s++) {
                uint64_t sel = codebook[s];
                uint32_t e = 0;
                for (uint32_t p = 0; p < 16; p++, sel >>= (kind ? 3 : 2)) e += E[p][sel & (V - 1)];
                if (e < best_err) { best_err = e; best = s; }
            }
            best_index[b] = best;
            if (!(kind && values_accum)) break;
        }
        uint32_t* t = tot + (size_t)best_index[b] * 16 * V;
        for (uint32_t p = 0; p < 16; p++) for (uint32_t s = 0; s < V; s++) t[p * V + s] += E[p][s];
        used[best_index[b]] = 1;
    }
    for (uint32_t i = 0; i < K; i++) {
        uint64_t out = 0;
        for (uint32_t p = 0; p < 16; p++) {
            const uint32_t* e = tot + ((size_t)i * 16 + p) * V;
            uint32_t s;
            if (!kind) { const uint32_t s03 = e[3] < e[0] ? 3 : 0, s12 = e[2] < e[1] ? 2 : 1; s = e[s12] < e[s03] ? s12 : s03; }
            else {
                const uint32_t s07 = e[7] < e[0] ? 7 : 0, s12 = e[2] < e[1] ? 2 : 1, s34 = e[4] < e[3] ? 4 : 3, s56 = e[6] < e[5] ? 6 : 5;
                const uint32_t s02 = e[s12] < e[s07] ? s12 : s07, s36 = e[s56] < e[s34] ? s56 : s34;
                s = e[s36] < e[s02] ? s36 : s02;
            }
            out |= (uint64_t)s << (p * (kind ? 3 : 2));
        }
        refined[i] = out;
    }
    free(tot);
}
