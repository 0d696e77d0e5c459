/* oracle/port/dxt1_port.c -- TEST INFRASTRUCTURE ONLY (see oracle_port.h).
 *
 * Sequential restatement of crnlib::dxt1_endpoint_optimizer (reference crnlib/crn_dxt1.cpp) with
 * endpoint caching disabled.  Floating-point expressions keep the reference's operand types and
 * evaluation order (float vs double, no contraction: build with -ffp-contract=off) because the
 * candidate endpoints are derived from them; everything that decides acceptance is integer.
 *
 * Two reference mechanisms are "skip only" and are kept here in the same role:
 *   - the per-channel lower-bound gate (crn_dxt1.cpp:1316-1322) is a true lower bound of the error;
 *   - m_solutions_tried (crn_dxt1.cpp:1323-1330) suppresses re-evaluation of an (low,high) pair,
 *     which could never pass the strict '<' acceptance a second time.
 * The projection sort of m_evaluated_colors (crn_dxt1.cpp:798-816) only changes the early-out point
 * of the hc evaluators, never their result, and is not restated.
 */
#include "oracle_port.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint8_t c[4]; uint32_t w; } ucolor;   /* unique_color, crn_dxt1.h:134 */
typedef struct { float v[3]; } vec3;

typedef struct {
    uint16_t lo, hi;
    uint64_t err;
    int alpha_block, alt_round, enforce, enforced_sel;
    uint8_t* sel;            /* per unique colour */
} solution;

typedef struct {
    const op_dxt1_params* p;
    const uint8_t* pixels;
    uint32_t n;
    ucolor* uc;
    uint32_t U;
    uint32_t total_w;
    int has_transparent;
    int evaluate_hc, perceptual;
    vec3 *norm, *normw;
    vec3 mean, meanw, axis;
    uint64_t rlo[32], rhi[32], glo[64], ghi[64], blo[32], bhi[32];
    solution best, trial;
    uint8_t* trial_sel;
    /* colour -> unique index map (open addressing) */
    uint32_t* map_key; int32_t* map_val; uint32_t map_mask;
    /* tried-solution set */
    uint32_t* tried; uint32_t tried_cap, tried_cnt;
} opt;

/* ---- ryg single-colour tables, crn_ryg_dxt.cpp:58-100, :517-535 ---- */
static uint8_t OMatch5[256][2], OMatch6[256][2], OMatch5_3[256][2], OMatch6_3[256][2];
static int g_tables_ready = 0;
static void prepare_opt_table(uint8_t (*tab)[2], int size, int three)
{
    for (int i = 0; i < 256; i++) {
        int best = 256;
        for (int mn = 0; mn < size; mn++)
            for (int mx = 0; mx < size; mx++) {
                int mine = size == 32 ? (mn << 3) | (mn >> 2) : (mn << 2) | (mn >> 4);
                int maxe = size == 32 ? (mx << 3) | (mx >> 2) : (mx << 2) | (mx >> 4);
                int v = three ? ((mine + maxe) >> 1) : ((maxe * 2 + mine) / 3);
                int err = abs(v - i) + ((abs(maxe - mine) * 8) >> 8);
                if (err < best) { tab[i][0] = (uint8_t)mx; tab[i][1] = (uint8_t)mn; best = err; }
            }
    }
}
static void init_tables(void)
{
    if (g_tables_ready) return;
    prepare_opt_table(OMatch5, 32, 0);
    prepare_opt_table(OMatch6, 64, 0);
    prepare_opt_table(OMatch5_3, 32, 1);
    prepare_opt_table(OMatch6_3, 64, 1);
    g_tables_ready = 1;
}
const uint8_t* op_omatch_table(int which)   /* exposed so tests can pin the device tables */
{
    init_tables();
    return which == 0 ? &OMatch5[0][0] : which == 1 ? &OMatch6[0][0] : which == 2 ? &OMatch5_3[0][0] : &OMatch6_3[0][0];
}

/* ---- 565 helpers, crn_dxt.cpp:142-182 ---- */
static uint16_t pack565(unsigned r, unsigned g, unsigned b, int scaled)
{
    if (scaled) { r = (r * 31u + 127u) / 255u; g = (g * 63u + 127u) / 255u; b = (b * 31u + 127u) / 255u; }
    if (r > 31) r = 31; if (g > 63) g = 63; if (b > 31) b = 31;
    return (uint16_t)(b | (g << 5) | (r << 11));
}
static void unpack565(uint16_t c, int scaled, int* r, int* g, int* b)
{
    int bb = c & 31, gg = (c >> 5) & 63, rr = (c >> 11) & 31;
    if (scaled) { bb = (bb << 3) | (bb >> 2); gg = (gg << 2) | (gg >> 4); rr = (rr << 3) | (rr >> 2); }
    *r = rr; *g = gg; *b = bb;
}
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
/* x86-64 cvttsd2si semantics for the reference's static_cast<int>(double) */
static int d2i(double x) { return (x > -2147483649.0 && x < 2147483648.0) ? (int)x : (int)0x80000000; }

/* ---- colour distance, crn_color.h:660-746 and crn_dxt1.cpp:1342-1368 ---- */
static uint32_t cdist(const opt* o, int perceptual, const uint8_t* a, int r, int g, int b)
{
    int dr = (int)a[0] - r, dg = (int)a[1] - g, db = (int)a[2] - b;
    if (perceptual) return (uint32_t)(8 * dr * dr + 25 * dg * dg + db * db);
    if (o->p->grayscale_sampling) {
        /* color::RGB_to_Y, crn_color.h:788-796 */
        int y0 = ((int)a[0] * 19595 + (int)a[1] * 38470 + (int)a[2] * 7471 + 32768) >> 16;
        int y1 = (r * 19595 + g * 38470 + b * 7471 + 32768) >> 16;
        int yd = y0 - y1;
        return (uint32_t)(yd * yd);
    }
    return (uint32_t)(dr * dr + dg * dg + db * db);
}

/* ---- tried-solution set ---- */
static int tried_insert(opt* o, uint32_t key)
{
    if ((o->tried_cnt + 1) * 2 > o->tried_cap) {
        uint32_t ncap = o->tried_cap ? o->tried_cap * 2 : 1024;
        uint32_t* nt = (uint32_t*)malloc(sizeof(uint32_t) * 2 * ncap);
        memset(nt, 0, sizeof(uint32_t) * 2 * ncap);
        for (uint32_t i = 0; i < o->tried_cap; i++)
            if (o->tried[2 * i + 1]) {
                uint32_t k = o->tried[2 * i], h = (k * 2654435761u) & (ncap - 1);
                while (nt[2 * h + 1]) h = (h + 1) & (ncap - 1);
                nt[2 * h] = k; nt[2 * h + 1] = 1;
            }
        free(o->tried);
        o->tried = nt; o->tried_cap = ncap;
    }
    uint32_t h = (key * 2654435761u) & (o->tried_cap - 1);
    while (o->tried[2 * h + 1]) {
        if (o->tried[2 * h] == key) return 0;
        h = (h + 1) & (o->tried_cap - 1);
    }
    o->tried[2 * h] = key; o->tried[2 * h + 1] = 1; o->tried_cnt++;
    return 1;
}

static void accept_trial(opt* o)
{
    uint8_t* s = o->best.sel;
    o->best = o->trial;
    o->best.sel = s;
    memcpy(o->best.sel, o->trial.sel, o->U);
}

/* crn_dxt1.cpp:1370-1592 (uber) */
static int evaluate_uber(opt* o, uint16_t lo, uint16_t hi, int alt)
{
    solution* t = &o->trial;
    t->lo = lo; t->hi = hi; t->err = o->best.err; t->alpha_block = 0;
    unsigned first_bt = 0, last_bt = 1;
    if (o->p->pixels_have_alpha || o->p->force_alpha_blocks) first_bt = 1;
    else if (!o->p->use_alpha_blocks) last_bt = 0;
    int c[4][3];
    unpack565(lo, 1, &c[0][0], &c[0][1], &c[0][2]);
    unpack565(hi, 1, &c[1][0], &c[1][1], &c[1][2]);
    for (unsigned bt = first_bt; bt <= last_bt; bt++) {
        uint64_t te = 0;
        int ncol;
        if (!bt) {
            for (int k = 0; k < 3; k++) { c[2][k] = (c[0][k] * 2 + c[1][k] + alt) / 3; c[3][k] = (c[1][k] * 2 + c[0][k] + alt) / 3; }
            ncol = 4;
        } else {
            for (int k = 0; k < 3; k++) c[2][k] = (c[0][k] + c[1][k] + alt) >> 1;
            ncol = 3;
        }
        for (int i = (int)o->U - 1; i >= 0; i--) {
            uint32_t be = cdist(o, o->perceptual, o->uc[i].c, c[0][0], c[0][1], c[0][2]);
            unsigned bi = 0;
            for (int k = 1; k < ncol; k++) {
                uint32_t e = cdist(o, o->perceptual, o->uc[i].c, c[k][0], c[k][1], c[k][2]);
                if (e < be) { be = e; bi = (unsigned)k; }
            }
            te += be * (uint64_t)o->uc[i].w;
            if (te >= t->err) break;
            o->trial_sel[i] = (uint8_t)bi;
        }
        if (te < t->err) {
            t->err = te; t->alpha_block = (bt != 0);
            memcpy(t->sel, o->trial_sel, o->U);
            t->alt_round = alt;
        }
    }
    t->enforce = !t->alpha_block && t->lo == t->hi;
    if (t->enforce) {
        unsigned s;
        if ((t->lo & 31) != 31) { t->lo++; s = 1; } else { t->hi--; s = 0; }
        memset(t->sel, (int)s, o->U);
        t->enforced_sel = (int)s;
    }
    if (t->err < o->best.err) { accept_trial(o); return 1; }
    return 0;
}

/* crn_dxt1.cpp:1594-1757 (fast: selector by projection onto the endpoint axis) */
static int evaluate_fast(opt* o, uint16_t lo, uint16_t hi, int alt)
{
    solution* t = &o->trial;
    t->lo = lo; t->hi = hi; t->err = o->best.err; t->alpha_block = 0;
    unsigned first_bt = 0, last_bt = 1;
    if (o->p->pixels_have_alpha || o->p->force_alpha_blocks) first_bt = 1;
    else if (!o->p->use_alpha_blocks) last_bt = 0;
    int c[4][3];
    unpack565(lo, 1, &c[0][0], &c[0][1], &c[0][2]);
    unpack565(hi, 1, &c[1][0], &c[1][1], &c[1][2]);
    int vr = c[1][0] - c[0][0], vg = c[1][1] - c[0][1], vb = c[1][2] - c[0][2];
    if (o->perceptual) { vr *= 8; vg *= 24; }
    int stops[4];
    stops[0] = c[0][0] * vr + c[0][1] * vg + c[0][2] * vb;
    stops[1] = c[1][0] * vr + c[1][1] * vg + c[1][2] * vb;
    int dirr = vr * 2, dirg = vg * 2, dirb = vb * 2;
    for (unsigned bt = first_bt; bt <= last_bt; bt++) {
        uint64_t te = 0;
        if (!bt) {
            for (int k = 0; k < 3; k++) { c[2][k] = (c[0][k] * 2 + c[1][k] + alt) / 3; c[3][k] = (c[1][k] * 2 + c[0][k] + alt) / 3; }
            stops[2] = c[2][0] * vr + c[2][1] * vg + c[2][2] * vb;
            stops[3] = c[3][0] * vr + c[3][1] * vg + c[3][2] * vb;
            int c0Point = stops[1] + stops[3], halfPoint = stops[3] + stops[2], c3Point = stops[2] + stops[0];
            for (int i = (int)o->U - 1; i >= 0; i--) {
                const uint8_t* q = o->uc[i].c;
                int dot = q[0] * dirr + q[1] * dirg + q[2] * dirb;
                unsigned bi = dot >= halfPoint ? (dot < c0Point ? 3u : 1u) : (dot < c3Point ? 0u : 2u);
                uint32_t be = cdist(o, o->perceptual, q, c[bi][0], c[bi][1], c[bi][2]);
                te += be * (uint64_t)o->uc[i].w;
                if (te >= t->err) break;
                o->trial_sel[i] = (uint8_t)bi;
            }
        } else {
            for (int k = 0; k < 3; k++) c[2][k] = (c[0][k] + c[1][k] + alt) >> 1;
            stops[2] = c[2][0] * vr + c[2][1] * vg + c[2][2] * vb;
            int c02Point = stops[0] + stops[2], c21Point = stops[2] + stops[1];
            for (int i = (int)o->U - 1; i >= 0; i--) {
                const uint8_t* q = o->uc[i].c;
                int dot = q[0] * dirr + q[1] * dirg + q[2] * dirb;
                unsigned bi = dot < c02Point ? 0u : (dot < c21Point ? 2u : 1u);
                uint32_t be = cdist(o, o->perceptual, q, c[bi][0], c[bi][1], c[bi][2]);
                te += be * (uint64_t)o->uc[i].w;
                if (te >= t->err) break;
                o->trial_sel[i] = (uint8_t)bi;
            }
        }
        if (te < t->err) {
            t->err = te; t->alpha_block = (bt != 0);
            memcpy(t->sel, o->trial_sel, o->U);
        }
    }
    if (!t->alpha_block && t->lo == t->hi) {
        unsigned s;
        if ((t->lo & 31) != 31) { t->lo++; s = 1; } else { t->hi--; s = 0; }
        memset(t->sel, (int)s, o->U);
    }
    if (t->err < o->best.err) {
        /* the fast evaluator leaves m_alternate_rounding / m_enforce_selector of the trial untouched
           (crn_dxt1.cpp:1722-1754); with no hc path active they are never read. */
        accept_trial(o);
        return 1;
    }
    return 0;
}

/* crn_dxt1.cpp:1759-1835 (hc: 4-colour only, no selector bookkeeping) */
static int evaluate_hc(opt* o, uint16_t lo, uint16_t hi, int alt)
{
    int c[4][3];
    unpack565(lo, 1, &c[0][0], &c[0][1], &c[0][2]);
    unpack565(hi, 1, &c[1][0], &c[1][1], &c[1][2]);
    for (int k = 0; k < 3; k++) { c[2][k] = (c[0][k] * 2 + c[1][k] + alt) / 3; c[3][k] = (c[1][k] * 2 + c[0][k] + alt) / 3; }
    uint64_t err = 0;
    for (uint32_t i = 0; i < o->U; i++) {
        uint32_t be = cdist(o, o->perceptual, o->uc[i].c, c[0][0], c[0][1], c[0][2]);
        for (int k = 1; k < 4; k++) {
            uint32_t e = cdist(o, o->perceptual, o->uc[i].c, c[k][0], c[k][1], c[k][2]);
            if (e < be) be = e;
        }
        err += be * (uint64_t)o->uc[i].w;
        if (err >= o->best.err) break;
    }
    if (err >= o->best.err) return 0;
    o->best.lo = lo; o->best.hi = hi; o->best.err = err; o->best.alpha_block = 0; o->best.alt_round = alt;
    o->best.enforce = lo == hi;
    if (o->best.enforce) {
        if ((o->best.lo & 31) != 31) { o->best.lo++; o->best.enforced_sel = 1; }
        else { o->best.hi--; o->best.enforced_sel = 0; }
    }
    return 1;
}

/* crn_dxt1.cpp:1312-1340 */
static int evaluate(opt* o, uint16_t lo, uint16_t hi, int alt)
{
    int r0, g0, b0, r1, g1, b1;
    unpack565(lo, 0, &r0, &g0, &b0);
    unpack565(hi, 0, &r1, &g1, &b1);
    uint64_t re = r0 < r1 ? o->rlo[r0] + o->rhi[r1] : o->rhi[r0] + o->rlo[r1];
    uint64_t ge = g0 < g1 ? o->glo[g0] + o->ghi[g1] : o->ghi[g0] + o->glo[g1];
    uint64_t be = b0 < b1 ? o->blo[b0] + o->bhi[b1] : o->bhi[b0] + o->blo[b1];
    if (re + ge + be >= o->best.err) return 0;
    if (!alt && !tried_insert(o, (uint32_t)lo | ((uint32_t)hi << 16))) return 0;
    if (o->evaluate_hc) return evaluate_hc(o, lo, hi, alt);
    if (o->p->quality >= 3) return evaluate_uber(o, lo, hi, alt);
    return evaluate_fast(o, lo, hi, alt);
}
static int evaluate_canon(opt* o, uint16_t lo, uint16_t hi)
{   /* dxt1_solution_coordinates::canonicalize, crn_dxt1.h:78-85 */
    if (lo < hi) { uint16_t t = lo; lo = hi; hi = t; }
    return evaluate(o, lo, hi, 0);
}

/* crn_dxt1.cpp:1845-1869 */
static void compute_selectors(opt* o)
{
    if (!o->evaluate_hc) return;
    if (o->best.enforce) { memset(o->best.sel, o->best.enforced_sel, o->U); return; }
    int c[4][3];
    unpack565(o->best.lo, 1, &c[0][0], &c[0][1], &c[0][2]);
    unpack565(o->best.hi, 1, &c[1][0], &c[1][1], &c[1][2]);
    int alt = o->best.alt_round;
    for (int k = 0; k < 3; k++) { c[2][k] = (c[0][k] * 2 + c[1][k] + alt) / 3; c[3][k] = (c[1][k] * 2 + c[0][k] + alt) / 3; }
    for (uint32_t i = 0; i < o->U; i++) {
        uint32_t e0 = cdist(o, o->perceptual, o->uc[i].c, c[0][0], c[0][1], c[0][2]);
        uint32_t e1 = cdist(o, o->perceptual, o->uc[i].c, c[1][0], c[1][1], c[1][2]);
        uint32_t e2 = cdist(o, o->perceptual, o->uc[i].c, c[2][0], c[2][1], c[2][2]);
        uint32_t e3 = cdist(o, o->perceptual, o->uc[i].c, c[3][0], c[3][1], c[3][2]);
        uint32_t e01 = e0 < e1 ? e0 : e1, e23 = e2 < e3 ? e2 : e3;
        o->best.sel[i] = (uint8_t)(e01 <= e23 ? (e01 == e0 ? 0 : 1) : (e23 == e2 ? 2 : 3));
    }
}

/* crn_dxt1.cpp:525-698 */
static int refine_solution(opt* o, int level)
{
    compute_selectors(o);
    static const int w1Tab[4] = { 3, 0, 2, 1 };
    static const int prods_0[4] = { 0x00, 0x00, 0x02, 0x02 };
    static const int prods_1[4] = { 0x00, 0x09, 0x01, 0x04 };
    static const int prods_2[4] = { 0x09, 0x00, 0x04, 0x01 };
    double akku_0 = 0, akku_1 = 0, akku_2 = 0;
    double At1_r = 0, At1_g = 0, At1_b = 0, At2_r = 0, At2_g = 0, At2_b = 0;
    for (uint32_t i = 0; i < o->U; i++) {
        const double weight = o->uc[i].w;
        double r = o->uc[i].c[0] * weight, g = o->uc[i].c[1] * weight, b = o->uc[i].c[2] * weight;
        int step = o->best.sel[i] ^ 1;
        int w1 = w1Tab[step];
        akku_0 += prods_0[step] * weight;
        akku_1 += prods_1[step] * weight;
        akku_2 += prods_2[step] * weight;
        At1_r += w1 * r; At1_g += w1 * g; At1_b += w1 * b;
        At2_r += r; At2_g += g; At2_b += b;
    }
    At2_r = 3 * At2_r - At1_r; At2_g = 3 * At2_g - At1_g; At2_b = 3 * At2_b - At1_b;
    double xx = akku_2, yy = akku_1, xy = akku_0;
    double t = xx * yy - xy * xy;
    if (!yy || !xx || (fabs(t) < .0000125f)) return 0;
    double frb = (3.0f * 31.0f / 255.0f) / t;
    double fg = frb * (63.0f / 31.0f);
    int e0[3], e1[3];
    e0[0] = clampi(d2i((At1_r * yy - At2_r * xy) * frb + 0.5f), 0, 31);
    e0[1] = clampi(d2i((At1_g * yy - At2_g * xy) * fg + 0.5f), 0, 63);
    e0[2] = clampi(d2i((At1_b * yy - At2_b * xy) * frb + 0.5f), 0, 31);
    e1[0] = clampi(d2i((At2_r * xx - At1_r * xy) * frb + 0.5f), 0, 31);
    e1[1] = clampi(d2i((At2_g * xx - At1_g * xy) * fg + 0.5f), 0, 63);
    e1[2] = clampi(d2i((At2_b * xx - At1_b * xy) * frb + 0.5f), 0, 31);
    int improved = 0;
    if (level == 0) {
        /* max16 = e0, min16 = e1 ; coordinates(min16, max16).canonicalize() */
        uint16_t mx = (uint16_t)((e0[0] << 11) | (e0[1] << 5) | e0[2]);
        uint16_t mn = (uint16_t)((e1[0] << 11) | (e1[1] << 5) | e1[2]);
        improved |= evaluate_canon(o, mn, mx);
    } else if (level == 1) {
        for (int i = 0; i < 2; i++)
            for (int rr = -1; rr <= 1; rr++)
                for (int gr = -1; gr <= 1; gr++)
                    for (int br = -1; br <= 1; br++) {
                        int c0[3] = { e0[0], e0[1], e0[2] }, c1[3] = { e1[0], e1[1], e1[2] };
                        int* c = i ? c1 : c0;
                        c[0] = clampi(c[0] + rr, 0, 31); c[1] = clampi(c[1] + gr, 0, 63); c[2] = clampi(c[2] + br, 0, 31);
                        improved |= evaluate_canon(o, pack565(c0[0], c0[1], c0[2], 0), pack565(c1[0], c1[1], c1[2], 0));
                    }
    } else {
        /* level 2 is never requested by optimize_endpoints / try_median4 (crn_dxt1.cpp:1026, :1305) */
        for (int orr = -1; orr <= 1; orr++) for (int ogr = -1; ogr <= 1; ogr++) for (int obr = -1; obr <= 1; obr++) {
            int c0[3] = { clampi(e0[0] + orr, 0, 31), clampi(e0[1] + ogr, 0, 63), clampi(e0[2] + obr, 0, 31) };
            int c1[3] = { e1[0], e1[1], e1[2] };
            for (int rr = -1; rr <= 1; rr++) for (int gr = -1; gr <= 1; gr++) for (int br = -1; br <= 1; br++) {
                c1[0] = clampi(c1[0] + rr, 0, 31); c1[1] = clampi(c1[1] + gr, 0, 63); c1[2] = clampi(c1[2] + br, 0, 31);
                improved |= evaluate_canon(o, pack565(c0[0], c0[1], c0[2], 0), pack565(c1[0], c1[1], c1[2], 0));
            }
        }
    }
    return improved;
}

/* fast_random, crn_rand.cpp:310-405 */
typedef struct { uint32_t jsr, jcong; } fast_random;
static uint32_t fr_u32(fast_random* r)
{
    r->jsr ^= (r->jsr << 17); r->jsr ^= (r->jsr >> 13); r->jsr ^= (r->jsr << 5);
    r->jcong = 69069u * r->jcong + 1234567u;
    return r->jsr ^ r->jcong;
}
static float fr_frand(fast_random* r, float l, float h)
{
    if (l >= h) return l;
    const double cNorm = 1.0 / (double)0x100000000ULL;
    float v = (float)(l + (h - l) * (fr_u32(r) * cNorm));
    return v < l ? l : (v > h ? h : v);
}

static float sqdist3(const vec3* a, const vec3* b)
{
    float d2 = 0;
    for (int i = 0; i < 3; i++) { float d = a->v[i] - b->v[i]; d2 += d * d; }
    return d2;
}

/* crn_dxt1.cpp:1181-1308 */
static int try_median4(opt* o, const vec3* low_color, const vec3* high_color)
{
    vec3 means[4];
    if (o->U <= 4) {
        for (uint32_t i = 0; i < 4; i++) {
            int idx = (int)o->U - 1 < (int)i ? (int)o->U - 1 : (int)i;
            means[i] = o->norm[idx];
        }
    } else {
        for (int k = 0; k < 3; k++) { means[0].v[k] = low_color->v[k] - o->mean.v[k]; means[3].v[k] = high_color->v[k] - o->mean.v[k]; }
        for (int k = 0; k < 3; k++) {
            means[1].v[k] = means[0].v[k] + (means[3].v[k] - means[0].v[k]) * (1.0f / 3.0f);
            means[2].v[k] = means[0].v[k] + (means[3].v[k] - means[0].v[k]) * (2.0f / 3.0f);
        }
        fast_random rm = { 0xABCD917Au, 0x17F3DEADu };
        const uint32_t cMaxIters = 8;
        uint32_t reassign_rover = 0;
        float prev_total_dist = 1.0e+37f;
        for (uint32_t iter = 0; iter < cMaxIters; iter++) {
            vec3 new_means[4]; float new_weights[4];
            memset(new_means, 0, sizeof(new_means)); memset(new_weights, 0, sizeof(new_weights));
            float total_dist = 0;
            for (uint32_t i = 0; i < o->U; i++) {
                const vec3* v = &o->norm[i];
                float best_dist = sqdist3(&means[0], v);
                int best_index = 0;
                for (int j = 1; j < 4; j++) {
                    float dist = sqdist3(&means[j], v);
                    if (dist < best_dist) { best_dist = dist; best_index = j; }
                }
                total_dist += best_dist;
                float fw = (float)o->uc[i].w;
                for (int k = 0; k < 3; k++) new_means[best_index].v[k] += v->v[k] * fw;
                new_weights[best_index] += fw;
            }
            uint32_t highest_index = 0; float highest_weight = 0; int empty_cell = 0;
            for (uint32_t j = 0; j < 4; j++) {
                if (new_weights[j] > 0.0f) {
                    for (int k = 0; k < 3; k++) means[j].v[k] = new_means[j].v[k] / new_weights[j];
                    if (new_weights[j] > highest_weight) { highest_weight = new_weights[j]; highest_index = j; }
                } else empty_cell = 1;
            }
            if (!empty_cell) {
                if (fabsf(total_dist - prev_total_dist) < .00001f) break;
                prev_total_dist = total_dist;
            } else prev_total_dist = 1.0e+37f;
            if (empty_cell && iter != cMaxIters - 1) {
                const uint32_t ri = (highest_index + reassign_rover) & 3;
                reassign_rover++;
                for (uint32_t j = 0; j < 4; j++)
                    if (new_weights[j] == 0.0f) {
                        means[j] = means[ri];
                        for (int k = 0; k < 3; k++) means[j].v[k] += fr_frand(&rm, -.00196f, .00196f);
                    }
            }
        }
    }
    int improved = 0;
    for (uint32_t i = 0; i < 3; i++)
        for (uint32_t j = i + 1; j < 4; j++) {
            vec3 v0, v1;
            for (int k = 0; k < 3; k++) { v0.v[k] = means[i].v[k] + o->mean.v[k]; v1.v[k] = means[j].v[k] + o->mean.v[k]; }
            /* color_quad_u8(int...) clamps to 0..255, pack_color(.., false) clamps to 31/63/31 */
            int a0 = clampi((int)floorf(.5f + v0.v[0] * 31.0f), 0, 255), a1 = clampi((int)floorf(.5f + v0.v[1] * 63.0f), 0, 255), a2 = clampi((int)floorf(.5f + v0.v[2] * 31.0f), 0, 255);
            int b0 = clampi((int)floorf(.5f + v1.v[0] * 31.0f), 0, 255), b1 = clampi((int)floorf(.5f + v1.v[1] * 63.0f), 0, 255), b2 = clampi((int)floorf(.5f + v1.v[2] * 31.0f), 0, 255);
            improved |= evaluate_canon(o, pack565(a0, a1, a2, 0), pack565(b0, b1, b2, 0));
        }
    improved |= refine_solution(o, o->p->quality == 4 ? 1 : 0);
    return improved;
}

/* crn_dxt1.cpp:93-153 */
static int try_average_block_as_solid(opt* o)
{
    uint64_t tot_r = 0, tot_g = 0, tot_b = 0;
    uint32_t total_weight = 0;
    for (uint32_t i = 0; i < o->U; i++) {
        uint32_t w = o->uc[i].w;
        total_weight += w;
        tot_r += o->uc[i].c[0] * (uint64_t)w; tot_g += o->uc[i].c[1] * (uint64_t)w; tot_b += o->uc[i].c[2] * (uint64_t)w;
    }
    const uint32_t half = total_weight >> 1;
    uint32_t ar = (uint32_t)((tot_r + half) / total_weight), ag = (uint32_t)((tot_g + half) / total_weight), ab = (uint32_t)((tot_b + half) / total_weight);
    int improved = evaluate(o, (uint16_t)((OMatch5[ar][0] << 11) | (OMatch6[ag][0] << 5) | OMatch5[ab][0]),
                            (uint16_t)((OMatch5[ar][1] << 11) | (OMatch6[ag][1] << 5) | OMatch5[ab][1]), 0);
    if (o->p->use_alpha_blocks && o->best.err)
        improved |= evaluate(o, (uint16_t)((OMatch5_3[ar][0] << 11) | (OMatch6_3[ag][0] << 5) | OMatch5_3[ab][0]),
                             (uint16_t)((OMatch5_3[ar][1] << 11) | (OMatch6_3[ag][1] << 5) | OMatch5_3[ab][1]), 0);
    if (o->p->quality == 4) {
        for (uint32_t i = 0; i < o->U; i++) {
            uint32_t r = o->uc[i].c[0], g = o->uc[i].c[1], b = o->uc[i].c[2];
            if (r == ar && g == ag && b == ab) continue;
            improved |= evaluate(o, (uint16_t)((OMatch5[r][0] << 11) | (OMatch6[g][0] << 5) | OMatch5[b][0]),
                                 (uint16_t)((OMatch5[r][1] << 11) | (OMatch6[g][1] << 5) | OMatch5[b][1]), 0);
            if (o->p->use_alpha_blocks && o->best.err)
                improved |= evaluate(o, (uint16_t)((OMatch5_3[r][0] << 11) | (OMatch6_3[g][0] << 5) | OMatch5_3[b][0]),
                                     (uint16_t)((OMatch5_3[r][1] << 11) | (OMatch6_3[g][1] << 5) | OMatch5_3[b][1]), 0);
        }
    }
    return improved;
}

/* crn_dxt1.cpp:369-413 */
static void comp_errors(opt* o, uint32_t comp, uint64_t (*error)[256], uint64_t* best_remaining)
{
    uint64_t W[4] = { 0 }, WP2[4] = { 0 }, WPP[4] = { 0 };
    for (uint32_t i = 0; i < o->U; i++) {
        uint32_t p = o->uc[i].c[comp], w = o->uc[i].w;
        uint8_t s = o->best.sel[i];
        W[s] += (uint64_t)(int64_t)w;
        WP2[s] += (uint64_t)((int64_t)w * p * 2);
        WPP[s] += (uint64_t)((int64_t)w * p * p);
    }
    const uint32_t limit = comp == 1 ? 64 : 32;
    for (uint32_t s = 0; s < 2; s++) {
        uint64_t be = error[s][0] = WPP[s];
        for (uint32_t c = 1; c < limit; c++) {
            uint8_t p = (uint8_t)(comp == 1 ? (c << 2 | c >> 4) : (c << 3 | c >> 2));
            error[s][c] = W[s] * p * p - WP2[s] * p + WPP[s];
            if (error[s][c] < be) be = error[s][c];
        }
        best_remaining[s] = be;
    }
    for (uint32_t s = 2; s < 4; s++) {
        uint64_t be = error[s][0] = WPP[s], d = W[s] - WP2[s], dd = W[s] << 1, e = WPP[s] + d;
        for (uint32_t p = 1; p < 256; p++, d += dd, e += d) {
            error[s][p] = e;
            if (e < be) be = e;
        }
        best_remaining[s] = be;
    }
    for (uint32_t s = 3; s; s--) best_remaining[s - 1] += best_remaining[s];
}

/* crn_dxt1.cpp:415-486 */
static void optimize_endpoint_comps(opt* o)
{
    compute_selectors(o);
    if (o->best.alpha_block || !o->best.err) return;
    int sl[3], sh[3];
    unpack565(o->best.lo, 1, &sl[0], &sl[1], &sl[2]);
    unpack565(o->best.hi, 1, &sh[0], &sh[1], &sh[2]);
    static uint64_t error[4][256];
    uint64_t brem[4];
    for (uint32_t comp = 0; comp < 3; comp++) {
        uint8_t p0 = (uint8_t)sl[comp], p1 = (uint8_t)sh[comp];
        int low[3], high[3];
        unpack565(o->best.lo, 0, &low[0], &low[1], &low[2]);
        unpack565(o->best.hi, 0, &high[0], &high[1], &high[2]);
        comp_errors(o, comp, error, brem);
        uint64_t best_error = error[0][low[comp]] + error[1][high[comp]] + error[2][(p0 * 2 + p1) / 3] + error[3][(p0 + p1 * 2) / 3];
        if (brem[0] >= best_error) continue;
        const uint32_t limit = comp == 1 ? 64 : 32;
        for (uint32_t c0 = 0; c0 < limit; c0++) {
            uint64_t e0 = error[0][c0];
            if (e0 + brem[1] >= best_error) continue;
            low[comp] = (int)c0;
            uint16_t packed_low = pack565(low[0], low[1], low[2], 0);
            p0 = (uint8_t)(comp == 1 ? (c0 << 2 | c0 >> 4) : (c0 << 3 | c0 >> 2));
            for (uint32_t c1 = 0; c1 < limit; c1++) {
                uint64_t e = e0 + error[1][c1];
                if (e + brem[2] >= best_error) continue;
                p1 = (uint8_t)(comp == 1 ? (c1 << 2 | c1 >> 4) : (c1 << 3 | c1 >> 2));
                e += error[2][(p0 * 2 + p1) / 3];
                if (e + brem[3] >= best_error) continue;
                e += error[3][(p0 + p1 * 2) / 3];
                if (e >= best_error) continue;
                high[comp] = (int)c1;
                if (!evaluate(o, packed_low, pack565(high[0], high[1], high[2], 0), 0)) continue;
                if (!o->best.err) return;
                compute_selectors(o);
                comp_errors(o, comp, error, brem);
                best_error = error[0][c0] + error[1][c1] + error[2][(p0 * 2 + p1) / 3] + error[3][(p0 + p1 * 2) / 3];
                e0 = error[0][c0];
                if (e0 + brem[1] >= best_error) break;
            }
        }
    }
}

/* crn_dxt1.cpp:1871-1882 */
static void lerp_color(const uint8_t* a, const uint8_t* b, float f, int rounding, uint8_t* out)
{
    float r = rounding ? 1.0f : 0.0f;
    for (int k = 0; k < 3; k++) {
        float fa = a[k], fb = b[k];
        out[k] = (uint8_t)clampi((int)(r + (fa + (fb - fa) * f)), 0, 255);
    }
    out[3] = 255;
}

/* crn_dxt1.cpp:1886-1997 */
static void try_combinatorial_encoding(opt* o)
{
    if (o->U < 2 || o->U > 4) return;
    uint8_t tmp[64][4];
    uint32_t nt = o->U;
    for (uint32_t i = 0; i < o->U; i++) memcpy(tmp[i], o->uc[i].c, 4);
    if (nt == 2) {
        static const float f2[10] = { 2.0f, 3.0f, .5f, 1.5f, -1.0f, 2.0f, -.5f, .5f, -2.0f, -1.0f };
        for (uint32_t k = 0; k < 2; k++)
            for (uint32_t q = 0; q < 2; q++) {
                const uint32_t r = q ^ 1;
                for (int m = 0; m < 10; m++) lerp_color(tmp[q], tmp[r], f2[m], (int)k, tmp[nt++]);
            }
    } else if (nt == 3) {
        for (uint32_t i = 0; i <= 2; i++)
            for (uint32_t j = 0; j <= 2; j++) {
                if (i == j) continue;
                lerp_color(tmp[i], tmp[j], 1.5f, 1, tmp[nt++]);
                lerp_color(tmp[i], tmp[j], 2.0f / 3.0f, 1, tmp[nt++]);
                lerp_color(tmp[i], tmp[j], 1.0f / 3.0f, 1, tmp[nt++]);
                lerp_color(tmp[i], tmp[j], -.5f, 1, tmp[nt++]);
            }
    }
    uint16_t packed[64]; uint32_t np = 0;
    for (uint32_t i = 0; i < nt; i++) {
        uint16_t pc = pack565(tmp[i][0], tmp[i][1], tmp[i][2], 1);
        uint32_t j;
        for (j = 0; j < np; j++) if (packed[j] == pc) break;
        if (j == np) packed[np++] = pc;
    }
    /* note: `i < size() - 1` is unsigned in the reference; np >= 1 always holds here */
    for (uint32_t i = 0; o->best.err && i + 1 < np; i++)
        for (uint32_t j = i + 1; o->best.err && j < np; j++)
            evaluate(o, packed[i], packed[j], 0);
    uint64_t error = o->best.err;
    if (error) o->best.err = 1;
    for (uint32_t i = 0; o->best.err && i + 1 < np; i++)
        for (uint32_t j = i + 1; o->best.err && j < np; j++)
            evaluate(o, packed[i], packed[j], 1);
    if (o->best.err) o->best.err = error;
}

static int are_selectors_all_equal(const opt* o)
{
    if (!o->U) return 0;
    for (uint32_t i = 1; i < o->U; i++) if (o->best.sel[i] != o->best.sel[0]) return 0;
    return 1;
}

static int32_t map_find(const opt* o, uint32_t key)
{
    uint32_t h = (key * 2654435761u) & o->map_mask;
    while (o->map_val[h] >= 0) {
        if (o->map_key[h] == key) return o->map_val[h];
        h = (h + 1) & o->map_mask;
    }
    return -1;
}

/* crn_dxt1.cpp:263-365 */
static void return_solution(opt* o, op_dxt1_result* r, uint8_t* selectors)
{
    compute_selectors(o);
    int invert;
    if (o->best.alpha_block) invert = o->best.lo > o->best.hi;
    else invert = o->best.lo < o->best.hi;
    if (invert) { r->low = o->best.hi; r->high = o->best.lo; }
    else { r->low = o->best.lo; r->high = o->best.hi; }
    static const uint8_t invNull[4] = { 0, 1, 2, 3 }, invAlpha[4] = { 1, 0, 2, 3 }, invColor[4] = { 1, 0, 3, 2 };
    const uint8_t* inv = invNull;
    if (invert) inv = o->best.alpha_block ? invAlpha : invColor;
    const uint32_t alpha_thresh = o->p->pixels_have_alpha ? (o->p->alpha_threshold << 24) : 0;
    for (uint32_t i = 0; i < o->n; i++) {
        uint32_t c;
        memcpy(&c, o->pixels + 4 * i, 4);
        uint8_t s = 3;
        if (c >= alpha_thresh) {
            c |= 0xFF000000u;
            s = inv[o->best.sel[map_find(o, c)]];
        }
        selectors[i] = s;
    }
    r->alpha_block = (uint8_t)o->best.alpha_block;
    r->error = o->best.err;
}

/* crn_dxt1.cpp:703-1067 */
static void optimize_endpoints(opt* o, vec3 low_color, vec3 high_color, op_dxt1_result* res, uint8_t* selectors)
{
    static const int16_t fast_tab[] = { 0, 1, 2, 3 }, normal_tab[] = { 0, 1, 3, 5, 7 },
                         better_tab[] = { 0, 1, 2, 3, 5, 9, 15, 19, 27, 43 },
                         uber_tab[] = { 0, 1, 2, 3, 5, 7, 9, 10, 13, 15, 19, 27, 43, 59, 91 };
    const vec3 orig_low = low_color, orig_high = high_color;
    uint32_t num_passes, probe_range;
    const int16_t* tab;
    float dist_per_trial = .015625f;
    switch (o->p->quality) {
    case 0: tab = fast_tab; probe_range = 4; dist_per_trial = .027063293f; num_passes = 1; break;
    case 1: tab = fast_tab; probe_range = 4; dist_per_trial = .027063293f; num_passes = 2; break;
    case 2: tab = normal_tab; probe_range = 5; dist_per_trial = .027063293f; num_passes = 2; break;
    case 3: tab = better_tab; probe_range = 10; num_passes = 2; break;
    default: tab = uber_tab; probe_range = 15; num_passes = 4; break;
    }
    if (o->p->quality >= 3) try_median4(o, &orig_low, &orig_high);

    uint32_t probe_low[31], probe_high[31];
    vec3 spa[2];
    for (int k = 0; k < 3; k++) spa[1].v[k] = o->axis.v[k] * dist_per_trial;
    spa[1].v[0] *= 31.0f; spa[1].v[1] *= 63.0f; spa[1].v[2] *= 31.0f;
    for (int k = 0; k < 3; k++) spa[0].v[k] = -spa[1].v[k];
    static const float lim[3] = { 31.0f, 63.0f, 31.0f };
    for (int k = 0; k < 3; k++) {
        float a = low_color.v[k] * lim[k]; low_color.v[k] = a < 0.0f ? 0.0f : (a > lim[k] ? lim[k] : a);
        float b = high_color.v[k] * lim[k]; high_color.v[k] = b < 0.0f ? 0.0f : (b > lim[k] ? lim[k] : b);
    }
    for (uint32_t pass = 0; pass < num_passes; pass++) {
        if (pass) {
            int r, g, b;
            unpack565(o->best.lo, 0, &r, &g, &b); low_color.v[0] = (float)r; low_color.v[1] = (float)g; low_color.v[2] = (float)b;
            unpack565(o->best.hi, 0, &r, &g, &b); high_color.v[0] = (float)r; high_color.v[1] = (float)g; high_color.v[2] = (float)b;
        }
        const uint64_t prev_best_error = o->best.err;
        if (!prev_best_error) break;
        uint32_t nlow = 0, nhigh = 0;
        for (int which = 0; which < 2; which++) {
            int prev[2] = { -1, -1 };
            const vec3* base = which ? &high_color : &low_color;
            vec3 init;
            for (int k = 0; k < 3; k++) init.v[k] = base->v[k] + .5f;
            for (uint32_t i = 0; i < probe_range; i++) {
                const int ls = i ? 0 : 1;
                int x = tab[i];
                for (int s = ls; s < 2; s++) {
                    vec3 pc;
                    for (int k = 0; k < 3; k++) pc.v[k] = init.v[k] + spa[s].v[k] * (float)x;
                    int r = clampi((int)floorf(pc.v[0]), 0, 31), g = clampi((int)floorf(pc.v[1]), 0, 63), b = clampi((int)floorf(pc.v[2]), 0, 31);
                    int packed = b | (g << 5) | (r << 11);
                    if (packed != prev[s]) {
                        if (which) probe_high[nhigh++] = (uint32_t)packed; else probe_low[nlow++] = (uint32_t)packed;
                        prev[s] = packed;
                    }
                }
            }
        }
        for (uint32_t i = 0; i < nlow; i++)
            for (uint32_t j = 0; j < nhigh; j++)
                evaluate_canon(o, (uint16_t)probe_low[i], (uint16_t)probe_high[j]);
        if (o->p->quality >= 2) {
            for (int which = 0; which < 2; which++) {
                int cr, cg, cb;
                unpack565(which ? o->best.hi : o->best.lo, 0, &cr, &cg, &cb);
                for (int z = -1; z <= 1; z++) for (int y = -1; y <= 1; y++) for (int x = -1; x <= 1; x++) {
                    /* g_adjacency order (crn_dxt1.cpp:489-522): x fastest, then y, then z, centre skipped */
                    if (!x && !y && !z) continue;
                    int r = cr + x; if (r < 0 || r > 31) continue;
                    int g = cg + y; if (g < 0 || g > 63) continue;
                    int b = cb + z; if (b < 0 || b > 31) continue;
                    if (which) evaluate_canon(o, o->best.lo, pack565(r, g, b, 0));
                    else evaluate_canon(o, pack565(r, g, b, 0), o->best.hi);
                }
                if (o->p->quality == 4) {
                    int c[3];
                    unpack565(which ? o->best.hi : o->best.lo, 0, &c[0], &c[1], &c[2]);
                    for (int a = 0; a < 3; a++) {
                        int limit = a == 1 ? 63 : 31;
                        for (int s = -2; s <= 2; s += 4) {
                            int q = c[a] + s;
                            if (q < 0 || q > limit) continue;
                            int cc[3] = { c[0], c[1], c[2] };
                            cc[a] = q;
                            if (which) evaluate_canon(o, o->best.lo, pack565(cc[0], cc[1], cc[2], 0));
                            else evaluate_canon(o, pack565(cc[0], cc[1], cc[2], 0), o->best.hi);
                        }
                    }
                }
            }
        }
        if (!o->best.err || (pass && o->best.err == prev_best_error)) break;
        if (o->p->quality >= 4) refine_solution(o, 1);
    }
    if (o->p->quality >= 2) {
        if (o->best.err && !o->p->pixels_have_alpha) {
            int choose_solid = 0;
            if (are_selectors_all_equal(o)) choose_solid = try_average_block_as_solid(o);
            if (!choose_solid && o->p->quality == 4) optimize_endpoint_comps(o);
        }
        if (o->p->quality == 4 && o->best.err) try_combinatorial_encoding(o);
    }
    return_solution(o, res, selectors);
}

/* crn_dxt1.cpp:155-189 */
static void compute_vectors(opt* o, const vec3* pw)
{
    memset(&o->mean, 0, sizeof(vec3)); memset(&o->meanw, 0, sizeof(vec3));
    for (uint32_t i = 0; i < o->U; i++) {
        const uint8_t* c = o->uc[i].c;
        const uint32_t w = o->uc[i].w;
        vec3 nc, ncw;
        nc.v[0] = c[0] * 1.0f / 255.0f; nc.v[1] = c[1] * 1.0f / 255.0f; nc.v[2] = c[2] * 1.0f / 255.0f;
        for (int k = 0; k < 3; k++) ncw.v[k] = pw->v[k] * nc.v[k];
        o->norm[i] = nc; o->normw[i] = ncw;
        for (int k = 0; k < 3; k++) { o->mean.v[k] += nc.v[k] * (float)w; o->meanw.v[k] += ncw.v[k] * (float)w; }
    }
    if (o->total_w) {
        float inv = 1.0f / o->total_w;
        for (int k = 0; k < 3; k++) { o->mean.v[k] *= inv; o->meanw.v[k] *= inv; }
    }
    for (uint32_t i = 0; i < o->U; i++)
        for (int k = 0; k < 3; k++) { o->norm[i].v[k] -= o->mean.v[k]; o->normw[i].v[k] -= o->meanw.v[k]; }
}

/* crn_dxt1.cpp:192-256 */
static void compute_pca(opt* o, vec3* axis, const vec3* cols, const vec3* def)
{
    double cov[6] = { 0, 0, 0, 0, 0, 0 };
    for (uint32_t i = 0; i < o->U; i++) {
        float r = cols[i].v[0], g = cols[i].v[1], b = cols[i].v[2];
        if (o->uc[i].w > 1) {
            const double weight = o->uc[i].w;
            cov[0] += r * r * weight; cov[1] += r * g * weight; cov[2] += r * b * weight;
            cov[3] += g * g * weight; cov[4] += g * b * weight; cov[5] += b * b * weight;
        } else {
            cov[0] += r * r; cov[1] += r * g; cov[2] += r * b; cov[3] += g * g; cov[4] += g * b; cov[5] += b * b;
        }
    }
    double vfr = .9f, vfg = 1.0f, vfb = .7f;
    for (uint32_t iter = 0; iter < 8; iter++) {
        double r = vfr * cov[0] + vfg * cov[1] + vfb * cov[2];
        double g = vfr * cov[1] + vfg * cov[3] + vfb * cov[4];
        double b = vfr * cov[2] + vfg * cov[4] + vfb * cov[5];
        double m = fabs(r) > fabs(g) ? fabs(r) : fabs(g);
        m = m > fabs(b) ? m : fabs(b);
        if (m > 1e-10) { m = 1.0f / m; r *= m; g *= m; b *= m; }
        double delta = (vfr - r) * (vfr - r) + (vfg - g) * (vfg - g) + (vfb - b) * (vfb - b);
        vfr = r; vfg = g; vfb = b;
        if (iter > 2 && delta < 1e-8) break;
    }
    double len = vfr * vfr + vfg * vfg + vfb * vfb;
    if (len < 1e-10) *axis = *def;
    else {
        len = 1.0f / sqrt(len);
        axis->v[0] = (float)(vfr * len); axis->v[1] = (float)(vfg * len); axis->v[2] = (float)(vfb * len);
    }
}

/* intersection::ray_aabb with the unit cube, crn_intersect.h:44-132.  Returns 1 on cSuccess. */
static int ray_unit_cube(vec3* coord, const vec3* org, const vec3* dir)
{
    int quadrant[3], inside = 1;
    float plane[3];
    for (int i = 0; i < 3; i++) {
        if (org->v[i] < 0.0f) { quadrant[i] = 1; plane[i] = 0.0f; inside = 0; }
        else if (org->v[i] > 1.0f) { quadrant[i] = 0; plane[i] = 1.0f; inside = 0; }
        else quadrant[i] = 2;
    }
    if (inside) { *coord = *org; return 0; /* cInside != cSuccess */ }
    float max_t[3];
    for (int i = 0; i < 3; i++) {
        if (quadrant[i] != 2 && dir->v[i] != 0.0f) max_t[i] = (plane[i] - org->v[i]) / dir->v[i];
        else max_t[i] = -1.0f;
    }
    int which = 0;
    for (int i = 1; i < 3; i++) if (max_t[which] < max_t[i]) which = i;
    if (max_t[which] < 0.0f) return 0;
    for (int i = 0; i < 3; i++) {
        if (i != which) {
            coord->v[i] = org->v[i] + max_t[which] * dir->v[i];
            if (coord->v[i] < 0.0f || coord->v[i] > 1.0f) return 0;
        } else coord->v[i] = plane[i];
    }
    return 1;
}

/* crn_dxt1.cpp:1069-1178 */
static void handle_multicolor_block(opt* o, op_dxt1_result* res, uint8_t* selectors)
{
    uint32_t num_passes = 1;
    vec3 pw = { { 1.0f, 1.0f, 1.0f } };
    if (o->perceptual) {
        float ave_redness = 0, ave_blueness = 0, ave_l = 0;
        for (uint32_t i = 0; i < o->U; i++) {
            const uint8_t* c = o->uc[i].c;
            int l = (c[0] + c[1] + c[2] + 1) / 3;
            float fl = (float)l;
            float scale = (float)o->uc[i].w / (1.0f > fl ? 1.0f : fl);
            ave_redness += scale * c[0];
            ave_blueness += scale * c[2];
            ave_l += l;
        }
        ave_redness /= o->total_w; ave_blueness /= o->total_w; ave_l /= o->total_w;
        ave_l = ave_l * 16.0f / 255.0f;
        ave_l = 1.0f < ave_l ? 1.0f : ave_l;
        float mx = ave_redness > ave_blueness ? ave_redness : ave_blueness;
        float sat = mx * 1.0f / 3.0f;
        sat = sat < 0.0f ? 0.0f : (sat > 1.0f ? 1.0f : sat);
        float p = ave_l * powf(sat, 2.75f);
        if (p >= 1.0f) num_passes = 1;
        else {
            num_passes = 2;
            static const float base[3] = { .212f, .72f, .072f };
            for (int k = 0; k < 3; k++) pw.v[k] = base[k] + (pw.v[k] - base[k]) * p;
        }
    }
    for (uint32_t pass = 0; pass < num_passes; pass++) {
        compute_vectors(o, &pw);
        static const vec3 def = { { .2837149f, 0.9540631f, 0.096277453f } };
        compute_pca(o, &o->axis, o->normw, &def);
        for (int k = 0; k < 3; k++) o->axis.v[k] /= pw.v[k];
        {   /* vec::normalize, crn_vec.h:674-689 */
            double n = o->axis.v[0] * o->axis.v[0];
            n += o->axis.v[1] * o->axis.v[1];
            n += o->axis.v[2] * o->axis.v[2];
            if (n != 0) { float s = (float)(1.0f / sqrt(n)); for (int k = 0; k < 3; k++) o->axis.v[k] *= s; }
        }
        if (num_passes > 1) {
            if (fabsf(o->axis.v[0]) >= .795f) { pw.v[0] = .424f; pw.v[1] = .6f; pw.v[2] = .072f; }
            else if (fabsf(o->axis.v[2]) >= .795f) { pw.v[0] = .212f; pw.v[1] = .6f; pw.v[2] = .212f; }
            else break;
        }
    }
    float l = 1e+9f, h = -1e+9f;
    for (uint32_t i = 0; i < o->U; i++) {
        float d = o->norm[i].v[0] * o->axis.v[0];
        d += o->norm[i].v[1] * o->axis.v[1];
        d += o->norm[i].v[2] * o->axis.v[2];
        l = l < d ? l : d;
        h = h > d ? h : d;
    }
    vec3 low, high;
    for (int k = 0; k < 3; k++) { low.v[k] = o->mean.v[k] + o->axis.v[k] * l; high.v[k] = o->mean.v[k] + o->axis.v[k] * h; }
    int in_low = 1, in_high = 1;
    for (int k = 0; k < 3; k++) {
        if (low.v[k] < 0.0f || low.v[k] > 1.0f) in_low = 0;
        if (high.v[k] < 0.0f || high.v[k] > 1.0f) in_high = 0;
    }
    if (!in_low) { vec3 coord; if (ray_unit_cube(&coord, &low, &o->axis)) low = coord; }
    if (!in_high) {
        vec3 coord, neg;
        for (int k = 0; k < 3; k++) neg.v[k] = -o->axis.v[k];
        if (ray_unit_cube(&coord, &high, &neg)) high = coord;
    }
    optimize_endpoints(o, low, high, res, selectors);
}

/* crn_dxt1.cpp:2081-2232 */
static void compute_internal(opt* o, op_dxt1_result* res, uint8_t* selectors)
{
    const op_dxt1_params* p = o->p;
    o->evaluate_hc = p->quality == 4 && !p->pixels_have_alpha && !p->force_alpha_blocks && !p->use_alpha_blocks && !p->grayscale_sampling;
    o->perceptual = p->perceptual && !p->grayscale_sampling;
    o->U = 0; o->total_w = 0; o->tried_cnt = 0;
    if (o->tried) memset(o->tried, 0, sizeof(uint32_t) * 2 * o->tried_cap);
    for (uint32_t i = 0; i <= o->map_mask; i++) o->map_val[i] = -1;
    o->best.lo = o->best.hi = 0; o->best.err = UINT64_MAX; o->best.alpha_block = 0;
    o->best.alt_round = 0; o->best.enforce = 0; o->best.enforced_sel = 0;
    for (uint32_t i = 0; i < o->n; i++) {
        const uint8_t* px = o->pixels + 4 * i;
        if (!p->pixels_have_alpha || px[3] >= p->alpha_threshold) {
            uint32_t key;
            memcpy(&key, px, 4);
            key |= 0xFF000000u;
            uint32_t h = (key * 2654435761u) & o->map_mask;
            while (o->map_val[h] >= 0 && o->map_key[h] != key) h = (h + 1) & o->map_mask;
            if (o->map_val[h] < 0) {
                o->map_key[h] = key; o->map_val[h] = (int32_t)o->U;
                memcpy(o->uc[o->U].c, px, 3); o->uc[o->U].c[3] = 255; o->uc[o->U].w = 1;
                o->U++;
            } else o->uc[o->map_val[h]].w++;
            o->total_w++;
        }
    }
    o->has_transparent = o->total_w != o->n;
    /* channel lower-bound tables, crn_dxt1.cpp:2134-2203 */
    uint64_t pw_[3][64] = { { 0 } }, pc_[3][64] = { { 0 } }, ps_[3][64] = { { 0 } };
    for (uint32_t i = 0; i < o->U; i++) {
        const ucolor* c = &o->uc[i];
        uint8_t R = c->c[0], r = (uint8_t)((R >> 3) + ((R & 7) > (R >> 5) ? 1 : 0));
        uint8_t G = c->c[1], g = (uint8_t)((G >> 2) + ((G & 3) > (G >> 6) ? 1 : 0));
        uint8_t B = c->c[2], b = (uint8_t)((B >> 3) + ((B & 7) > (B >> 5) ? 1 : 0));
        pw_[0][r] += c->w; pc_[0][r] += (uint64_t)c->w * R; ps_[0][r] += (uint64_t)c->w * R * R;
        pw_[1][g] += c->w; pc_[1][g] += (uint64_t)c->w * G; ps_[1][g] += (uint64_t)c->w * G * G;
        pw_[2][b] += c->w; pc_[2][b] += (uint64_t)c->w * B; ps_[2][b] += (uint64_t)c->w * B * B;
    }
    if (o->perceptual) {
        for (int c = 0; c < 32; c++) { pw_[0][c] *= 8; pc_[0][c] *= 8; ps_[0][c] *= 8; }
        for (int c = 0; c < 64; c++) { pw_[1][c] *= 25; pc_[1][c] *= 25; ps_[1][c] *= 25; }
    }
    for (int ch = 0; ch < 3; ch++) {
        int n = ch == 1 ? 64 : 32;
        for (int c = 1; c < n; c++) { pw_[ch][c] += pw_[ch][c - 1]; pc_[ch][c] += pc_[ch][c - 1]; ps_[ch][c] += ps_[ch][c - 1]; }
    }
    for (int c = 0; c < 32; c++) {
        uint64_t C = (uint8_t)(c << 3 | c >> 2);
        o->rlo[c] = ps_[0][c] + C * C * pw_[0][c] - 2 * C * pc_[0][c];
        o->rhi[c] = ps_[0][31] + C * C * pw_[0][31] - 2 * C * pc_[0][31] - o->rlo[c];
        o->blo[c] = ps_[2][c] + C * C * pw_[2][c] - 2 * C * pc_[2][c];
        o->bhi[c] = ps_[2][31] + C * C * pw_[2][31] - 2 * C * pc_[2][31] - o->blo[c];
    }
    for (int c = 0; c < 64; c++) {
        uint64_t C = (uint8_t)(c << 2 | c >> 4);
        o->glo[c] = ps_[1][c] + C * C * pw_[1][c] - 2 * C * pc_[1][c];
        o->ghi[c] = ps_[1][63] + C * C * pw_[1][63] - 2 * C * pc_[1][63] - o->glo[c];
    }
    if (!o->U) {
        res->low = 0; res->high = 0; res->alpha_block = 1;
        memset(selectors, 3, o->n);
        /* m_error is left untouched by the reference here; report 0 */
        res->error = 0;
    } else if (o->U == 1 && !o->has_transparent) {
        int r = o->uc[0].c[0], g = o->uc[0].c[1], b = o->uc[0].c[2];
        evaluate(o, (uint16_t)((OMatch5[r][0] << 11) | (OMatch6[g][0] << 5) | OMatch5[b][0]),
                 (uint16_t)((OMatch5[r][1] << 11) | (OMatch6[g][1] << 5) | OMatch5[b][1]), 0);
        if (p->use_alpha_blocks && o->best.err)
            evaluate(o, (uint16_t)((OMatch5_3[r][0] << 11) | (OMatch6_3[g][0] << 5) | OMatch5_3[b][0]),
                     (uint16_t)((OMatch5_3[r][1] << 11) | (OMatch6_3[g][1] << 5) | OMatch5_3[b][1]), 0);
        return_solution(o, res, selectors);
    } else {
        handle_multicolor_block(o, res, selectors);
    }
}

static void opt_alloc(opt* o, uint32_t n)
{
    memset(o, 0, sizeof(*o));
    o->uc = (ucolor*)malloc(sizeof(ucolor) * (n + 1));
    o->norm = (vec3*)malloc(sizeof(vec3) * (n + 1));
    o->normw = (vec3*)malloc(sizeof(vec3) * (n + 1));
    o->best.sel = (uint8_t*)calloc(n + 1, 1);
    o->trial.sel = (uint8_t*)calloc(n + 1, 1);
    o->trial_sel = (uint8_t*)calloc(n + 1, 1);
    uint32_t cap = 16;
    while (cap < 2 * n + 2) cap <<= 1;
    o->map_mask = cap - 1;
    o->map_key = (uint32_t*)malloc(sizeof(uint32_t) * cap);
    o->map_val = (int32_t*)malloc(sizeof(int32_t) * cap);
}
static void opt_free(opt* o)
{
    free(o->uc); free(o->norm); free(o->normw); free(o->best.sel); free(o->trial.sel); free(o->trial_sel);
    free(o->map_key); free(o->map_val); free(o->tried);
}

static void run_once(const uint8_t* pixels, uint32_t n, const op_dxt1_params* p, op_dxt1_result* r, uint8_t* selectors)
{
    opt o;
    opt_alloc(&o, n);
    o.p = p; o.pixels = pixels; o.n = n;
    compute_internal(&o, r, selectors);
    opt_free(&o);
}

/* 3-colour palette of an encoded block, dxt1_block::get_block_colors3 (crn_dxt.cpp:234-260) */
static void block_colors3(uint16_t c0, uint16_t c1, int (*c)[3])
{
    unpack565(c0, 1, &c[0][0], &c[0][1], &c[0][2]);
    unpack565(c1, 1, &c[1][0], &c[1][1], &c[1][2]);
    for (int k = 0; k < 3; k++) { c[2][k] = (c[0][k] + c[1][k]) >> 1; c[3][k] = 0; }
}

int op_dxt1_optimize(const uint8_t* pixels, uint32_t n, const op_dxt1_params* p, op_dxt1_result* r, uint8_t* selectors)
{
    if (!pixels) return 0;
    init_tables();
    run_once(pixels, n, p, r, selectors);
    /* try_alpha_as_black_optimization, crn_dxt1.cpp:2001-2079, :2241-2244 */
    if (p->use_alpha_blocks && p->transparent_for_black && !p->pixels_have_alpha) {
        uint32_t dark = 0, uniq_dark = 0, uniq = 0;
        /* the reference counts UNIQUE colours; recount on unique set */
        {
            opt o; opt_alloc(&o, n); o.p = p; o.pixels = pixels; o.n = n;
            op_dxt1_result tmp; uint8_t* ts = (uint8_t*)malloc(n ? n : 1);
            compute_internal(&o, &tmp, ts);
            uniq = o.U;
            for (uint32_t i = 0; i < o.U; i++) if (o.uc[i].c[0] <= 4 && o.uc[i].c[1] <= 4 && o.uc[i].c[2] <= 4) uniq_dark++;
            free(ts); opt_free(&o);
        }
        (void)dark;
        if (!uniq_dark || uniq_dark == uniq) return 1;
        uint8_t* tc = (uint8_t*)malloc(4 * (size_t)n);
        memcpy(tc, pixels, 4 * (size_t)n);
        for (uint32_t i = 0; i < n; i++) if (tc[4 * i] <= 4 && tc[4 * i + 1] <= 4 && tc[4 * i + 2] <= 4) tc[4 * i + 3] = 0;
        op_dxt1_params tp = *p;
        tp.pixels_have_alpha = 1;
        op_dxt1_result tr; uint8_t* tsel = (uint8_t*)malloc(n);
        run_once(tc, n, &tp, &tr, tsel);
        int c[4][3];
        block_colors3(tr.low, tr.high, c);
        int perceptual = p->perceptual && !p->grayscale_sampling;
        opt dummy; memset(&dummy, 0, sizeof(dummy)); dummy.p = p;
        uint64_t te = 0;
        for (uint32_t i = 0; i < n; i++) te += cdist(&dummy, perceptual, tc + 4 * i, c[tsel[i]][0], c[tsel[i]][1], c[tsel[i]][2]);
        if (te < r->error) {
            r->error = te; r->low = tr.low; r->high = tr.high; r->alpha_block = 1;
            memcpy(selectors, tsel, n);
        }
        free(tc); free(tsel);
    }
    return 1;
}
