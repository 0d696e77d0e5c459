/* oracle/port/dxt5_port.c -- TEST INFRASTRUCTURE ONLY (see oracle_port.h).
 * Sequential restatement of crnlib::dxt5_endpoint_optimizer (reference crnlib/crn_dxt5a.cpp:40-262),
 * including the m_flags symmetric-dedup bitmap (:107-144) and every early-out. */
#include "oracle_port.h"
#include <stdlib.h>
#include <string.h>
#include <limits.h>

typedef struct {
    uint32_t U;
    uint8_t val[256];
    uint32_t wgt[256];
    uint8_t trial[256], best[256];
    int both;
    uint64_t err;
    uint8_t first, second, block_type;
} d5;

static void values8(uint32_t* v, uint32_t l, uint32_t h)
{   /* crn_dxt.cpp:418-430 */
    v[0] = l; v[1] = h; v[2] = (l * 6 + h) / 7; v[3] = (l * 5 + h * 2) / 7; v[4] = (l * 4 + h * 3) / 7;
    v[5] = (l * 3 + h * 4) / 7; v[6] = (l * 2 + h * 5) / 7; v[7] = (l + h * 6) / 7;
}
static void values6(uint32_t* v, uint32_t l, uint32_t h)
{   /* crn_dxt.cpp:404-416 */
    v[0] = l; v[1] = h; v[2] = (l * 4 + h) / 5; v[3] = (l * 3 + h * 2) / 5; v[4] = (l * 2 + h * 3) / 5; v[5] = (l + h * 4) / 5;
    v[6] = 0; v[7] = 255;
}

/* crn_dxt5a.cpp:198-262 */
static void evaluate(d5* o, uint32_t l, uint32_t h)
{
    for (uint32_t bt = 0; bt < (o->both ? 2u : 1u); bt++) {
        uint32_t sv[8];
        if (!bt) values8(sv, l, h); else values6(sv, l, h);
        uint64_t te = 0;
        for (uint32_t i = 0; i < o->U; i++) {
            const uint32_t val = o->val[i], weight = o->wgt[i];
            uint32_t bse = UINT_MAX, bs = 0;
            for (uint32_t j = 0; j < 8; j++) {
                uint32_t d = val - sv[j];
                uint32_t se = d * d * weight;          /* the reference's int product, same bits */
                if (se < bse) { bse = se; bs = j; if (!bse) break; }
            }
            o->trial[i] = (uint8_t)bs;
            te += bse;
            if (te > o->err) break;
        }
        if (te < o->err) {
            o->err = te; o->first = (uint8_t)l; o->second = (uint8_t)h; o->block_type = (uint8_t)bt;
            memcpy(o->best, o->trial, o->U);
            if (!te) break;
        }
    }
}

int op_dxt5_optimize(const uint8_t* pixels, uint32_t n, uint32_t comp, uint32_t quality, uint32_t both,
                     uint8_t* first, uint8_t* second, uint8_t* selectors, uint64_t* error, uint8_t* block_type)
{
    if (!n || !pixels) return 0;
    d5* o = (d5*)calloc(1, sizeof(d5));
    int map[256];
    for (int i = 0; i < 256; i++) map[i] = -1;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t a = pixels[4 * i + comp];
        if (map[a] < 0) { map[a] = (int)o->U; o->val[o->U] = (uint8_t)a; o->wgt[o->U] = 0; o->U++; }
        o->wgt[map[a]]++;
    }
    o->both = both != 0;
    if (o->U == 1) {
        *block_type = 0; *error = 0; *first = *second = o->val[0];
        memset(selectors, 0, n);
        free(o);
        return 1;
    }
    o->err = UINT64_MAX;
    for (uint32_t i = 0; i + 1 < o->U; i++)
        for (uint32_t j = i + 1; j < o->U; j++)
            evaluate(o, o->val[i], o->val[j]);
    if (quality >= 3 && o->err) {
        uint8_t* flags = (uint8_t*)calloc(65536 / 8, 1);
        const int P = quality == 4 ? 16 : 8;
        for (int ld = -P; ld <= P; ld++) {
            const int l = o->first + ld;
            if (l < 0) continue; else if (l > 255) break;
            for (int hd = -P; hd <= P; hd++) {
                const int h = o->second + hd;
                if (h < 0) continue; else if (h > 255) break;
                uint32_t b0 = (uint32_t)l * 256 + (uint32_t)h, b1 = (uint32_t)h * 256 + (uint32_t)l;
                if ((flags[b0 >> 3] >> (b0 & 7) & 1) || (flags[b1 >> 3] >> (b1 & 7) & 1)) continue;
                flags[b0 >> 3] |= (uint8_t)(1u << (b0 & 7));
                evaluate(o, (uint32_t)l, (uint32_t)h);
            }
        }
        free(flags);
    }
    static const uint8_t six_inv[8] = { 1, 0, 5, 4, 3, 2, 6, 7 }, eight_inv[8] = { 1, 0, 7, 6, 5, 4, 3, 2 };  /* crn_dxt.cpp:41-42 */
    if (o->first == o->second) memset(o->best, 0, o->U);
    else if (o->block_type) {
        if (o->first > o->second) {
            uint8_t t = o->first; o->first = o->second; o->second = t;
            for (uint32_t i = 0; i < o->U; i++) o->best[i] = six_inv[o->best[i]];
        }
    } else if (o->first <= o->second) {
        uint8_t t = o->first; o->first = o->second; o->second = t;
        for (uint32_t i = 0; i < o->U; i++) o->best[i] = eight_inv[o->best[i]];
    }
    for (uint32_t i = 0; i < n; i++) selectors[i] = o->best[map[pixels[4 * i + comp]]];
    *first = o->first; *second = o->second; *error = o->err; *block_type = o->block_type;
    free(o);
    return 1;
}
