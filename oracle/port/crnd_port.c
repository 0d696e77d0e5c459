/* oracle/port/crnd_port.c -- TEST INFRASTRUCTURE ONLY (see oracle_port.h).
 *
 * Sequential restatement of the CRN -> DXTn transcoder (reference inc/crn_decomp.h): header crack
 * (:2657-2670, inc/crn_defs.h:286-341), static Huffman model receive (:3044-3123), canonical decoder
 * (:2150-2326, :3186-3253; restated as a plain canonical-code length search, which yields the same
 * symbol for every valid stream), palette decode (:3715-3851) and the per-level block loops for
 * DXT1 / DXT5 (+swizzled variants) / DXN / DXT5A (:3944-4223).  ETC formats are out of scope.
 */
#include "oracle_port.h"
#include <stdlib.h>
#include <string.h>

typedef struct {
    const uint8_t* p; uint32_t size; uint64_t bitpos;   /* logical MSB-first stream, zero padded past the end */
} bitrd;

static uint32_t peek16(const bitrd* b)
{
    uint64_t byte = b->bitpos >> 3;
    uint32_t v = 0;
    for (int i = 0; i < 4; i++) v = (v << 8) | (byte + i < b->size ? b->p[byte + i] : 0);
    return (v >> (16 - (b->bitpos & 7))) & 0xffff;
}
static uint32_t getbits(bitrd* b, uint32_t n)
{
    uint32_t r = 0;
    while (n) {   /* at most 16 per step */
        uint32_t k = n > 16 ? n - 16 : n;
        r = (r << k) | (peek16(b) >> (16 - k));
        b->bitpos += k; n -= k;
    }
    return r;
}

typedef struct {
    uint32_t nsyms;
    uint8_t* len;                 /* code size per symbol */
    uint16_t* sorted;             /* symbols sorted by (length, index) */
    uint32_t first_code[18], first_idx[18], count[18];
} hmodel;

static void hm_free(hmodel* m) { free(m->len); free(m->sorted); memset(m, 0, sizeof(*m)); }

static int hm_prepare(hmodel* m)
{   /* canonical code assignment, crn_decomp.h:2150-2235 */
    memset(m->count, 0, sizeof(m->count));
    for (uint32_t i = 0; i < m->nsyms; i++) if (m->len[i]) m->count[m->len[i]]++;
    uint32_t code = 0, idx = 0, pos[18];
    for (uint32_t l = 1; l <= 16; l++) {
        m->first_code[l] = code; m->first_idx[l] = idx; pos[l] = idx;
        code += m->count[l]; idx += m->count[l];
        code <<= 1;
    }
    m->sorted = (uint16_t*)malloc(sizeof(uint16_t) * (idx ? idx : 1));
    for (uint32_t i = 0; i < m->nsyms; i++) if (m->len[i]) m->sorted[pos[m->len[i]]++] = (uint16_t)i;
    return 1;
}
static uint32_t hm_decode(const hmodel* m, bitrd* b)
{
    uint32_t k = peek16(b);
    for (uint32_t l = 1; l <= 16; l++) {
        if (!m->count[l]) continue;
        uint32_t c = k >> (16 - l);
        if (c >= m->first_code[l] && c - m->first_code[l] < m->count[l]) {
            b->bitpos += l;
            return m->sorted[m->first_idx[l] + (c - m->first_code[l])];
        }
    }
    return 0;   /* corrupted stream: the reference returns symbol 0 without consuming bits */
}

/* decode_receive_static_data_model, crn_decomp.h:3044-3123 */
static int hm_receive(hmodel* m, bitrd* b)
{
    static const uint8_t order[21] = { 17, 18, 19, 20, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15, 16 };
    memset(m, 0, sizeof(*m));
    uint32_t total = getbits(b, 14);
    if (!total) return 1;
    m->nsyms = total;
    m->len = (uint8_t*)calloc(total, 1);
    uint32_t ncl = getbits(b, 5);
    if (ncl < 1 || ncl > 21) return 0;
    hmodel dm; memset(&dm, 0, sizeof(dm));
    dm.nsyms = 21; dm.len = (uint8_t*)calloc(21, 1);
    for (uint32_t i = 0; i < ncl; i++) dm.len[order[i]] = (uint8_t)getbits(b, 3);
    hm_prepare(&dm);
    uint32_t ofs = 0;
    int ok = 1;
    while (ofs < total) {
        uint32_t rem = total - ofs, code = hm_decode(&dm, b);
        if (code <= 16) m->len[ofs++] = (uint8_t)code;
        else if (code == 17) { uint32_t n = getbits(b, 3) + 3; if (n > rem) { ok = 0; break; } ofs += n; }
        else if (code == 18) { uint32_t n = getbits(b, 7) + 11; if (n > rem) { ok = 0; break; } ofs += n; }
        else {
            uint32_t n = code == 19 ? getbits(b, 2) + 3 : getbits(b, 6) + 7;
            if (!ofs || n > rem) { ok = 0; break; }
            uint8_t prev = m->len[ofs - 1];
            if (!prev) { ok = 0; break; }
            for (uint32_t e = ofs + n; ofs < e;) m->len[ofs++] = prev;
        }
    }
    hm_free(&dm);
    if (!ok || ofs != total) return 0;
    return hm_prepare(m);
}

static uint32_t be(const uint8_t* p, int n) { uint32_t v = 0; for (int i = 0; i < n; i++) v = (v << 8) | p[i]; return v; }

typedef struct op_crnd {
    const uint8_t* data; uint32_t size;
    uint32_t width, height, levels, faces, format;
    uint32_t pal_ofs[4], pal_size[4], pal_num[4];   /* colour endpoints, colour selectors, alpha endpoints, alpha selectors */
    uint32_t tables_ofs, tables_size;
    uint32_t level_ofs[16];
    hmodel ref_dm, ep_dm[2], sel_dm[2];
    uint32_t* color_endpoints; uint32_t* color_selectors;
    uint16_t* alpha_endpoints; uint16_t* alpha_selectors;
} op_crnd;

static int parse_header(op_crnd* c, const uint8_t* d, uint32_t size)
{   /* crnd_get_header, crn_decomp.h:2657-2670; layout inc/crn_defs.h:286-341 (74 bytes + 4 per extra level) */
    if (!d || size < 74) return 0;
    if (be(d, 2) != (('H' << 8) | 'x')) return 0;
    if (be(d + 2, 2) < 74 || size < be(d + 6, 4)) return 0;
    c->data = d; c->size = size;
    c->width = be(d + 12, 2); c->height = be(d + 14, 2); c->levels = d[16]; c->faces = d[17]; c->format = d[18];
    for (int i = 0; i < 4; i++) {
        const uint8_t* q = d + 33 + 8 * i;
        c->pal_ofs[i] = be(q, 3); c->pal_size[i] = be(q + 3, 3); c->pal_num[i] = be(q + 6, 2);
    }
    c->tables_size = be(d + 65, 2); c->tables_ofs = be(d + 67, 3);
    if (c->levels < 1 || c->levels > 16) return 0;
    for (uint32_t i = 0; i < c->levels; i++) c->level_ofs[i] = be(d + 70 + 4 * i, 4);
    return 1;
}

int op_crnd_info(const uint8_t* data, uint32_t size, uint32_t* out)
{
    op_crnd c; memset(&c, 0, sizeof(c));
    if (!parse_header(&c, data, size)) return 0;
    out[0] = c.width; out[1] = c.height; out[2] = c.levels; out[3] = c.faces;
    out[4] = (c.format == 0 || c.format == 9) ? 8 : 16; out[5] = c.format; out[6] = be(data + 25, 4); out[7] = be(data + 29, 4);
    return 1;
}

void op_crnd_end(op_crnd* c)
{
    if (!c) return;
    hm_free(&c->ref_dm); hm_free(&c->ep_dm[0]); hm_free(&c->ep_dm[1]); hm_free(&c->sel_dm[0]); hm_free(&c->sel_dm[1]);
    free(c->color_endpoints); free(c->color_selectors); free(c->alpha_endpoints); free(c->alpha_selectors);
    free(c);
}

op_crnd* op_crnd_begin(const uint8_t* data, uint32_t size)
{
    op_crnd* c = (op_crnd*)calloc(1, sizeof(op_crnd));
    if (!parse_header(c, data, size)) { free(c); return NULL; }
    if (c->format >= 10) { free(c); return NULL; }            /* ETC: out of scope */
    bitrd b = { data + c->tables_ofs, c->tables_size, 0 };
    /* init_tables, :3662-3692 */
    if (!hm_receive(&c->ref_dm, &b)) goto fail;
    if (!c->pal_num[0] && !c->pal_num[2]) goto fail;
    if (c->pal_num[0]) { if (!hm_receive(&c->ep_dm[0], &b) || !hm_receive(&c->sel_dm[0], &b)) goto fail; }
    if (c->pal_num[2]) { if (!hm_receive(&c->ep_dm[1], &b) || !hm_receive(&c->sel_dm[1], &b)) goto fail; }
    if (c->pal_num[0]) {
        /* decode_color_endpoints, :3715-3761 */
        hmodel dm[2];
        bitrd p = { data + c->pal_ofs[0], c->pal_size[0], 0 };
        if (!hm_receive(&dm[0], &p) || !hm_receive(&dm[1], &p)) goto fail;
        c->color_endpoints = (uint32_t*)malloc(4 * c->pal_num[0]);
        uint32_t a = 0, bb = 0, cc = 0, d = 0, e = 0, f = 0;
        for (uint32_t i = 0; i < c->pal_num[0]; i++) {
            a = (a + hm_decode(&dm[0], &p)) & 31; bb = (bb + hm_decode(&dm[1], &p)) & 63; cc = (cc + hm_decode(&dm[0], &p)) & 31;
            d = (d + hm_decode(&dm[0], &p)) & 31; e = (e + hm_decode(&dm[1], &p)) & 63; f = (f + hm_decode(&dm[0], &p)) & 31;
            c->color_endpoints[i] = cc | (bb << 5) | (a << 11) | (f << 16) | (e << 21) | (d << 27);
        }
        hm_free(&dm[0]); hm_free(&dm[1]);
        /* decode_color_selectors, :3763-3798 */
        hmodel sm;
        bitrd q = { data + c->pal_ofs[1], c->pal_size[1], 0 };
        if (!hm_receive(&sm, &q)) goto fail;
        c->color_selectors = (uint32_t*)malloc(4 * (c->pal_num[1] ? c->pal_num[1] : 1));
        uint32_t s = 0;
        for (uint32_t i = 0; i < c->pal_num[1]; i++) {
            for (uint32_t j = 0; j < 32; j += 4) s ^= hm_decode(&sm, &q) << j;
            c->color_selectors[i] = ((s ^ s << 1) & 0xAAAAAAAAu) | (s >> 1 & 0x55555555u);
        }
        hm_free(&sm);
    }
    if (c->pal_num[2]) {
        /* decode_alpha_endpoints, :3800-3827 */
        hmodel dm;
        bitrd p = { data + c->pal_ofs[2], c->pal_size[2], 0 };
        if (!hm_receive(&dm, &p)) goto fail;
        c->alpha_endpoints = (uint16_t*)malloc(2 * c->pal_num[2]);
        uint32_t a = 0, bb = 0;
        for (uint32_t i = 0; i < c->pal_num[2]; i++) {
            a = (a + hm_decode(&dm, &p)) & 255; bb = (bb + hm_decode(&dm, &p)) & 255;
            c->alpha_endpoints[i] = (uint16_t)(a | (bb << 8));
        }
        hm_free(&dm);
        /* decode_alpha_selectors, :3829-3851 */
        static const uint8_t from_linear[8] = { 0, 2, 3, 4, 5, 6, 7, 1 };   /* g_dxt5_from_linear */
        uint8_t fl[64];
        for (uint32_t i = 0; i < 64; i++) fl[i] = (uint8_t)(from_linear[i & 7] | from_linear[i >> 3] << 3);
        hmodel sm;
        bitrd q = { data + c->pal_ofs[3], c->pal_size[3], 0 };
        if (!hm_receive(&sm, &q)) goto fail;
        c->alpha_selectors = (uint16_t*)malloc(6 * (c->pal_num[3] ? c->pal_num[3] : 1));
        uint32_t s0l = 0, s1l = 0;
        for (uint32_t i = 0; i < c->pal_num[3] * 3;) {
            uint32_t s0 = 0, s1 = 0;
            for (uint32_t j = 0; j < 24; j += 6) { s0l ^= hm_decode(&sm, &q) << j; s0 |= (uint32_t)fl[s0l >> j & 0x3F] << j; }
            for (uint32_t j = 0; j < 24; j += 6) { s1l ^= hm_decode(&sm, &q) << j; s1 |= (uint32_t)fl[s1l >> j & 0x3F] << j; }
            c->alpha_selectors[i++] = (uint16_t)s0;
            c->alpha_selectors[i++] = (uint16_t)(s0 >> 16 | s1 << 8);
            c->alpha_selectors[i++] = (uint16_t)(s1 >> 8);
        }
        hm_free(&sm);
    }
    return c;
fail:
    op_crnd_end(c);
    return NULL;
}

/* unpack_level + unpack_dxt1/dxt5/dxn/dxt5a, crn_decomp.h:3552-3619, :3944-4223 */
int op_crnd_unpack_level(op_crnd* c, void** dst, uint32_t dst_size, uint32_t row_pitch, uint32_t level)
{
    if (!c || level >= c->levels) return 0;
    uint32_t cur = c->level_ofs[level], next = level + 1 < c->levels ? c->level_ofs[level + 1] : c->size;
    uint32_t w = c->width >> level; if (!w) w = 1;
    uint32_t h = c->height >> level; if (!h) h = 1;
    const uint32_t bx = (w + 3) >> 2, by = (h + 3) >> 2;
    const uint32_t fmt = c->format;
    const uint32_t bs = (fmt == 0 || fmt == 9) ? 8 : 16;
    uint32_t minpitch = bs * bx;
    if (!row_pitch) row_pitch = minpitch;
    else if (row_pitch < minpitch || (row_pitch & 3)) return 0;
    if (dst_size < row_pitch * by) return 0;
    bitrd b = { c->data + cur, next - cur, 0 };
    const int has_color = fmt <= 6, has_a0 = fmt != 0, has_a1 = fmt == 7 || fmt == 8;
    const int is_dxn = has_a1;
    const uint32_t W = (bx + 1) & ~1u, H = (by + 1) & ~1u;
    typedef struct { uint16_t ref, ce, a0, a1; } bbuf;
    bbuf* buf = (bbuf*)calloc(W, sizeof(bbuf));
    uint32_t ce = 0, a0 = 0, a1 = 0;
    uint8_t group = 0;
    const uint32_t nce = c->pal_num[0], nae = c->pal_num[2];
    for (uint32_t f = 0; f < c->faces; f++) {
        for (uint32_t y = 0; y < H; y++) {
            int visible = y < by;
            for (uint32_t x = 0; x < W; x++) {
                visible = visible && x < bx;
                if (!(y & 1) && !(x & 1)) group = (uint8_t)hm_decode(&c->ref_dm, &b);
                uint8_t r;
                if (y & 1) r = (uint8_t)buf[x].ref;
                else { r = group & 3; group >>= 2; buf[x].ref = group & 3; group >>= 2; }
                if (!r) {
                    if (has_color && !is_dxn) { ce += hm_decode(&c->ep_dm[0], &b); if (ce >= nce) ce -= nce; buf[x].ce = (uint16_t)ce; }
                    if (has_a0) { a0 += hm_decode(&c->ep_dm[1], &b); if (a0 >= nae) a0 -= nae; buf[x].a0 = (uint16_t)a0; }
                    if (has_a1) { a1 += hm_decode(&c->ep_dm[1], &b); if (a1 >= nae) a1 -= nae; buf[x].a1 = (uint16_t)a1; }
                } else if (r == 1) { buf[x].ce = (uint16_t)ce; buf[x].a0 = (uint16_t)a0; buf[x].a1 = (uint16_t)a1; }
                else { ce = buf[x].ce; a0 = buf[x].a0; a1 = buf[x].a1; }
                uint32_t cs = 0, s0 = 0, s1 = 0;
                if (has_color && !is_dxn) cs = hm_decode(&c->sel_dm[0], &b);
                if (has_a0) s0 = hm_decode(&c->sel_dm[1], &b);
                if (has_a1) s1 = hm_decode(&c->sel_dm[1], &b);
                if (visible) {
                    uint32_t* o = (uint32_t*)((uint8_t*)dst[f] + (size_t)y * row_pitch + (size_t)x * bs);
                    if (fmt == 0) { o[0] = c->color_endpoints[ce]; o[1] = c->color_selectors[cs]; }
                    else {
                        const uint16_t* as0 = &c->alpha_selectors[s0 * 3];
                        o[0] = c->alpha_endpoints[a0] | ((uint32_t)as0[0] << 16);
                        o[1] = as0[1] | ((uint32_t)as0[2] << 16);
                        if (is_dxn) {
                            const uint16_t* as1 = &c->alpha_selectors[s1 * 3];
                            o[2] = c->alpha_endpoints[a1] | ((uint32_t)as1[0] << 16);
                            o[3] = as1[1] | ((uint32_t)as1[2] << 16);
                        } else if (fmt != 9) { o[2] = c->color_endpoints[ce]; o[3] = c->color_selectors[cs]; }
                    }
                }
            }
        }
    }
    free(buf);
    return 1;
}
