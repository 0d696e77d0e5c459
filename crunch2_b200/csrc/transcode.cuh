// transcode.cuh -- CRN -> DXTn transcoder kernels for sm_100a.
//
// Replaces crnd::crn_unpacker::decode_palettes / unpack_level (+ symbol_codec::decode), reference
// inc/crn_decomp.h:3694-3851, :3552-3619, :3944-4223, :3186-3253.
//
// What the format allows.  This CRN revision carries ONE byte-aligned Huffman bitstream per mip level
// (all faces concatenated, no chunks, no resync markers); which of five code tables applies
// to the next symbol depends on the block position and on reference symbols decoded up to a row
// earlier.  The entropy decode of one level is therefore a serial dependency chain; the parallelism
// the format offers is (file x level).  The kernel maps one WARP to one level stream:
//   * lane 0 walks the bitstream (64-bit MSB-first window in registers, 11-bit first-level lookup in
//     shared memory, canonical-code search for longer codes) and resolves the left / top endpoint
//     references, 32 visible blocks at a time, into a small shared-memory index batch;
//   * then all 32 lanes expand the batch: palette gathers (<= 144 KB per file, L1/L2 resident) and one
//     coalesced 8/16-byte store per lane -- the bandwidth-bound half of the job.
// One CTA serves one file (16 warps = up to 16 levels decode concurrently, sharing the file's tables).
#pragma once
#include "warp_util.cuh"

namespace crn {

constexpr int kHuffLookupBits = 11;
constexpr int kHuffLookupSize = 1 << kHuffLookupBits;
constexpr int kTranscodeWarps = 16;
constexpr uint32_t kHuffLong = 0xFFFFFFFFu;

// Canonical Huffman decoder tables of one model, built on the host (see HuffHost in crn_b200.cu).
struct HuffModelDev {
    uint32_t lookup[kHuffLookupSize];   // top-11-bit prefix -> sym | len << 16, or kHuffLong
    uint32_t limit[17];                 // limit[l] = (first_code[l] + count[l]) << (16 - l): exclusive, left-justified
    uint32_t first_code[17];
    uint32_t first_idx[17];
    uint32_t sorted_ofs;                // into the file's uint16 sorted-symbol pool
    uint32_t nsyms;
    uint32_t pad[2];
};

enum { kDmRef = 0, kDmColorEp = 1, kDmColorSel = 2, kDmAlphaEp = 3, kDmAlphaSel = 4, kNumBlockModels = 5,
       kDmPalCe0 = 5, kDmPalCe1 = 6, kDmPalCs = 7, kDmPalAe = 8, kDmPalAs = 9, kNumModels = 10 };

struct LevelStream {
    uint32_t src_ofs, src_size;         // level bitstream within the file bytes
    uint32_t blocks_x, blocks_y;        // visible blocks
    uint32_t row_pitch;                 // bytes
    uint32_t rowbuf_ofs;                // into the file's row-buffer pool (entries)
    uint32_t active;
    uint32_t pad;
    unsigned long long dst[6];          // device pointers, one per face
};

struct TranscodeFile {                  // everything the kernels need for one .crn, in device memory
    const uint8_t* bytes;               // file image, zero padded by 16 bytes
    const HuffModelDev* models;         // kNumModels
    const uint16_t* sorted_pool;
    uint32_t* color_endpoints;          // lo565 | hi565 << 16
    uint32_t* color_selectors;
    uint16_t* alpha_endpoints;          // lo | hi << 8
    uint16_t* alpha_selectors;          // 3 x uint16 per entry
    uint2* rowbuf_pool;                 // per level: padded-width entries {ref | ce << 16, a0 | a1 << 16}
    uint32_t num_color_endpoints, num_color_selectors, num_alpha_endpoints, num_alpha_selectors;
    uint32_t pal_data_ofs[4], pal_data_bit[4], pal_size_end[4];   // first symbol of each palette stream (byte, bit) and segment end
    uint32_t format, faces;
    LevelStream levels[16];
};

// MSB-first bit window: `buf` holds `cnt` valid bits left-justified.
struct BitWindow {
    const uint8_t* p;
    uint32_t pos, end;                  // next byte to fetch, one past the last valid byte
    unsigned long long buf;
    int cnt;
};
__device__ __forceinline__ void bw_refill(BitWindow& w)
{
    while (w.cnt <= 56) {
        const unsigned long long b = w.pos < w.end ? w.p[w.pos] : 0ull;   // zero padding past the end (crn_decomp.h:3168-3170)
        w.pos++;
        w.buf |= b << (56 - w.cnt);
        w.cnt += 8;
    }
}
__device__ __forceinline__ void bw_init(BitWindow& w, const uint8_t* p, uint32_t ofs, uint32_t end, uint32_t bit)
{
    w.p = p; w.pos = ofs; w.end = end; w.buf = 0; w.cnt = 0;
    bw_refill(w);
    w.buf <<= bit; w.cnt -= (int)bit;
    bw_refill(w);
}
__device__ __forceinline__ uint32_t bw_decode(BitWindow& w, const uint32_t* lookup, const HuffModelDev* m, const uint16_t* pool)
{
    const uint32_t t = lookup[(uint32_t)(w.buf >> (64 - kHuffLookupBits))];
    uint32_t sym, len;
    if (t != kHuffLong) { sym = t & 0xffffu; len = t >> 16; }
    else {
        const uint32_t k = (uint32_t)(w.buf >> 48);
        len = kHuffLookupBits + 1;
        while (len < 16 && k >= m->limit[len]) len++;
        sym = pool[m->sorted_ofs + m->first_idx[len] + ((k >> (16 - len)) - m->first_code[len])];
    }
    w.buf <<= len; w.cnt -= (int)len;
    if (w.cnt <= 32) bw_refill(w);
    return sym;
}

// ---- palettes (decode_palettes, crn_decomp.h:3694-3851): one warp (lane 0) per palette ----------
__global__ void __launch_bounds__(128) transcode_palettes_kernel(const TranscodeFile* __restrict__ files)
{
    const TranscodeFile& f = files[blockIdx.x];
    const unsigned warp = threadIdx.x >> 5;
    if (lane_id() != 0) return;
    const HuffModelDev* M = f.models;
    BitWindow w;
    if (warp == 0 && f.num_color_endpoints) {   // :3715-3761
        bw_init(w, f.bytes, f.pal_data_ofs[0], f.pal_size_end[0], f.pal_data_bit[0]);
        uint32_t a = 0, b = 0, c = 0, d = 0, e = 0, g = 0;
        const HuffModelDev* m0 = &M[kDmPalCe0]; const HuffModelDev* m1 = &M[kDmPalCe1];
        for (uint32_t i = 0; i < f.num_color_endpoints; i++) {
            a = (a + bw_decode(w, m0->lookup, m0, f.sorted_pool)) & 31;
            b = (b + bw_decode(w, m1->lookup, m1, f.sorted_pool)) & 63;
            c = (c + bw_decode(w, m0->lookup, m0, f.sorted_pool)) & 31;
            d = (d + bw_decode(w, m0->lookup, m0, f.sorted_pool)) & 31;
            e = (e + bw_decode(w, m1->lookup, m1, f.sorted_pool)) & 63;
            g = (g + bw_decode(w, m0->lookup, m0, f.sorted_pool)) & 31;
            f.color_endpoints[i] = c | (b << 5) | (a << 11) | (g << 16) | (e << 21) | (d << 27);
        }
    } else if (warp == 1 && f.num_color_endpoints) {   // :3763-3798
        bw_init(w, f.bytes, f.pal_data_ofs[1], f.pal_size_end[1], f.pal_data_bit[1]);
        const HuffModelDev* m = &M[kDmPalCs];
        uint32_t s = 0;
        for (uint32_t i = 0; i < f.num_color_selectors; i++) {
            for (uint32_t j = 0; j < 32; j += 4) s ^= bw_decode(w, m->lookup, m, f.sorted_pool) << j;
            f.color_selectors[i] = ((s ^ s << 1) & 0xAAAAAAAAu) | (s >> 1 & 0x55555555u);
        }
    } else if (warp == 2 && f.num_alpha_endpoints) {   // :3800-3827
        bw_init(w, f.bytes, f.pal_data_ofs[2], f.pal_size_end[2], f.pal_data_bit[2]);
        const HuffModelDev* m = &M[kDmPalAe];
        uint32_t a = 0, b = 0;
        for (uint32_t i = 0; i < f.num_alpha_endpoints; i++) {
            a = (a + bw_decode(w, m->lookup, m, f.sorted_pool)) & 255;
            b = (b + bw_decode(w, m->lookup, m, f.sorted_pool)) & 255;
            f.alpha_endpoints[i] = (uint16_t)(a | (b << 8));
        }
    } else if (warp == 3 && f.num_alpha_endpoints) {   // :3829-3851
        bw_init(w, f.bytes, f.pal_data_ofs[3], f.pal_size_end[3], f.pal_data_bit[3]);
        const HuffModelDev* m = &M[kDmPalAs];
        uint32_t s0l = 0, s1l = 0;
        for (uint32_t i = 0; i < f.num_alpha_selectors; i++) {
            uint32_t s0 = 0, s1 = 0;
            for (uint32_t j = 0; j < 24; j += 6) {
                s0l ^= bw_decode(w, m->lookup, m, f.sorted_pool) << j;
                const uint32_t v = s0l >> j & 0x3F;   // two linear 3-bit selectors -> DXT5 order {0,2,3,4,5,6,7,1} (g_dxt5_from_linear)
                const uint32_t lo = v & 7, hi = v >> 3;
                s0 |= ((lo == 0 ? 0u : (lo == 7 ? 1u : lo + 1)) | (hi == 0 ? 0u : (hi == 7 ? 1u : hi + 1)) << 3) << j;
            }
            for (uint32_t j = 0; j < 24; j += 6) {
                s1l ^= bw_decode(w, m->lookup, m, f.sorted_pool) << j;
                const uint32_t v = s1l >> j & 0x3F;
                const uint32_t lo = v & 7, hi = v >> 3;
                s1 |= ((lo == 0 ? 0u : (lo == 7 ? 1u : lo + 1)) | (hi == 0 ? 0u : (hi == 7 ? 1u : hi + 1)) << 3) << j;
            }
            f.alpha_selectors[3 * i + 0] = (uint16_t)s0;
            f.alpha_selectors[3 * i + 1] = (uint16_t)(s0 >> 16 | s1 << 8);
            f.alpha_selectors[3 * i + 2] = (uint16_t)(s1 >> 8);
        }
    }
}

// ---- levels (unpack_level, crn_decomp.h:3552-3619, :3944-4223) ----------------------------------
struct BlockBatch {                     // per-warp shared memory: 32 visible blocks awaiting expansion
    uint32_t out_ofs[32];               // byte offset within the face
    uint16_t ce[32], cs[32], a0[32], s0[32], a1[32], s1[32];
    uint8_t face[32];
    uint32_t n, done;
};

struct TranscodeSmem {
    uint32_t lookup[kNumBlockModels][kHuffLookupSize];
    BlockBatch batch[kTranscodeWarps];
};

__global__ void __launch_bounds__(kTranscodeWarps * 32) transcode_levels_kernel(const TranscodeFile* __restrict__ files)
{
    CRN_DYN_SMEM(TranscodeSmem, sm);
    const TranscodeFile& f = files[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < (uint32_t)(kNumBlockModels * kHuffLookupSize); i += blockDim.x)
        sm->lookup[i / kHuffLookupSize][i % kHuffLookupSize] = f.models[i / kHuffLookupSize].lookup[i % kHuffLookupSize];
    __syncthreads();
    const unsigned warp = threadIdx.x >> 5, lane = lane_id();
    const LevelStream& ls = f.levels[warp];
    if (!ls.active) return;
    BlockBatch* bb = &sm->batch[warp];
    const uint32_t fmt = f.format;
    const bool is_dxn = fmt == 7 || fmt == 8;
    const bool has_color = fmt <= 6, has_a0 = fmt != 0;
    const uint32_t bs = (fmt == 0 || fmt == 9) ? 8u : 16u;
    const uint32_t W = (ls.blocks_x + 1) & ~1u, H = (ls.blocks_y + 1) & ~1u;
    uint2* rowbuf = f.rowbuf_pool + ls.rowbuf_ofs;

    // lane-0 decoder state
    BitWindow w;
    uint32_t ce = 0, a0 = 0, a1 = 0, group = 0, x = 0, y = 0, face = 0;
    if (lane == 0) {
        bw_init(w, f.bytes, ls.src_ofs, ls.src_ofs + ls.src_size, 0);
        for (uint32_t i = 0; i < W; i++) rowbuf[i] = make_uint2(0u, 0u);
    }
    const uint32_t nce = f.num_color_endpoints, nae = f.num_alpha_endpoints;
    const HuffModelDev* M = f.models;
    for (;;) {
        if (lane == 0) {
            uint32_t n = 0;
            while (face < f.faces && n < 32) {
                const bool visible = y < ls.blocks_y && x < ls.blocks_x;
                if (!(y & 1) && !(x & 1)) group = bw_decode(w, sm->lookup[kDmRef], &M[kDmRef], f.sorted_pool);
                uint2 rb = rowbuf[x];
                uint32_t r;
                if (y & 1) r = rb.x & 0xffffu;
                else { r = group & 3; group >>= 2; rb.x = (rb.x & 0xffff0000u) | (group & 3); group >>= 2; }
                if (!r) {
                    if (has_color) { ce += bw_decode(w, sm->lookup[kDmColorEp], &M[kDmColorEp], f.sorted_pool); if (ce >= nce) ce -= nce; }
                    if (has_a0) { a0 += bw_decode(w, sm->lookup[kDmAlphaEp], &M[kDmAlphaEp], f.sorted_pool); if (a0 >= nae) a0 -= nae; }
                    if (is_dxn) { a1 += bw_decode(w, sm->lookup[kDmAlphaEp], &M[kDmAlphaEp], f.sorted_pool); if (a1 >= nae) a1 -= nae; }
                } else if (r == 2) { ce = rb.x >> 16; a0 = rb.y & 0xffffu; a1 = rb.y >> 16; }
                // r == 0 and r == 1 both leave the running indices in the row buffer (:3986-3990)
                rb.x = (rb.x & 0xffffu) | (ce << 16); rb.y = a0 | (a1 << 16);
                rowbuf[x] = rb;
                uint32_t cs = 0, s0 = 0, s1 = 0;
                if (has_color) cs = bw_decode(w, sm->lookup[kDmColorSel], &M[kDmColorSel], f.sorted_pool);
                if (has_a0) s0 = bw_decode(w, sm->lookup[kDmAlphaSel], &M[kDmAlphaSel], f.sorted_pool);
                if (is_dxn) s1 = bw_decode(w, sm->lookup[kDmAlphaSel], &M[kDmAlphaSel], f.sorted_pool);
                if (visible) {
                    bb->out_ofs[n] = y * ls.row_pitch + x * bs;
                    bb->ce[n] = (uint16_t)ce; bb->cs[n] = (uint16_t)cs; bb->a0[n] = (uint16_t)a0; bb->s0[n] = (uint16_t)s0;
                    bb->a1[n] = (uint16_t)a1; bb->s1[n] = (uint16_t)s1; bb->face[n] = (uint8_t)face;
                    n++;
                }
                if (++x == W) { x = 0; if (++y == H) { y = 0; face++; } }
            }
            bb->n = n; bb->done = face >= f.faces;
        }
        __syncwarp();
        const uint32_t n = bb->n, done = bb->done;
        if (lane < n) {
            uint32_t* o = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(ls.dst[bb->face[lane]]) + bb->out_ofs[lane]);
            uint32_t q0 = 0, q1 = 0, q2 = 0, q3 = 0;
            if (has_a0) {
                const uint16_t* as0 = f.alpha_selectors + 3u * bb->s0[lane];
                q0 = f.alpha_endpoints[bb->a0[lane]] | ((uint32_t)as0[0] << 16);
                q1 = as0[1] | ((uint32_t)as0[2] << 16);
                if (is_dxn) {
                    const uint16_t* as1 = f.alpha_selectors + 3u * bb->s1[lane];
                    q2 = f.alpha_endpoints[bb->a1[lane]] | ((uint32_t)as1[0] << 16);
                    q3 = as1[1] | ((uint32_t)as1[2] << 16);
                } else if (has_color) { q2 = f.color_endpoints[bb->ce[lane]]; q3 = f.color_selectors[bb->cs[lane]]; }
            } else { q0 = f.color_endpoints[bb->ce[lane]]; q1 = f.color_selectors[bb->cs[lane]]; }
            if (bs == 8) { o[0] = q0; o[1] = q1; }
            else { o[0] = q0; o[1] = q1; o[2] = q2; o[3] = q3; }
        }
        __syncwarp();
        if (done) break;
    }
}

}  // namespace crn
