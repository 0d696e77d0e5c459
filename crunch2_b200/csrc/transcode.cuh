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
    uint32_t nsorted;                   // symbols with a code (entries of this model in the pool)
    uint32_t pad;
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
    uint2* rowbuf_pool;                 // fallback row buffers for files too wide for shared memory: rowbuf_total uint2 values, then rowbuf_total reference bytes
    uint32_t rowbuf_total, pad0;
    uint32_t num_color_endpoints, num_color_selectors, num_alpha_endpoints, num_alpha_selectors;
    uint32_t pal_data_ofs[4], pal_data_bit[4], pal_size_end[4];   // first symbol of each palette stream (byte, bit) and segment end
    uint32_t format, faces;
    LevelStream levels[16];
};

// MSB-first bit window: `buf` holds `cnt` valid bits left-justified.  Words are fetched as aligned 32-bit
// loads (byte-swapped) and the NEXT word is always already in flight (`nextw`), so the refill never sits
// on the symbol-to-symbol dependency chain.  Reads past `end` see zeros (crn_decomp.h:3168-3170); the
// file image is padded so the aligned loads themselves never leave the allocation.
struct BitWindow {
    const uint32_t* words;              // 4-byte aligned base of the file image
    uint32_t wi;                        // index of the word held in nextw
    uint32_t end_byte;                  // one past the last valid byte of this stream
    uint32_t nextw;                     // big-endian value of words[wi], masked to the stream end
    unsigned long long buf;
    int cnt;
};
__device__ __forceinline__ uint32_t bw_fetch(const BitWindow& w, uint32_t wi)
{
    const uint32_t first = wi * 4;
    if (first >= w.end_byte) return 0u; // past the end: zeros, and NO load (a truncated / corrupt stream keeps "decoding" here)
    uint32_t v = __byte_perm(w.words[wi], 0, 0x0123);
    if (first + 4 > w.end_byte) v &= 0xFFFFFFFFu << (32 - 8 * (w.end_byte - first));   // zero the bytes past the end
    return v;
}
__device__ __forceinline__ void bw_refill(BitWindow& w)
{   // precondition: cnt <= 32
    w.buf |= (unsigned long long)w.nextw << (32 - w.cnt);
    w.cnt += 32;
    w.wi++;
    w.nextw = bw_fetch(w, w.wi);
}
__device__ __forceinline__ void bw_init(BitWindow& w, const uint8_t* p, uint32_t ofs, uint32_t end, uint32_t bit)
{
    w.words = reinterpret_cast<const uint32_t*>(p);    // the file image is 256-byte aligned
    w.end_byte = end;
    w.wi = ofs >> 2;
    w.nextw = bw_fetch(w, w.wi);
    w.buf = 0; w.cnt = 0;
    bw_refill(w);
    const uint32_t skip = (ofs & 3) * 8 + bit;           // leading bits that belong to earlier data
    w.buf <<= skip; w.cnt -= (int)skip;
    if (w.cnt <= 32) bw_refill(w);
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
        // incomplete codes (corrupt files) can index past the model's symbols: clamp (the reference returns symbol 0, crn_decomp.h:3239)
        const uint32_t idx = m->first_idx[len] + ((k >> (16 - len)) - m->first_code[len]);
        sym = idx < m->nsorted ? pool[m->sorted_ofs + idx] : 0u;
    }
    w.buf <<= len; w.cnt -= (int)len;
    if (w.cnt <= 32) bw_refill(w);
    return sym;
}

// Codes longer than the 11-bit lookup: the five left-justified limits of lengths 12..16 and the matching
// symbol-pool bases, kept in shared memory per block model.  The LENGTH comes from register compares, so
// the bit-parsing chain never waits for the symbol-pool load (only reference symbols feed back into
// parsing, and their alphabet is 256 symbols).
struct LongCodes {
    uint32_t limit[5];                  // lengths 12..16
    int32_t base[5];                    // sorted_ofs + first_idx[len] - first_code[len]
    int32_t pool_lo, pool_hi;           // this model's slice of the pool [lo, hi): indices outside it (corrupt files) read as symbol 0
};
__device__ __forceinline__ void longcodes_fill(LongCodes& lc, const HuffModelDev& hm, int j)
{   // one thread per (model, j = 0..4)
    const int len = 12 + j;
    lc.limit[j] = hm.limit[len];
    lc.base[j] = (int32_t)(hm.sorted_ofs + hm.first_idx[len]) - (int32_t)hm.first_code[len];
    if (j == 0) { lc.pool_lo = (int32_t)hm.sorted_ofs; lc.pool_hi = (int32_t)(hm.sorted_ofs + hm.nsorted); }
}
__device__ __forceinline__ uint32_t bw_decode_fast(BitWindow& w, const uint32_t* lookup, const LongCodes* lc, const uint16_t* pool)
{
    const uint32_t t = lookup[(uint32_t)(w.buf >> (64 - kHuffLookupBits))];
    uint32_t sym, len;
    if (t != kHuffLong) { sym = t & 0xffffu; len = t >> 16; }
    else {
        const uint32_t k = (uint32_t)(w.buf >> 48);
        const uint32_t i = (k >= lc->limit[0]) + (k >= lc->limit[1]) + (k >= lc->limit[2]) + (k >= lc->limit[3]);
        len = 12 + i;
        const int32_t idx = lc->base[i] + (int32_t)(k >> (4 - i));
        sym = (idx >= lc->pool_lo && idx < lc->pool_hi) ? pool[idx] : 0u;
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
// A single thread issues dependent instructions ~4-6 cycles apart, so what limits one level stream is
// the NUMBER OF INSTRUCTIONS on lane 0's path, not memory.  Lane 0 therefore does only what is serial by
// construction -- symbol lengths and the reference symbols that select the next code table -- and drops
// raw symbols into a batch of up to 32 blocks of one block row.  Everything else is done by all lanes:
// the running endpoint indices (":3983-3995": idx = (idx + delta) mod N, "left" keeps it, "top" reloads
// it from the row above) are a segmented prefix sum over the batch, then palette gathers and stores.
struct BlockBatch {                     // per-warp shared memory
    uint16_t dce[32], da0[32], da1[32]; // endpoint index deltas (valid where ref == 0)
    uint16_t cs[32], s0[32], s1[32];    // selector indices
    uint8_t ref[32];                    // endpoint reference 0 new / 1 left / 2 top
    uint32_t n, done, x0, y, face;
};

constexpr uint32_t kRowbufSmemEntries = 6144;      // every level of a <= 8192-wide texture (sum of padded widths < 2 * 2048 + 32)

struct TranscodeSmem {
    uint32_t lookup[kNumBlockModels][kHuffLookupSize];
    LongCodes longc[kNumBlockModels];
    BlockBatch batch[kTranscodeWarps];
    uint32_t row_entries;               // columns the launch reserved behind this struct (<= kRowbufSmemEntries; sized by the widest file of the
                                        // batch, so that a batch of small textures fits more CTAs per SM)
    uint32_t pad[3];
    uint2 rowval[1];                    // row_entries x { ce | a0 << 16, a1 } (resolved indices of the row above), then row_entries reference
                                        // bytes (the odd row's references, delivered by the even row's group symbol)
};
__host__ __device__ inline size_t transcode_smem_bytes(uint32_t row_entries) { return sizeof(TranscodeSmem) + (size_t)row_entries * 9 + 16; }

// segmented running-index scan over the batch for one component
__device__ __forceinline__ uint32_t resolve_indices(uint32_t ref, uint32_t delta, uint32_t top, uint32_t carry, uint32_t n_pal, uint32_t n)
{
    const unsigned lane = lane_id();
    uint32_t s = (ref == 0 && lane < n) ? delta : 0u;
#pragma unroll
    for (int ofs = 1; ofs < 32; ofs <<= 1) {
        const uint32_t v = __shfl_up_sync(CRN_FULL_MASK, s, ofs);
        if ((int)lane >= ofs) s += v;
    }
    const unsigned resets = __ballot_sync(CRN_FULL_MASK, ref == 2 && lane < n);
    const unsigned below = resets & (0xFFFFFFFFu >> (31 - lane));
    const int h = below ? 31 - __clz((int)below) : 0;
    const uint32_t top_h = __shfl_sync(CRN_FULL_MASK, top, h);
    const uint32_t s_h = __shfl_sync(CRN_FULL_MASK, s, h);
    const uint32_t v = below ? top_h + (s - s_h) : carry + s;
    return v % n_pal;
}

template <bool HAS_COLOR, bool HAS_A0, bool IS_DXN>
__device__ __forceinline__ void transcode_level(TranscodeSmem* sm, const TranscodeFile& f, const LevelStream& ls, BlockBatch* bb)
{
    const unsigned lane = lane_id();
    const uint32_t faces = f.faces;
    const uint32_t bs = ((HAS_COLOR && HAS_A0) || IS_DXN) ? 16u : 8u;   // DXT5 / DXN: two elements, DXT1 / DXT5A: one
    const uint32_t W = (ls.blocks_x + 1) & ~1u, H = (ls.blocks_y + 1) & ~1u;
    const uint32_t bxv = ls.blocks_x, byv = ls.blocks_y, pitch = ls.row_pitch;
    const bool in_smem = ls.rowbuf_ofs + W <= sm->row_entries;
    uint2* rowval = in_smem ? &sm->rowval[ls.rowbuf_ofs] : f.rowbuf_pool + ls.rowbuf_ofs;
    uint8_t* rowref = in_smem ? reinterpret_cast<uint8_t*>(&sm->rowval[sm->row_entries]) + ls.rowbuf_ofs : reinterpret_cast<uint8_t*>(f.rowbuf_pool + f.rowbuf_total) + ls.rowbuf_ofs;
    const uint16_t* pool = f.sorted_pool;
    const uint32_t* ce_pal = f.color_endpoints; const uint32_t* cs_pal = f.color_selectors;
    const uint16_t* ae_pal = f.alpha_endpoints; const uint16_t* as_pal = f.alpha_selectors;
    const uint32_t nce = f.num_color_endpoints, nae = f.num_alpha_endpoints;
    for (uint32_t i = lane; i < W; i += 32) { rowval[i] = make_uint2(0u, 0u); rowref[i] = 0; }
    __syncwarp();

    BitWindow w;
    uint32_t group = 0, x = 0, y = 0, face = 0;
    uint32_t ce = 0, a0 = 0, a1 = 0;                     // running indices (warp-uniform carry)
    if (lane == 0) bw_init(w, f.bytes, ls.src_ofs, ls.src_ofs + ls.src_size, 0);
    for (;;) {
        if (lane == 0) {
            const uint32_t n = min(32u, W - x);
            bb->x0 = x; bb->y = y; bb->face = face; bb->n = n;
            const bool odd = y & 1;
            for (uint32_t i = 0; i < n; i++, x++) {
                uint32_t r;
                if (odd) r = rowref[x];
                else {
                    if (!(x & 1)) group = bw_decode_fast(w, sm->lookup[kDmRef], &sm->longc[kDmRef], pool);
                    r = group & 3; rowref[x] = (uint8_t)((group >> 2) & 3); group >>= 4;
                }
                bb->ref[i] = (uint8_t)r;
                if (!r) {
                    if (HAS_COLOR) bb->dce[i] = (uint16_t)bw_decode_fast(w, sm->lookup[kDmColorEp], &sm->longc[kDmColorEp], pool);
                    if (HAS_A0) bb->da0[i] = (uint16_t)bw_decode_fast(w, sm->lookup[kDmAlphaEp], &sm->longc[kDmAlphaEp], pool);
                    if (IS_DXN) bb->da1[i] = (uint16_t)bw_decode_fast(w, sm->lookup[kDmAlphaEp], &sm->longc[kDmAlphaEp], pool);
                }
                if (HAS_COLOR) bb->cs[i] = (uint16_t)bw_decode_fast(w, sm->lookup[kDmColorSel], &sm->longc[kDmColorSel], pool);
                if (HAS_A0) bb->s0[i] = (uint16_t)bw_decode_fast(w, sm->lookup[kDmAlphaSel], &sm->longc[kDmAlphaSel], pool);
                if (IS_DXN) bb->s1[i] = (uint16_t)bw_decode_fast(w, sm->lookup[kDmAlphaSel], &sm->longc[kDmAlphaSel], pool);
            }
            if (x == W) { x = 0; if (++y == H) { y = 0; face++; } }
            bb->done = face >= faces;
        }
        __syncwarp();
        const uint32_t n = bb->n, done = bb->done, bx0 = bb->x0, by0 = bb->y, bf = bb->face;
        {
            const uint32_t col = bx0 + (lane < n ? lane : 0u);
            const uint32_t ref = lane < n ? bb->ref[lane] : 1u;
            const uint2 top = lane < n ? rowval[col] : make_uint2(0u, 0u);      // idle lanes must not read the column lane 0 rewrites below
            uint32_t vce = 0, va0 = 0, va1 = 0;
            if (HAS_COLOR) { vce = resolve_indices(ref, bb->dce[lane], top.x & 0xffffu, ce, nce, n); ce = __shfl_sync(CRN_FULL_MASK, vce, (int)n - 1); }
            if (HAS_A0) { va0 = resolve_indices(ref, bb->da0[lane], top.x >> 16, a0, nae, n); a0 = __shfl_sync(CRN_FULL_MASK, va0, (int)n - 1); }
            if (IS_DXN) { va1 = resolve_indices(ref, bb->da1[lane], top.y, a1, nae, n); a1 = __shfl_sync(CRN_FULL_MASK, va1, (int)n - 1); }
            if (lane < n) {
                rowval[col] = make_uint2(vce | (va0 << 16), va1);
                if (by0 < byv && col < bxv) {
                    uint32_t* o = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(ls.dst[bf]) + (size_t)by0 * pitch + (size_t)col * bs);
                    uint32_t q0 = 0, q1 = 0, q2 = 0, q3 = 0;
                    if (HAS_A0) {
                        const uint16_t* as0 = as_pal + 3u * bb->s0[lane];
                        q0 = ae_pal[va0] | ((uint32_t)as0[0] << 16);
                        q1 = as0[1] | ((uint32_t)as0[2] << 16);
                        if (IS_DXN) {
                            const uint16_t* as1 = as_pal + 3u * bb->s1[lane];
                            q2 = ae_pal[va1] | ((uint32_t)as1[0] << 16);
                            q3 = as1[1] | ((uint32_t)as1[2] << 16);
                        } else if (HAS_COLOR) { q2 = ce_pal[vce]; q3 = cs_pal[bb->cs[lane]]; }
                    } else { q0 = ce_pal[vce]; q1 = cs_pal[bb->cs[lane]]; }
                    if (bs == 8) { o[0] = q0; o[1] = q1; }
                    else { o[0] = q0; o[1] = q1; o[2] = q2; o[3] = q3; }
                }
            }
        }
        __syncwarp();
        if (done) break;
    }
}

__global__ void __launch_bounds__(kTranscodeWarps * 32) transcode_levels_kernel(const TranscodeFile* __restrict__ files, uint32_t row_entries)
{
    CRN_DYN_SMEM(TranscodeSmem, sm);
    const TranscodeFile& f = files[blockIdx.x];
    if (threadIdx.x == 0) sm->row_entries = row_entries;
    for (uint32_t i = threadIdx.x; i < (uint32_t)(kNumBlockModels * kHuffLookupSize); i += blockDim.x)
        sm->lookup[i / kHuffLookupSize][i % kHuffLookupSize] = f.models[i / kHuffLookupSize].lookup[i % kHuffLookupSize];
    if (threadIdx.x < kNumBlockModels * 5) longcodes_fill(sm->longc[threadIdx.x / 5], f.models[threadIdx.x / 5], threadIdx.x % 5);
    __syncthreads();
    const unsigned warp = threadIdx.x >> 5;
    const LevelStream ls = f.levels[warp];              // by value: registers, not repeated global loads
    if (!ls.active) return;
    BlockBatch* bb = &sm->batch[warp];
    const uint32_t fmt = f.format;
    if (fmt == 0) transcode_level<true, false, false>(sm, f, ls, bb);
    else if (fmt == 9) transcode_level<false, true, false>(sm, f, ls, bb);
    else if (fmt == 7 || fmt == 8) transcode_level<false, true, true>(sm, f, ls, bb);
    else transcode_level<true, true, false>(sm, f, ls, bb);
}

}  // namespace crn
