// transcode_streams.cuh -- CRN -> DXTn transcoding of LARGE BATCHES, one LANE per level stream.
//
// crnd::crn_unpacker::unpack_level (inc/crn_decomp.h:3944-4223) is a serial dependency chain per level stream; the
// parallelism a batch of files offers is (file x level).  transcode.cuh gives each stream a warp (lane 0 parses, 31 lanes
// wait) and transcode_wide.cuh gives it two CTAs -- right for ONE big file, but a batch of thousands of streams then keeps
// only ~150 of them in flight.  Here every LANE owns a stream and runs the whole serial loop itself, so one warp
// instruction advances 32 streams and an SM holds 512 streams:
//   * code LENGTHS -- all that the bit-position chain depends on -- come from the canonical code's left-justified limits
//     (16 words per model, kept per lane in shared memory, bank = lane): a 4-ary search, two rounds of three independent
//     shared loads, no table in HBM on the chain;
//   * symbol VALUES (palette indices, endpoint deltas) are fetched off that chain from the file's 11-bit lookup / sorted
//     symbol pool (L2); only the reference-group symbol feeds back into parsing (its pool slice is <= 512 bytes per file);
//   * the row buffer (crn_decomp.h "m_block_buffer") lives in the file's row-buffer pool, read one column ahead;
//   * 8 / 16 bytes stored per block per lane; a lane's consecutive stores fill its sectors in L2.
// Streams are sorted by (format, size) on the host so the lanes of a warp finish together.  Bit-exact with crnd_unpack_level.
#pragma once
#include "transcode_wide.cuh"

namespace crn {

constexpr int kStreamThreads = 128;

struct StreamDesc {
    const TranscodeFile* file;
    uint32_t slot;                           // index into file->levels
    uint32_t pad;
};

struct StreamSmem {
    uint32_t limit[5][16][kStreamThreads];   // [slot R E0 E1 S0 S1][length - 1][thread]: left-justified exclusive limits, monotone
    int32_t ref_base[16][kStreamThreads];    // reference model: pool index of a code of length l = ref_base[l - 1] + (k >> (16 - l))
};

// smallest l in 1..16 with k < limit[l] (16 when none: corrupt stream); lim = &limit[m][0][tid]
__device__ __forceinline__ uint32_t stream_code_len(const uint32_t* lim, uint32_t k)
{
    const uint32_t a = lim[3 * kStreamThreads], b = lim[7 * kStreamThreads], c = lim[11 * kStreamThreads];
    const uint32_t q = (k >= a) + (k >= b) + (k >= c);
    const uint32_t* p = lim + 4 * q * kStreamThreads;
    const uint32_t x0 = p[0], x1 = p[kStreamThreads], x2 = p[2 * kStreamThreads];
    return 4 * q + 1 + (k >= x0) + (k >= x1) + (k >= x2);
}

// value of the code of length `len` whose 16-bit window is k (off the parsing chain)
__device__ __forceinline__ uint32_t stream_symbol(const HuffModelDev* __restrict__ m, const uint16_t* __restrict__ pool, uint32_t k, uint32_t len)
{
    if (len <= (uint32_t)kHuffLookupBits) {
        const uint32_t t = m->lookup[k >> (16 - kHuffLookupBits)];
        if (t != kHuffLong) return t & 0xffffu;
    }
    const uint32_t idx = m->first_idx[len] + ((k >> (16 - len)) - m->first_code[len]);
    return idx < m->nsorted ? pool[m->sorted_ofs + idx] : 0u;
}

__device__ __forceinline__ void stream_consume(BitWindow& w, uint32_t len)
{
    w.buf <<= len; w.cnt -= (int)len;
    if (w.cnt <= 32) bw_refill(w);
}

__global__ void __launch_bounds__(kStreamThreads) transcode_streams_kernel(const StreamDesc* __restrict__ streams, uint32_t nstreams)
{
    CRN_DYN_SMEM(StreamSmem, sm);
    const unsigned tid = threadIdx.x;
    const uint32_t sid = blockIdx.x * kStreamThreads + tid;
    if (sid >= nstreams) return;                         // no barriers below: lanes are independent
    const StreamDesc sd = streams[sid];
    const TranscodeFile& f = *sd.file;
    const LevelStream& ls = f.levels[sd.slot];
    const uint32_t fmt = f.format;
    uint32_t ne, ns, e_model[2], s_model[2];
    wide_format_slots(fmt, ne, ns, e_model, s_model);
    const bool two = ne == 2;
    const HuffModelDev* __restrict__ M = f.models;
    {
        const uint32_t ids[5] = { (uint32_t)kDmRef, e_model[0], e_model[1], s_model[0], s_model[1] };
#pragma unroll
        for (int m = 0; m < 5; m++)
            for (int l = 1; l <= 16; l++) sm->limit[m][l - 1][tid] = M[ids[m]].limit[l];
        const HuffModelDev& hr = M[kDmRef];
        for (int l = 1; l <= 16; l++) sm->ref_base[l - 1][tid] = (int32_t)(hr.sorted_ofs + hr.first_idx[l]) - (int32_t)hr.first_code[l];
    }
    const uint32_t* limR = &sm->limit[0][0][tid];
    const uint32_t* limE0 = &sm->limit[1][0][tid]; const uint32_t* limE1 = &sm->limit[2][0][tid];
    const uint32_t* limS0 = &sm->limit[3][0][tid]; const uint32_t* limS1 = &sm->limit[4][0][tid];
    const int32_t* refb = &sm->ref_base[0][tid];
    const HuffModelDev* mE0 = &M[e_model[0]]; const HuffModelDev* mE1 = &M[e_model[1]];
    const HuffModelDev* mS0 = &M[s_model[0]]; const HuffModelDev* mS1 = &M[s_model[1]];
    const uint16_t* __restrict__ pool = f.sorted_pool;
    const int32_t ref_lo = (int32_t)M[kDmRef].sorted_ofs, ref_hi = ref_lo + (int32_t)M[kDmRef].nsorted;
    // component c of the stream and its palette: DXT1 colour | DXT5A alpha | DXN alpha, alpha | DXT5 colour, alpha
    const bool c0_color = fmt != 9 && fmt != 7 && fmt != 8;
    const uint32_t n0 = c0_color ? f.num_color_endpoints : f.num_alpha_endpoints, n1 = f.num_alpha_endpoints;
    const uint32_t* __restrict__ ce_pal = f.color_endpoints; const uint32_t* __restrict__ cs_pal = f.color_selectors;
    const uint16_t* __restrict__ ae_pal = f.alpha_endpoints; const uint16_t* __restrict__ as_pal = f.alpha_selectors;
    const uint32_t bxv = ls.blocks_x, byv = ls.blocks_y, pitch = ls.row_pitch;
    const uint32_t W = (bxv + 1) & ~1u, H = (byv + 1) & ~1u;
    const uint32_t bs = two ? 16u : 8u;
    // row buffer: x = resolved index of component 0 | component 1 << 16; the odd row's references as bytes behind the pool
    uint32_t* rowval = reinterpret_cast<uint32_t*>(f.rowbuf_pool + ls.rowbuf_ofs);
    uint8_t* rowref = reinterpret_cast<uint8_t*>(f.rowbuf_pool + f.rowbuf_total) + ls.rowbuf_ofs;
    for (uint32_t x = 0; x < W; x++) { rowval[2 * x] = 0u; rowref[x] = 0; }

    BitWindow w;
    bw_init(w, f.bytes, ls.src_ofs, ls.src_ofs + ls.src_size, 0);
    uint32_t cur0 = 0, cur1 = 0, group = 0;
    const uint32_t faces = f.faces;
    for (uint32_t face = 0; face < faces; face++) {
        uint8_t* dst_face = reinterpret_cast<uint8_t*>(ls.dst[face]);
        for (uint32_t y = 0; y < H; y++) {
            const bool odd = y & 1, row_visible = y < byv;
            uint8_t* dst_row = dst_face + (size_t)y * pitch;
            uint32_t nrv = rowval[0], nrr = rowref[0];            // one column ahead: the odd rows' reference decides what is parsed next
            for (uint32_t x = 0; x < W; x++) {
                const uint32_t rv = nrv, rr = nrr;
                if (x + 1 < W) { nrv = rowval[2 * (x + 1)]; nrr = rowref[x + 1]; }
                uint32_t r;
                if (odd) r = rr;
                else {
                    if (!(x & 1)) {
                        const uint32_t k = (uint32_t)(w.buf >> 48);
                        const uint32_t len = stream_code_len(limR, k);
                        const int32_t pi = refb[(len - 1) * kStreamThreads] + (int32_t)(k >> (16 - len));
                        group = (pi >= ref_lo && pi < ref_hi) ? pool[pi] : 0u;
                        stream_consume(w, len);
                    }
                    r = group & 3u;
                    rowref[x] = (uint8_t)((group >> 2) & 3u);
                    group >>= 4;
                }
                if (r == 0) {
                    const uint32_t k0 = (uint32_t)(w.buf >> 48);
                    const uint32_t l0 = stream_code_len(limE0, k0);
                    stream_consume(w, l0);
                    uint32_t k1 = 0, l1 = 0;
                    if (two) { k1 = (uint32_t)(w.buf >> 48); l1 = stream_code_len(limE1, k1); stream_consume(w, l1); }
                    cur0 += stream_symbol(mE0, pool, k0, l0);
                    if (cur0 >= n0) cur0 -= n0;
                    if (two) { cur1 += stream_symbol(mE1, pool, k1, l1); if (cur1 >= n1) cur1 -= n1; }
                } else if (r == 2) { cur0 = rv & 0xffffu; cur1 = rv >> 16; }
                rowval[2 * x] = cur0 | (cur1 << 16);
                const uint32_t ks0 = (uint32_t)(w.buf >> 48);
                const uint32_t ls0 = stream_code_len(limS0, ks0);
                stream_consume(w, ls0);
                uint32_t ks1 = 0, ls1 = 0;
                if (two) { ks1 = (uint32_t)(w.buf >> 48); ls1 = stream_code_len(limS1, ks1); stream_consume(w, ls1); }
                if (row_visible && x < bxv) {
                    const uint32_t sel0 = stream_symbol(mS0, pool, ks0, ls0);
                    uint32_t* o = reinterpret_cast<uint32_t*>(dst_row + (size_t)x * bs);
                    if (!two) {
                        if (c0_color) { o[0] = ce_pal[cur0]; o[1] = cs_pal[sel0]; }
                        else {
                            const uint16_t* as0 = as_pal + 3u * sel0;
                            o[0] = ae_pal[cur0] | ((uint32_t)as0[0] << 16);
                            o[1] = as0[1] | ((uint32_t)as0[2] << 16);
                        }
                    } else {
                        const uint32_t sel1 = stream_symbol(mS1, pool, ks1, ls1);
                        uint32_t q0, q1, q2, q3;
                        if (c0_color) {                           // DXT5: stream order colour, alpha; block order alpha, colour
                            const uint16_t* as0 = as_pal + 3u * sel1;
                            q0 = ae_pal[cur1] | ((uint32_t)as0[0] << 16);
                            q1 = as0[1] | ((uint32_t)as0[2] << 16);
                            q2 = ce_pal[cur0]; q3 = cs_pal[sel0];
                        } else {                                  // DXN
                            const uint16_t* as0 = as_pal + 3u * sel0; const uint16_t* as1 = as_pal + 3u * sel1;
                            q0 = ae_pal[cur0] | ((uint32_t)as0[0] << 16);
                            q1 = as0[1] | ((uint32_t)as0[2] << 16);
                            q2 = ae_pal[cur1] | ((uint32_t)as1[0] << 16);
                            q3 = as1[1] | ((uint32_t)as1[2] << 16);
                        }
                        if ((reinterpret_cast<uintptr_t>(o) & 15) == 0) *reinterpret_cast<uint4*>(o) = make_uint4(q0, q1, q2, q3);
                        else { o[0] = q0; o[1] = q1; o[2] = q2; o[3] = q3; }
                    }
                }
            }
        }
    }
}

}  // namespace crn
