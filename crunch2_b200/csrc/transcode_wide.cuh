// transcode_wide.cuh -- CRN -> DXTn transcoding of LARGE mip levels, second generation.
//
// crnd::crn_unpacker::unpack_level (inc/crn_decomp.h:3944-4223) reads one Huffman bitstream per level; which code
// table applies next depends on reference symbols decoded earlier, so the position of block k+1 is only known after
// block k has been parsed.  transcode.cuh accepts that and parses with one lane (~500 cycles per block).  Here the
// serial part is reduced to ONE shared-memory lookup per TWO blocks, and everything else is spread over the GPU:
//
//   A  transcode_tables_kernel (all SMs): for EVERY bit offset o of the level, tabulate where parsing would be after
//      the next pair of blocks if a pair started at o:
//          Gd[o], Gg[o]   even block rows: a reference-group symbol g (two bits per block: this row's pair and the
//                         pair below it) followed by two blocks whose layout g selects; Gd = bits consumed
//          PP[t][o]       odd block rows: two blocks whose layout t comes from the group symbol of the row above
//      A block is [endpoint deltas, only if its reference is 0][selector indices]; code LENGTHS are all that matter
//      here.  Built in four shared-memory stages (symbol lengths -> one block -> two blocks -> group).
//   B  transcode_walk_kernel (one CTA per level): thread 0 follows  o <- o + Gd[o]  /  o <- o + PP[t][o]  through
//      2048-offset windows that the other warps copy from HBM into shared memory one window ahead, and records the
//      bit offset of every pair.  Unrolled four steps per window check: ~30 cycles per pair.
//   C  transcode_resolve_kernel (one CTA per level): row by row, 1024 threads decode the symbol VALUES of a whole
//      block row in parallel from the recorded offsets, resolve the running endpoint indices (idx = (idx + delta) mod N;
//      "left" keeps, "top" reloads from the row above, :3983-3995) with a CTA-wide segmented scan, gather the palettes
//      and store 8/16 bytes per block.
// Bit-exact with crnd_unpack_level; levels below kWideMinBlocks keep the warp-per-level kernel of transcode.cuh.
#pragma once
#include "transcode.cuh"

namespace crn {

constexpr int kWideT = 2048;                 // bit offsets per tile of kernel A
constexpr int kWideMG = 160;                 // group entries past the tile (the two-group stage looks <= 144 ahead)
constexpr int kWideMP = kWideMG + 32, kWideMB = kWideMP + 64, kWideML = kWideMB + 64;   // margins of the pair / block / length stages
constexpr int kWideTilesPerCta = 16;
constexpr int kWideThreadsA = 512;
constexpr int kWideWB = 4096;                // window of the walker (bit offsets)
constexpr int kWideMW = 1536;                // window margin: four unchecked steps of <= 288 bits, one checked
constexpr int kWideWin = kWideWB + kWideMW;
constexpr int kWideTabBytes = 10;            // table bytes per bit offset: PP[4], Gd, Gt (u8), Gd2, Gt2 (u16)
constexpr int kWideThreadsB = 1024;           // walker CTA: thread 0 walks, warps 1.. copy windows (same launch as the resolver)
constexpr int kWideThreadsC = 1024;
constexpr uint32_t kWideMaxW = 4096;         // padded block columns the resolve kernel's row buffers hold
constexpr uint32_t kWideRowWords = 6144;     // bitstream words of one block row staged in shared memory (longer rows read HBM)

struct WideLevel {
    const TranscodeFile* file;
    uint32_t slot;                           // index into file->levels
    uint32_t nbits;                          // 8 * src_size
    uint32_t W, H;                           // padded block columns / rows (even)
    uint32_t nrows;                          // faces * H
    uint32_t npairs;                         // nrows * W / 2
    uint32_t stride;                         // entries per table array
    uint32_t ntiles;                         // kWideT-offset tiles covering nbits
    uint8_t* tab;                            // PP[0..3], Gd, Gt: six arrays of `stride` bytes, then Gd2, Gt2: two of 2 * stride
    uint32_t* pair_ofs;                      // npairs bit offsets relative to the level's first bit
    uint32_t* progress;                      // block rows the walker has completed (the resolver runs behind it)
    uint32_t first_cta, num_ctas;            // kernel A: CTAs [first_cta, first_cta + num_ctas) build this level
    uint32_t ne, ns;                         // endpoint / selector symbols per block (1 or 2)
    uint32_t e_model[2], s_model[2];         // code table of each symbol slot
};

__host__ __device__ inline void wide_format_slots(uint32_t fmt, uint32_t& ne, uint32_t& ns, uint32_t* e_model, uint32_t* s_model)
{
    if (fmt == 0) { ne = ns = 1; e_model[0] = e_model[1] = kDmColorEp; s_model[0] = s_model[1] = kDmColorSel; }
    else if (fmt == 9) { ne = ns = 1; e_model[0] = e_model[1] = kDmAlphaEp; s_model[0] = s_model[1] = kDmAlphaSel; }
    else if (fmt == 7 || fmt == 8) { ne = ns = 2; e_model[0] = e_model[1] = kDmAlphaEp; s_model[0] = s_model[1] = kDmAlphaSel; }
    else { ne = ns = 2; e_model[0] = kDmColorEp; e_model[1] = kDmAlphaEp; s_model[0] = kDmColorSel; s_model[1] = kDmAlphaSel; }
}

// ---- A: transition tables ---------------------------------------------------------------------------------
struct WideSmemA {
    uint8_t len11[4][kHuffLookupSize];       // code length by 11-bit prefix for E0, E1, S0, S1 (0 = longer than 11)
    uint16_t ref11[kHuffLookupSize];         // reference model: sym | len << 8, 0xffff = longer than 11
    uint32_t limit[5][4];                    // left-justified limits of lengths 12..15 per model slot (E0 E1 S0 S1 R)
    int32_t ref_base[5];                     // symbol-pool bases of the reference model's long codes
    int32_t ref_lo, ref_hi;                  // the reference model's slice of the pool (indices outside: symbol 0)
    uint8_t bytes[(kWideT + kWideML) / 8 + 8];
    uint8_t L[4][kWideT + kWideML];
    uint8_t B[2][kWideT + kWideMB];          // [0] selectors only, [1] endpoints + selectors
    uint8_t PP[4][kWideT + kWideMP];         // [t0 + 2 * t1]
    uint8_t Gd[kWideT + kWideMG], Gt[kWideT + kWideMG];
    uint16_t Gd2[kWideT], Gt2[kWideT];
};

__device__ __forceinline__ uint32_t wide_peek16(const uint8_t* bytes, uint32_t idx, uint32_t bit0)
{   // 16 bits starting at tile-relative bit idx; bytes[0] holds the byte of tile bit -bit0
    const uint32_t o = idx + bit0, b = o >> 3;
    const uint32_t v = ((uint32_t)bytes[b] << 16) | ((uint32_t)bytes[b + 1] << 8) | bytes[b + 2];
    return (v >> (8 - (o & 7))) & 0xffffu;
}

__global__ void __launch_bounds__(kWideThreadsA) transcode_tables_kernel(const WideLevel* __restrict__ levels, uint32_t nlevels)
{
    __shared__ WideSmemA sm;
    uint32_t lo = 0, hi = nlevels;                                   // last level whose first_cta <= blockIdx.x
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (levels[mid].first_cta <= blockIdx.x) lo = mid; else hi = mid; }
    const WideLevel& wl = levels[lo];
    const TranscodeFile& f = *wl.file;
    const LevelStream& ls = f.levels[wl.slot];
    const unsigned tid = threadIdx.x;
    const uint32_t models[5] = { wl.e_model[0], wl.e_model[1], wl.s_model[0], wl.s_model[1], (uint32_t)kDmRef };
    for (uint32_t i = tid; i < (uint32_t)kHuffLookupSize; i += blockDim.x) {
#pragma unroll
        for (int s = 0; s < 4; s++) { const uint32_t t = f.models[models[s]].lookup[i]; sm.len11[s][i] = t == kHuffLong ? (uint8_t)0 : (uint8_t)(t >> 16); }
        const uint32_t t = f.models[kDmRef].lookup[i];
        sm.ref11[i] = t == kHuffLong ? (uint16_t)0xffffu : (uint16_t)((t & 0xffu) | ((t >> 16) << 8));
    }
    if (tid < 20) sm.limit[tid / 4][tid % 4] = f.models[models[tid / 4]].limit[12 + tid % 4];
    if (tid < 5) { const HuffModelDev& hm = f.models[kDmRef]; sm.ref_base[tid] = (int32_t)(hm.sorted_ofs + hm.first_idx[12 + tid]) - (int32_t)hm.first_code[12 + tid]; }
    if (tid == 5) { const HuffModelDev& hm = f.models[kDmRef]; sm.ref_lo = (int32_t)hm.sorted_ofs; sm.ref_hi = (int32_t)(hm.sorted_ofs + hm.nsorted); }
    const uint32_t ne = wl.ne, ns = wl.ns;
    const uint32_t tile0 = (blockIdx.x - wl.first_cta) * kWideTilesPerCta;
    for (uint32_t tile = tile0; tile < tile0 + kWideTilesPerCta && tile < wl.ntiles; tile++) {
        const uint32_t base_bit = tile * kWideT, base_byte = base_bit >> 3;      // kWideT is a multiple of 8
        __syncthreads();
        for (uint32_t i = tid; i < sizeof(sm.bytes); i += blockDim.x) {
            const uint32_t b = base_byte + i;
            sm.bytes[i] = b < ls.src_size ? f.bytes[ls.src_ofs + b] : (uint8_t)0;     // zeros past the end (crn_decomp.h:3168-3170)
        }
        __syncthreads();
        // stage 1: code length of each symbol slot at every offset
        for (uint32_t idx = tid; idx < (uint32_t)(kWideT + kWideML); idx += blockDim.x) {
            const uint32_t k = wide_peek16(sm.bytes, idx, 0);
#pragma unroll
            for (int s = 0; s < 4; s++) {
                uint32_t len = sm.len11[s][k >> 5];
                if (!len) len = 12 + (k >= sm.limit[s][0]) + (k >= sm.limit[s][1]) + (k >= sm.limit[s][2]) + (k >= sm.limit[s][3]);
                sm.L[s][idx] = (uint8_t)len;
            }
        }
        __syncthreads();
        // stage 2: one block
        for (uint32_t idx = tid; idx < (uint32_t)(kWideT + kWideMB); idx += blockDim.x) {
            uint32_t s = sm.L[2][idx];
            if (ns == 2) s += sm.L[3][idx + s];
            sm.B[0][idx] = (uint8_t)s;
            uint32_t e = sm.L[0][idx];
            if (ne == 2) e += sm.L[1][idx + e];
            uint32_t s2 = sm.L[2][idx + e];
            if (ns == 2) s2 += sm.L[3][idx + e + s2];
            sm.B[1][idx] = (uint8_t)(e + s2);
        }
        __syncthreads();
        // stage 3: two blocks
        for (uint32_t idx = tid; idx < (uint32_t)(kWideT + kWideMP); idx += blockDim.x) {
            const uint32_t d0 = sm.B[0][idx], d1 = sm.B[1][idx];
            sm.PP[0][idx] = (uint8_t)(d0 + sm.B[0][idx + d0]);
            sm.PP[1][idx] = (uint8_t)(d1 + sm.B[0][idx + d1]);
            sm.PP[2][idx] = (uint8_t)(d0 + sm.B[1][idx + d0]);
            sm.PP[3][idx] = (uint8_t)(d1 + sm.B[1][idx + d1]);
        }
        __syncthreads();
        // stage 4: reference group symbol + the pair it describes; Gt = layout of the pair below (bit 0 / 1: its first /
        // second block carries endpoint deltas)
        for (uint32_t idx = tid; idx < (uint32_t)(kWideT + kWideMG); idx += blockDim.x) {
            const uint32_t k = wide_peek16(sm.bytes, idx, 0);
            const uint32_t t = sm.ref11[k >> 5];
            uint32_t len, g;
            if (t != 0xffffu) { g = t & 0xffu; len = t >> 8; }
            else {
                const uint32_t i = (k >= sm.limit[4][0]) + (k >= sm.limit[4][1]) + (k >= sm.limit[4][2]) + (k >= sm.limit[4][3]);
                len = 12 + i;
                const int32_t pi = sm.ref_base[i] + (int32_t)(k >> (4 - i));
                g = (pi >= sm.ref_lo && pi < sm.ref_hi) ? (f.sorted_pool[pi] & 0xffu) : 0u;
            }
            const uint32_t t0 = (g & 3u) == 0u, t1 = ((g >> 4) & 3u) == 0u;
            sm.Gd[idx] = (uint8_t)(len + sm.PP[t0 + 2 * t1][idx + len]);
            sm.Gt[idx] = (uint8_t)((((g >> 2) & 3u) == 0u) | ((((g >> 6) & 3u) == 0u) << 1));
        }
        __syncthreads();
        // stage 5: two groups
        for (uint32_t idx = tid; idx < (uint32_t)kWideT; idx += blockDim.x) {
            const uint32_t d = sm.Gd[idx];
            sm.Gd2[idx] = (uint16_t)(d + sm.Gd[idx + d]);
            sm.Gt2[idx] = (uint16_t)(sm.Gt[idx] | ((uint32_t)sm.Gt[idx + d] << 8));
        }
        __syncthreads();
        // write out (word copies; every array offset is a multiple of 4)
        for (int a = 0; a < 6; a++) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(a < 4 ? sm.PP[a] : (a == 4 ? sm.Gd : sm.Gt));
            uint32_t* dst = reinterpret_cast<uint32_t*>(wl.tab + (size_t)a * wl.stride + base_bit);
            for (uint32_t i = tid; i < (uint32_t)kWideT / 4; i += blockDim.x) dst[i] = src[i];
        }
        for (int a = 0; a < 2; a++) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(a ? sm.Gt2 : sm.Gd2);
            uint32_t* dst = reinterpret_cast<uint32_t*>(wl.tab + (size_t)(6 + 2 * a) * wl.stride + (size_t)base_bit * 2);
            for (uint32_t i = tid; i < (uint32_t)kWideT / 2; i += blockDim.x) dst[i] = src[i];
        }
    }
    // the walker's windows read a little past the last tile: the level's last CTA zeroes that tail of every array
    if (tile0 + kWideTilesPerCta >= wl.ntiles) {
        const uint32_t built = wl.ntiles * kWideT, tail = wl.stride - built;          // both multiples of 4
        for (int a = 0; a < 6; a++) {
            uint32_t* dst = reinterpret_cast<uint32_t*>(wl.tab + (size_t)a * wl.stride + built);
            for (uint32_t i = tid; i < tail / 4; i += blockDim.x) dst[i] = 0u;
        }
        for (int a = 0; a < 2; a++) {
            uint32_t* dst = reinterpret_cast<uint32_t*>(wl.tab + (size_t)(6 + 2 * a) * wl.stride + (size_t)built * 2);
            for (uint32_t i = tid; i < tail / 2; i += blockDim.x) dst[i] = 0u;
        }
    }
}

// ---- B: the walk ------------------------------------------------------------------------------------------------
// pair offsets and the progress counter are written by the walker CTA while the resolver CTA of the same launch reads them:
// L2-coherent loads only.
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t wide_ld_cg(const uint32_t* p) { return __ldcg(p); }
__device__ __forceinline__ uint32_t wide_ld_volatile(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void wide_backoff() { __nanosleep(256); }
#else
static inline uint32_t wide_ld_cg(const uint32_t* p) { return *p; }
static inline uint32_t wide_ld_volatile(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
static inline void wide_backoff() {}
#endif

// Shared-memory accessors of the walker.  nvcc: 32-bit shared-window addresses + inline PTX, so that "biased base + bit
// offset" is ONE register add and the table offsets fold into the instruction's immediate; emulator: plain pointers.
#ifdef __CUDACC__
typedef uint32_t wide_addr;
__device__ __forceinline__ wide_addr wide_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t wide_ld8(wide_addr a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t wide_ld16(wide_addr a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void wide_st8(wide_addr a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void wide_st16(wide_addr a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
#else
typedef uintptr_t wide_addr;
static inline wide_addr wide_smem(const void* p) { return reinterpret_cast<uintptr_t>(p); }
static inline uint32_t wide_ld8(wide_addr a) { return *reinterpret_cast<const uint8_t*>(a); }
static inline uint32_t wide_ld16(wide_addr a) { return *reinterpret_cast<const uint16_t*>(a); }
static inline void wide_st8(wide_addr a, uint32_t v) { *reinterpret_cast<uint8_t*>(a) = (uint8_t)v; }
static inline void wide_st16(wide_addr a, uint32_t v) { *reinterpret_cast<uint16_t*>(a) = (uint16_t)v; }
#endif

struct __align__(16) WideSmemB {
    uint8_t tab[2][kWideTabBytes * kWideWin];   // per window: PP[4], Gd, Gt (kWideWin bytes each), Gd2, Gt2 (2 * kWideWin each)
    uint16_t rt[kWideMaxW / 4 + 8];             // layouts of the pairs below, two pairs per entry (+ read-ahead slack)
    uint32_t done[2];                           // end-of-walk flag of epoch e in done[e & 1]: the walker may be one epoch ahead of a thread still reading
};

__device__ __forceinline__ void wide_walk(WideSmemB* smp, const WideLevel& wl, int pipe)
{
    WideSmemB& sm = *smp;
    const unsigned tid = threadIdx.x;
    constexpr uint32_t WIN = kWideWin;
    // Window copies are software-pipelined through registers: the loads of window e + 2 are issued in epoch e and stored to
    // shared memory in epoch e + 1 (when they have long arrived), so the copy never makes the walker wait at the barrier.
    constexpr uint32_t NV = kWideTabBytes * WIN / 16, NCOPY = kWideThreadsB - 32, MAXV = (NV + NCOPY - 1) / NCOPY;
    uint4 regs[MAXV];
    const uint8_t* __restrict__ tabg = wl.tab;
    const uint32_t stride = wl.stride;
    auto load_window = [&](uint32_t e) {
        const size_t base = (size_t)e * kWideWB;
#pragma unroll
        for (uint32_t k = 0; k < MAXV; k++) {
            const uint32_t v = (tid - 32) + k * NCOPY;
            if (v < NV) {
                const uint32_t b = v * 16;
                const uint8_t* src;
                if (b < 6 * WIN) src = tabg + (size_t)(b / WIN) * stride + base + (b % WIN);
                else { const uint32_t b2 = b - 6 * WIN; src = tabg + (size_t)(6 + 2 * (b2 / (2 * WIN))) * stride + 2 * base + (b2 % (2 * WIN)); }
                regs[k] = *reinterpret_cast<const uint4*>(src);
            }
        }
    };
    auto store_window = [&](uint32_t e) {
#pragma unroll
        for (uint32_t k = 0; k < MAXV; k++) {
            const uint32_t v = (tid - 32) + k * NCOPY;
            if (v < NV) *reinterpret_cast<uint4*>(&sm.tab[e & 1][v * 16]) = regs[k];
        }
    };
    if (tid == 0) { sm.done[0] = 0; sm.done[1] = 0; }
    if (tid >= 32) { load_window(0); store_window(0); if (pipe) load_window(1); }
    __syncthreads();
    // The walker is ONE thread: what bounds it is the number of instructions it issues per pair (a lone warp issues one
    // every ~5 cycles), so the loops below keep to: table load(s), one add on the absolute bit offset `o` (tables are
    // addressed through pointers biased by the window base), one store through a running output pointer.
    const uint32_t half_w = wl.W >> 1, nrows = wl.nrows, nbits = wl.nbits;
    uint32_t o = 0, x = 0, row = 0;
    uint32_t* __restrict__ outp = wl.pair_ofs;
    for (uint32_t e = 0;; e++) {
        if (tid == 0 && e * (uint32_t)kWideWB > nbits) {           // corrupt stream: past the end, nothing left to follow
            uint32_t* const last = wl.pair_ofs + wl.npairs;
            while (outp < last) *outp++ = nbits;
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(wl.progress) = nrows;
            sm.done[e & 1] = 1;
        } else if (tid == 0) {
            const uint32_t wb = e * kWideWB, wend = wb + kWideWB;
            const wide_addr tb = wide_smem(sm.tab[e & 1]) - wb;            // tb + o = &PP[0][o - wb]
            const wide_addr tb2 = wide_smem(sm.tab[e & 1]) - 2 * (wide_addr)wb;   // tb2 + 2 * o + 6 * WIN = &Gd2[o - wb]
            const wide_addr rt0 = wide_smem(sm.rt);
            while (o < wend && row < nrows) {
                uint32_t left = half_w - x;                     // pairs left in this block row
                if (!(row & 1)) {
                    // two groups (four blocks) per step; x is even here, so the two layouts are one 16-bit store
                    wide_addr rtp = rt0 + x;
                    while (left >= 8 && o < wend) {
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const wide_addr a2 = tb2 + 2 * (wide_addr)o;
                            const uint32_t d2 = wide_ld16(a2 + 6 * WIN), d1 = wide_ld8(tb + o + 4 * WIN);
                            wide_st16(rtp + 2 * u, wide_ld16(a2 + 8 * WIN));
                            outp[2 * u] = o;
                            outp[2 * u + 1] = o + d1;
                            o += d2;
                        }
                        outp += 8; rtp += 8; left -= 8;
                    }
                    while (left >= 2 && o < wend) {
                        const wide_addr a2 = tb2 + 2 * (wide_addr)o;
                        const uint32_t d2 = wide_ld16(a2 + 6 * WIN), d1 = wide_ld8(tb + o + 4 * WIN);
                        wide_st16(rtp, wide_ld16(a2 + 8 * WIN));
                        outp[0] = o;
                        outp[1] = o + d1;
                        o += d2;
                        outp += 2; rtp += 2; left -= 2;
                    }
                    if (left == 1 && o < wend) {
                        wide_st8(rt0 + half_w - 1, wide_ld8(tb + o + 5 * WIN));
                        *outp++ = o;
                        o += wide_ld8(tb + o + 4 * WIN);
                        left = 0;
                    }
                } else {
                    // layouts are fetched one iteration ahead: nothing but "base + o -> load -> o +=" sits on the chain
                    wide_addr rtp = rt0 + x;
                    uint32_t t0 = wide_ld8(rtp), t1 = wide_ld8(rtp + 1), t2 = wide_ld8(rtp + 2), t3 = wide_ld8(rtp + 3);
                    while (left >= 4 && o < wend) {
                        const wide_addr b0 = tb + t0 * WIN, b1 = tb + t1 * WIN, b2 = tb + t2 * WIN, b3 = tb + t3 * WIN;
                        rtp += 4;
                        t0 = wide_ld8(rtp); t1 = wide_ld8(rtp + 1); t2 = wide_ld8(rtp + 2); t3 = wide_ld8(rtp + 3);
                        outp[0] = o; o += wide_ld8(b0 + o);
                        outp[1] = o; o += wide_ld8(b1 + o);
                        outp[2] = o; o += wide_ld8(b2 + o);
                        outp[3] = o; o += wide_ld8(b3 + o);
                        outp += 4; left -= 4;
                    }
                    while (left && o < wend) {
                        *outp++ = o;
                        o += wide_ld8(tb + wide_ld8(rtp) * WIN + o);
                        rtp++; left--;
                    }
                }
                x = half_w - left;
                if (!left) {
                    x = 0; row++;
                    __threadfence();                            // the row's offsets before the count
                    *reinterpret_cast<volatile uint32_t*>(wl.progress) = row;
                }
            }
            if (row >= nrows) sm.done[e & 1] = 1;
        } else if (tid >= 32) {
            if (pipe) { store_window(e + 1); load_window(e + 2); }
            else { load_window(e + 1); store_window(e + 1); }
        }
        __syncthreads();
        if (sm.done[e & 1]) break;
    }
}

// ---- C: values, running indices, palettes, stores --------------------------------------------------------------
struct WideSmemC {
    uint32_t lookup[kNumBlockModels][kHuffLookupSize];
    LongCodes longc[kNumBlockModels];
    uint16_t delta[3][kWideMaxW];            // endpoint index deltas (valid where ref == 0)
    uint16_t sel[3][kWideMaxW];              // selector indices
    uint16_t rowval[3][kWideMaxW];           // resolved endpoint indices of the row above
    uint8_t ref[kWideMaxW];                  // this row's references
    uint8_t rowref[kWideMaxW];               // the odd row's references, delivered by the even row's group symbols
    uint32_t warp_r[32], warp_v[2][32];
    uint32_t carry[2][2];                    // [row parity][component]: running indices at the start of the row
    uint32_t rowbits[2][kWideRowWords];      // the bitstream words of this block row / the next one
};

// CTA-wide scan of "running index" updates.  Element = (reset, v[NC]): reset ? idx[c] <- v[c] : idx[c] <- idx[c] + v[c]; the
// reference flag of a block is shared by its components, so one scan carries all of them.  Every thread owns K consecutive
// columns, folded into one element first.  Returns the EXCLUSIVE prefix of the thread's first column.  Two barriers; the
// caller's end-of-row barrier separates the reuse of the warp totals.
template <int NC>
__device__ __forceinline__ void wide_scan(WideSmemC* sm, uint32_t r, const uint32_t (&v)[NC], uint32_t& ex_r, uint32_t (&ex_v)[NC])
{
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t ir = r, iv[NC];                                   // inclusive within the warp
#pragma unroll
    for (int c = 0; c < NC; c++) iv[c] = v[c];
#pragma unroll
    for (int ofs = 1; ofs < 32; ofs <<= 1) {
        const uint32_t pr = __shfl_up_sync(CRN_FULL_MASK, ir, ofs);
        uint32_t pv[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) pv[c] = __shfl_up_sync(CRN_FULL_MASK, iv[c], ofs);
        if ((int)lane >= ofs && !ir) {
#pragma unroll
            for (int c = 0; c < NC; c++) iv[c] += pv[c];
            ir = pr;
        }
    }
    if (lane == 31) {
        sm->warp_r[warp] = ir;
#pragma unroll
        for (int c = 0; c < NC; c++) sm->warp_v[c][warp] = iv[c];
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t wr = sm->warp_r[lane], wv[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) wv[c] = sm->warp_v[c][lane];
#pragma unroll
        for (int ofs = 1; ofs < 32; ofs <<= 1) {
            const uint32_t pr = __shfl_up_sync(CRN_FULL_MASK, wr, ofs);
            uint32_t pv[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) pv[c] = __shfl_up_sync(CRN_FULL_MASK, wv[c], ofs);
            if ((int)lane >= ofs && !wr) {
#pragma unroll
                for (int c = 0; c < NC; c++) wv[c] += pv[c];
                wr = pr;
            }
        }
        sm->warp_r[lane] = wr;                                 // inclusive over warps
#pragma unroll
        for (int c = 0; c < NC; c++) sm->warp_v[c][lane] = wv[c];
    }
    __syncthreads();
    // exclusive prefix of this thread = (warps before) then (lanes before)
    uint32_t lr = __shfl_up_sync(CRN_FULL_MASK, ir, 1);
    if (lane == 0) lr = 0;
    const uint32_t br = warp ? sm->warp_r[warp - 1] : 0u;
    ex_r = lr ? 1u : br;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        uint32_t lv = __shfl_up_sync(CRN_FULL_MASK, iv[c], 1);
        if (lane == 0) lv = 0;
        const uint32_t bv = warp ? sm->warp_v[c][warp - 1] : 0u;
        ex_v[c] = lr ? lv : bv + lv;
    }
}

template <bool HAS_COLOR, bool HAS_A0, bool IS_DXN>
__device__ __forceinline__ void wide_resolve_level(WideSmemC* sm, const WideLevel& wl, const TranscodeFile& f, const LevelStream& ls)
{
    constexpr int NC = IS_DXN ? 2 : ((HAS_COLOR ? 1 : 0) + (HAS_A0 ? 1 : 0));   // components in stream order
    constexpr uint32_t bs = ((HAS_COLOR && HAS_A0) || IS_DXN) ? 16u : 8u;
    const unsigned tid = threadIdx.x;
    const uint32_t W = wl.W, H = wl.H, half_w = W >> 1;
    const uint32_t bxv = ls.blocks_x, byv = ls.blocks_y, pitch = ls.row_pitch;
    const uint32_t src_ofs = ls.src_ofs, src_size = ls.src_size, nbits = wl.nbits;
    const uint32_t* pair_ofs = wl.pair_ofs;
    const uint16_t* __restrict__ pool = f.sorted_pool;
    const uint32_t* __restrict__ ce_pal = f.color_endpoints; const uint32_t* __restrict__ cs_pal = f.color_selectors;
    const uint16_t* __restrict__ ae_pal = f.alpha_endpoints; const uint16_t* __restrict__ as_pal = f.alpha_selectors;
    unsigned long long dst_face[6];
#pragma unroll
    for (int i = 0; i < 6; i++) dst_face[i] = ls.dst[i];
    // component c of the stream: its endpoint / selector model and palette size
    const int e_model[2] = { HAS_COLOR ? kDmColorEp : kDmAlphaEp, kDmAlphaEp };
    const int s_model[2] = { HAS_COLOR ? kDmColorSel : kDmAlphaSel, kDmAlphaSel };
    const uint32_t n_pal[2] = { HAS_COLOR ? f.num_color_endpoints : f.num_alpha_endpoints, f.num_alpha_endpoints };
    const uint32_t K = (W + blockDim.x - 1) / blockDim.x;      // columns per thread in the scan (<= 4)
    for (uint32_t i = tid; i < W; i += blockDim.x) { sm->rowref[i] = 0; for (int c = 0; c < 3; c++) sm->rowval[c][i] = 0; }
    if (tid < 4) sm->carry[tid >> 1][tid & 1] = 0;
    // Row pipeline: while row r is decoded from shared memory, the bitstream words and pair offsets of row r + 1 are
    // already in flight into registers (they are stored at the end of the row), and the first offset of row r + 3 is
    // being fetched, so no row waits for HBM.
    constexpr uint32_t KW = kWideRowWords / kWideThreadsC, KP = kWideMaxW / 2 / kWideThreadsC;
    const uint32_t nrows = wl.nrows;
    const uint32_t* __restrict__ words = reinterpret_cast<const uint32_t*>(f.bytes);
    const uint32_t last_word = (src_ofs + src_size + 12) >> 2;            // the file image is padded by 16 bytes
    auto row_start = [&](uint32_t r) -> uint32_t { return r < nrows ? wide_ld_cg(pair_ofs + (size_t)r * half_w) : nbits; };
    // rows [0, need) must be complete before the resolver touches row `need - 4`'s look-ahead
    auto wait_rows = [&](uint32_t need) {
        if (need > nrows) need = nrows;
        if (tid == 0) while (wide_ld_volatile(wl.progress) < need) wide_backoff();
        __syncthreads();
    };
    wait_rows(4);
    auto row_words = [&](uint32_t s_begin, uint32_t s_end, uint32_t& w0, uint32_t& nw) {
        w0 = (src_ofs + (s_begin >> 3)) >> 2;
        nw = ((src_ofs + (s_end >> 3) + 20 + 3) >> 2) - w0;
    };
    uint32_t s0 = row_start(0), s1 = row_start(1), s2 = row_start(2), s3 = 0;
    uint32_t po[KP], npo[KP], pre[KW];
#pragma unroll
    for (uint32_t k = 0; k < KP; k++) { const uint32_t p = tid + k * kWideThreadsC; po[k] = p < half_w ? wide_ld_cg(pair_ofs + p) : 0u; npo[k] = 0; }
    {
        uint32_t w0, nw;
        row_words(s0, s1, w0, nw);
        if (nw <= kWideRowWords)
            for (uint32_t i = tid; i < nw; i += blockDim.x) sm->rowbits[0][i] = w0 + i <= last_word ? words[w0 + i] : 0u;
    }
    __syncthreads();
    for (uint32_t row = 0; row < nrows; row++) {
        const uint32_t y = row % H, face = row / H;
        const bool odd = y & 1;
        wait_rows(row + 5);
        uint32_t w0c, nwc, w0n, nwn;
        row_words(s0, s1, w0c, nwc);
        row_words(s1, s2, w0n, nwn);
        const bool staged = nwc <= kWideRowWords, stage_next = row + 1 < nrows && nwn <= kWideRowWords;
        s3 = row_start(row + 3);
#pragma unroll
        for (uint32_t k = 0; k < KW; k++) {
            const uint32_t i = tid + k * kWideThreadsC;
            pre[k] = (stage_next && i < nwn && w0n + i <= last_word) ? words[w0n + i] : 0u;
        }
        if (row + 1 < nrows) {
#pragma unroll
            for (uint32_t k = 0; k < KP; k++) { const uint32_t p = tid + k * kWideThreadsC; if (p < half_w) npo[k] = wide_ld_cg(pair_ofs + (size_t)(row + 1) * half_w + p); }
        }
        // 1. symbol values of the whole block row, one pair per thread
        const uint8_t* src_base = staged ? reinterpret_cast<const uint8_t*>(sm->rowbits[row & 1]) : f.bytes;
        const uint32_t src_shift = staged ? 4 * w0c : 0u;                       // byte address of src_base[0] within the file image
#pragma unroll
        for (uint32_t kp = 0; kp < KP; kp++) {
            const uint32_t p = tid + kp * kWideThreadsC;
            if (p >= half_w) break;
            const uint32_t o = po[kp];
            BitWindow w;
            bw_init(w, src_base, src_ofs + (o >> 3) - src_shift, src_ofs + src_size - src_shift, o & 7);
            uint32_t r0, r1;
            if (odd) { r0 = sm->rowref[2 * p]; r1 = sm->rowref[2 * p + 1]; }
            else {
                const uint32_t g = bw_decode_fast(w, sm->lookup[kDmRef], &sm->longc[kDmRef], pool);
                r0 = g & 3; r1 = (g >> 4) & 3;
                sm->rowref[2 * p] = (uint8_t)((g >> 2) & 3); sm->rowref[2 * p + 1] = (uint8_t)((g >> 6) & 3);
            }
#pragma unroll
            for (int b = 0; b < 2; b++) {
                const uint32_t x = 2 * p + b, r = b ? r1 : r0;
                sm->ref[x] = (uint8_t)r;
                if (!r) {
#pragma unroll
                    for (int c = 0; c < NC; c++) sm->delta[c][x] = (uint16_t)bw_decode_fast(w, sm->lookup[e_model[c]], &sm->longc[e_model[c]], pool);
                }
#pragma unroll
                for (int c = 0; c < NC; c++) sm->sel[c][x] = (uint16_t)bw_decode_fast(w, sm->lookup[s_model[c]], &sm->longc[s_model[c]], pool);
            }
        }
        __syncthreads();
        // 2. running endpoint indices: fold this thread's K columns, scan, then apply
        const uint32_t x0 = tid * K;
        {
            uint32_t er = 0, ev[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) ev[c] = 0;
            for (uint32_t k = 0; k < K; k++) {
                const uint32_t x = x0 + k;
                if (x >= W) break;
                const uint32_t r = sm->ref[x];
                if (r == 2) {
                    er = 1;
#pragma unroll
                    for (int c = 0; c < NC; c++) ev[c] = sm->rowval[c][x];
                } else if (r == 0) {
#pragma unroll
                    for (int c = 0; c < NC; c++) ev[c] += sm->delta[c][x];
                }
            }
            uint32_t pr, pv[NC], cur[NC];
            wide_scan<NC>(sm, er, ev, pr, pv);
#pragma unroll
            for (int c = 0; c < NC; c++) cur[c] = pr ? pv[c] : sm->carry[row & 1][c] + pv[c];
            for (uint32_t k = 0; k < K; k++) {
                const uint32_t x = x0 + k;
                if (x >= W) break;
                const uint32_t r = sm->ref[x];
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    if (r == 2) cur[c] = sm->rowval[c][x];
                    else if (r == 0) cur[c] += sm->delta[c][x];
                    cur[c] %= n_pal[c];
                    sm->rowval[c][x] = (uint16_t)cur[c];
                }
            }
            if (x0 < W && x0 + K >= W) {                       // the thread that owns the last column hands the indices on
#pragma unroll
                for (int c = 0; c < NC; c++) sm->carry[(row + 1) & 1][c] = cur[c];
            }
        }
        __syncthreads();
        // 3. palettes and stores
        if (y < byv) {
            uint8_t* dst_row = reinterpret_cast<uint8_t*>(dst_face[face]) + (size_t)y * pitch;
            const bool vec_ok = bs == 16 && ((reinterpret_cast<uintptr_t>(dst_row) & 15) == 0);
            for (uint32_t x = tid; x < bxv; x += blockDim.x) {
                uint32_t q0, q1, q2 = 0, q3 = 0;
                if (HAS_A0 || IS_DXN) {
                    const int ca = (HAS_COLOR && !IS_DXN) ? 1 : 0;                 // first alpha component in stream order
                    const uint16_t* as0 = as_pal + 3u * sm->sel[ca][x];
                    q0 = ae_pal[sm->rowval[ca][x]] | ((uint32_t)as0[0] << 16);
                    q1 = as0[1] | ((uint32_t)as0[2] << 16);
                    if (IS_DXN) {
                        const uint16_t* as1 = as_pal + 3u * sm->sel[1][x];
                        q2 = ae_pal[sm->rowval[1][x]] | ((uint32_t)as1[0] << 16);
                        q3 = as1[1] | ((uint32_t)as1[2] << 16);
                    } else if (HAS_COLOR) { q2 = ce_pal[sm->rowval[0][x]]; q3 = cs_pal[sm->sel[0][x]]; }
                } else { q0 = ce_pal[sm->rowval[0][x]]; q1 = cs_pal[sm->sel[0][x]]; }
                uint32_t* o = reinterpret_cast<uint32_t*>(dst_row + (size_t)x * bs);
                if (bs == 8) { o[0] = q0; o[1] = q1; }
                else if (vec_ok) *reinterpret_cast<uint4*>(o) = make_uint4(q0, q1, q2, q3);
                else { o[0] = q0; o[1] = q1; o[2] = q2; o[3] = q3; }
            }
        }
        if (stage_next) {
#pragma unroll
            for (uint32_t k = 0; k < KW; k++) { const uint32_t i = tid + k * kWideThreadsC; if (i < nwn) sm->rowbits[(row + 1) & 1][i] = pre[k]; }
        }
#pragma unroll
        for (uint32_t k = 0; k < KP; k++) po[k] = npo[k];
        s0 = s1; s1 = s2; s2 = s3;
        __syncthreads();
    }
}

// One launch, two CTAs per level: the even CTA walks, the odd CTA resolves the rows the walker has finished (at most
// 32 CTAs, all resident, so the consumer can poll the producer's row counter).
union WideSmemBC { WideSmemB b; WideSmemC c; };

__device__ __forceinline__ void wide_resolve(WideSmemC* sm, const WideLevel& wl)
{
    const TranscodeFile& f = *wl.file;
    for (uint32_t i = threadIdx.x; i < (uint32_t)(kNumBlockModels * kHuffLookupSize); i += blockDim.x)
        sm->lookup[i / kHuffLookupSize][i % kHuffLookupSize] = f.models[i / kHuffLookupSize].lookup[i % kHuffLookupSize];
    if (threadIdx.x < kNumBlockModels * 5) longcodes_fill(sm->longc[threadIdx.x / 5], f.models[threadIdx.x / 5], threadIdx.x % 5);
    __syncthreads();
    const LevelStream& ls = f.levels[wl.slot];
    const uint32_t fmt = f.format;
    if (fmt == 0) wide_resolve_level<true, false, false>(sm, wl, f, ls);
    else if (fmt == 9) wide_resolve_level<false, true, false>(sm, wl, f, ls);
    else if (fmt == 7 || fmt == 8) wide_resolve_level<false, true, true>(sm, wl, f, ls);
    else wide_resolve_level<true, true, false>(sm, wl, f, ls);
}

__global__ void __launch_bounds__(kWideThreadsC) transcode_walk_resolve_kernel(const WideLevel* __restrict__ levels, int pipe)
{
    CRN_DYN_SMEM(WideSmemBC, smu);
    const WideLevel& wl = levels[blockIdx.x >> 1];
    if (!(blockIdx.x & 1)) wide_walk(&smu->b, wl, pipe);
    else wide_resolve(&smu->c, wl);
}

// Batches with more levels than SM pairs: the same two roles as two launches (no co-residency needed; the resolver
// finds every row counter already at its final value).
__global__ void __launch_bounds__(kWideThreadsC) transcode_walk_kernel(const WideLevel* __restrict__ levels, int pipe)
{
    CRN_DYN_SMEM(WideSmemBC, smu);
    wide_walk(&smu->b, levels[blockIdx.x], pipe);
}
__global__ void __launch_bounds__(kWideThreadsC) transcode_resolve_kernel(const WideLevel* __restrict__ levels)
{
    CRN_DYN_SMEM(WideSmemBC, smu);
    wide_resolve(&smu->c, levels[blockIdx.x]);
}

}  // namespace crn
