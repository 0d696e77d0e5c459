// crn_writer.h -- the .CRN writer back-end: host C++, no device code.
//
// From dxt_hc's output (four palettes, per-block endpoint / selector indices, reference flags) to the bytes of a
// .crn file.  Follows the reference's crn_comp (crnlib/crn_comp.cpp):
//   palette ordering         optimize_color / optimize_alpha            :990-1058, :1285-1354
//     greedy chains          sort_color_endpoints / sort_alpha_endpoints :767-798, :1075-1104
//     transition-aware       remap_color_endpoints / remap_alpha_endpoints :800-878, :1106-1165
//     trial costing          optimize_*_endpoints_task                   :880-933, :1167-1234
//     selectors              optimize_color_selectors / _alpha_selectors :935-988, :1236-1283
//   palette coding           pack_color_endpoints ... pack_alpha_selectors :43-123, :158-293
//   block coding, 2 passes   pack_blocks + compress_internal             :295-422, :1515-1611
//   models + file assembly   pack_data_models, create_comp_data          :1356-1496
// and crn_symbol_codec.cpp (:365-480 model init, :844-1018 model transmit, :1373-1412 bit output).
// The four endpoint-ordering trials run on host threads like the reference's task pool.  The transition histogram
// is kept sparse (one adjacency list per palette entry) instead of the reference's dense n x n table: the
// arithmetic (32-bit wrapping products included) and every tie rule are the reference's, so the orderings are too.
// Huffman code lengths are optimal and length-limited; the contract is "decodes to the blocks the reference's writer
// would have coded, size within coding slack", and on every vector of tests/test_crn_writer_cpu.py the file is in fact
// byte-identical to the one the reference's crn_compress writes from the same palettes and indices.
#pragma once
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

namespace crnw {

struct Level { uint32_t first_block, num_blocks, block_width; };

struct Input {
    uint32_t crn_format, width, height, num_levels, num_faces, userdata0, userdata1;
    const Level* levels;
    uint32_t num_blocks;
    const uint16_t* endpoint_indices;   // n x 4: color, alpha0, alpha1, reference
    const uint16_t* selector_indices;   // n x 4: color, alpha0, alpha1, -
    const uint32_t* color_endpoints; uint32_t n_color_endpoints;
    const uint32_t* alpha_endpoints; uint32_t n_alpha_endpoints;
    const uint32_t* color_selectors; uint32_t n_color_selectors;
    const uint64_t* alpha_selectors; uint32_t n_alpha_selectors;
    bool has_color, has_alpha0, has_alpha1;
    // Optional: the colour palette orderings computed elsewhere (csrc/writer_kernels.cuh through crn_gpu_compress_crn).  Fills remap[0..3]
    // (trial 0 = greedy chain, 1..3 = the weighted chains with similarity bases `base`) and sel_remap exactly as the host loops below would;
    // returns false to decline (the host loops run).
    struct ColorOrderHook {
        void* user;
        bool (*run)(void* user, const uint32_t* ep_lo, const uint32_t* ep_hi, uint32_t n, const uint32_t* row_start, const uint32_t* col, const uint32_t* cnt,
                    uint32_t selected, const uint32_t base[3], const uint32_t* selectors, uint32_t n_sel, uint16_t* remap4, uint16_t* sel_remap);
    };
    const ColorOrderHook* color_order_hook;
};

// ---- bit output: MSB first, 7 zero bits of padding, whole bytes (crn_symbol_codec.cpp:1373-1412) ----------
struct BitWriter {
    std::vector<uint8_t> bytes;
    uint64_t acc = 0; int nacc = 0; uint64_t total = 0;
    bool simulate = false;
    void put(uint32_t v, uint32_t nbits)
    {
        if (!nbits) return;
        total += nbits;
        if (simulate) return;
        acc = (acc << nbits) | (v & ((nbits >= 32) ? 0xFFFFFFFFu : ((1u << nbits) - 1u)));
        nacc += (int)nbits;
        while (nacc >= 8) { bytes.push_back((uint8_t)(acc >> (nacc - 8))); nacc -= 8; }
    }
    void append(const BitWriter& o)       // o's bits (unfinished: whole bytes + its pending bits) after ours
    {
        if (nacc == 0) { bytes.insert(bytes.end(), o.bytes.begin(), o.bytes.end()); total += 8ull * o.bytes.size(); }
        else {
            const int sft = nacc;
            uint32_t carry = (uint32_t)(acc & ((1u << sft) - 1u));
            const size_t at = bytes.size();
            bytes.resize(at + o.bytes.size());
            uint8_t* dst = bytes.data() + at;
            for (size_t k = 0; k < o.bytes.size(); k++) { const uint32_t v = (carry << 8) | o.bytes[k]; dst[k] = (uint8_t)(v >> sft); carry = v & ((1u << sft) - 1u); }
            acc = carry; total += 8ull * o.bytes.size();
        }
        if (o.nacc) put((uint32_t)(o.acc & ((1u << o.nacc) - 1u)), (uint32_t)o.nacc);
    }
    void finish() { if (!simulate) { uint64_t t = total; put(0, 7); total = t; nacc = 0; acc = 0; } }
};

// ---- static Huffman model -----------------------------------------------------------------------------
struct Model {
    std::vector<uint8_t> len;
    std::vector<uint16_t> code;
    uint32_t size() const { return (uint32_t)len.size(); }
};

// Optimal prefix-code lengths of 16-bit frequencies, limited to max_len by moving Kraft weight from the longest
// codes (the usual package: fold, repair the Kraft sum, hand lengths back in frequency order).
inline void code_lengths(const uint16_t* freq, uint32_t n, uint32_t max_len, uint8_t* out)
{
    struct Sym { uint32_t f; uint32_t s; };
    std::vector<Sym> used;
    for (uint32_t i = 0; i < n; i++) { out[i] = 0; if (freq[i]) used.push_back({freq[i], i}); }
    if (used.empty()) return;
    if (used.size() == 1) { out[used[0].s] = 1; return; }
    std::sort(used.begin(), used.end(), [](const Sym& a, const Sym& b) { return a.f != b.f ? a.f < b.f : a.s < b.s; });
    const uint32_t m = (uint32_t)used.size();
    // two-queue Huffman over the sorted leaves; parent links give the depths
    std::vector<uint64_t> w(2 * m - 1);
    std::vector<uint32_t> parent(2 * m - 1, 0);
    for (uint32_t i = 0; i < m; i++) w[i] = used[i].f;
    uint32_t leaf = 0, node = m, next = m;
    auto take = [&]() -> uint32_t {
        if (leaf < m && (node >= next || w[leaf] <= w[node])) return leaf++;
        return node++;
    };
    while (next < 2 * m - 1) {
        uint32_t a = take(), b = take();
        w[next] = w[a] + w[b]; parent[a] = parent[b] = next; next++;
    }
    std::vector<uint32_t> depth(2 * m - 1, 0);
    for (int i = (int)(2 * m - 3); i >= 0; i--) depth[i] = depth[parent[i]] + 1;
    uint32_t count[64] = {0};
    uint32_t deepest = 0;
    for (uint32_t i = 0; i < m; i++) { uint32_t d = std::min<uint32_t>(depth[i], 63); count[d]++; deepest = std::max(deepest, d); }
    if (deepest > max_len) {
        for (uint32_t d = max_len + 1; d < 64; d++) { count[max_len] += count[d]; count[d] = 0; }
        uint64_t kraft = 0;
        for (uint32_t d = 1; d <= max_len; d++) kraft += (uint64_t)count[d] << (max_len - d);
        while (kraft > (1ull << max_len)) {
            count[max_len]--;
            for (uint32_t d = max_len - 1; d >= 1; d--)
                if (count[d]) { count[d]--; count[d + 1] += 2; break; }
            kraft--;
        }
    }
    // rarest symbols take the longest codes
    uint32_t i = 0;
    for (uint32_t d = std::min<uint32_t>(deepest, max_len); d >= 1; d--)
        for (uint32_t k = 0; k < count[d]; k++) out[used[i++].s] = (uint8_t)d;
}

// static_huffman_data_model::init for 32-bit histograms: frequencies scaled into 16 bits first
// (crn_symbol_codec.cpp:430-480); canonical codes by (length, symbol) (inc/crn_decomp.h:2161-2235).
inline bool build_model(const uint32_t* hist, uint32_t n, uint32_t max_len, Model& m)
{
    uint32_t max_freq = 0;
    for (uint32_t i = 0; i < n; i++) max_freq = std::max(max_freq, hist[i]);
    if (!max_freq) return false;
    std::vector<uint16_t> f16(n, 0);
    for (uint32_t i = 0; i < n; i++) {
        uint32_t f = hist[i];
        if (!f) continue;
        if (max_freq <= 0xFFFFu) f16[i] = (uint16_t)f;
        else {
            uint64_t fl = (((uint64_t)f << 16) - f + (max_freq >> 1)) / max_freq;
            f16[i] = (uint16_t)(fl < 1 ? 1 : fl);
        }
    }
    m.len.assign(n, 0); m.code.assign(n, 0);
    code_lengths(f16.data(), n, max_len, m.len.data());
    uint32_t count[17] = {0}, first[18] = {0};
    for (uint32_t i = 0; i < n; i++) count[m.len[i]]++;
    count[0] = 0;
    uint32_t c = 0;
    for (uint32_t l = 1; l <= 16; l++) { first[l] = c; c = (c + count[l]) << 1; }
    for (uint32_t i = 0; i < n; i++) if (m.len[i]) m.code[i] = (uint16_t)first[m.len[i]]++;
    return true;
}

inline void encode(BitWriter& bw, const Model& m, uint32_t sym) { bw.put(m.code[sym], m.len[sym]); }

// encode_transmit_static_huffman_data_model (crn_symbol_codec.cpp:844-1018): 14-bit used-symbol count, 5-bit number
// of code-length codes, their 3-bit lengths in the fixed order, then the run-length coded lengths.
inline void transmit_model(BitWriter& bw, const Model& m)
{
    static const uint8_t order[21] = {17, 18, 19, 20, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15, 16};
    uint32_t total = 0;
    for (uint32_t i = m.size(); i > 0; i--) if (m.len[i - 1]) { total = i; break; }
    bw.put(total, 14);
    if (!total) return;
    struct Tok { uint8_t sym, extra_bits; uint16_t extra; };
    std::vector<Tok> toks;
    for (uint32_t i = 0; i < total;) {
        const uint8_t v = m.len[i];
        uint32_t j = i;
        while (j < total && m.len[j] == v) j++;
        uint32_t run = j - i;
        if (!v) {
            while (run >= 11) { uint32_t r = std::min(run, 138u); toks.push_back({18, 7, (uint16_t)(r - 11)}); run -= r; }
            if (run >= 3) { toks.push_back({17, 3, (uint16_t)(run - 3)}); run = 0; }
            while (run--) toks.push_back({0, 0, 0});
        } else {
            toks.push_back({v, 0, 0}); run--;
            while (run >= 7) { uint32_t r = std::min(run, 70u); toks.push_back({20, 6, (uint16_t)(r - 7)}); run -= r; }
            if (run >= 3) { toks.push_back({19, 2, (uint16_t)(run - 3)}); run = 0; }
            while (run--) toks.push_back({v, 0, 0});
        }
        i = j;
    }
    uint32_t hist[21] = {0};
    for (const Tok& t : toks) hist[t.sym]++;
    Model cl;
    build_model(hist, 21, 7, cl);
    uint32_t ncl = 21;
    while (ncl > 0 && !cl.len[order[ncl - 1]]) ncl--;
    bw.put(ncl, 5);
    for (uint32_t k = 0; k < ncl; k++) bw.put(cl.len[order[k]], 3);
    for (const Tok& t : toks) {
        encode(bw, cl, t.sym);
        if (t.extra_bits) bw.put(t.extra, t.extra_bits);
    }
}

inline uint64_t transmit_cost(const Model& m)
{
    BitWriter bw; bw.simulate = true;
    transmit_model(bw, m);
    return bw.total;
}

// crc16 of crnlib/crn_checksum.cpp (the header and data checksums crnd_validate_file verifies): CCITT polynomial, one dependent chain of ~8
// operations per byte in the reference's form (13 ms for a 3.7 MB file).  Same function eight bytes at a time: the register only reaches the
// first two bytes of a group, so the state after the group is the XOR of eight table entries, T[k][x] = state after byte x and k zero bytes.
inline uint16_t crc16_byte(uint16_t crc, uint8_t b)
{
    const uint16_t q = (uint16_t)(b ^ (crc >> 8));
    crc = (uint16_t)(crc << 8);
    uint16_t r = (uint16_t)((q >> 4) ^ q);
    crc ^= r; r = (uint16_t)(r << 5); crc ^= r; r = (uint16_t)(r << 7); crc ^= r;
    return crc;
}
struct Crc16Tables {
    uint16_t t[8][256];
    Crc16Tables()
    {
        for (uint32_t x = 0; x < 256; x++) {
            uint16_t c = crc16_byte(0, (uint8_t)x);
            t[0][x] = c;
            for (int k = 1; k < 8; k++) { c = crc16_byte(c, 0); t[k][x] = c; }
        }
    }
};
inline uint16_t crc16(const uint8_t* p, size_t n)
{
    static const Crc16Tables T;
    uint16_t crc = 0xFFFF;
    size_t i = 0;
    for (; i + 8 <= n; i += 8)
        crc = (uint16_t)(T.t[7][(uint8_t)((crc >> 8) ^ p[i])] ^ T.t[6][(uint8_t)((crc & 0xFF) ^ p[i + 1])] ^ T.t[5][p[i + 2]] ^ T.t[4][p[i + 3]] ^
                         T.t[3][p[i + 4]] ^ T.t[2][p[i + 5]] ^ T.t[1][p[i + 6]] ^ T.t[0][p[i + 7]]);
    for (; i < n; i++) crc = crc16_byte(crc, p[i]);
    return (uint16_t)~crc;
}

// ---- palette coding -------------------------------------------------------------------------------------
inline void unpack565(uint32_t c, bool scaled, int out[3])
{
    uint32_t b = c & 31u, g = (c >> 5) & 63u, r = (c >> 11) & 31u;
    if (scaled) { b = (b << 3) | (b >> 2); g = (g << 2) | (g >> 4); r = (r << 3) | (r >> 2); }
    out[0] = (int)r; out[1] = (int)g; out[2] = (int)b;
}

inline bool pack_color_endpoints(const Input& in, const std::vector<uint16_t>& remap, std::vector<uint8_t>& out)
{
    const uint32_t n = in.n_color_endpoints;
    std::vector<uint32_t> ordered(n);
    for (uint32_t i = 0; i < n; i++) ordered[remap[i]] = in.color_endpoints[i];
    static const int limit[6] = {31, 63, 31, 31, 63, 31};
    uint32_t hist5[32] = {0}, hist6[64] = {0};
    std::vector<uint8_t> syms((size_t)n * 6);
    int prev[6] = {0, 0, 0, 0, 0, 0};
    for (uint32_t i = 0; i < n; i++) {
        int cur[6];
        unpack565(ordered[i] & 0xFFFFu, false, cur); unpack565(ordered[i] >> 16, false, cur + 3);
        for (int k = 0; k < 6; k++) {
            const int s = (cur[k] - prev[k]) & limit[k];
            syms[(size_t)i * 6 + k] = (uint8_t)s;
            (k % 3 == 1 ? hist6 : hist5)[s]++;
            prev[k] = cur[k];
        }
    }
    Model m5, m6;
    if (!build_model(hist5, 32, 15, m5) || !build_model(hist6, 64, 15, m6)) return false;
    BitWriter bw;
    transmit_model(bw, m5); transmit_model(bw, m6);
    for (size_t i = 0; i < syms.size(); i++) encode(bw, (i % 3 == 1) ? m6 : m5, syms[i]);
    bw.finish();
    out.swap(bw.bytes);
    return true;
}

inline bool pack_alpha_endpoints(const Input& in, const std::vector<uint16_t>& remap, std::vector<uint8_t>& out)
{
    const uint32_t n = in.n_alpha_endpoints;
    std::vector<uint32_t> ordered(n);
    for (uint32_t i = 0; i < n; i++) ordered[remap[i]] = in.alpha_endpoints[i];
    uint32_t hist[256] = {0};
    std::vector<uint8_t> syms((size_t)n * 2);
    int prev[2] = {0, 0};
    for (uint32_t i = 0; i < n; i++)
        for (int j = 0; j < 2; j++) {
            const int cur = (int)((ordered[i] >> (8 * j)) & 0xFFu);
            const int s = (cur - prev[j]) & 255;
            syms[(size_t)i * 2 + j] = (uint8_t)s; hist[s]++; prev[j] = cur;
        }
    Model m;
    if (!build_model(hist, 256, 15, m)) return false;
    BitWriter bw;
    transmit_model(bw, m);
    for (uint8_t s : syms) encode(bw, m, s);
    bw.finish();
    out.swap(bw.bytes);
    return true;
}

// XOR against the previous entry, 8 symbols of `bits` bits per entry, low bits first (crn_comp.cpp:231-293)
template <typename T>
inline bool pack_selectors(const T* selectors, uint32_t n, uint32_t bits, const std::vector<uint16_t>& remap, std::vector<uint8_t>& out)
{
    std::vector<T> ordered(n);
    for (uint32_t i = 0; i < n; i++) ordered[remap[i]] = selectors[i];
    const uint32_t nsyms = 1u << bits;
    std::vector<uint32_t> hist(nsyms, 0);
    std::vector<uint8_t> syms((size_t)n * 8);
    T prev = 0;
    for (uint32_t i = 0; i < n; i++) {
        T x = prev ^ ordered[i];
        prev = ordered[i];
        for (int c = 0; c < 8; c++, x >>= bits) { const uint8_t s = (uint8_t)(x & (nsyms - 1)); syms[(size_t)i * 8 + c] = s; hist[s]++; }
    }
    Model m;
    if (!build_model(hist.data(), nsyms, 15, m)) return false;
    BitWriter bw;
    transmit_model(bw, m);
    for (uint8_t s : syms) encode(bw, m, s);
    bw.finish();
    out.swap(bw.bytes);
    return true;
}

// ---- orderings ------------------------------------------------------------------------------------------
struct ColorEp { int lo[3], hi[3]; };
inline uint32_t dist3(const int a[3], const int b[3]) { int dr = a[0] - b[0], dg = a[1] - b[1], db = a[2] - b[2]; return (uint32_t)(dr * dr + dg * dg + db * db); }
inline uint32_t ep_dist(const ColorEp& a, const ColorEp& b) { return dist3(a.lo, b.lo) + dist3(a.hi, b.hi); }

// Greedy nearest-neighbour chain from `start`, candidates kept in an array with swap-with-last removal and the first
// minimum winning: the shape of sort_color_endpoints / sort_alpha_endpoints / optimize_*_selectors.
template <typename E, typename D>
inline void greedy_chain(const E* items, uint32_t n, E start, D dist, std::vector<uint16_t>& remap)
{
    remap.resize(n);
    std::vector<E> pool(items, items + n);
    std::vector<uint16_t> idx(n);
    for (uint32_t i = 0; i < n; i++) idx[i] = (uint16_t)i;
    E cur = start;
    for (uint32_t left = n; left;) {
        uint32_t best = 0, best_err = 0xFFFFFFFFu;
        for (uint32_t i = 0; i < left; i++) {
            const uint32_t e = dist(pool[i], cur);
            if (e < best_err) { best_err = e; best = i; }
        }
        cur = pool[best];
        remap[idx[best]] = (uint16_t)(n - left);
        left--;
        pool[best] = pool[left]; idx[best] = idx[left];
    }
}

// Symmetric transition counts between palette entries, as adjacency lists (the reference's dense hist[n*n])
struct Transitions {
    std::vector<uint32_t> row_start;     // n + 1
    std::vector<uint32_t> col, cnt;
    std::vector<uint32_t> sum;
    // range(b0, b1, f) must call f(i, j) once per directed transition of blocks [b0, b1) (a block's transitions depend only on itself and its
    // predecessor), the same ones every time it is called.  Blocks are cut into chunks that run on host threads: per-chunk row counts, a
    // chunk-major prefix inside every row, per-chunk fills, then the rows are merged (duplicates counted, columns sorted) in parallel.
    template <typename Range>
    void build(uint32_t n, uint32_t num_blocks, Range&& range)
    {
        uint32_t nchunk = num_blocks >= (1u << 16) ? 8u : 1u;
        if (const char* e = getenv("CRN_B200_WRITER_CHUNKS")) { const int v = atoi(e); if (v >= 1 && v <= 64 && (uint32_t)v <= num_blocks) nchunk = (uint32_t)v; }   // tests force the chunked path on small inputs
        auto chunk = [&](uint32_t t, uint32_t& b0, uint32_t& b1) { b0 = (uint32_t)((uint64_t)num_blocks * t / nchunk); b1 = (uint32_t)((uint64_t)num_blocks * (t + 1) / nchunk); };
        auto run = [&](auto fn) {
            std::vector<std::thread> th;
            for (uint32_t t = 1; t < nchunk; t++) th.emplace_back(fn, t);
            fn(0u);
            for (std::thread& x : th) x.join();
        };
        std::vector<std::vector<uint32_t>> cnt_t(nchunk, std::vector<uint32_t>(n, 0u));
        run([&](uint32_t t) { uint32_t b0, b1; chunk(t, b0, b1); std::vector<uint32_t>& c = cnt_t[t]; range(b0, b1, [&](uint32_t i, uint32_t) { c[i]++; }); });
        std::vector<uint32_t> start(n + 1, 0);
        for (uint32_t i = 0; i < n; i++) {
            uint32_t run_ofs = start[i];
            for (uint32_t t = 0; t < nchunk; t++) { const uint32_t v = cnt_t[t][i]; cnt_t[t][i] = run_ofs; run_ofs += v; }     // now: the chunk's first slot in row i
            start[i + 1] = run_ofs;
        }
        std::vector<uint32_t> cols(start[n]);
        run([&](uint32_t t) { uint32_t b0, b1; chunk(t, b0, b1); std::vector<uint32_t>& f = cnt_t[t]; range(b0, b1, [&](uint32_t i, uint32_t j) { cols[f[i]++] = j; }); });
        // merge every row: rows dealt to the threads in contiguous runs of equal transition count
        row_start.assign(n + 1, 0); sum.assign(n, 0);
        std::vector<std::vector<uint32_t>> col_t(nchunk), cntv_t(nchunk);
        std::vector<uint32_t> row_lo(nchunk + 1, n);
        row_lo[0] = 0;
        for (uint32_t t = 1, i = 0; t < nchunk; t++) { const uint64_t want = (uint64_t)start[n] * t / nchunk; while (i < n && start[i] < want) i++; row_lo[t] = i; }
        std::vector<uint32_t> row_len(n, 0);
        run([&](uint32_t t) {
            std::vector<uint32_t> seen(n, 0), touched;
            std::vector<uint32_t>& oc = col_t[t]; std::vector<uint32_t>& ov = cntv_t[t];
            for (uint32_t i = row_lo[t]; i < row_lo[t + 1]; i++) {
                touched.clear();
                for (uint32_t k = start[i]; k < start[i + 1]; k++) if (!seen[cols[k]]++) touched.push_back(cols[k]);
                std::sort(touched.begin(), touched.end());
                for (uint32_t j : touched) { oc.push_back(j); ov.push_back(seen[j]); seen[j] = 0; }
                sum[i] = start[i + 1] - start[i];
                row_len[i] = (uint32_t)touched.size();
            }
        });
        for (uint32_t i = 0; i < n; i++) row_start[i + 1] = row_start[i] + row_len[i];
        col.clear(); cnt.clear();
        col.reserve(row_start[n]); cnt.reserve(row_start[n]);
        for (uint32_t t = 0; t < nchunk; t++) { col.insert(col.end(), col_t[t].begin(), col_t[t].end()); cnt.insert(cnt.end(), cntv_t[t].begin(), cntv_t[t].end()); }
    }
    uint16_t busiest() const
    {
        uint16_t sel = 0; uint32_t best = 0;
        for (uint32_t i = 0; i < sum.size(); i++) if (best < sum[i]) { best = sum[i]; sel = (uint16_t)i; }
        return sel;
    }
};

// Σ scale·frequency over the chosen chain, from its front and from its back (crn_comp.cpp:850-855, :1143-1148):
// the entry at distance p from the front weighs L - 2p while that is positive, L = chain length - 1.
inline void chain_pull(const Transitions& T, uint32_t row, const std::vector<int>& pos, int front, int back, uint32_t& pull_front, uint32_t& pull_back)
{
    pull_front = pull_back = 0;
    const int L = back - front;
    for (uint32_t k = T.row_start[row]; k < T.row_start[row + 1]; k++) {
        const int at = pos[T.col[k]];
        if (at < 0) continue;
        const int p = at - front, q = back - at;
        if (L - 2 * p > 0) pull_front += (uint32_t)(L - 2 * p) * T.cnt[k];
        if (L - 2 * q > 0) pull_back += (uint32_t)(L - 2 * q) * T.cnt[k];
    }
}

inline void remap_color_endpoints(const ColorEp* eps, const Transitions& T, uint32_t n, uint16_t selected, float weight, std::vector<uint16_t>& remap)
{
    struct Node { uint32_t index, front_sim, back_sim; ColorEp e; };
    remap.resize(n);
    std::vector<Node> remaining(n);
    for (uint32_t i = 0; i < n; i++) { remaining[i].index = i; remaining[i].front_sim = remaining[i].back_sim = 0; remaining[i].e = eps[i]; }
    std::vector<uint32_t> freq(n, 0);            // Node::frequency, by palette index
    std::vector<int> pos(n, -1);                 // slot in the chosen chain
    std::vector<uint16_t> chosen(2 * (size_t)n + 1);
    uint32_t remaining_count = n;
    int front = (int)n, back = (int)n;
    chosen[front] = selected; pos[selected] = front;
    ColorEp front_e = remaining[selected].e, back_e = front_e;
    bool front_updated = true, back_updated = true;
    remaining[selected] = remaining[--remaining_count];
    uint32_t row = selected;
    const uint32_t base = (uint32_t)(4000 * (1.0f + weight));
    uint32_t normalizer = 0;
    while (remaining_count) {
        for (uint32_t k = T.row_start[row]; k < T.row_start[row + 1]; k++) freq[T.col[k]] += T.cnt[k];
        uint64_t best_value = 0;
        uint32_t best_i = 0;
        for (uint32_t i = 0; i < remaining_count; i++) {
            Node& nd = remaining[i];
            if (front_updated) nd.front_sim = base - std::min<uint32_t>(4000u, ep_dist(nd.e, front_e));
            if (back_updated) nd.back_sim = base - std::min<uint32_t>(4000u, ep_dist(nd.e, back_e));
            const uint64_t value = (uint32_t)(std::max(nd.front_sim, nd.back_sim) * (freq[nd.index] + normalizer) + 1u);   // 32-bit product, as the reference's
            if (value > best_value || (value == best_value && nd.index < selected)) { best_value = value; best_i = i; selected = (uint16_t)nd.index; }
        }
        row = selected;
        uint32_t pull_front, pull_back;
        chain_pull(T, row, pos, front, back, pull_front, pull_back);
        front_updated = back_updated = false;
        Node& bn = remaining[best_i];
        normalizer = freq[bn.index] << 3;
        if ((uint64_t)bn.front_sim * pull_front > (uint64_t)bn.back_sim * pull_back) { chosen[--front] = selected; pos[selected] = front; front_e = bn.e; front_updated = true; }
        else { chosen[++back] = selected; pos[selected] = back; back_e = bn.e; back_updated = true; }
        bn = remaining[--remaining_count];
    }
    for (int i = front; i <= back; i++) remap[chosen[i]] = (uint16_t)(i - front);
}

struct AlphaEp { uint8_t lo, hi; };
inline uint32_t alpha_dist(const AlphaEp& a, const AlphaEp& b) { int d0 = (int)a.lo - b.lo, d1 = (int)a.hi - b.hi; return (uint32_t)(d0 * d0 + d1 * d1); }

inline void remap_alpha_endpoints(const AlphaEp* eps, const Transitions& T, uint32_t n, uint16_t selected, float weight, std::vector<uint16_t>& remap)
{
    remap.resize(n);
    std::vector<uint16_t> remaining;
    std::vector<uint32_t> total_freq(n, 0);
    std::vector<int> pos(n, -1);
    std::vector<uint16_t> chosen(2 * (size_t)n + 1);
    int front = (int)n, back = (int)n;
    chosen[front] = selected; pos[selected] = front;
    for (uint32_t i = 0; i < n; i++) if (i != selected) remaining.push_back((uint16_t)i);
    for (uint32_t k = T.row_start[selected]; k < T.row_start[selected + 1]; k++) if (T.col[k] != selected) total_freq[T.col[k]] = T.cnt[k];
    const uint32_t base = (uint32_t)(1000 * (1.0f + weight));
    uint32_t normalizer = 0;
    while (!remaining.empty()) {
        const AlphaEp& ef = eps[chosen[front]];
        const AlphaEp& eb = eps[chosen[back]];
        uint32_t sel_i = 0;
        uint64_t best_value = 0, sel_sim_front = 0, sel_sim_back = 0;
        for (uint32_t i = 0; i < remaining.size(); i++) {
            const uint32_t r = remaining[i];
            const uint64_t sf = base - std::min<uint32_t>(alpha_dist(eps[r], ef), 1000u);
            const uint64_t sb = base - std::min<uint32_t>(alpha_dist(eps[r], eb), 1000u);
            const uint64_t value = std::max(sf, sb) * (uint32_t)(total_freq[r] + normalizer) + 1;
            if (value > best_value) { best_value = value; sel_i = i; sel_sim_front = sf; sel_sim_back = sb; }
        }
        selected = remaining[sel_i];
        normalizer = total_freq[selected];
        uint32_t pull_front, pull_back;
        chain_pull(T, selected, pos, front, back, pull_front, pull_back);
        if (sel_sim_front * pull_front > sel_sim_back * pull_back) { chosen[--front] = selected; pos[selected] = front; }
        else { chosen[++back] = selected; pos[selected] = back; }
        remaining.erase(remaining.begin() + sel_i);
        for (uint32_t k = T.row_start[selected]; k < T.row_start[selected + 1]; k++) if (pos[T.col[k]] < 0) total_freq[T.col[k]] += T.cnt[k];
    }
    for (int i = front; i <= back; i++) remap[chosen[i]] = (uint16_t)(i - front);
}

// ---- the writer ------------------------------------------------------------------------------------------
struct Writer {
    const Input& in;
    std::vector<uint16_t> ep_remap[2], sel_remap[2];          // [0] colour, [1] alpha
    std::vector<uint8_t> packed_ep[2], packed_sel[2], packed_models;
    std::vector<std::vector<uint8_t>> packed_levels;
    explicit Writer(const Input& i) : in(i) {}

    bool coded(uint32_t b) const { return in.endpoint_indices[(size_t)b * 4 + 3] == 0; }

    // cost of one endpoint ordering: palette bytes + delta symbols under their own model + the model (crn_comp.cpp:897-931, :1185-1233)
    uint64_t trial_bits(int comp, const std::vector<uint16_t>& remap, const std::vector<uint8_t>& packed) const
    {
        const uint32_t n = comp ? in.n_alpha_endpoints : in.n_color_endpoints;
        std::vector<uint32_t> hist(n, 0);
        for (uint32_t l = 0; l < in.num_levels; l++) {
            uint32_t run[2] = {0, 0};
            for (uint32_t b = in.levels[l].first_block, e = b + in.levels[l].num_blocks; b < e; b++) {
                const int first = comp ? 1 : 0, last = comp ? (in.has_alpha1 ? 2 : 1) : 0;
                for (int c = first; c <= last; c++) {
                    const uint32_t index = remap[in.endpoint_indices[(size_t)b * 4 + c]];
                    if (coded(b)) { int sym = (int)index - (int)run[c - first]; hist[sym < 0 ? sym + (int)n : sym]++; }
                    run[c - first] = index;
                }
            }
        }
        Model m;
        uint32_t bits = (uint32_t)(packed.size() << 3);       // 32-bit like the reference's total_bits
        if (build_model(hist.data(), n, 16, m)) {
            for (uint32_t s = 0; s < n; s++) bits += hist[s] * m.len[s];
            bits += (uint32_t)transmit_cost(m);
        }
        return bits;
    }

    bool order_color()
    {
        const uint32_t n = in.n_color_endpoints;
        const auto tb0 = std::chrono::steady_clock::now();
        Transitions T;
        T.build(n, in.num_blocks, [&](uint32_t b0, uint32_t b1, auto&& emit) {
            uint32_t prev = b0 ? in.endpoint_indices[(size_t)(b0 - 1) * 4] : 0u;
            for (uint32_t b = b0; b < b1; b++) {
                const uint32_t i = in.endpoint_indices[(size_t)b * 4];
                if (coded(b) && i != prev) { emit(i, prev); emit(prev, i); }
                prev = i;
            }
        });
        if (getenv("CRN_B200_TRACE")) fprintf(stderr, "[crn_writer] colour transitions: %zu distinct, build %.1f ms\n", T.col.size(), std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb0).count());
        const uint16_t selected = T.busiest();
        std::vector<ColorEp> eps(n);
        for (uint32_t i = 0; i < n; i++) { unpack565(in.color_endpoints[i] & 0xFFFFu, true, eps[i].lo); unpack565(in.color_endpoints[i] >> 16, true, eps[i].hi); }
        static const float weights[4] = {0.0f, 0.0f, 1.0f / 6.0f, 0.5f};
        std::vector<uint16_t> remap[4]; std::vector<uint8_t> packed[4]; uint64_t bits[4]; bool ok[4] = {false, false, false, false};
        bool sel_ok = true;
        bool ordered = false;                                                     // all five orderings already done by the hook
        if (in.color_order_hook && n <= 8192 && in.n_color_selectors <= 8192) {
            std::vector<uint32_t> lo(n), hi(n);
            for (uint32_t i = 0; i < n; i++) {
                lo[i] = (uint32_t)eps[i].lo[0] | ((uint32_t)eps[i].lo[1] << 8) | ((uint32_t)eps[i].lo[2] << 16);
                hi[i] = (uint32_t)eps[i].hi[0] | ((uint32_t)eps[i].hi[1] << 8) | ((uint32_t)eps[i].hi[2] << 16);
            }
            const uint32_t base[3] = { (uint32_t)(4000 * (1.0f + weights[1])), (uint32_t)(4000 * (1.0f + weights[2])), (uint32_t)(4000 * (1.0f + weights[3])) };
            std::vector<uint16_t> all((size_t)4 * n);
            sel_remap[0].resize(in.n_color_selectors);
            const auto th0 = std::chrono::steady_clock::now();
            ordered = in.color_order_hook->run(in.color_order_hook->user, lo.data(), hi.data(), n, T.row_start.data(), T.col.data(), T.cnt.data(), selected, base,
                                               in.color_selectors, in.n_color_selectors, all.data(), sel_remap[0].data());
            if (ordered) for (int t = 0; t < 4; t++) remap[t].assign(all.begin() + (size_t)t * n, all.begin() + (size_t)(t + 1) * n);
            if (getenv("CRN_B200_TRACE")) fprintf(stderr, "[crn_writer] colour orderings on the device: %.1f ms%s\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - th0).count(), ordered ? "" : " (declined)");
        }
        std::vector<std::thread> pool;
        pool.emplace_back([&]() {                                                 // independent of the endpoint order
            sel_ok = ordered ? pack_selectors<uint32_t>(in.color_selectors, in.n_color_selectors, 4, sel_remap[0], packed_sel[0]) : order_color_selectors();
        });
        for (int t = 0; t < 4; t++)
            pool.emplace_back([&, t]() {
                const auto tt0 = std::chrono::steady_clock::now();
                if (ordered) { }
                else if (t) remap_color_endpoints(eps.data(), T, n, selected, weights[t], remap[t]);
                else {
                    ColorEp zero; memset(&zero, 0, sizeof(zero));
                    greedy_chain(eps.data(), n, zero, [](const ColorEp& a, const ColorEp& b) { return ep_dist(a, b); }, remap[0]);
                }
                const auto tt1 = std::chrono::steady_clock::now();
                ok[t] = pack_color_endpoints(in, remap[t], packed[t]);
                if (ok[t]) bits[t] = trial_bits(0, remap[t], packed[t]);
                if (getenv("CRN_B200_TRACE")) fprintf(stderr, "[crn_writer] colour trial %d: order %.1f ms, cost %.1f ms\n", t, std::chrono::duration<double, std::milli>(tt1 - tt0).count(), std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tt1).count());
            });
        for (auto& th : pool) th.join();
        uint64_t best = 0xFFFFFFFFull;
        for (int t = 0; t < 4; t++) {
            if (!ok[t]) return false;
            if (bits[t] < best) { best = bits[t]; ep_remap[0].swap(remap[t]); packed_ep[0].swap(packed[t]); }
        }
        return sel_ok;
    }

    bool order_color_selectors()
    {   // per-pixel selector distance d[] = {0, 5, 14, 10} on the XOR of 2-bit selectors (crn_comp.cpp:941-953)
        static const uint8_t d[4] = {0, 5, 14, 10};
        std::vector<uint8_t> D8(0x10000);
        uint8_t D4[256];
        for (uint32_t i = 0; i < 256; i++) D4[i] = (uint8_t)(d[(i ^ (i >> 4)) & 3] + d[((i >> 2) ^ (i >> 6)) & 3]);
        for (uint32_t i = 0; i < 0x10000; i++) D8[i] = (uint8_t)(D4[((i >> 8) & 0xF0) | ((i >> 4) & 0xF)] + D4[((i >> 4) & 0xF0) | (i & 0xF)]);
        const uint8_t* T8 = D8.data();
        greedy_chain(in.color_selectors, in.n_color_selectors, (uint32_t)0,
                     [T8](uint32_t s, uint32_t ref) {
                         return (uint32_t)T8[((s >> 16) & 0xFF00) | ((ref >> 24) & 0xFF)] + T8[((s >> 8) & 0xFF00) | ((ref >> 16) & 0xFF)] +
                                T8[(s & 0xFF00) | ((ref >> 8) & 0xFF)] + T8[((s << 8) & 0xFF00) | (ref & 0xFF)];
                     },
                     sel_remap[0]);
        return pack_selectors<uint32_t>(in.color_selectors, in.n_color_selectors, 4, sel_remap[0], packed_sel[0]);
    }

    bool order_alpha_selectors()
    {   // d[] = {0, 2, 3, 3, 5, 5, 4, 4} on the XOR of 3-bit selectors, two pixels per table look-up (crn_comp.cpp:1242-1249)
        static const uint8_t d[8] = {0, 2, 3, 3, 5, 5, 4, 4};
        std::vector<uint8_t> D6(0x1000);
        for (uint32_t i = 0; i < 0x1000; i++) D6[i] = (uint8_t)(d[(i ^ (i >> 6)) & 7] + d[((i >> 3) ^ (i >> 9)) & 7]);
        const uint8_t* T6 = D6.data();
        greedy_chain(in.alpha_selectors, in.n_alpha_selectors, (uint64_t)0,
                     [T6](uint64_t s, uint64_t ref) {
                         uint32_t e = 0;
                         uint64_t a = s << 6;
                         for (int j = 0; j < 8; j++, a >>= 6, ref >>= 6) e += T6[(a & 0xFC0) | (ref & 0x3F)];
                         return e;
                     },
                     sel_remap[1]);
        return pack_selectors<uint64_t>(in.alpha_selectors, in.n_alpha_selectors, 6, sel_remap[1], packed_sel[1]);
    }

    bool order_alpha()
    {
        const uint32_t n = in.n_alpha_endpoints;
        Transitions T;
        T.build(n, in.num_blocks, [&](uint32_t b0, uint32_t b1, auto&& emit) {
            uint32_t prev[2] = {0, 0};
            if (b0) { prev[0] = in.endpoint_indices[(size_t)(b0 - 1) * 4 + 1]; prev[1] = in.endpoint_indices[(size_t)(b0 - 1) * 4 + 2]; }
            for (uint32_t b = b0; b < b1; b++) {
                const uint32_t i0 = in.endpoint_indices[(size_t)b * 4 + 1], i1 = in.endpoint_indices[(size_t)b * 4 + 2];
                if (coded(b)) {
                    if (in.has_alpha0 && i0 != prev[0]) { emit(i0, prev[0]); emit(prev[0], i0); }
                    if (in.has_alpha1 && i1 != prev[1]) { emit(i1, prev[1]); emit(prev[1], i1); }
                }
                prev[0] = i0; prev[1] = i1;
            }
        });
        const uint16_t selected = T.busiest();
        std::vector<AlphaEp> eps(n);
        for (uint32_t i = 0; i < n; i++) { eps[i].lo = (uint8_t)(in.alpha_endpoints[i] & 0xFF); eps[i].hi = (uint8_t)((in.alpha_endpoints[i] >> 8) & 0xFF); }
        static const float weights[4] = {0.0f, 0.0f, 1.0f / 6.0f, 0.5f};
        std::vector<uint16_t> remap[4]; std::vector<uint8_t> packed[4]; uint64_t bits[4]; bool ok[4] = {false, false, false, false};
        bool sel_ok = true;
        std::vector<std::thread> pool;
        pool.emplace_back([&]() { sel_ok = order_alpha_selectors(); });
        for (int t = 0; t < 4; t++)
            pool.emplace_back([&, t]() {
                if (t) remap_alpha_endpoints(eps.data(), T, n, selected, weights[t], remap[t]);
                else {
                    AlphaEp zero = {0, 0};
                    greedy_chain(eps.data(), n, zero, [](const AlphaEp& a, const AlphaEp& b) { return alpha_dist(a, b); }, remap[0]);
                }
                ok[t] = pack_alpha_endpoints(in, remap[t], packed[t]);
                if (ok[t]) bits[t] = trial_bits(1, remap[t], packed[t]);
            });
        for (auto& th : pool) th.join();
        uint64_t best = 0xFFFFFFFFull;
        for (int t = 0; t < 4; t++) {
            if (!ok[t]) return false;
            if (bits[t] < best) { best = bits[t]; ep_remap[1].swap(remap[t]); packed_ep[1].swap(packed[t]); }
        }
        return sel_ok;
    }

    // One walk over a level in stream order (crn_comp.cpp:349-420): with models == nullptr it fills the histograms.
    struct Stats { std::vector<uint32_t> ref, ep[2], sel[2]; };
    struct Models { Model ref, ep[2], sel[2]; };
    // Rows [by0, by1) of a level (by0 even).  The running endpoint index a block is coded against is simply the previous
    // block's index (crn_comp.cpp:392-408 updates it on every block), so row ranges are independent of each other.
    void walk_rows(uint32_t l, uint32_t by0, uint32_t by1, Stats* st, const Models* md, BitWriter* bw) const
    {
        const Level& lv = in.levels[l];
        const bool has[3] = {in.has_color, in.has_alpha0, in.has_alpha1};
        uint32_t run[3] = {0, 0, 0};
        const uint32_t W = lv.block_width;
        const uint16_t* E = in.endpoint_indices;
        uint32_t b = lv.first_block + by0 * W;
        if (by0)
            for (int c = 0; c < 3; c++) if (has[c]) run[c] = ep_remap[c ? 1 : 0][E[(size_t)(b - 1) * 4 + c]];
        for (uint32_t by = by0, e = lv.first_block + by1 * W; b < e; by++)
            for (uint32_t bx = 0; bx < W; bx++, b++) {
                if (!(by & 1) && !(bx & 1)) {
                    const uint32_t g = (E[(size_t)b * 4 + 3] & 3u) | ((E[(size_t)(b + W) * 4 + 3] & 3u) << 2) | ((E[(size_t)(b + 1) * 4 + 3] & 3u) << 4) |
                                       ((E[(size_t)(b + W + 1) * 4 + 3] & 3u) << 6);
                    if (bw) encode(*bw, md->ref, g); else st->ref[g]++;
                }
                for (int c = 0; c < 3; c++) {
                    if (!has[c]) continue;
                    const int k = c ? 1 : 0;
                    const uint32_t n = k ? in.n_alpha_endpoints : in.n_color_endpoints;
                    const uint32_t index = ep_remap[k][E[(size_t)b * 4 + c]];
                    if (coded(b)) {
                        int sym = (int)index - (int)run[c];
                        if (sym < 0) sym += (int)n;
                        if (bw) encode(*bw, md->ep[k], (uint32_t)sym); else st->ep[k][sym]++;
                    }
                    run[c] = index;
                }
                for (int c = 0; c < 3; c++) {
                    if (!has[c]) continue;
                    const int k = c ? 1 : 0;
                    const uint32_t index = sel_remap[k][in.selector_indices[(size_t)b * 4 + c]];
                    if (bw) encode(*bw, md->sel[k], index); else st->sel[k][index]++;
                }
            }
    }

    static void put_be(uint8_t* p, uint64_t v, int n) { for (int i = 0; i < n; i++) p[i] = (uint8_t)(v >> (8 * (n - 1 - i))); }

    bool write(std::vector<uint8_t>& file)
    {
        const bool has_alpha = in.has_alpha0 || in.has_alpha1;
        for (uint32_t l = 0; l < in.num_levels; l++) {
            const Level& lv = in.levels[l];
            if (!lv.block_width || (lv.block_width & 1) || lv.num_blocks % (2 * lv.block_width) || (uint64_t)lv.first_block + lv.num_blocks > in.num_blocks) return false;
        }
        const bool trace = getenv("CRN_B200_TRACE") != nullptr;
        auto now = []() { return std::chrono::steady_clock::now(); };
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        const auto t0 = now();
        // colour and alpha orderings are independent: run them side by side
        bool ok_c = true, ok_a = true;
        std::thread ta;
        if (has_alpha) ta = std::thread([&]() { ok_a = order_alpha(); });
        if (in.has_color) ok_c = order_color();
        if (has_alpha) ta.join();
        if (!ok_c || !ok_a) return false;

        const auto t1 = now();
        Stats st;
        st.ref.assign(256, 0);
        if (in.has_color) { st.ep[0].assign(in.n_color_endpoints, 0); st.sel[0].assign(in.n_color_selectors, 0); }
        if (has_alpha) { st.ep[1].assign(in.n_alpha_endpoints, 0); st.sel[1].assign(in.n_alpha_selectors, 0); }
        // both passes over the blocks run as row-range tasks on host threads (per-thread histograms, per-task bit buffers)
        struct Task { uint32_t level, by0, by1; };
        std::vector<Task> tasks;
        for (uint32_t l = 0; l < in.num_levels; l++) {
            const uint32_t W = in.levels[l].block_width, rows = in.levels[l].num_blocks / W;
            const uint32_t step = std::max(2u, (65536u / W) & ~1u);
            for (uint32_t y = 0; y < rows; y += step) tasks.push_back({l, y, std::min(rows, y + step)});
        }
        const uint32_t nthreads = std::max(1u, std::min<uint32_t>({(uint32_t)tasks.size(), std::max(1u, std::thread::hardware_concurrency()), 16u}));
        auto run_tasks = [&](auto&& body) {
            std::atomic<uint32_t> next{0};
            std::vector<std::thread> pool;
            for (uint32_t t = 0; t < nthreads; t++)
                pool.emplace_back([&, t]() { for (uint32_t i; (i = next.fetch_add(1)) < tasks.size();) body(t, i); });
            for (auto& th : pool) th.join();
        };
        {
            std::vector<Stats> part(nthreads, st);
            run_tasks([&](uint32_t t, uint32_t i) { walk_rows(tasks[i].level, tasks[i].by0, tasks[i].by1, &part[t], nullptr, nullptr); });
            auto merge = [](std::vector<uint32_t>& a, const std::vector<uint32_t>& b) { for (size_t k = 0; k < a.size(); k++) a[k] += b[k]; };
            for (const Stats& ps : part) { merge(st.ref, ps.ref); for (int k = 0; k < 2; k++) { merge(st.ep[k], ps.ep[k]); merge(st.sel[k], ps.sel[k]); } }
        }
        Models md;
        if (!build_model(st.ref.data(), 256, 16, md.ref)) return false;
        for (int k = 0; k < 2; k++) {
            if (!st.ep[k].empty()) build_model(st.ep[k].data(), (uint32_t)st.ep[k].size(), 16, md.ep[k]);      // all-reference streams leave the model empty
            if (!st.sel[k].empty() && !build_model(st.sel[k].data(), (uint32_t)st.sel[k].size(), 16, md.sel[k])) return false;
        }
        for (int k = 0; k < 2; k++) if (!st.ep[k].empty() && md.ep[k].len.empty()) { md.ep[k].len.assign(st.ep[k].size(), 0); md.ep[k].code.assign(st.ep[k].size(), 0); }
        const auto t2 = now();
        packed_levels.resize(in.num_levels);
        {
            std::vector<BitWriter> piece(tasks.size());
            run_tasks([&](uint32_t, uint32_t i) {
                piece[i].bytes.reserve((size_t)(tasks[i].by1 - tasks[i].by0) * in.levels[tasks[i].level].block_width);
                walk_rows(tasks[i].level, tasks[i].by0, tasks[i].by1, nullptr, &md, &piece[i]);
            });
            size_t i = 0;
            for (uint32_t l = 0; l < in.num_levels; l++) {
                BitWriter bw;
                bw.bytes.swap(piece[i].bytes); bw.acc = piece[i].acc; bw.nacc = piece[i].nacc; bw.total = piece[i].total;
                for (i++; i < tasks.size() && tasks[i].level == l; i++) bw.append(piece[i]);
                bw.finish();
                packed_levels[l].swap(bw.bytes);
            }
        }
        {
            BitWriter bw;
            transmit_model(bw, md.ref);
            for (int k = 0; k < 2; k++) {
                if (md.ep[k].size()) transmit_model(bw, md.ep[k]);
                if (md.sel[k].size()) transmit_model(bw, md.sel[k]);
            }
            bw.finish();
            packed_models.swap(bw.bytes);
        }
        const auto t3 = now();
        if (trace) fprintf(stderr, "[crn_writer] orderings %.1f ms, histogram pass %.1f ms, coding pass %.1f ms\n", ms(t0, t1), ms(t1, t2), ms(t2, t3));
        // file assembly (crn_comp.cpp:1414-1496; header of inc/crn_defs.h:286-341, big-endian fields)
        const uint32_t header_size = 70 + 4 * in.num_levels;
        file.assign(header_size, 0);
        struct Pal { uint32_t ofs, size, num; } pal[4] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        auto append = [&](const std::vector<uint8_t>& v) { uint32_t o = (uint32_t)file.size(); file.insert(file.end(), v.begin(), v.end()); return o; };
        if (!packed_ep[0].empty()) { pal[0] = {0, (uint32_t)packed_ep[0].size(), in.n_color_endpoints}; pal[0].ofs = append(packed_ep[0]); }
        if (!packed_sel[0].empty()) { pal[1] = {0, (uint32_t)packed_sel[0].size(), in.n_color_selectors}; pal[1].ofs = append(packed_sel[0]); }
        if (!packed_ep[1].empty()) { pal[2] = {0, (uint32_t)packed_ep[1].size(), in.n_alpha_endpoints}; pal[2].ofs = append(packed_ep[1]); }
        if (!packed_sel[1].empty()) { pal[3] = {0, (uint32_t)packed_sel[1].size(), in.n_alpha_selectors}; pal[3].ofs = append(packed_sel[1]); }
        const uint32_t tables_ofs = append(packed_models);
        if (packed_models.size() > 0xFFFFu) return false;
        std::vector<uint32_t> level_ofs(in.num_levels);
        for (uint32_t l = 0; l < in.num_levels; l++) level_ofs[l] = append(packed_levels[l]);
        uint8_t* h = file.data();
        put_be(h + 0, ('H' << 8) | 'x', 2);
        put_be(h + 2, header_size, 2);
        put_be(h + 6, file.size(), 4);
        put_be(h + 10, crc16(h + header_size, file.size() - header_size), 2);
        put_be(h + 12, in.width, 2); put_be(h + 14, in.height, 2);
        h[16] = (uint8_t)in.num_levels; h[17] = (uint8_t)in.num_faces; h[18] = (uint8_t)in.crn_format;
        put_be(h + 19, 0, 2);                       // m_flags
        put_be(h + 21, 0, 4);                       // m_reserved
        put_be(h + 25, in.userdata0, 4); put_be(h + 29, in.userdata1, 4);
        for (int k = 0; k < 4; k++) { put_be(h + 33 + 8 * k, pal[k].ofs, 3); put_be(h + 36 + 8 * k, pal[k].size, 3); put_be(h + 39 + 8 * k, pal[k].num, 2); }
        put_be(h + 65, packed_models.size(), 2); put_be(h + 67, tables_ofs, 3);
        for (uint32_t l = 0; l < in.num_levels; l++) put_be(h + 70 + 4 * l, level_ofs[l], 4);
        put_be(h + 4, crc16(h + 6, header_size - 6), 2);
        return true;
    }
};

}  // namespace crnw
