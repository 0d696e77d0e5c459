// transcode_host.h -- host side of the CRN -> DXTn transcoder: header crack, Huffman model receive and
// decoder-table construction (the control-plane part of crnd_unpack_begin, reference
// inc/crn_decomp.h:2657-2670, :3044-3123, :2150-2326, :3662-3692).  Everything that touches block or
// palette data runs on the device (transcode.cuh).
#pragma once
#include "transcode.cuh"
#include <vector>

namespace crn {

struct HostBits {                      // MSB-first reader over a byte range, zero padded past the end
    const uint8_t* p; uint32_t size; uint64_t pos;
    uint32_t peek16() const
    {
        const uint64_t byte = pos >> 3;
        uint32_t v = 0;
        for (int i = 0; i < 3; i++) v = (v << 8) | (byte + i < size ? p[byte + i] : 0u);
        return (v >> (8 - (pos & 7))) & 0xffffu;
    }
    uint32_t get(uint32_t n)
    {
        uint32_t r = 0;
        while (n) { const uint32_t k = n > 16 ? n - 16 : n; r = (r << k) | (peek16() >> (16 - k)); pos += k; n -= k; }
        return r;
    }
};

struct HostModel {
    std::vector<uint8_t> len;
    std::vector<uint16_t> sorted;
    uint32_t first_code[17], first_idx[17], count[17];
    HostModel() { for (int l = 0; l <= 16; l++) count[l] = first_code[l] = first_idx[l] = 0; }
    bool build()
    {   // canonical codes: shorter first, ties by symbol index (crn_decomp.h:2161-2235)
        for (int l = 0; l <= 16; l++) count[l] = first_code[l] = first_idx[l] = 0;
        for (size_t i = 0; i < len.size(); i++) { if (len[i] > 16) return false; if (len[i]) count[len[i]]++; }
        uint32_t code = 0, idx = 0, pos[17];
        for (int l = 1; l <= 16; l++) {
            first_code[l] = code; first_idx[l] = idx; pos[l] = idx;
            code += count[l]; idx += count[l];
            if (code > (1u << l)) return false;          // over-subscribed code: corrupt file
            code <<= 1;
        }
        sorted.assign(idx, 0);
        for (size_t i = 0; i < len.size(); i++) if (len[i]) sorted[pos[len[i]]++] = (uint16_t)i;
        return true;
    }
    uint32_t decode(HostBits& b) const
    {
        const uint32_t k = b.peek16();
        for (int l = 1; l <= 16; l++) {
            if (!count[l]) continue;
            const uint32_t c = k >> (16 - l);
            if (c >= first_code[l] && c - first_code[l] < count[l]) { b.pos += l; return sorted[first_idx[l] + (c - first_code[l])]; }
        }
        return 0;
    }
    // decode_receive_static_data_model (crn_decomp.h:3044-3123)
    bool receive(HostBits& b)
    {
        static const uint8_t order[21] = { 17, 18, 19, 20, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15, 16 };
        len.clear(); sorted.clear();
        const uint32_t total = b.get(14);
        if (!total) return build();
        len.assign(total, 0);
        const uint32_t ncl = b.get(5);
        if (ncl < 1 || ncl > 21) return false;
        HostModel dm;
        dm.len.assign(21, 0);
        for (uint32_t i = 0; i < ncl; i++) dm.len[order[i]] = (uint8_t)b.get(3);
        if (!dm.build()) return false;
        uint32_t ofs = 0;
        while (ofs < total) {
            const uint32_t rem = total - ofs, code = dm.decode(b);
            if (code <= 16) len[ofs++] = (uint8_t)code;
            else if (code == 17) { const uint32_t n = b.get(3) + 3; if (n > rem) return false; ofs += n; }
            else if (code == 18) { const uint32_t n = b.get(7) + 11; if (n > rem) return false; ofs += n; }
            else {
                const uint32_t n = code == 19 ? b.get(2) + 3 : b.get(6) + 7;
                if (!ofs || n > rem || !len[ofs - 1]) return false;
                const uint8_t prev = len[ofs - 1];
                for (const uint32_t e = ofs + n; ofs < e;) len[ofs++] = prev;
            }
        }
        return build();
    }
    // device form: 11-bit first-level lookup + canonical limits for the longer codes
    void to_device(HuffModelDev& d, std::vector<uint16_t>& pool) const
    {
        memset(&d, 0, sizeof(d));
        for (int i = 0; i < kHuffLookupSize; i++) d.lookup[i] = kHuffLong;
        d.sorted_ofs = (uint32_t)pool.size();
        d.nsyms = (uint32_t)len.size();
        d.nsorted = (uint32_t)sorted.size();
        pool.insert(pool.end(), sorted.begin(), sorted.end());
        for (int l = 1; l <= 16; l++) {
            d.first_code[l] = first_code[l]; d.first_idx[l] = first_idx[l];
            d.limit[l] = count[l] ? (first_code[l] + count[l]) << (16 - l) : 0;
            if (l <= kHuffLookupBits)
                for (uint32_t c = 0; c < count[l]; c++) {
                    const uint32_t code = first_code[l] + c, sym = sorted[first_idx[l] + c];
                    const uint32_t base = code << (kHuffLookupBits - l);
                    for (uint32_t j = 0; j < (1u << (kHuffLookupBits - l)); j++) d.lookup[base + j] = sym | ((uint32_t)l << 16);
                }
        }
        // make limit[] monotone so the device search "first l with k < limit[l]" skips unused lengths
        for (int l = 2; l <= 16; l++) if (!count[l]) d.limit[l] = d.limit[l - 1] > 0 ? d.limit[l - 1] : 0;
    }
};

inline uint32_t be_n(const uint8_t* p, int n) { uint32_t v = 0; for (int i = 0; i < n; i++) v = (v << 8) | p[i]; return v; }

struct CrnHeaderInfo {
    uint32_t width, height, levels, faces, format, userdata0, userdata1;
    uint32_t pal_ofs[4], pal_size[4], pal_num[4], tables_ofs, tables_size, level_ofs[16], data_size;
};

// crnd_get_header / crnd_get_texture_info (crn_decomp.h:2657-2670, :2737-2760; layout inc/crn_defs.h:286-341)
inline bool crn_parse_header(const uint8_t* d, uint32_t size, CrnHeaderInfo& h)
{
    if (!d || size < 74) return false;
    if (be_n(d, 2) != (('H' << 8) | 'x')) return false;
    if (be_n(d + 2, 2) < 74 || size < be_n(d + 2, 2) || size < be_n(d + 6, 4)) return false;   // header_size and data_size lie within the buffer
    h.data_size = be_n(d + 6, 4);
    h.width = be_n(d + 12, 2); h.height = be_n(d + 14, 2); h.levels = d[16]; h.faces = d[17]; h.format = d[18];
    h.userdata0 = be_n(d + 25, 4); h.userdata1 = be_n(d + 29, 4);
    for (int i = 0; i < 4; i++) { const uint8_t* q = d + 33 + 8 * i; h.pal_ofs[i] = be_n(q, 3); h.pal_size[i] = be_n(q + 3, 3); h.pal_num[i] = be_n(q + 6, 2); }
    h.tables_size = be_n(d + 65, 2); h.tables_ofs = be_n(d + 67, 3);
    if (h.levels < 1 || h.levels > 16 || (h.faces != 1 && h.faces != 6) || !h.width || !h.height) return false;
    if (be_n(d + 2, 2) < 70 + 4 * h.levels) return false;
    for (uint32_t i = 0; i < h.levels; i++) {
        h.level_ofs[i] = be_n(d + 70 + 4 * i, 4);
        if (h.level_ofs[i] >= size) return false;
    }
    for (int i = 0; i < 4; i++) if ((uint64_t)h.pal_ofs[i] + h.pal_size[i] > size) return false;
    if ((uint64_t)h.tables_ofs + h.tables_size > size) return false;
    return true;
}

}  // namespace crn
