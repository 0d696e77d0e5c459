// unpack_kernels.cuh -- DXTn blocks -> RGBA8 pixels for sm_100a.
//
// Replaces crnlib::dxt_image::unpack / get_block_pixels (reference crnlib/crn_dxt_image.cpp:495-567, :1094-1190), the
// decode half of crn_decompress_dds_to_images (crnlib/crnlib.cpp:293-333).  Bandwidth bound: 8 or 16 bytes in and 64 bytes
// out per block.  One thread per block: one 8/16-byte load, four 16-byte row stores; consecutive threads take consecutive
// blocks of a block row, so a warp writes 512 contiguous bytes per pixel row.
// Reference semantics kept: the colour palette is 3-colour whenever color0 <= color1, for every format (crn_dxt.cpp:291-301);
// its alpha only reaches the pixel for DXT1 / DXT1A; channels a format does not carry are 0, alpha 255 (the reference's
// scratch block starts as (0, 0, 0, 255), crn_dxt_image.cpp:503-507).
#pragma once
#include "dxt5a_opt.cuh"

namespace crn {

// The kernel is issue-bound (86 % issue-active, profiles/r1z_ncu_unpack_mip_summaries.txt), so the per-pixel work is kept to a
// handful of instructions: selectors are peeled off 32-bit words, the alpha value is picked by one PRMT and merged into the
// pixel by a second one.
__device__ __forceinline__ void unpack_color_element(unsigned long long e, bool keep_alpha, unsigned (&px)[16])
{
    const unsigned c0 = (unsigned)(e & 0xffff), c1 = (unsigned)((e >> 16) & 0xffff);
    unsigned r0 = (c0 >> 11) & 31, g0 = (c0 >> 5) & 63, b0 = c0 & 31, r1 = (c1 >> 11) & 31, g1 = (c1 >> 5) & 63, b1 = c1 & 31;
    r0 = (r0 << 3) | (r0 >> 2); g0 = (g0 << 2) | (g0 >> 4); b0 = (b0 << 3) | (b0 >> 2);
    r1 = (r1 << 3) | (r1 >> 2); g1 = (g1 << 2) | (g1 >> 4); b1 = (b1 << 3) | (b1 >> 2);
    unsigned p0, p1, p2, p3;
    p0 = r0 | (g0 << 8) | (b0 << 16) | 0xff000000u;
    p1 = r1 | (g1 << 8) | (b1 << 16) | 0xff000000u;
    if (c0 > c1) {
        p2 = ((r0 * 2 + r1) / 3) | (((g0 * 2 + g1) / 3) << 8) | (((b0 * 2 + b1) / 3) << 16) | 0xff000000u;
        p3 = ((r1 * 2 + r0) / 3) | (((g1 * 2 + g0) / 3) << 8) | (((b1 * 2 + b0) / 3) << 16) | 0xff000000u;
    } else {
        p2 = ((r0 + r1) >> 1) | (((g0 + g1) >> 1) << 8) | (((b0 + b1) >> 1) << 16) | 0xff000000u;
        p3 = 0;
    }
    const unsigned sel = (unsigned)(e >> 32);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const unsigned lo = (sel >> (2 * i)) & 1u, hi = (sel >> (2 * i + 1)) & 1u;
        const unsigned a = lo ? p1 : p0, b = lo ? p3 : p2;
        const unsigned v = hi ? b : a;
        px[i] = keep_alpha ? v : __byte_perm(v, px[i], 0x7210);          // colour bytes from v, alpha byte kept
    }
}

__device__ __forceinline__ void unpack_alpha_element(unsigned long long e, unsigned comp, unsigned (&px)[16])
{
    const unsigned l = (unsigned)(e & 0xff), h = (unsigned)((e >> 8) & 0xff);
    unsigned v[8];
    if (l > h) dxt5a_values8(l, h, v); else dxt5a_values6(l, h, v);
    // the eight values as bytes of two registers: one PRMT per pixel picks value[selector], a second one drops it into channel `comp`
    const unsigned lo4 = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24), hi4 = v[4] | (v[5] << 8) | (v[6] << 16) | (v[7] << 24);
    const unsigned s_lo = (unsigned)(e >> 16) & 0xffffffu, s_hi = (unsigned)(e >> 40);       // pixels 0-7 / 8-15, 3 bits each
    const unsigned merge = comp == 0 ? 0x3214u : (comp == 1 ? 0x3240u : (comp == 2 ? 0x3410u : 0x4210u));
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const unsigned s = ((i < 8 ? s_lo : s_hi) >> (3 * (i & 7))) & 7u;
        const unsigned a = __byte_perm(lo4, hi4, s);                                          // byte 0 = value[s]
        px[i] = __byte_perm(px[i], a, merge);
    }
}

__device__ __forceinline__ void unpack_dxt3_alpha(unsigned long long e, unsigned (&px)[16])
{   // dxt3_block::get_alpha(x, y, scaled = true), crn_dxt.cpp:350-366
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const unsigned a = (unsigned)(e >> (4 * i)) & 15u;
        px[i] = (px[i] & 0x00ffffffu) | (((a << 4) | a) << 24);
    }
}

// format: crn_gpu_format.  blocks: blocks_x * blocks_y elements of 8 / 16 bytes, row-major.
__global__ void __launch_bounds__(256)
unpack_blocks_kernel(const unsigned long long* __restrict__ blocks, uint32_t format, uint32_t width, uint32_t height, uint32_t blocks_x, uint32_t total_blocks,
                     uint8_t* __restrict__ rgba, uint32_t pitch)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= total_blocks) return;
    const uint32_t bx = b % blocks_x, by = b / blocks_x;
    unsigned px[16];
#pragma unroll
    for (int i = 0; i < 16; i++) px[i] = 0xff000000u;
    if (format == 0 || format == 1) unpack_color_element(blocks[b], true, px);
    else if (format == 4) unpack_alpha_element(blocks[b], 3, px);
    else {
        const ulonglong2 e = reinterpret_cast<const ulonglong2*>(blocks)[b];
        if (format == 2) { unpack_dxt3_alpha(e.x, px); unpack_color_element(e.y, false, px); }
        else if (format == 3) { unpack_alpha_element(e.x, 3, px); unpack_color_element(e.y, false, px); }
        else { unpack_alpha_element(e.x, format == 5 ? 0 : 1, px); unpack_alpha_element(e.y, format == 5 ? 1 : 0, px); }
    }
    const uint32_t x0 = bx * 4, y0 = by * 4;
    if (x0 + 4 <= width && y0 + 4 <= height && !(pitch & 15u) && !((size_t)rgba & 15u)) {
#pragma unroll
        for (int y = 0; y < 4; y++)
            *reinterpret_cast<uint4*>(rgba + (size_t)(y0 + y) * pitch + (size_t)x0 * 4) = make_uint4(px[4 * y], px[4 * y + 1], px[4 * y + 2], px[4 * y + 3]);
    } else {
        for (uint32_t y = 0; y < 4 && y0 + y < height; y++)
            for (uint32_t x = 0; x < 4 && x0 + x < width; x++)
                *reinterpret_cast<unsigned*>(rgba + (size_t)(y0 + y) * pitch + (size_t)(x0 + x) * 4) = px[4 * y + x];
    }
}

}  // namespace crn
