// dds_kernels.cuh -- the pixel-side kernels of the .dds container edge for sm_100a.
//
//   pixel_convert_kernel       image_utils::convert_image (reference crnlib/crn_image_utils.cpp:1181-1380) for the conversions the
//                              DXT paths use: "cooking" RGBA into the swizzled DXT5 layouts before compression (crn_dds_comp.cpp:90-104,
//                              crn_comp.cpp:440-452) and "uncooking" after decode (mip_level::uncook_image, crn_mipmapped_texture.cpp:351-356;
//                              DXN gets its z back).  Elementwise, in place, one thread per pixel.
//   dxt1_has_alpha_kernel      dxt_image::has_alpha for DXT1 (crnlib/crn_dxt_image.cpp:586-611): any 3-colour block using selector 3,
//                              which is what turns a DXT1 .dds into DXT1A on load (read_dds_internal, crn_mipmapped_texture.cpp:773-775).
//   dds_raw_pixels_kernel      the uncompressed branch of read_dds_internal (:787-832): little-endian pixels of 1-4 bytes, channel bit
//                              masks -> RGBA8 with the reference's rounding (bits * 255 + (mask >> 1)) / mask.
// All bit-exact (integer arithmetic; regen_z's float expression is evaluated with the reference's operand order, -fmad=false).
#pragma once
#include "launch.h"

namespace crn {

enum PixelConversion { kConvToCCxY = 1, kConvFromCCxY = 2, kConvToxGxR = 3, kConvFromxGxR = 4, kConvToxGBR = 5, kConvFromxGBR = 6,
                       kConvToAGBR = 7, kConvFromAGBR = 8, kConvXYtoXYZ = 9, kConvYtoA = 10, kConvRenormNormalMap = 11 };

__device__ __forceinline__ uint32_t clamp_u8(int v) { return v < 0 ? 0u : (v > 255 ? 255u : (uint32_t)v); }

// regen_z (crn_image_utils.cpp:1160-1180)
__device__ __forceinline__ uint32_t regen_z(uint32_t x, uint32_t y)
{
    float vx = ((float)x - 128.0f) * 1.0f / 127.0f, vy = ((float)y - 128.0f) * 1.0f / 127.0f;
    vx = vx < -1.0f ? -1.0f : (vx > 1.0f ? 1.0f : vx);
    vy = vy < -1.0f ? -1.0f : (vy > 1.0f ? 1.0f : vy);
    float t = 1.0f - vx * vx - vy * vy;
    t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
    float vz = sqrtf(t);
    vz = vz * 127.0f + 128.0f;
    if (vz < 128.0f) vz -= .5f; else vz += .5f;
    return clamp_u8((int)vz);
}

// image_utils::renorm_normal_map (crn_image_utils.cpp:369-428), one texel
__device__ __forceinline__ uint32_t renorm_pixel(uint32_t p)
{
    uint32_t c[3] = { p & 255u, (p >> 8) & 255u, (p >> 16) & 255u };
    const uint32_t a = p >> 24;
    if (c[0] == 128 && c[1] == 128 && c[2] == 128) return p;
    float v[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float t = (float)c[i];
        t *= 1.0f / 255.0f; t *= 2.0f; t -= 1.0f;
        v[i] = t < -1.0f ? -1.0f : (t > 1.0f ? 1.0f : t);
    }
    float n = v[0] * v[0]; n += v[1] * v[1]; n += v[2] * v[2];          // vec::norm, in component order
    const float length = sqrtf(n);
    if (length < .077f) { c[0] = c[1] = c[2] = 128; }
    else if (fabsf(length - 1.0f) > .077f) {
        if (length != 0.0f) { v[0] /= length; v[1] /= length; v[2] /= length; }
#pragma unroll
        for (int i = 0; i < 3; i++) {
            float t = floorf((v[i] + 1.0f) * .5f * 255.0f + .5f);
            t = t < 0.0f ? 0.0f : (t > 255.0f ? 255.0f : t);
            c[i] = (uint32_t)t;
        }
        if (c[0] == 128 && c[1] == 128) c[2] = c[2] < 128 ? 0u : 255u;
    }
    return c[0] | (c[1] << 8) | (c[2] << 16) | (a << 24);
}

__device__ __forceinline__ uint32_t convert_pixel(uint32_t p, uint32_t conv)
{
    if (conv == kConvRenormNormalMap) return renorm_pixel(p);
    const int r = p & 255, g = (p >> 8) & 255, b = (p >> 16) & 255, a = p >> 24;
    uint32_t dr, dg, db, da;
    switch (conv) {
    case kConvToCCxY: {          // color::RGB_to_YCC, biases 123 / 125 (crn_color.h:800-807)
        da = (uint32_t)((r * 19595 + g * 38470 + b * 7471 + 32768) >> 16) & 255u;
        dr = clamp_u8(123 + ((r * -11059 + g * -21709 + b * 32768 + 32768) >> 16));
        dg = clamp_u8(125 + ((r * 32768 + g * -27439 + b * -5329 + 32768) >> 16));
        db = 0;
        break;
    }
    case kConvFromCCxY: {        // color::YCC_to_RGB (crn_color.h:811-820)
        const int y = a, cb = r - 123, cr = g - 125;
        dr = clamp_u8(y + ((91881 * cr + 32768) >> 16));
        dg = clamp_u8(y + ((-46802 * cr + -22554 * cb + 32768) >> 16));
        db = clamp_u8(y + ((116130 * cb + 32768) >> 16));
        da = 255;
        break;
    }
    case kConvToxGxR: dr = 0; dg = g; db = 0; da = r; break;
    case kConvFromxGxR: dr = a; dg = g; db = regen_z(a, g); da = 255; break;
    case kConvToxGBR: dr = 0; dg = g; db = b; da = r; break;
    case kConvFromxGBR: dr = a; dg = g; db = b; da = 255; break;
    case kConvToAGBR: dr = a; dg = g; db = b; da = r; break;
    case kConvFromAGBR: dr = a; dg = g; db = b; da = r; break;
    case kConvXYtoXYZ: dr = r; dg = g; db = regen_z(r, g); da = 255; break;
    case kConvYtoA: dr = r; dg = g; db = b; da = (uint32_t)((r * 19595 + g * 38470 + b * 7471 + 32768) >> 16) & 255u; break;   // image::set_alpha_to_luma (crn_image.h:303-315)
    default: return p;
    }
    return dr | (dg << 8) | (db << 16) | (da << 24);
}

__global__ void __launch_bounds__(256) pixel_convert_kernel(uint8_t* __restrict__ rgba, uint32_t width, uint32_t height, uint32_t pitch, uint32_t conv)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)width * height) return;
    const uint32_t x = (uint32_t)(i % width), y = (uint32_t)(i / width);
    uint32_t* p = reinterpret_cast<uint32_t*>(rgba + (size_t)y * pitch) + x;
    *p = convert_pixel(*p, conv);
}

__global__ void __launch_bounds__(256) dxt1_has_alpha_kernel(const unsigned long long* __restrict__ blocks, uint32_t n, uint32_t* __restrict__ flag)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long e = blocks[i];
    const uint32_t c0 = (uint32_t)(e & 0xffff), c1 = (uint32_t)((e >> 16) & 0xffff), sel = (uint32_t)(e >> 32);
    if (c0 <= c1 && ((sel & (sel >> 1)) & 0x55555555u)) *flag = 1u;       // some 2-bit selector equals 3
}

struct DdsRawFormat { uint32_t bytes_per_pixel, mask_ofs[4], mask_size[4], luminance; };

__global__ void __launch_bounds__(256) dds_raw_pixels_kernel(const uint8_t* __restrict__ src, uint32_t line_pitch, uint32_t width, uint32_t height, DdsRawFormat f,
                                                             uint32_t* __restrict__ dst)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)width * height) return;
    const uint32_t x = (uint32_t)(i % width), y = (uint32_t)(i / width);
    const uint8_t* px = src + (size_t)y * line_pitch + (size_t)x * f.bytes_per_pixel;
    uint32_t c = 0;
    for (uint32_t l = 0; l < f.bytes_per_pixel; l++) c |= (uint32_t)px[l] << (l * 8);
    uint32_t q[4] = { 0, 0, 0, 255 };
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (!f.mask_size[k]) continue;
        const uint32_t mask = f.mask_size[k] >= 32 ? 0xffffffffu : (1u << f.mask_size[k]) - 1u;
        const uint32_t bits = (c >> f.mask_ofs[k]) & mask;
        q[k] = ((bits * 255u + (mask >> 1)) / mask) & 255u;                // set_component: the reference stores a uint8
    }
    if (f.luminance) { q[1] = q[0]; q[2] = q[0]; }
    dst[i] = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
}

}  // namespace crn
