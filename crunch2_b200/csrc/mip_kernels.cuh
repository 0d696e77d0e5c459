// mip_kernels.cuh -- separable image resampling / mip-chain generation for sm_100a.
//
// Replaces image_utils::resample in its task-pool form (reference crnlib/crn_image_utils.cpp:668-875 ->
// threaded_resampler::resample_x_task / resample_y_task, crnlib/crn_threaded_resampler.cpp:64-270), which is what
// mipmapped_texture::generate_mipmaps (crnlib/crn_mipmapped_texture.cpp:2140-2220) runs once per mip level, always from
// level 0.  Contributor lists (Resampler::make_clist, crnlib/crn_resampler.cpp:119-420) and the two gamma tables are
// built on the host with the reference's arithmetic (mip_host.h); the kernels apply them with the reference's float
// operation order: horizontal pass first, per output sample a running sum over its contributors in list order; vertical
// pass = first contributor line scaled, the others multiplied and added in order (a single contributor is copied
// unscaled, :201-204); clamp to [0, 1]; 8-bit conversion.  Built with -fmad=false, so every sum rounds like the
// reference's SSE code: results are bit-exact.
#pragma once
#include "launch.h"

namespace crn {

struct MipTables {
    const float* to_linear;          // 256 floats: pow(i / 255, gamma) (sRGB) -- nullptr: v * (1 / 255)
    const uint8_t* to_srgb;          // 8192 bytes (sRGB) -- nullptr: (int)(255 v + .5)
};

// tmp[src_y][dst_x] = sum_k src[src_y][pixel_k] * weight_k   (float4 per sample; unused components stay 0)
__global__ void __launch_bounds__(256)
mip_resample_x_kernel(const uint8_t* __restrict__ src, uint32_t src_pitch, uint32_t src_h, uint32_t dst_w, int num_comps, int srgb,
                      const uint32_t* __restrict__ c_off, const uint32_t* __restrict__ c_pix, const float* __restrict__ c_wgt,
                      const float* __restrict__ to_linear, float4* __restrict__ tmp)
{
    __shared__ float lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = srgb ? to_linear[i] : (float)i * (1.0f / 255.0f);
    __syncthreads();
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dst_w || y >= src_h) return;
    const uint32_t* row = reinterpret_cast<const uint32_t*>(src + (size_t)y * src_pitch);
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    for (uint32_t k = c_off[x], e = c_off[x + 1]; k < e; k++) {
        const uint32_t p = row[c_pix[k]];
        const float w = c_wgt[k];
        s0 += lut[p & 0xff] * w;
        s1 += lut[(p >> 8) & 0xff] * w;
        s2 += lut[(p >> 16) & 0xff] * w;
        if (num_comps > 3) s3 += ((float)(p >> 24) * (1.0f / 255.0f)) * w;          // alpha is always linear (:797-800)
    }
    tmp[(size_t)y * dst_w + x] = make_float4(s0, s1, s2, s3);
}

__device__ __forceinline__ float mip_clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
__device__ __forceinline__ unsigned mip_to_u8(float v, int srgb, const uint8_t* __restrict__ to_srgb)
{
    if (!srgb) { int c = (int)(255.0f * v + .5f); return (unsigned)(c < 0 ? 0 : (c > 255 ? 255 : c)); }
    int j = (int)(8192 * v + .5f);
    j = j < 0 ? 0 : (j >= 8192 ? 8191 : j);
    return to_srgb[j];
}

__global__ void __launch_bounds__(256)
mip_resample_y_kernel(const float4* __restrict__ tmp, uint32_t dst_w, uint32_t dst_h, int num_comps, int srgb,
                      const uint32_t* __restrict__ c_off, const uint32_t* __restrict__ c_pix, const float* __restrict__ c_wgt,
                      const uint8_t* __restrict__ to_srgb, uint8_t* __restrict__ dst, uint32_t dst_pitch)
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dst_w || y >= dst_h) return;
    const uint32_t k0 = c_off[y], k1 = c_off[y + 1];
    float4 s;
    if (k1 - k0 == 1) s = tmp[(size_t)c_pix[k0] * dst_w + x];
    else {
        s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t k = k0; k < k1; k++) {
            const float4 p = tmp[(size_t)c_pix[k] * dst_w + x];
            const float w = c_wgt[k];
            if (k == k0) { s.x = p.x * w; s.y = p.y * w; s.z = p.z * w; s.w = p.w * w; }
            else { s.x += p.x * w; s.y += p.y * w; s.z += p.z * w; s.w += p.w * w; }
        }
    }
    unsigned out = mip_to_u8(mip_clamp01(s.x), srgb, to_srgb) | (mip_to_u8(mip_clamp01(s.y), srgb, to_srgb) << 8) | (mip_to_u8(mip_clamp01(s.z), srgb, to_srgb) << 16);
    out |= num_comps > 3 ? (mip_to_u8(mip_clamp01(s.w), 0, to_srgb) << 24) : 0xff000000u;
    *reinterpret_cast<unsigned*>(dst + (size_t)y * dst_pitch + (size_t)x * 4) = out;
}

}  // namespace crn
