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

__device__ __forceinline__ float mip_clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
__device__ __forceinline__ unsigned mip_to_u8(float v, int srgb, const uint8_t* __restrict__ to_srgb)
{
    if (!srgb) { int c = (int)(255.0f * v + .5f); return (unsigned)(c < 0 ? 0 : (c > 255 ? 255 : c)); }
    int j = (int)(8192 * v + .5f);
    j = j < 0 ? 0 : (j >= 8192 ? 8191 : j);
    return to_srgb[j];
}

// tmp[src_y][dst_x][c] = sum_k src[src_y][pixel_k][c] * weight_k.  One thread per (row, output sample, component), component
// fastest, flattened over the whole intermediate image so that warps stay full when the level is only a few samples wide.
// Each thread runs its own sum in contributor order (the reference's order).
__global__ void __launch_bounds__(256)
mip_resample_x_kernel(const uint8_t* __restrict__ src, uint32_t src_pitch, uint32_t src_h, uint32_t dst_w, int num_comps, int srgb,
                      const uint32_t* __restrict__ c_off, const uint32_t* __restrict__ c_pix, const float* __restrict__ c_wgt,
                      const float* __restrict__ to_linear, float* __restrict__ tmp)
{
    __shared__ float lut[2][256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        lut[1][i] = (float)i * (1.0f / 255.0f);                       // alpha and non-sRGB channels (:797-800)
        lut[0][i] = srgb ? to_linear[i] : lut[1][i];
    }
    __syncthreads();
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x, total = (size_t)src_h * dst_w * 4;
    if (id >= total) return;
    const unsigned c = (unsigned)(id & 3);
    const uint32_t x = (uint32_t)((id >> 2) % dst_w), y = (uint32_t)((id >> 2) / dst_w);
    float s = 0.0f;
    if ((int)c < num_comps) {
        const uint32_t* row = reinterpret_cast<const uint32_t*>(src + (size_t)y * src_pitch);
        const float* l = lut[c == 3 ? 1 : 0];
        const unsigned sh = 8 * c;
        // the loads of the next taps do not depend on the running sum: unrolled so they are in flight while the adds retire in order
#pragma unroll 8
        for (uint32_t k = c_off[x], e = c_off[x + 1]; k < e; k++) s += l[(row[c_pix[k]] >> sh) & 0xff] * c_wgt[k];
    }
    tmp[id] = s;
}

// out[dst_y][dst_x][c]: first contributor line scaled, the rest multiplied and added in order; a single contributor is copied.
__global__ void __launch_bounds__(256)
mip_resample_y_kernel(const float* __restrict__ tmp, uint32_t dst_w, uint32_t dst_h, int num_comps, int srgb,
                      const uint32_t* __restrict__ c_off, const uint32_t* __restrict__ c_pix, const float* __restrict__ c_wgt,
                      const uint8_t* __restrict__ to_srgb, uint8_t* __restrict__ dst, uint32_t dst_pitch)
{
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x, total = (size_t)dst_h * dst_w * 4;
    if (id >= total) return;
    const unsigned c = (unsigned)(id & 3);
    const uint32_t x = (uint32_t)((id >> 2) % dst_w), y = (uint32_t)((id >> 2) / dst_w);
    unsigned out = 255;
    if ((int)c < num_comps) {
        const uint32_t k0 = c_off[y], k1 = c_off[y + 1];
        const size_t col = (size_t)x * 4 + c, stride = (size_t)dst_w * 4;
        float s;
        if (k1 - k0 == 1) s = tmp[(size_t)c_pix[k0] * stride + col];
        else {
            s = tmp[(size_t)c_pix[k0] * stride + col] * c_wgt[k0];
#pragma unroll 8
            for (uint32_t k = k0 + 1; k < k1; k++) s += tmp[(size_t)c_pix[k] * stride + col] * c_wgt[k];
        }
        out = mip_to_u8(mip_clamp01(s), (srgb && c != 3) ? 1 : 0, to_srgb);
    }
    dst[(size_t)y * dst_pitch + (size_t)x * 4 + c] = (uint8_t)out;
}

}  // namespace crn
