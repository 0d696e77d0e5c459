// crn_b200.cu -- C-ABI entry points of libcrn_b200.so (see include/crn_b200.h).
//
// One translation unit: the kernels live in the .cuh files included below.  Built by nvcc for
// sm_100a (the product) and, for the CPU-side tests only, by g++ against tests/cusim (SIMT emulator).
#include "../../include/crn_b200.h"
#include "launch.h"
#include "pack_kernels.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>

struct crn_gpu_ctx {
    int device;
    cudaStream_t stream;
    int sm_count;
    uint64_t launches;
    char err[256];
    // reusable device staging for the *_host entry points
    void* d_in; size_t d_in_cap;
    void* d_out; size_t d_out_cap;
    void* d_state; size_t d_state_cap;   // Dxt1BlockState scratch of the colour phase kernels
};

namespace {

int set_err(crn_gpu_ctx* ctx, int code, const char* what, cudaError_t ce = cudaSuccess)
{
    if (ctx) {
        if (ce != cudaSuccess) snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what, cudaGetErrorString(ce));
        else snprintf(ctx->err, sizeof(ctx->err), "%s", what);
    }
    return code;
}

#define CRN_CUDA(ctx, call)                                                          \
    do {                                                                             \
        cudaError_t ce_ = (call);                                                    \
        if (ce_ != cudaSuccess) return set_err((ctx), CRN_GPU_ERR_CUDA, #call, ce_); \
    } while (0)

int ensure(crn_gpu_ctx* ctx, void** p, size_t* cap, size_t need)
{
    if (*cap >= need) return CRN_GPU_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    cudaError_t ce = cudaMalloc(p, need);
    if (ce != cudaSuccess) return set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "cudaMalloc", ce);
    *cap = need;
    return CRN_GPU_OK;
}

int grid_for(const crn_gpu_ctx* ctx, uint32_t total_blocks, int warps_per_cta, int ctas_per_sm)
{
    const uint32_t need = (total_blocks + warps_per_cta - 1) / warps_per_cta;
    const uint32_t cap = (uint32_t)(ctx->sm_count * ctas_per_sm);
    uint32_t g = need < cap ? need : cap;
    return (int)(g ? g : 1);
}

}  // namespace

extern "C" {

uint32_t crn_gpu_abi_version(void) { return CRN_B200_ABI_VERSION; }

int crn_gpu_is_native(void)
{
#ifdef __CUDACC__
    return 1;
#else
    return 0;
#endif
}

int crn_gpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int crn_gpu_create(int device, crn_gpu_ctx** out_ctx)
{
    if (!out_ctx) return CRN_GPU_ERR_BAD_PARAM;
    *out_ctx = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return CRN_GPU_ERR_NO_DEVICE;
    crn_gpu_ctx* ctx = new (std::nothrow) crn_gpu_ctx();
    if (!ctx) return CRN_GPU_ERR_NO_MEMORY;
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return CRN_GPU_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return CRN_GPU_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return CRN_GPU_ERR_CUDA; }
    *out_ctx = ctx;
    return CRN_GPU_OK;
}

void crn_gpu_destroy(crn_gpu_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->d_in) cudaFree(ctx->d_in);
    if (ctx->d_out) cudaFree(ctx->d_out);
    if (ctx->d_state) cudaFree(ctx->d_state);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* crn_gpu_last_error(const crn_gpu_ctx* ctx) { return ctx ? ctx->err : "null context"; }
void* crn_gpu_stream(crn_gpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t crn_gpu_launch_count(const crn_gpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

int crn_gpu_synchronize(crn_gpu_ctx* ctx)
{
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CRN_GPU_OK;
}

void crn_gpu_default_pack_params(crn_gpu_pack_params* p)
{   // defaults of crn_comp_params::clear() (inc/crnlib.h:239-273) as seen by the block packer
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->struct_size = sizeof(*p);
    p->dxt_quality = 4;
    p->perceptual = 1;
    p->use_both_block_types = 1;
    p->dxt1a_alpha_threshold = 128;
}

uint32_t crn_gpu_bytes_per_block(uint32_t format)
{
    switch (format) {
    case CRN_GPU_FMT_DXT1: case CRN_GPU_FMT_DXT1A: case CRN_GPU_FMT_DXT5A: return 8;
    case CRN_GPU_FMT_DXT3: case CRN_GPU_FMT_DXT5: case CRN_GPU_FMT_DXN_XY: case CRN_GPU_FMT_DXN_YX: return 16;
    default: return 0;
    }
}

int crn_gpu_pack_image(crn_gpu_ctx* ctx, uint32_t format, const crn_gpu_pack_params* params,
                       const void* d_rgba, uint32_t width, uint32_t height, uint32_t pitch_bytes, void* d_out)
{
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (!params || params->struct_size != sizeof(crn_gpu_pack_params) || !d_rgba || !d_out || !width || !height ||
        pitch_bytes < width * 4u || (pitch_bytes & 3u))
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_pack_image: bad argument");
    const uint32_t bpb = crn_gpu_bytes_per_block(format);
    if (!bpb) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_pack_image: unknown format");
    if (params->dxt_quality > 4 || params->dxt1a_alpha_threshold > 255)
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_pack_image: parameter out of range");
    const bool has_color = format == CRN_GPU_FMT_DXT1 || format == CRN_GPU_FMT_DXT1A || format == CRN_GPU_FMT_DXT3 || format == CRN_GPU_FMT_DXT5;
    if (has_color && params->dxt_quality < 3)
        return set_err(ctx, CRN_GPU_ERR_UNSUPPORTED, "crn_gpu_pack_image: colour blocks need dxt_quality better (3) or uber (4)");
    if (has_color && params->use_transparent_indices_for_black)
        return set_err(ctx, CRN_GPU_ERR_UNSUPPORTED, "crn_gpu_pack_image: use_transparent_indices_for_black is not implemented");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));

    crn::ImageView img;
    img.rgba = static_cast<const uint8_t*>(d_rgba);
    img.width = width; img.height = height; img.pitch = pitch_bytes;
    img.blocks_x = (width + 3) >> 2; img.blocks_y = (height + 3) >> 2;
    const uint32_t total = img.blocks_x * img.blocks_y;
    uint8_t* out = static_cast<uint8_t*>(d_out);
    const int q = (int)params->dxt_quality;
    const int both = params->use_both_block_types ? 1 : 0;
    const int threads = crn::kPackWarpsPerCta * 32;
    const int grid = grid_for(ctx, total, crn::kPackWarpsPerCta, 8);

    auto launch_alpha = [&](uint32_t comp, uint32_t ofs) {
        CRN_LAUNCH(crn::pack_alpha_element_kernel, grid, threads, 0, ctx->stream, img, comp, q, both, out, bpb, ofs);
        ctx->launches++;
    };
    int color_rc = CRN_GPU_OK;
    auto launch_color = [&](uint32_t ofs) {
        crn::Dxt1Params dp;
        dp.quality = q;
        dp.perceptual = params->perceptual ? 1 : 0;
        dp.pixels_have_alpha = 0;
        // crn_dxt_image.cpp:1463-1473: 3-colour blocks only for DXT1 / DXT1A
        dp.use_alpha_blocks = (format == CRN_GPU_FMT_DXT1 || format == CRN_GPU_FMT_DXT1A) ? both : 0;
        dp.force_alpha_blocks = 0;
        dp.grayscale_sampling = params->grayscale_sampling ? 1 : 0;
        dp.alpha_threshold = params->dxt1a_alpha_threshold;
        const int dxt1a = format == CRN_GPU_FMT_DXT1A;
        // five phase kernels per chunk of blocks; the per-block state lives in ctx->d_state between them
        const uint32_t chunk_cap = 1u << 18;
        const uint32_t chunk = total < chunk_cap ? total : chunk_cap;
        color_rc = ensure(ctx, &ctx->d_state, &ctx->d_state_cap, (size_t)chunk * sizeof(crn::Dxt1BlockState));
        if (color_rc) return;
        crn::Dxt1BlockState* st = static_cast<crn::Dxt1BlockState*>(ctx->d_state);
        for (uint32_t first = 0; first < total; first += chunk) {
            const uint32_t count = total - first < chunk ? total - first : chunk;
            const int g = grid_for(ctx, count, crn::kPackWarpsPerCta, 8);
            CRN_LAUNCH(crn::pack_color_phase_kernel<0>, g, threads, 0, ctx->stream, img, dp, dxt1a, st, first, count, out, bpb, ofs);
            CRN_LAUNCH(crn::pack_color_phase_kernel<1>, g, threads, 0, ctx->stream, img, dp, dxt1a, st, first, count, out, bpb, ofs);
            CRN_LAUNCH(crn::pack_color_phase_kernel<2>, g, threads, 0, ctx->stream, img, dp, dxt1a, st, first, count, out, bpb, ofs);
            CRN_LAUNCH(crn::pack_color_phase_kernel<3>, g, threads, 0, ctx->stream, img, dp, dxt1a, st, first, count, out, bpb, ofs);
            CRN_LAUNCH(crn::pack_color_phase_kernel<4>, g, threads, 0, ctx->stream, img, dp, dxt1a, st, first, count, out, bpb, ofs);
            ctx->launches += 5;
        }
    };

    switch (format) {
    case CRN_GPU_FMT_DXT1: case CRN_GPU_FMT_DXT1A: launch_color(0); break;
    case CRN_GPU_FMT_DXT3: {
        const int g3 = (int)((total + 255) / 256);
        CRN_LAUNCH(crn::pack_dxt3_alpha_kernel, g3 ? g3 : 1, 256, 0, ctx->stream, img, 3u, out, bpb, 0u);
        ctx->launches++;
        launch_color(8);
        break;
    }
    case CRN_GPU_FMT_DXT5: launch_alpha(3, 0); launch_color(8); break;
    case CRN_GPU_FMT_DXT5A: launch_alpha(3, 0); break;
    case CRN_GPU_FMT_DXN_XY: launch_alpha(0, 0); launch_alpha(1, 8); break;
    case CRN_GPU_FMT_DXN_YX: launch_alpha(1, 0); launch_alpha(0, 8); break;
    default: break;
    }
    if (color_rc) return color_rc;
    CRN_CUDA(ctx, cudaGetLastError());
    return CRN_GPU_OK;
}

int crn_gpu_pack_image_host(crn_gpu_ctx* ctx, uint32_t format, const crn_gpu_pack_params* params,
                            const void* h_rgba, uint32_t width, uint32_t height, uint32_t pitch_bytes, void* h_out)
{
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (!h_rgba || !h_out || !width || !height) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_pack_image_host: bad argument");
    const uint32_t bpb = crn_gpu_bytes_per_block(format);
    if (!bpb) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_pack_image_host: unknown format");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t in_bytes = (size_t)pitch_bytes * height;
    const size_t out_bytes = (size_t)((width + 3) >> 2) * ((height + 3) >> 2) * bpb;
    int rc = ensure(ctx, &ctx->d_in, &ctx->d_in_cap, in_bytes);
    if (rc) return rc;
    rc = ensure(ctx, &ctx->d_out, &ctx->d_out_cap, out_bytes);
    if (rc) return rc;
    CRN_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, h_rgba, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = crn_gpu_pack_image(ctx, format, params, ctx->d_in, width, height, pitch_bytes, ctx->d_out);
    if (rc) return rc;
    CRN_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CRN_GPU_OK;
}

}  // extern "C"
