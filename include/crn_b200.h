/* crn_b200.h -- C ABI of the B200-native crnlib hot path (libcrn_b200.so).
 *
 * This is the seam SURVEY.md section 8(b) describes: the reference's C++ host code (crnlib) keeps its
 * public API (inc/crnlib.h, inc/crn_defs.h) and calls DOWN into these entry points instead of its
 * pthread task-pool loops.  Plain pointers and sizes only; no C++ or torch types; every function
 * returns 0 on success or a negative crn_gpu_status and never throws.  Pointers named d_* are device
 * memory of the context's GPU, h_* are host memory (pinned or pageable).
 *
 * There is no CPU fallback: without a CUDA device every call fails with CRN_GPU_ERR_NO_DEVICE.
 */
#ifndef CRN_B200_H
#define CRN_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRN_B200_ABI_VERSION 1

#if defined(CRN_B200_BUILD) && defined(__GNUC__)
#define CRN_API __attribute__((visibility("default")))
#else
#define CRN_API
#endif

typedef enum crn_gpu_status {
    CRN_GPU_OK = 0,
    CRN_GPU_ERR_NO_DEVICE = -1,
    CRN_GPU_ERR_BAD_PARAM = -2,
    CRN_GPU_ERR_CUDA = -3,
    CRN_GPU_ERR_UNSUPPORTED = -4,
    CRN_GPU_ERR_NO_MEMORY = -5,
    CRN_GPU_ERR_BAD_DATA = -6
} crn_gpu_status;

/* Block formats.  Numbering follows crnlib::dxt_format (reference crnlib/crn_dxt.h:56-76) so the
 * reference's dxt_image can pass its m_format straight through. */
typedef enum crn_gpu_format {
    CRN_GPU_FMT_DXT1 = 0,
    CRN_GPU_FMT_DXT1A = 1,
    CRN_GPU_FMT_DXT3 = 2,
    CRN_GPU_FMT_DXT5 = 3,
    CRN_GPU_FMT_DXT5A = 4,
    CRN_GPU_FMT_DXN_XY = 5,   /* element 0 = R (comp 0), element 1 = G (comp 1) */
    CRN_GPU_FMT_DXN_YX = 6    /* element 0 = G (comp 1), element 1 = R (comp 0) */
} crn_gpu_format;

/* Mirrors crnlib::dxt_image::pack_params (reference crnlib/crn_dxt_image.h:166-228), the per-image
 * knobs of the block-by-block path.  Endpoint caching does not exist here: results are those of the
 * reference with cCRNCompFlagDisableEndpointCaching (the only thread-count independent mode). */
typedef struct crn_gpu_pack_params {
    uint32_t struct_size;               /* sizeof(crn_gpu_pack_params) */
    uint32_t dxt_quality;               /* crn_dxt_quality: 0 superfast .. 4 uber (inc/crnlib.h:160-170) */
    uint32_t perceptual;                /* cCRNCompFlagPerceptual */
    uint32_t use_both_block_types;      /* cCRNCompFlagUseBothBlockTypes */
    uint32_t dxt1a_alpha_threshold;     /* default 128 */
    uint32_t use_transparent_indices_for_black;
    uint32_t grayscale_sampling;
    uint32_t reserved[5];
} crn_gpu_pack_params;

typedef struct crn_gpu_ctx crn_gpu_ctx;

/* Library / device ---------------------------------------------------------------------------- */
CRN_API uint32_t crn_gpu_abi_version(void);
/* 1 when the library was compiled by nvcc for sm_100a, 0 for the g++ SIMT-emulation test build. */
CRN_API int crn_gpu_is_native(void);
CRN_API int crn_gpu_device_count(void);
CRN_API int crn_gpu_create(int device, crn_gpu_ctx** out_ctx);
CRN_API void crn_gpu_destroy(crn_gpu_ctx* ctx);
CRN_API const char* crn_gpu_last_error(const crn_gpu_ctx* ctx);
/* The context's cudaStream_t (as void*), so a caller can order its own copies / events on it. */
CRN_API void* crn_gpu_stream(crn_gpu_ctx* ctx);
CRN_API int crn_gpu_synchronize(crn_gpu_ctx* ctx);
/* Kernels launched through this context since creation (bench.py reports it as gpu_launches). */
CRN_API uint64_t crn_gpu_launch_count(const crn_gpu_ctx* ctx);
CRN_API void crn_gpu_default_pack_params(crn_gpu_pack_params* p);
CRN_API uint32_t crn_gpu_bytes_per_block(uint32_t format);

/* Block-by-block packing ------------------------------------------------
 * Replaces dxt_image::init / init_task / set_block_pixels for the CRN compressor
 * (reference crnlib/crn_dxt_image.cpp:283-349, :447-493, :1427-1541) together with the optimisers
 * it calls (crn_dxt1.cpp:2234, crn_dxt5a.cpp:40).
 * d_rgba: row-major RGBA8 (r first), `pitch_bytes` per row; blocks are gathered with edge clamping.
 * d_out: ((w+3)/4)*((h+3)/4) blocks of crn_gpu_bytes_per_block(format), row-major, alpha element first.
 * Asynchronous on the context's stream. */
CRN_API int crn_gpu_pack_image(crn_gpu_ctx* ctx, uint32_t format, const crn_gpu_pack_params* params,
                       const void* d_rgba, uint32_t width, uint32_t height, uint32_t pitch_bytes, void* d_out);
/* Same through host buffers: H2D copy, kernels, D2H copy, synchronised on return. */
CRN_API int crn_gpu_pack_image_host(crn_gpu_ctx* ctx, uint32_t format, const crn_gpu_pack_params* params,
                            const void* h_rgba, uint32_t width, uint32_t height, uint32_t pitch_bytes, void* h_out);

#ifdef __cplusplus
}
#endif
#endif /* CRN_B200_H */
