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

#define CRN_B200_ABI_VERSION 2

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
    CRN_GPU_ERR_BAD_DATA = -6,
    CRN_GPU_ERR_CANCELLED = -7          /* the progress callback returned 0 */
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
    uint32_t non_hierarchical;          /* clustered paths: qdxt1/qdxt5_params::m_hierarchical == false -- per-block training vectors, no adaptive tiles */
    uint32_t reserved[4];
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
/* Progress / cancel hook, the C form of crn_progress_callback_func (reference inc/crnlib.h:224-228).  The whole-call entry
 * points (crn_gpu_compress_dds / _crn / _mip_chain) invoke it on the CALLING thread between device phases with the
 * reference's phase numbering where it has one -- .CRN: (24, 25, 1, 1) at the end of each pass (crn_comp.cpp:1600);
 * .DDS: (0, 1, percent, 100) block-by-block (crn_dds_comp.cpp:130-134, :233-236), (0, 2, ..) init and (1, 2, ..) pack for the
 * clustered path (:136-146, :172-188).  Returning 0 abandons the call with CRN_GPU_ERR_CANCELLED.  fn = NULL removes it. */
typedef int (*crn_gpu_progress_fn)(uint32_t phase_index, uint32_t total_phases, uint32_t subphase_index, uint32_t total_subphases, void* user);
CRN_API void crn_gpu_set_progress(crn_gpu_ctx* ctx, crn_gpu_progress_fn fn, void* user);
/* cudaMalloc calls made by the buffer pool of this context (and its per-element child contexts) since creation.  Buffers of the clustered /
 * CRN paths are cached per context in size classes, so compressing texture after texture -- or trial after trial of a bitrate search -- stops
 * allocating after the first few calls: the counter stands still in steady state (tests assert it). */
CRN_API uint64_t crn_gpu_pool_mallocs(const crn_gpu_ctx* ctx);
/* Vector-quantiser flavour of the clustered-DDS path (crn_gpu_vq_clusterize, crn_gpu_qdxt_init / _pack and everything above them).
 * 0 (default): crnlib::clusterizer<V>'s algorithm with sums in a fixed parallel order, one launch per frontier (csrc/vq_fast.cuh) -- the
 *    tolerance class BASELINE.json states for clustered output (PSNR within 0.05 dB, bitrate within 1 %).  The N-pixel colour optimiser
 *    (crn_gpu_dxt1_optimize_clusters and its callers) likewise forms the O(U) float sums of clusters with more than 64 unique colours
 *    (mean, covariance, try_median4's 4-means) lane-parallel; every integer error sum stays exact.
 * 1: the reference's member-order float accumulations reproduced bit for bit (csrc/vq_kernels.cuh), so that the cluster assignment EQUALS
 *    the reference's -- several times slower; kept for verification.  The environment variable CRN_B200_VQ_EXACT=1 sets it at context creation. */
CRN_API void crn_gpu_set_vq_mode(crn_gpu_ctx* ctx, int exact_member_order);
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

/* Cluster ("N-pixel") endpoint optimisation -------------------------
 * Replaces the bodies of qdxt1::pack_endpoints_task / qdxt5::pack_endpoints_task (reference
 * crnlib/crn_qdxt1.cpp:471-699, crnlib/crn_qdxt5.cpp:452-576; the same per-cluster step is
 * dxt_hc::determine_color/alpha_endpoint_codebook_task, crnlib/crn_dxt_hc.cpp:663-754, :1014-1130):
 * for every cluster, the pixels of all member blocks are optimised as ONE N-pixel problem and the shared
 * endpoints + per-pixel selectors are written into every member block's 8-byte element.
 *   d_blocks_rgba    n_blocks x 16 RGBA8 pixels (dxt_pixel_block, pixel index 4y+x)
 *   d_cluster_offsets n_clusters+1 CSR offsets into d_cluster_blocks; total_member_blocks = offsets[n_clusters]
 *   d_out            element of block b is written at d_out + b*out_stride_bytes + out_offset_bytes
 *   d_cluster_endpoints / d_cluster_error  optional (may be NULL): low|high<<16 (colour) or first|second<<8
 *                    (alpha) and the optimiser's error per cluster
 * Colour: params->use_both_block_types selects DXT1 semantics (3-colour blocks allowed); pass 0 for the
 * colour element of DXT5.  dxt1a != 0 applies the alpha threshold as qdxt1 does.  Asynchronous. */
CRN_API int crn_gpu_dxt1_optimize_clusters(crn_gpu_ctx* ctx, const crn_gpu_pack_params* params, int dxt1a,
                                           const void* d_blocks_rgba, uint32_t n_blocks,
                                           const uint32_t* d_cluster_offsets, const uint32_t* d_cluster_blocks,
                                           uint32_t n_clusters, uint32_t total_member_blocks,
                                           void* d_out, uint32_t out_stride_bytes, uint32_t out_offset_bytes,
                                           uint32_t* d_cluster_endpoints, uint64_t* d_cluster_error);
CRN_API int crn_gpu_dxt5_optimize_clusters(crn_gpu_ctx* ctx, const crn_gpu_pack_params* params, uint32_t component,
                                           const void* d_blocks_rgba, uint32_t n_blocks,
                                           const uint32_t* d_cluster_offsets, const uint32_t* d_cluster_blocks,
                                           uint32_t n_clusters, uint32_t total_member_blocks,
                                           void* d_out, uint32_t out_stride_bytes, uint32_t out_offset_bytes,
                                           uint32_t* d_cluster_endpoints, uint64_t* d_cluster_error);

/* Clustered-DDS quantiser, tile analysis -----------------------------------
 * Replaces the training-vector half of qdxt1::init (kind 0; crnlib/crn_qdxt1.cpp:103-361) and qdxt5::init
 * (kind 1, channel `component`; crnlib/crn_qdxt5.cpp:103-330) for the hierarchical (adaptive tile) mode:
 * per 8x8 chunk the nine dxt_fast tile fits, the encoding choice, and one endpoint training vector +
 * weight per block.  Blocks are laid out level after level as mipmapped_texture::qdxt_pack_init does;
 * mips[i] = {first_block, block_width, block_height}.  d_vectors: 6 (kind 0: lo.rgb, hi.rgb) or 2 (kind 1)
 * bytes per block; d_weights: uint32 per block; d_chunk_encoding: optional, one byte per chunk. */
typedef struct crn_gpu_mip_desc { uint32_t first_block, block_width, block_height; } crn_gpu_mip_desc;
CRN_API int crn_gpu_qdxt_training(crn_gpu_ctx* ctx, uint32_t kind, uint32_t component, const void* d_blocks_rgba, uint32_t n_blocks,
                                  const crn_gpu_mip_desc* mips, uint32_t num_mips,
                                  void* d_vectors, uint32_t* d_weights, uint8_t* d_chunk_encoding);

/* Selector re-vote per selector cluster -------------------------------------------
 * Replaces qdxt1::optimize_selectors_task (crnlib/crn_qdxt1.cpp:714-865; kind 0) and
 * qdxt5::optimize_selectors_task (crnlib/crn_qdxt5.cpp:578-687; kind 1): within each cluster and block
 * category, every pixel position gets the selector minimising the error summed over the member blocks (each
 * with its own endpoints), written into every member.  d_elements is updated in place.  For kind 0,
 * params->perceptual selects the metric and dxt1a_alpha_threshold (0 = off) exempts 3-colour blocks that
 * contain transparent pixels; for kind 1, `component` is the source channel. */
CRN_API int crn_gpu_optimize_selectors(crn_gpu_ctx* ctx, uint32_t kind, const crn_gpu_pack_params* params, uint32_t component,
                                       const void* d_blocks_rgba, uint32_t n_blocks,
                                       const uint32_t* d_cluster_offsets, const uint32_t* d_cluster_blocks, uint32_t n_clusters,
                                       void* d_elements, uint32_t stride_bytes, uint32_t offset_bytes);

/* Top-down binary-split vector quantiser ------------------------------------------
 * Replaces crnlib::clusterizer<V>::generate_codebook + retrieve_clusters (crnlib/crn_clusterizer.h:65-167,
 * :301-332) and, with threaded != 0, crnlib::threaded_clusterizer<V>::create_clusters
 * (crnlib/crn_threaded_clusterizer.h:70-174), for V = vec2F / vec6F / vec16F with integer components
 * 0..255 (dims = 2, 6, 16).  d_vectors: dims bytes per vector; d_weights: uint32 per vector (both device).
 * max_codebook_size is generate_codebook's max_size (the per-call budget of create_clusters when threaded);
 * retrieve_max_clusters is retrieve_clusters' argument (0 = every leaf, which is what create_clusters does).
 * h_cluster_of (HOST, n entries) receives each vector's cluster index; clusters are numbered in the
 * reference's retrieval order and listing the members of a cluster by ascending vector index reproduces the
 * reference's member order.  The reference's member-order float accumulations (centroids, covariances) are
 * reproduced bit for bit, so the cluster assignment equals the reference's. */
CRN_API int crn_gpu_vq_clusterize(crn_gpu_ctx* ctx, uint32_t dims, const void* d_vectors, const uint32_t* d_weights, uint32_t n,
                                  uint32_t max_codebook_size, uint32_t retrieve_max_clusters, int threaded,
                                  uint32_t* h_cluster_of, uint32_t* num_clusters, uint32_t* codebook_size);

/* Clustered DDS compression (the `-quality N` DDS path) -------------------
 * Mirrors mipmapped_texture::qdxt_pack_init / qdxt_pack (crnlib/crn_mipmapped_texture.cpp:2310-2490,
 * :2492-2590) as dds_comp::convert_to_dxt drives them (crnlib/crn_dds_comp.cpp:156-212), i.e. qdxt1::init/pack
 * (crnlib/crn_qdxt1.cpp:75-446, :910-1030) and qdxt5::init/pack (crnlib/crn_qdxt5.cpp:76-426, :838-960):
 *   init : pixel blocks of all levels -> adaptive-tile analysis -> endpoint training vectors -> endpoint
 *          tree (clusterizer, 65535 leaves max) + the distinct dxt_fast selector count, per element;
 *   pack : quality_level 0..255 -> prune the endpoint tree, optimise one endpoint pair per cluster, build
 *          selector vectors, cluster them (threaded_clusterizer), re-vote the selectors per cluster.
 * levels: faces x mip levels in the reference's order (face-major); pixels RGBA8 on the device or, with
 * pixels_on_host != 0, on the host (copied during init).  The hierarchical (adaptive tile) mode of
 * cCRNCompFlagHierarchical -- crnlib's default -- is the one implemented.  Formats: DXT1, DXT1A, DXT5, DXT5A,
 * DXN_XY, DXN_YX (DXT3 is never clustered by the reference either: crn_dds_comp.cpp:158).
 * pack() writes every level back to back (alpha element first), blocks row-major; it may be called again with
 * another quality_level on the same state, as crnlib's bitrate search does.  Synchronous.
 * Parity: tolerance class (PSNR within 0.05 dB, LZMA size within 1 % of the reference's DDS at the same
 * settings); the reference itself is not bit-reproducible across thread counts on this path. */
typedef struct crn_gpu_qdxt crn_gpu_qdxt;
typedef struct crn_gpu_level_desc { const void* rgba; uint32_t width, height, pitch_bytes; } crn_gpu_level_desc;
typedef struct crn_gpu_qdxt_info {
    uint32_t struct_size;
    uint32_t n_blocks, num_elements;
    uint32_t endpoint_codebook_size[3];    /* clusterizer::get_codebook_size() per element */
    uint32_t max_selector_clusters[3];     /* distinct dxt_fast selectors + 128 */
    uint32_t endpoint_clusters[3];         /* of the last pack() */
    uint32_t selector_clusters[3];
    float endpoint_opt_ms[3];              /* device time of the per-cluster endpoint optimisation kernels of the last pack() */
    /* algorithmic work of the colour element's optimiser in the last pack(): candidate endpoint pairs evaluated, the unique
     * colours they ranged over (sum of U over the evaluations) and the palette entries P per evaluation: integer ops = colour_evals * (11 P + 1) */
    uint64_t opt_candidates[3], opt_colour_evals[3];
    uint32_t opt_palette_entries[3];
    uint32_t reserved;
} crn_gpu_qdxt_info;
CRN_API int crn_gpu_qdxt_init(crn_gpu_ctx* ctx, uint32_t format, const crn_gpu_pack_params* params,
                              const crn_gpu_level_desc* levels, uint32_t num_levels, int pixels_on_host, crn_gpu_qdxt** out);
CRN_API uint64_t crn_gpu_qdxt_output_size(const crn_gpu_qdxt* q);
CRN_API uint64_t crn_gpu_qdxt_level_offset(const crn_gpu_qdxt* q, uint32_t level);
CRN_API int crn_gpu_qdxt_pack(crn_gpu_qdxt* q, uint32_t quality_level, void* dst, int dst_on_host);
CRN_API int crn_gpu_qdxt_get_info(const crn_gpu_qdxt* q, crn_gpu_qdxt_info* info);
CRN_API void crn_gpu_qdxt_free(crn_gpu_qdxt* q);

/* dxt_hc building blocks -- stand-alone; the dxt_hc pipeline itself is not built yet ----
 * crn_gpu_refine_endpoints replaces crnlib::dxt_endpoint_refiner::refine (crnlib/crn_dxt_endpoint_refiner.cpp:36-301)
 * as dxt_hc::determine_color/alpha_endpoint_codebook_task call it per endpoint cluster (crnlib/crn_dxt_hc.cpp:739-753,
 * :1102-1129).  Clusters are CSR ranges over a pixel array: d_pixels_rgba (RGBA8) and d_selectors (one byte per pixel:
 * DXT1 selectors when dxt1_selectors != 0, else DXT5 alpha selectors of channel `component`), d_offsets[n_clusters+1].
 * d_error_to_beat: optional per-cluster params::m_error_to_beat (NULL = UINT64_MAX).  Outputs per cluster:
 * d_endpoints = m_low_color | m_high_color << 16, d_error = m_error, d_ok = refine()'s return value.  Bit-exact. */
CRN_API int crn_gpu_refine_endpoints(crn_gpu_ctx* ctx, int dxt1_selectors, int perceptual, uint32_t component,
                                     const void* d_pixels_rgba, const uint8_t* d_selectors, const uint32_t* d_offsets, uint32_t n_clusters,
                                     const uint64_t* d_error_to_beat, uint32_t* d_endpoints, uint64_t* d_error, uint8_t* d_ok);
/* crn_gpu_nearest_codebook replaces dxt_hc::determine_color_endpoint_clusters_task (dims 6, crnlib/crn_dxt_hc.cpp:836-886)
 * and determine_alpha_endpoint_clusters_task (dims 2, :1132-1163): for each of n float vectors the index of the first
 * codebook entry at minimum squared distance (float, summed in component order). */
CRN_API int crn_gpu_nearest_codebook(crn_gpu_ctx* ctx, uint32_t dims, const float* d_vectors, uint32_t n,
                                     const float* d_codebook, uint32_t codebook_size, uint32_t* d_out);

/* crn_gpu_assign_selectors replaces dxt_hc::create_color_selector_codebook_task (kind 0,
 * crnlib/crn_dxt_hc.cpp:1306-1360) / create_alpha_selector_codebook_task (kind 1, :1516-1586) and the re-vote tail of
 * create_color/alpha_selector_codebook (:1488-1503, :1702-1720): the exhaustive blocks x codebook search (first entry at
 * minimum summed error), the per-entry error tables, and the refined selectors.
 *   d_blocks_rgba   n_blocks x 16 RGBA8
 *   d_block_values  per block the palette it is matched against: kind 0: 4 RGBA8 colours (color_cluster::color_values,
 *                   16 bytes); kind 1: 8 alpha values (alpha_cluster::alpha_values as bytes, 8 bytes), channel `component`
 *   d_block_values_accum  kind 1, optional: the values whose error is accumulated (refined_alpha_values); NULL = same
 *   d_codebook      codebook_size selectors, uint64 each: 16 x 2 bits (kind 0, low 32 bits) or 16 x 3 bits (kind 1),
 *                   pixel p at bit 2p / 3p
 * Outputs: d_best_index[n_blocks], d_refined_codebook[codebook_size] (uint64), d_used[codebook_size].  Bit-exact. */
CRN_API int crn_gpu_assign_selectors(crn_gpu_ctx* ctx, uint32_t kind, int perceptual, uint32_t component,
                                     const void* d_blocks_rgba, uint32_t n_blocks, const void* d_block_values, const void* d_block_values_accum,
                                     const uint64_t* d_codebook, uint32_t codebook_size,
                                     uint32_t* d_best_index, uint64_t* d_refined_codebook, uint8_t* d_used);

/* DXTn blocks -> RGBA8 pixels ------------------------------------------------------------
 * Replaces crnlib::dxt_image::unpack (reference crnlib/crn_dxt_image.cpp:495-567 with get_block_pixels, :1094-1190), the
 * decode half of crn_decompress_dds_to_images (crnlib/crnlib.cpp:293-333) and of crn_decompress_block (:451-498).
 * d_blocks: ((w+3)/4)*((h+3)/4) blocks as crn_gpu_pack_image writes them; d_rgba: height rows of pitch_bytes.  Channels
 * the format does not carry are 0, alpha 255, like the reference's scratch block.  Bit-exact.  Asynchronous; the _host
 * variant copies in and out and synchronises. */
CRN_API int crn_gpu_unpack_image(crn_gpu_ctx* ctx, uint32_t format, const void* d_blocks, uint32_t width, uint32_t height, void* d_rgba, uint32_t pitch_bytes);
CRN_API int crn_gpu_unpack_image_host(crn_gpu_ctx* ctx, uint32_t format, const void* h_blocks, uint32_t width, uint32_t height, void* h_rgba, uint32_t pitch_bytes);

/* Resampling and mip-chain generation ---------------------------------------------------------
 * crn_gpu_resample replaces image_utils::resample in the form the public API runs it (params.m_multithreaded, i.e.
 * threaded_resampler: reference crnlib/crn_image_utils.cpp:668-875, crnlib/crn_threaded_resampler.cpp:64-350, contributor
 * lists from Resampler::make_clist, crnlib/crn_resampler.cpp:119-420); crn_gpu_generate_mipmaps replaces
 * mipmapped_texture::generate_mipmaps (crnlib/crn_mipmapped_texture.cpp:2140-2220): level l = max(1, w >> l) x max(1, h >> l),
 * every level resampled from level 0.  RGBA8 in and out, first_comp = 0.  Bit-exact (the contributor weights and gamma
 * tables are computed on the host with the reference's arithmetic and the same libm; the kernels keep its float order).
 *   filter        crn_mip_filter (inc/crnlib.h:438-446): 0 box, 1 tent, 2 lanczos4, 3 mitchell, 4 kaiser (crnlib/crn_resample_filters.cpp)
 *   num_comps     4: filter alpha too; 3: alpha of the output is 255 (the reference passes 4 when the image has valid alpha)
 *   d_mips        levels 1 .. n-1, tight (pitch = width * 4), one after the other; *num_levels = n including level 0 */
typedef struct crn_gpu_resample_params {
    uint32_t struct_size;               /* sizeof(crn_gpu_resample_params) */
    uint32_t filter;
    float filter_scale;                 /* crn_mipmap_params::m_blurriness (mips) / 1.0 (plain resize) */
    uint32_t srgb;                      /* m_gamma_filtering */
    float source_gamma;                 /* 2.2 */
    uint32_t wrapping;                  /* m_tiled */
    uint32_t num_comps;
    uint32_t renormalize;               /* m_renormalize: image_utils::renorm_normal_map on every resampled image */
    uint32_t reserved[2];
} crn_gpu_resample_params;
CRN_API void crn_gpu_default_resample_params(crn_gpu_resample_params* p);
CRN_API int crn_gpu_resample(crn_gpu_ctx* ctx, const crn_gpu_resample_params* params, const void* d_src, uint32_t src_width, uint32_t src_height, uint32_t src_pitch_bytes,
                             void* d_dst, uint32_t dst_width, uint32_t dst_height, uint32_t dst_pitch_bytes);
CRN_API uint32_t crn_gpu_mip_level_count(uint32_t width, uint32_t height, uint32_t min_mip_size, uint32_t max_levels);
CRN_API int crn_gpu_generate_mipmaps(crn_gpu_ctx* ctx, const crn_gpu_resample_params* params, const void* d_level0, uint32_t width, uint32_t height, uint32_t pitch_bytes,
                                     uint32_t min_mip_size, uint32_t max_levels, void* d_mips, uint64_t capacity, uint32_t* num_levels);
CRN_API int crn_gpu_generate_mipmaps_host(crn_gpu_ctx* ctx, const crn_gpu_resample_params* params, const void* h_level0, uint32_t width, uint32_t height, uint32_t pitch_bytes,
                                          uint32_t min_mip_size, uint32_t max_levels, void* h_mips, uint64_t capacity, uint32_t* num_levels);

/* Block gather: image -> contiguous [block][16] RGBA8 with edge clamp, the level padded to a multiple of
 * pad_pixels: 8 for crn_comp::quantize_images (reference crnlib/crn_comp.cpp:717-741, feeds crn_gpu_hc_compress), 4 for
 * mipmapped_texture::qdxt_pack_init (crnlib/crn_mipmapped_texture.cpp:2418-2472).  d_blocks: blocks_x * blocks_y * 64 bytes. */
CRN_API int crn_gpu_blockify(crn_gpu_ctx* ctx, const void* d_rgba, uint32_t width, uint32_t height, uint32_t pitch_bytes, uint32_t pad_pixels, void* d_blocks,
                             uint32_t* blocks_x, uint32_t* blocks_y);

/* dxt_hc pipeline --------------------------------------------------------------
 * crn_gpu_hc_compress replaces crnlib::dxt_hc::compress (reference crnlib/crn_dxt_hc.cpp:98-312; params mirror
 * dxt_hc::params, crnlib/crn_dxt_hc.h:103-172) as crn_comp::quantize_images calls it (crnlib/crn_comp.cpp:717-766) for
 * the DXT formats: from the [block][16] RGBA8 array of all levels and faces to the four palettes and the per-block
 * endpoint / selector indices + reference flags that the CRN writer codes.  Tile determination, tile palettizing, the
 * three tree quantisers, nearest-codebook assignment, the per-cluster optimiser + refiner, the per-block selector
 * assignment and the selector search / re-vote run on the device.  Tolerance-class, like the reference itself (its
 * output depends on the helper-thread count): same algorithm, float sums in a different (fixed, deterministic) order.
 *   levels[i]       first_block / num_blocks / block_width / weight as crn_comp sets them; block widths and the number
 *                   of block rows per face must be even (crn_comp pads levels to multiples of 8 pixels)
 *   blocks_rgba     num_blocks x 16 RGBA8, host (blocks_on_host != 0) or device memory
 * Results are host arrays owned by the returned object (release with crn_gpu_hc_free):
 *   endpoint_indices  num_blocks x 4 uint16: color, alpha0, alpha1, reference (0 none, 1 left, 2 top)
 *   selector_indices  num_blocks x 4 uint16: color, alpha0, alpha1, 0
 *   color_endpoints   uint32 low565 | high565 << 16;  alpha_endpoints uint32 first | second << 8
 *   color_selectors   uint32, pixel p at bits 2p (linear selector order);  alpha_selectors uint64, pixel p at bits 3p */
/* In-place all-gather over the ranks sharing one dxt_hc call: h_buf (host memory) holds nranks slices of bytes_per_rank
 * bytes, the caller's own slice is filled; on return (0 = success) every slice must be. */
typedef int (*crn_gpu_exchange_fn)(void* user, void* h_buf, uint64_t bytes_per_rank, uint32_t nranks);
typedef struct crn_gpu_hc_level { uint32_t first_block, num_blocks, block_width; float weight; } crn_gpu_hc_level;
typedef struct crn_gpu_hc_params {
    uint32_t struct_size;               /* sizeof(crn_gpu_hc_params) */
    uint32_t format;                    /* CRN_GPU_FMT_DXT1, DXT5, DXT5A, DXN_XY or DXN_YX */
    uint32_t num_blocks, num_levels, num_faces;
    crn_gpu_hc_level levels[16];
    uint32_t perceptual;
    uint32_t color_endpoint_codebook_size, color_selector_codebook_size, alpha_endpoint_codebook_size, alpha_selector_codebook_size;
    float adaptive_tile_color_psnr_derating, adaptive_tile_alpha_psnr_derating, adaptive_tile_color_alpha_weighting_ratio;
    uint32_t alpha_component_indices[2];
    /* One texture on several GPUs: the per-cluster endpoint optimisation + refinement (more than half of the call) is dealt
     * to shard_count ranks by cluster, largest first; every rank runs the rest of the pipeline itself and the per-cluster
     * results (16 bytes per cluster) are all-gathered through `exchange` -- NCCL / gloo on the caller's side.  Every rank
     * must make the same call on the same blocks; the result is identical on all ranks and identical to shard_count = 1. */
    uint32_t shard_rank, shard_count;   /* default 0, 1 */
    crn_gpu_exchange_fn exchange;       /* required when shard_count > 1 */
    void* exchange_user;
} crn_gpu_hc_params;
typedef struct crn_gpu_hc_info {
    uint32_t struct_size;               /* sizeof(crn_gpu_hc_info) */
    uint32_t num_blocks, num_tiles;
    uint32_t n_color_endpoints, n_alpha_endpoints, n_color_selectors, n_alpha_selectors;
    uint32_t vq_rounds[4];              /* device rounds of the four tree quantisers: colour / alpha endpoints, colour / alpha selectors */
    uint32_t unique_vectors[4];         /* their training-set sizes */
} crn_gpu_hc_info;
typedef struct crn_gpu_hc crn_gpu_hc;
CRN_API void crn_gpu_default_hc_params(crn_gpu_hc_params* p);
CRN_API int crn_gpu_hc_compress(crn_gpu_ctx* ctx, const crn_gpu_hc_params* params, const void* blocks_rgba, int blocks_on_host, crn_gpu_hc** out);
CRN_API int crn_gpu_hc_get_info(const crn_gpu_hc* hc, crn_gpu_hc_info* info);
CRN_API const uint16_t* crn_gpu_hc_endpoint_indices(const crn_gpu_hc* hc);
CRN_API const uint16_t* crn_gpu_hc_selector_indices(const crn_gpu_hc* hc);
CRN_API const uint32_t* crn_gpu_hc_color_endpoints(const crn_gpu_hc* hc);
CRN_API const uint32_t* crn_gpu_hc_alpha_endpoints(const crn_gpu_hc* hc);
CRN_API const uint32_t* crn_gpu_hc_color_selectors(const crn_gpu_hc* hc);
CRN_API const uint64_t* crn_gpu_hc_alpha_selectors(const crn_gpu_hc* hc);
CRN_API const uint8_t* crn_gpu_hc_block_encodings(const crn_gpu_hc* hc);      /* m_block_encodings, one byte per block */
CRN_API const uint32_t* crn_gpu_hc_tile_indices(const crn_gpu_hc* hc);        /* m_tile_indices */
CRN_API void crn_gpu_hc_free(crn_gpu_hc* hc);

/* .CRN writer back-end ---------------------------------------------------------------
 * Replaces what crn_comp does after quantize_images (reference crnlib/crn_comp.cpp): optimize_color / optimize_alpha
 * (:990-1058, :1285-1354: palette orderings, four endpoint trials costed by their coded size), pack_* (:43-293),
 * the two pack_blocks passes of compress_internal (:295-422, :1515-1611), pack_data_models and create_comp_data
 * (:1356-1496).  Host C++ (no device work, usable without a GPU): the orderings follow the reference step for step;
 * a file decodes (crnd_unpack_level) to exactly the blocks the reference's own writer would produce from the same
 * palettes and indices, and on every test vector it is byte-identical to the reference's file (tests/test_crn_writer_cpu.py).
 * crn_gpu_crn_params mirrors the crn_comp_params fields crn_comp reads for a .CRN (inc/crnlib.h:231-414). */
typedef struct crn_gpu_crn_params {
    uint32_t struct_size;               /* sizeof(crn_gpu_crn_params) */
    uint32_t crn_format;                /* crn_format: 0 DXT1, 2 DXT5, 7 DXN_XY, 8 DXN_YX, 9 DXT5A (others: CRN_GPU_ERR_UNSUPPORTED) */
    uint32_t width, height, levels, faces;
    uint32_t quality_level;             /* m_quality_level, 0..255 */
    uint32_t perceptual;                /* cCRNCompFlagPerceptual */
    uint32_t alpha_component;           /* m_alpha_component */
    uint32_t userdata0, userdata1;
    uint32_t palette_sizes[4];          /* colour endpoints, colour selectors, alpha endpoints, alpha selectors; any 0 = derive all
                                           from quality_level (otherwise cCRNCompFlagManualPaletteSizes) */
    float adaptive_tile_color_psnr_derating, adaptive_tile_alpha_psnr_derating;
    float target_bitrate;               /* m_target_bitrate in bits per texel; 0 = use quality_level (crn_gpu_compress_crn only) */
    /* One texture on several GPUs (crn_gpu_compress_crn only): passed to every crn_gpu_hc_compress of the call, see
     * crn_gpu_hc_params.  Every rank makes the same call on the same images and returns the same file. */
    uint32_t shard_rank, shard_count;   /* default 0, 1 (0 is read as 1) */
    crn_gpu_exchange_fn exchange;       /* required when shard_count > 1 */
    void* exchange_user;
} crn_gpu_crn_params;
CRN_API void crn_gpu_default_crn_params(crn_gpu_crn_params* p);
/* crn_comp::alias_images' level table + quantize_images' parameter derivation (crn_comp.cpp:458-466, :525-716): fills
 * every field of *hp (format, blocks, levels with weights min(12, 1.3^level), codebook sizes from the quality level). */
CRN_API int crn_gpu_crn_hc_params(const crn_gpu_crn_params* p, crn_gpu_hc_params* hp);
/* Palettes + indices (the arrays of crn_gpu_hc_*, or any source with that layout) -> .crn bytes.  *out_file is
 * malloc'ed by the library; release it with crn_gpu_free_file. */
CRN_API int crn_gpu_crn_write(const crn_gpu_crn_params* p, const crn_gpu_hc_params* hp, const uint16_t* endpoint_indices, const uint16_t* selector_indices,
                              const uint32_t* color_endpoints, uint32_t n_color_endpoints, const uint32_t* alpha_endpoints, uint32_t n_alpha_endpoints,
                              const uint32_t* color_selectors, uint32_t n_color_selectors, const uint64_t* alpha_selectors, uint32_t n_alpha_selectors,
                              void** out_file, uint32_t* out_size);
/* crn_compress to a .CRN (create_compressed_texture, crnlib/crn_texture_comp.cpp:60-275, over crn_comp::compress_pass,
 * crn_comp.cpp:1613-1656): h_images[face * levels + level] are host RGBA8 images of max(1, width >> level) x
 * max(1, height >> level), tight pitch (crn_comp_params::m_pImages).  Gathers the padded blocks on the device once, then per
 * pass runs crn_gpu_hc_compress and the writer.  target_bitrate > 0 runs the reference's interpolative quality search
 * (same bracket / interpolation / acceptance rules) over blocks that stay resident in HBM.  out_bitrate (optional) = file
 * bits / texels, out_quality (optional) = the quality level of the returned file. */
CRN_API int crn_gpu_compress_crn(crn_gpu_ctx* ctx, const crn_gpu_crn_params* p, const void* const* h_images, void** out_file, uint32_t* out_size, float* out_bitrate,
                                 uint32_t* out_quality);
CRN_API void crn_gpu_free_file(void* file);

/* CRN -> DXTn transcoding ---------------------------------------------------
 * Mirrors the crnd_* API of inc/crn_defs.h:139-221 (bodies in inc/crn_decomp.h): crnd_get_texture_info
 * (:2737), crnd_unpack_begin (:4404), crnd_unpack_level (:4441), crnd_unpack_end (:4478).  Same contract:
 * the context borrows nothing after begin() returns (the file bytes are copied to the device);
 * unpack_level writes blocks_y rows of blocks_x blocks per face, `row_pitch_in_bytes` apart (0 = tight,
 * otherwise >= blocks_x * bytes_per_block and a multiple of 4), dst_size_in_bytes >= pitch * blocks_y,
 * bit-for-bit what crnd_unpack_level produces.  Destination pointers are DEVICE memory. */
typedef struct crn_gpu_texture_info {
    uint32_t struct_size;          /* sizeof(crn_gpu_texture_info) */
    uint32_t width, height, levels, faces;
    uint32_t bytes_per_block;      /* 8 or 16 */
    uint32_t userdata0, userdata1;
    uint32_t format;               /* crn_format (inc/crnlib.h:72-112) */
} crn_gpu_texture_info;

typedef struct crn_gpu_texture crn_gpu_texture;

/* Host-only header crack; no device needed. */
CRN_API int crn_gpu_crnd_get_texture_info(const void* h_crn, uint32_t crn_size, crn_gpu_texture_info* info);
/* Parses header + Huffman models on the host, uploads the file, decodes the four palettes on the device. */
CRN_API int crn_gpu_crnd_unpack_begin(crn_gpu_ctx* ctx, const void* h_crn, uint32_t crn_size, crn_gpu_texture** out_tex);
/* One level, caller-chosen destination per face (1 or 6 device pointers).  Asynchronous. */
CRN_API int crn_gpu_crnd_unpack_level(crn_gpu_texture* tex, void* const* d_dst_faces, uint32_t dst_size_in_bytes,
                                      uint32_t row_pitch_in_bytes, uint32_t level_index);
/* Same with HOST destination pointers -- crnd_unpack_level's own contract (inc/crn_decomp.h:4441-4476: "cached or write
 * combined memory"): the level is transcoded into device scratch and each face copied out row by row at the caller's pitch.
 * Synchronous. */
CRN_API int crn_gpu_crnd_unpack_level_host(crn_gpu_texture* tex, void* const* h_dst_faces, uint32_t dst_size_in_bytes,
                                           uint32_t row_pitch_in_bytes, uint32_t level_index);
/* All levels in ONE launch (levels decode concurrently, one warp each) into a tightly packed device
 * buffer laid out level-major then face-major; crn_gpu_crnd_level_offset gives each face's offset. */
CRN_API uint64_t crn_gpu_crnd_total_size(const crn_gpu_texture* tex);
CRN_API uint64_t crn_gpu_crnd_level_offset(const crn_gpu_texture* tex, uint32_t level_index, uint32_t face_index);
CRN_API int crn_gpu_crnd_unpack_all_levels(crn_gpu_texture* tex, void* d_dst, uint64_t dst_capacity);
/* Same, then copies the result to host memory and synchronises. */
CRN_API int crn_gpu_crnd_unpack_all_levels_host(crn_gpu_texture* tex, void* h_dst, uint64_t dst_capacity);
/* Several textures of one context in ONE launch (one CTA per file): the batched form in which the
 * per-level serial entropy decode is amortised.  d_dst[i] receives texture i, tight layout. */
CRN_API int crn_gpu_crnd_unpack_batch(crn_gpu_ctx* ctx, crn_gpu_texture* const* textures, uint32_t count, void* const* d_dst,
                                      const uint64_t* dst_capacity);
CRN_API int crn_gpu_crnd_unpack_end(crn_gpu_texture* tex);

/* DDS container edge --------------------------------------------------------------------
 * crn_gpu_dds_header: the 128 bytes ("DDS " + DDSURFACEDESC2) mipmapped_texture::write_dds writes for a block-compressed
 * texture (crnlib/crn_mipmapped_texture.cpp:921-1084), from the crn_format of a .crn header; host only.
 * crn_gpu_crn_to_dds: crn_decompress_crn_to_dds (inc/crnlib.h:620, crnlib/crnlib.cpp:269-291) -- all levels transcoded on the
 * device, payload laid out faces outermost; byte-identical to the reference's file.  *out_file is malloc'ed by the library;
 * release it with crn_gpu_free_file. */
CRN_API int crn_gpu_dds_header(uint32_t crn_format, uint32_t width, uint32_t height, uint32_t levels, uint32_t faces, void* out_128_bytes);
CRN_API int crn_gpu_crn_to_dds(crn_gpu_ctx* ctx, const void* h_crn, uint32_t crn_size, void** out_file, uint32_t* out_size);
/* crn_compress(cCRNFileTypeDDS) (inc/crnlib.h:609; dds_comp::compress_init / convert_to_dxt / compress_pass,
 * crnlib/crn_dds_comp.cpp:148-289): quality_level 255 (or DXT3) packs block by block (crn_gpu_pack_image), anything lower runs
 * the clustered path (crn_gpu_qdxt_init + crn_gpu_qdxt_pack); the result is a complete .dds file (header + faces outermost).
 * h_images[face * levels + level]: host RGBA8, tight pitch.  DXT1 becomes DXT1A under the reference's rule (any alpha < 255,
 * both block types, dxt1a_for_transparency). */
typedef struct crn_gpu_dds_params {
    uint32_t struct_size;               /* sizeof(crn_gpu_dds_params) */
    uint32_t crn_format;                /* crn_format: 0 DXT1, 1 DXT3, 2 DXT5, 3 DXT5_CCxY, 4 DXT5_xGxR, 5 DXT5_xGBR, 6 DXT5_AGBR, 7 DXN_XY, 8 DXN_YX, 9 DXT5A */
    uint32_t width, height, levels, faces;
    uint32_t quality_level;             /* m_quality_level */
    uint32_t dxt1a_for_transparency;    /* cCRNCompFlagDXT1AForTransparency */
    crn_gpu_pack_params pack;           /* dxt_image::pack_params::init(crn_comp_params) (crnlib/crn_dxt_image.h:192-203) */
    float target_bitrate;               /* m_target_bitrate: LZMA-compressed bits per texel; 0 = use quality_level (crn_gpu_compress_dds_ex only) */
    uint32_t hierarchical;              /* cCRNCompFlagHierarchical (default 1; 0 is CRN_GPU_ERR_UNSUPPORTED below quality 255) */
    uint32_t reserved[2];
} crn_gpu_dds_params;
CRN_API void crn_gpu_default_dds_params(crn_gpu_dds_params* p);
CRN_API int crn_gpu_compress_dds(crn_gpu_ctx* ctx, const crn_gpu_dds_params* p, const void* const* h_images, void** out_file, uint32_t* out_size);
/* Same with crn_compress's two optional outputs (inc/crnlib.h:609) and the target-bitrate search of create_compressed_texture
 * (crnlib/crn_texture_comp.cpp:120-262) over dds_comp::compress_pass (crnlib/crn_dds_comp.cpp:254-306): the clustered state is built once
 * (qdxt_pack_init) and every trial is one qdxt_pack at that quality level, as in the reference.  out_bitrate = LZMA-compressed file bits /
 * texels of all faces and levels; the size comes from the system's liblzma (dlopen, same coder parameters as the reference's vendored LZMA
 * SDK at its default level) -- 0.0 when liblzma is absent, and a target bitrate is then CRN_GPU_ERR_UNSUPPORTED.  out_quality = the level
 * of the returned file.  Either output may be NULL. */
CRN_API int crn_gpu_compress_dds_ex(crn_gpu_ctx* ctx, const crn_gpu_dds_params* p, const void* const* h_images, void** out_file, uint32_t* out_size,
                                    float* out_bitrate, uint32_t* out_quality);
/* The size measurement alone (bytes, 0 = liblzma unavailable); host only. */
CRN_API uint64_t crn_gpu_lzma_size(const void* data, uint64_t size);

/* .dds in: crn_decompress_dds_to_images (inc/crnlib.h:634, crnlib/crnlib.cpp:293-333) = mipmapped_texture::read_dds
 * (crnlib/crn_mipmapped_texture.cpp:489-862) + unpack_from_dxt(uncook = true) (:2013-2031, mip_level::get_unpacked_image :358-371).
 * Block formats (DXT1 / DXT1A by the reference's scan for transparent texels, DXT3, DXT5 and its four swizzled variants, DXN_XY / ATI2,
 * DXT5A) are decoded on the device (unpack_blocks_kernel) and "uncooked" (swizzles undone, DXN z regenerated); the uncompressed layouts
 * the reference reads (8-32 bit RGB / luminance / alpha with channel bit masks) go through a mask-extraction kernel.  Bit-exact.
 * crn_gpu_dds_get_desc is host-only and reports the geometry plus `pixel_format` = the crnlib::pixel_format (inc/dds_defs.h:33-69) the
 * DECODED images carry (what crn_texture_desc::m_fmt_fourcc receives) -- for a DXT1 file under the assumption that no block is transparent;
 * crn_gpu_dds_to_images rewrites *desc (may be NULL) with the final answer.  h_images[level + levels * face] (crnlib.cpp:327): caller
 * memory of max(1, width >> level) * max(1, height >> level) * 4 bytes each, RGBA8 (r first). */
typedef struct crn_gpu_dds_desc {
    uint32_t struct_size;               /* sizeof(crn_gpu_dds_desc) */
    uint32_t width, height, levels, faces;
    uint32_t pixel_format;              /* of the decoded images: 'RGBx', 'RGBA', 'Lxxx', 'LxxA' or 'xxxA' */
    uint32_t file_format;               /* crnlib::pixel_format of the file (DXT1 ... or one of the above) */
    uint32_t block_format;              /* crn_gpu_format of the payload, 0xFFFFFFFF for an uncompressed file */
} crn_gpu_dds_desc;
CRN_API int crn_gpu_dds_get_desc(const void* h_dds, uint32_t dds_size, crn_gpu_dds_desc* desc);
CRN_API int crn_gpu_dds_to_images(crn_gpu_ctx* ctx, const void* h_dds, uint32_t dds_size, void* const* h_images, uint32_t num_images, crn_gpu_dds_desc* desc);
/* image_utils::convert_image (crnlib/crn_image_utils.cpp:1181-1380) on a device image, in place: conversion 1 To_CCxY, 2 From_CCxY,
 * 3 To_xGxR, 4 From_xGxR, 5 To_xGBR, 6 From_xGBR, 7 To_AGBR, 8 From_AGBR, 9 XY_to_XYZ.  Asynchronous. */
CRN_API int crn_gpu_convert_pixels(crn_gpu_ctx* ctx, void* d_rgba, uint32_t width, uint32_t height, uint32_t pitch_bytes, uint32_t conversion);

/* crn_compress with a crn_mipmap_params (inc/crnlib.h:614; create_texture_mipmaps in its generate mode,
 * crnlib/crn_texture_comp.cpp:352-575): level 0 of each face in, the chain from crn_gpu_generate_mipmaps, then
 * crn_gpu_compress_crn (file_type 0, `cp`) or crn_gpu_compress_dds (file_type 1, `dp`); the params' `levels` is replaced by the
 * generated count.  mip = NULL takes crn_mipmap_params' defaults (kaiser, gamma filtering 2.2, blurriness 0.9); num_comps 0 =
 * decide like the reference (4 when any source texel has alpha < 255, else 3).  min_mip_size / max_levels 0 = 1 / 16.
 * Cropping, clamping and rescaling of the source (crn_mipmap_params::m_window_*, m_clamp_*, m_scale_mode): crn_gpu_prepare_mip_source first. */
/* The source options of create_texture_mipmaps (crnlib/crn_texture_comp.cpp:392-540): crop to a window, clamp, rescale (crn_scale_mode), and
 * the resample to the resulting size (filter scale 1, never wrapping -- the reference clears m_wrapping there -- renormalised when asked; also
 * forced by renormalize && rtopmip at unchanged size).  Level 0 of each face in (host, tight pitch).  *out_width / *out_height = the size
 * compression continues with; out_faces[f] = a malloc'ed replacement image (crn_gpu_free_file) or NULL when the source is used as it is.
 * Like the reference, cropping a cubemap is skipped (with clamp_scale = 0 its clamp too) and every mip level but 0 is dropped by a crop or resize. */
typedef struct crn_gpu_mip_source_params {
    uint32_t struct_size;               /* sizeof(crn_gpu_mip_source_params) */
    uint32_t window_left, window_top, window_right, window_bottom;
    uint32_t clamp_width, clamp_height, clamp_scale;
    uint32_t scale_mode;                /* crn_scale_mode (inc/crnlib.h:454-466): 0 disabled, 1 absolute, 2 relative, 3 lower, 4 nearest, 5 next power of two */
    float scale_x, scale_y;
    uint32_t rtopmip;
    uint32_t reserved[4];
} crn_gpu_mip_source_params;
CRN_API int crn_gpu_prepare_mip_source(crn_gpu_ctx* ctx, const crn_gpu_mip_source_params* sp, const crn_gpu_resample_params* mip, uint32_t faces, uint32_t width, uint32_t height,
                                       const void* const* h_level0_faces, void** out_faces, uint32_t* out_width, uint32_t* out_height, uint32_t* out_changed);
CRN_API int crn_gpu_compress_mip_chain(crn_gpu_ctx* ctx, uint32_t file_type, const crn_gpu_crn_params* cp, const crn_gpu_dds_params* dp, const crn_gpu_resample_params* mip,
                                       uint32_t min_mip_size, uint32_t max_levels, const void* const* h_level0_faces, void** out_file, uint32_t* out_size);

#ifdef __cplusplus
}
#endif
#endif /* CRN_B200_H */
