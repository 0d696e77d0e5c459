// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" window onto the UNMODIFIED reference (crnlib 1.2.0), which
// oracle/Makefile compiles from the sources where they lie under /root/reference
// into oracle/_ref/liboracle_ref.so.  Only tests/, __graft_entry__.smoke() and
// bench.py's CPU-baseline legs may load that library; the product never does.
//
// Every entry point below forwards to a reference class / function without
// altering its arguments:
//   ref_dxt1_optimize   -> crnlib::dxt1_endpoint_optimizer::compute   (crnlib/crn_dxt1.cpp:2234)
//   ref_dxt5_optimize   -> crnlib::dxt5_endpoint_optimizer::compute   (crnlib/crn_dxt5a.cpp:40)
//   ref_pack_image      -> crnlib::dxt_image::init                    (crnlib/crn_dxt_image.cpp:447)
//   ref_compress        -> crn_compress                               (crnlib/crnlib.cpp:215)
//   ref_transcode_*     -> crnd::crnd_unpack_begin/level/end          (inc/crn_decomp.h:4404-4460)
//   ref_crn_to_dds      -> crn_decompress_crn_to_dds                  (crnlib/crnlib.cpp:269)
#include "crn_core.h"
#include "crn_dxt1.h"
#include "crn_dxt5a.h"
#include "crn_dxt_image.h"
#include "crn_image.h"
#include "crn_dxt_endpoint_refiner.h"
#include "crn_dxt_fast.h"
#include "crnlib.h"
#include "crn_defs.h"   // declarations only (crnd:: API); the bodies live in crn_decomp.cpp
#include <string.h>
#include <time.h>

using namespace crnlib;

#define SHIM_API extern "C" __attribute__((visibility("default")))

SHIM_API const char* ref_version(void) { return "crnlib-1.2.0-unmodified"; }

// One N-pixel DXT1 endpoint optimisation.  pixels: RGBA8 (r first).  Returns 1 on success.
SHIM_API int ref_dxt1_optimize(const uint8_t* pixels, uint32_t num_pixels, int quality, int perceptual,
                               int pixels_have_alpha, int use_alpha_blocks, uint32_t alpha_threshold,
                               int grayscale_sampling, int transparent_for_black, int force_alpha_blocks,
                               uint16_t* low, uint16_t* high, uint8_t* selectors, uint64_t* error, uint8_t* alpha_block)
{
    dxt1_endpoint_optimizer opt;
    dxt1_endpoint_optimizer::params p;
    dxt1_endpoint_optimizer::results r;
    p.m_pPixels = reinterpret_cast<const color_quad_u8*>(pixels);
    p.m_num_pixels = num_pixels;
    p.m_quality = (crn_dxt_quality)quality;
    p.m_perceptual = perceptual != 0;
    p.m_pixels_have_alpha = pixels_have_alpha != 0;
    p.m_use_alpha_blocks = use_alpha_blocks != 0;
    p.m_dxt1a_alpha_threshold = alpha_threshold;
    p.m_grayscale_sampling = grayscale_sampling != 0;
    p.m_use_transparent_indices_for_black = transparent_for_black != 0;
    p.m_force_alpha_blocks = force_alpha_blocks != 0;
    p.m_endpoint_caching = false;
    r.m_pSelectors = selectors;
    r.m_error = 0;
    r.m_low_color = r.m_high_color = 0;
    r.m_alpha_block = false;
    if (!opt.compute(p, r))
        return 0;
    *low = r.m_low_color;
    *high = r.m_high_color;
    *error = r.m_error;
    *alpha_block = r.m_alpha_block ? 1 : 0;
    return 1;
}

// Batch form: n_blocks consecutive groups of pixels_per_block RGBA8 pixels, a fresh optimiser state per
// group is NOT needed (caching disabled), so one optimiser instance is reused like the reference's block loop does.
SHIM_API int ref_dxt1_optimize_batch(const uint8_t* pixels, uint32_t n_groups, uint32_t pixels_per_group, int quality, int perceptual,
                                     int use_alpha_blocks, int dxt1a, uint32_t alpha_threshold, int transparent_for_black,
                                     uint16_t* low, uint16_t* high, uint8_t* selectors, uint64_t* error, uint8_t* alpha_block)
{
    dxt1_endpoint_optimizer opt;
    for (uint32_t g = 0; g < n_groups; g++)
    {
        const color_quad_u8* px = reinterpret_cast<const color_quad_u8*>(pixels) + (size_t)g * pixels_per_group;
        dxt1_endpoint_optimizer::params p;
        dxt1_endpoint_optimizer::results r;
        bool have_alpha = false;
        if (dxt1a)
            for (uint32_t i = 0; i < pixels_per_group; i++)
                if (px[i].a < alpha_threshold) { have_alpha = true; break; }
        p.m_pPixels = px;
        p.m_num_pixels = pixels_per_group;
        p.m_quality = (crn_dxt_quality)quality;
        p.m_perceptual = perceptual != 0;
        p.m_pixels_have_alpha = have_alpha;
        p.m_use_alpha_blocks = use_alpha_blocks != 0;
        p.m_dxt1a_alpha_threshold = alpha_threshold;
        p.m_use_transparent_indices_for_black = transparent_for_black != 0;
        p.m_endpoint_caching = false;
        r.m_pSelectors = selectors + (size_t)g * pixels_per_group;
        r.m_error = 0;
        if (!opt.compute(p, r))
            return 0;
        low[g] = r.m_low_color;
        high[g] = r.m_high_color;
        error[g] = r.m_error;
        alpha_block[g] = r.m_alpha_block ? 1 : 0;
    }
    return 1;
}

SHIM_API int ref_dxt5_optimize_batch(const uint8_t* pixels, uint32_t n_groups, uint32_t pixels_per_group, uint32_t comp_index,
                                     int quality, int use_both_block_types,
                                     uint8_t* first, uint8_t* second, uint8_t* selectors, uint64_t* error, uint8_t* block_type)
{
    dxt5_endpoint_optimizer opt;
    for (uint32_t g = 0; g < n_groups; g++)
    {
        dxt5_endpoint_optimizer::params p;
        dxt5_endpoint_optimizer::results r;
        p.m_pPixels = reinterpret_cast<const color_quad_u8*>(pixels) + (size_t)g * pixels_per_group;
        p.m_num_pixels = pixels_per_group;
        p.m_comp_index = comp_index;
        p.m_quality = (crn_dxt_quality)quality;
        p.m_use_both_block_types = use_both_block_types != 0;
        r.m_pSelectors = selectors + (size_t)g * pixels_per_group;
        if (!opt.compute(p, r))
            return 0;
        first[g] = r.m_first_endpoint;
        second[g] = r.m_second_endpoint;
        error[g] = r.m_error;
        block_type[g] = r.m_block_type;
    }
    return 1;
}

// dxt_endpoint_refiner::refine (crnlib/crn_dxt_endpoint_refiner.cpp:36)
SHIM_API int ref_refine_endpoints(const uint8_t* pixels, const uint8_t* selectors, uint32_t num_pixels, int dxt1_selectors,
                                  int perceptual, uint32_t alpha_comp_index, uint64_t error_to_beat, int block_type,
                                  uint32_t* low, uint32_t* high)
{
    dxt_endpoint_refiner refiner;
    dxt_endpoint_refiner::params p;
    dxt_endpoint_refiner::results r;
    p.m_pPixels = reinterpret_cast<const color_quad_u8*>(pixels);
    p.m_pSelectors = selectors;
    p.m_num_pixels = num_pixels;
    p.m_dxt1_selectors = dxt1_selectors != 0;
    p.m_perceptual = perceptual != 0;
    p.m_alpha_comp_index = alpha_comp_index;
    p.m_error_to_beat = error_to_beat;
    p.m_block_index = 0;
    (void)block_type;
    bool ok = refiner.refine(p, r);
    *low = r.m_low_color;
    *high = r.m_high_color;
    return ok ? 1 : 0;
}

// Whole-image block-by-block packing through dxt_image::init with endpoint caching disabled
// (the only thread-count independent configuration).  fmt is crnlib::dxt_format.
// out must hold blocks_x*blocks_y*(8|16) bytes.  Returns bytes written or 0.
SHIM_API uint32_t ref_pack_image(int fmt, const uint8_t* rgba, uint32_t width, uint32_t height, int quality, int perceptual,
                                 int use_both_block_types, int transparent_for_black, uint32_t alpha_threshold,
                                 int compressor, uint32_t helper_threads, uint8_t* out)
{
    image_u8 img;
    img.alias(const_cast<color_quad_u8*>(reinterpret_cast<const color_quad_u8*>(rgba)), width, height);
    dxt_image::pack_params pp;
    pp.m_quality = (crn_dxt_quality)quality;
    pp.m_perceptual = perceptual != 0;
    pp.m_use_both_block_types = use_both_block_types != 0;
    pp.m_use_transparent_indices_for_black = transparent_for_black != 0;
    pp.m_dxt1a_alpha_threshold = alpha_threshold;
    pp.m_compressor = (crn_dxt_compressor_type)compressor;
    pp.m_num_helper_threads = helper_threads;
    pp.m_endpoint_caching = false;
    dxt_image d;
    if (!d.init((dxt_format)fmt, img, pp))
        return 0;
    uint32_t n = d.get_size_in_bytes();
    memcpy(out, d.get_element_ptr(), n);
    return n;
}

// crn_compress through the public API.  images[f*levels+l] -> RGBA8 of level l (max(1,w>>l) x max(1,h>>l)).
// Returns a malloc'd copy (free with ref_free) and its size; NULL on failure.
SHIM_API void* ref_compress(int file_type, int format, uint32_t width, uint32_t height, uint32_t faces, uint32_t levels,
                            const uint32_t* const* images, uint32_t flags, uint32_t quality_level, float target_bitrate,
                            int dxt_quality, uint32_t helper_threads, uint32_t alpha_component,
                            uint32_t* out_size, uint32_t* actual_quality, float* actual_bitrate, int want_bitrate)
{
    crn_comp_params cp;
    cp.m_file_type = (crn_file_type)file_type;
    cp.m_format = (crn_format)format;
    cp.m_width = width;
    cp.m_height = height;
    cp.m_faces = faces;
    cp.m_levels = levels;
    cp.m_flags = flags;
    cp.m_quality_level = quality_level;
    cp.m_target_bitrate = target_bitrate;
    cp.m_dxt_quality = (crn_dxt_quality)dxt_quality;
    cp.m_num_helper_threads = helper_threads;
    cp.m_alpha_component = alpha_component;
    for (uint32_t f = 0; f < faces; f++)
        for (uint32_t l = 0; l < levels; l++)
            cp.m_pImages[f][l] = images[f * levels + l];
    crn_uint32 size = 0, aq = 0;
    float ab = 0.0f;
    void* p = crn_compress(cp, size, &aq, want_bitrate ? &ab : NULL);
    *out_size = 0;
    if (!p)
        return NULL;
    void* q = malloc(size);
    memcpy(q, p, size);
    crn_free_block(p);
    *out_size = size;
    if (actual_quality) *actual_quality = aq;
    if (actual_bitrate) *actual_bitrate = ab;
    return q;
}

// crn_compress with a crn_mipmap_params (inc/crnlib.h:614): level 0 in, the reference generates the chain (defaults of
// crn_mipmap_params::clear(): kaiser, gamma filtering 2.2, blurriness 0.9, down to 1x1) and compresses it.
SHIM_API void* ref_compress_mip_chain(int file_type, int format, uint32_t width, uint32_t height, const uint32_t* level0, uint32_t flags,
                                      uint32_t quality_level, int dxt_quality, uint32_t helper_threads, uint32_t* out_size)
{
    crn_comp_params cp;
    cp.m_file_type = (crn_file_type)file_type;
    cp.m_format = (crn_format)format;
    cp.m_width = width; cp.m_height = height; cp.m_faces = 1; cp.m_levels = 1;
    cp.m_flags = flags;
    cp.m_quality_level = quality_level;
    cp.m_dxt_quality = (crn_dxt_quality)dxt_quality;
    cp.m_num_helper_threads = helper_threads;
    cp.m_pImages[0][0] = level0;
    crn_mipmap_params mp;
    crn_uint32 size = 0;
    void* p = crn_compress(cp, mp, size, NULL, NULL);
    *out_size = 0;
    if (!p) return NULL;
    void* q = malloc(size);
    memcpy(q, p, size);
    crn_free_block(p);
    *out_size = size;
    return q;
}

// crn_compress(comp_params, mipmap_params) with a caller-built crn_mipmap_params (same layout as inc/crnlib.h:471-574): the crop / clamp /
// rescale / renormalise options of create_texture_mipmaps
SHIM_API void* ref_compress_mip_params(int file_type, int format, uint32_t width, uint32_t height, uint32_t levels, const uint32_t* const* images, uint32_t flags,
                                       uint32_t quality_level, uint32_t helper_threads, const void* mipmap_params, uint32_t* out_size)
{
    crn_comp_params cp;
    cp.m_file_type = (crn_file_type)file_type;
    cp.m_format = (crn_format)format;
    cp.m_width = width; cp.m_height = height; cp.m_faces = 1; cp.m_levels = levels;
    cp.m_flags = flags;
    cp.m_quality_level = quality_level;
    cp.m_num_helper_threads = helper_threads;
    for (uint32_t l = 0; l < levels; l++) cp.m_pImages[0][l] = images[l];
    crn_mipmap_params mp;
    memcpy(&mp, mipmap_params, sizeof(mp));
    mp.m_size_of_obj = sizeof(mp);
    crn_uint32 size = 0;
    void* p = crn_compress(cp, mp, size, NULL, NULL);
    *out_size = 0;
    if (!p) return NULL;
    void* q = malloc(size);
    memcpy(q, p, size);
    crn_free_block(p);
    *out_size = size;
    return q;
}

SHIM_API void ref_free(void* p) { free(p); }

SHIM_API void* ref_crn_to_dds(const void* crn, uint32_t crn_size, uint32_t* dds_size)
{
    crn_uint32 sz = crn_size;
    void* p = crn_decompress_crn_to_dds(crn, sz);
    *dds_size = 0;
    if (!p)
        return NULL;
    void* q = malloc(sz);
    memcpy(q, p, sz);
    crn_free_block(p);
    *dds_size = sz;
    return q;
}

// Texture info: out[0..7] = width,height,levels,faces,bytes_per_block,format,userdata0,userdata1
SHIM_API int ref_crn_info(const void* crn, uint32_t crn_size, uint32_t* out)
{
    crnd::crn_texture_info ti;
    if (!crnd::crnd_get_texture_info(crn, crn_size, &ti))
        return 0;
    out[0] = ti.m_width; out[1] = ti.m_height; out[2] = ti.m_levels; out[3] = ti.m_faces;
    out[4] = ti.m_bytes_per_block; out[5] = (uint32_t)ti.m_format; out[6] = ti.m_userdata0; out[7] = ti.m_userdata1;
    return 1;
}

SHIM_API void* ref_transcode_begin(const void* crn, uint32_t crn_size) { return crnd::crnd_unpack_begin(crn, crn_size); }
SHIM_API int ref_transcode_level(void* ctx, void** face_ptrs, uint32_t dst_size, uint32_t row_pitch, uint32_t level)
{
    return crnd::crnd_unpack_level(ctx, face_ptrs, dst_size, row_pitch, level) ? 1 : 0;
}
SHIM_API void ref_transcode_end(void* ctx) { crnd::crnd_unpack_end(ctx); }

// Timed whole-file transcode: unpacks every level `repeats` times into caller memory laid out
// level-major / face-major, tightly pitched.  Returns seconds spent inside crnd_unpack_level only.
SHIM_API double ref_transcode_all(const void* crn, uint32_t crn_size, uint8_t* out, uint64_t out_capacity, uint32_t repeats)
{
    crnd::crn_texture_info ti;
    if (!crnd::crnd_get_texture_info(crn, crn_size, &ti))
        return -1.0;
    void* ctx = crnd::crnd_unpack_begin(crn, crn_size);
    if (!ctx)
        return -1.0;
    double total = 0.0;
    for (uint32_t rep = 0; rep < repeats; rep++)
    {
        uint64_t ofs = 0;
        for (uint32_t l = 0; l < ti.m_levels; l++)
        {
            uint32_t w = ti.m_width >> l; if (!w) w = 1;
            uint32_t h = ti.m_height >> l; if (!h) h = 1;
            uint32_t bx = (w + 3) >> 2, by = (h + 3) >> 2;
            uint32_t pitch = bx * ti.m_bytes_per_block;
            uint32_t face_size = pitch * by;
            void* faces[6];
            for (uint32_t f = 0; f < ti.m_faces; f++)
            {
                if (ofs + face_size > out_capacity) { crnd::crnd_unpack_end(ctx); return -2.0; }
                faces[f] = out + ofs;
                ofs += face_size;
            }
            struct timespec t0, t1;
            clock_gettime(CLOCK_MONOTONIC, &t0);
            bool ok = crnd::crnd_unpack_level(ctx, faces, face_size, pitch, l);
            clock_gettime(CLOCK_MONOTONIC, &t1);
            if (!ok) { crnd::crnd_unpack_end(ctx); return -3.0; }
            total += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
        }
    }
    crnd::crnd_unpack_end(ctx);
    return total;
}

// dxt_fast primitives used by the clustered-DDS quantisers (crnlib/crn_dxt_fast.cpp:725, :788, :855)
SHIM_API void ref_fast_color_block(uint32_t n, const uint8_t* pixels, uint32_t* low16, uint32_t* high16, uint8_t* selectors)
{
    uint lo, hi;
    dxt_fast::compress_color_block(n, reinterpret_cast<const color_quad_u8*>(pixels), lo, hi, selectors);
    *low16 = lo; *high16 = hi;
}
SHIM_API void ref_fast_alpha_block(uint32_t n, const uint8_t* pixels, uint32_t comp, uint32_t* low8, uint32_t* high8, uint8_t* selectors)
{
    uint lo, hi;
    dxt_fast::compress_alpha_block(n, reinterpret_cast<const color_quad_u8*>(pixels), lo, hi, selectors, comp);
    *low8 = lo; *high8 = hi;
}
SHIM_API void ref_find_representative_colors(uint32_t n, const uint8_t* pixels, uint8_t* lo, uint8_t* hi)
{
    color_quad_u8 l, h;
    dxt_fast::find_representative_colors(n, reinterpret_cast<const color_quad_u8*>(pixels), l, h);
    lo[0] = l.r; lo[1] = l.g; lo[2] = l.b; hi[0] = h.r; hi[1] = h.g; hi[2] = h.b;
}

// clusterizer<vecNF>::generate_codebook + retrieve_clusters (crnlib/crn_clusterizer.h:65, :301) and
// threaded_clusterizer<shim_vec16F>::create_clusters (crnlib/crn_threaded_clusterizer.h:70).
#include "crn_clusterizer.h"
#include "crn_threaded_clusterizer.h"
typedef vec<6, float> shim_vec6F;
typedef vec<16, float> shim_vec16F;
template <typename V>
static int run_clusterizer(uint32_t n, const float* vecs, const uint32_t* weights, uint32_t max_size, uint32_t retrieve,
                           uint32_t* cluster_of, uint32_t* num_clusters, uint32_t* codebook_size)
{
    clusterizer<V> c;
    c.reserve_training_vecs(n);
    for (uint32_t i = 0; i < n; i++) {
        V v;
        for (uint32_t d = 0; d < V::num_elements; d++) v[d] = vecs[(size_t)i * V::num_elements + d];
        c.add_training_vec(v, weights[i]);
    }
    if (!c.generate_codebook(max_size)) return 0;
    *codebook_size = c.get_codebook_size();
    crnlib::vector<crnlib::vector<uint> > clusters;
    c.retrieve_clusters(retrieve ? retrieve : c.get_codebook_size(), clusters);
    *num_clusters = clusters.size();
    for (uint32_t k = 0; k < clusters.size(); k++)
        for (uint32_t j = 0; j < clusters[k].size(); j++) cluster_of[clusters[k][j]] = k;
    return 1;
}
SHIM_API int ref_clusterizer(uint32_t dims, uint32_t n, const float* vecs, const uint32_t* weights, uint32_t max_size, uint32_t retrieve,
                             uint32_t* cluster_of, uint32_t* num_clusters, uint32_t* codebook_size)
{
    if (dims == 2) return run_clusterizer<vec2F>(n, vecs, weights, max_size, retrieve, cluster_of, num_clusters, codebook_size);
    if (dims == 6) return run_clusterizer<shim_vec6F>(n, vecs, weights, max_size, retrieve, cluster_of, num_clusters, codebook_size);
    if (dims == 16) return run_clusterizer<shim_vec16F>(n, vecs, weights, max_size, retrieve, cluster_of, num_clusters, codebook_size);
    return 0;
}
SHIM_API int ref_threaded_clusterizer16(uint32_t n, const float* vecs, const uint32_t* weights, uint32_t max_clusters, uint32_t threads,
                                        uint32_t* cluster_of, uint32_t* num_clusters)
{
    task_pool pool;
    if (!pool.init(threads)) return 0;
    threaded_clusterizer<shim_vec16F> tc(pool);
    threaded_clusterizer<shim_vec16F>::weighted_vec_array wv(n);
    for (uint32_t i = 0; i < n; i++) {
        for (uint32_t d = 0; d < 16; d++) wv[i].m_vec[d] = vecs[(size_t)i * 16 + d];
        wv[i].m_weight = weights[i];
    }
    crnlib::vector<crnlib::vector<uint> > clusters;
    if (!tc.create_clusters(wv, max_clusters, clusters, NULL, NULL)) return 0;
    *num_clusters = clusters.size();
    for (uint32_t k = 0; k < clusters.size(); k++)
        for (uint32_t j = 0; j < clusters[k].size(); j++) cluster_of[clusters[k][j]] = k;
    return 1;
}

// ref_refine -> crnlib::dxt_endpoint_refiner::refine (crnlib/crn_dxt_endpoint_refiner.cpp:36), one call per cluster.
// offsets: n_clusters + 1 CSR offsets into pixels / selectors.  ok[c] = refine()'s return value.
SHIM_API void ref_refine(int dxt1_selectors, int perceptual, uint32_t comp, const uint8_t* pixels, const uint8_t* selectors,
                         const uint32_t* offsets, uint32_t n_clusters, const uint64_t* error_to_beat,
                         uint32_t* low, uint32_t* high, uint64_t* error, uint8_t* ok)
{
    dxt_endpoint_refiner refiner;
    for (uint32_t c = 0; c < n_clusters; c++)
    {
        dxt_endpoint_refiner::params p;
        dxt_endpoint_refiner::results r;
        p.m_pPixels = reinterpret_cast<const color_quad_u8*>(pixels) + offsets[c];
        p.m_num_pixels = offsets[c + 1] - offsets[c];
        p.m_pSelectors = selectors + offsets[c];
        p.m_alpha_comp_index = comp;
        p.m_error_to_beat = error_to_beat ? error_to_beat[c] : cUINT64_MAX;
        p.m_dxt1_selectors = dxt1_selectors != 0;
        p.m_perceptual = perceptual != 0;
        r.m_low_color = 0; r.m_high_color = 0; r.m_error = cUINT64_MAX;
        ok[c] = refiner.refine(p, r) ? 1 : 0;
        low[c] = r.m_low_color; high[c] = r.m_high_color; error[c] = r.m_error;
    }
}

// ---- dxt_hc private tasks, for pinning the oracle port (tests only) ---------------------------------------------------
// The functions below call UNMODIFIED private member tasks of crnlib::dxt_hc on a hand-filled object.  Access control is
// lifted for this translation unit only (the class layout does not depend on it); every header dxt_hc.h pulls in is
// included normally first so that nothing but dxt_hc itself is affected.
#include "crn_tree_clusterizer.h"
#include "crn_threading.h"
#include "crn_dxt_hc_common.h"
#include "crn_dxt_endpoint_refiner.h"
#include <queue>
#define private public
#include "crn_dxt_hc.h"
#undef private

namespace {
struct shim_color_selector_details { shim_color_selector_details() { memset(this, 0, sizeof(*this)); } uint error[16][4]; bool used; };   // layout of crn_dxt_hc.cpp:1296-1304
struct shim_alpha_selector_details { shim_alpha_selector_details() { memset(this, 0, sizeof(*this)); } uint error[16][8]; bool used; };   // crn_dxt_hc.cpp:1506-1514
}

// ref_hc_assign_selectors -> dxt_hc::create_color_selector_codebook_task (kind 0, crn_dxt_hc.cpp:1306) /
// create_alpha_selector_codebook_task (kind 1, :1516) over all blocks with one task.  Block b is matched against its own
// palette: values = 4 RGBA8 colours (kind 0) or 8 alpha values (kind 1); values_accum optional (kind 1: refined values).
// n <= 65535 (cluster indices are uint16).  errors: K x 16 x (4|8) uint32 out.
SHIM_API int ref_hc_assign_selectors(int kind, int perceptual, uint32_t comp, const uint8_t* blocks, uint32_t n, const uint8_t* values, const uint8_t* values_accum,
                                     const uint64_t* codebook, uint32_t K, uint32_t* best_index, uint32_t* errors, uint8_t* used)
{
    if (n > 65535) return 0;
    task_pool tp;
    if (!tp.init(0)) return 0;
    dxt_hc hc;
    hc.m_pTask_pool = &tp;
    hc.m_num_blocks = n;
    hc.m_has_subblocks = false;
    hc.m_blocks = (color_quad_u8(*)[16])blocks;
    hc.m_params.m_perceptual = perceptual != 0;
    hc.m_endpoint_indices.resize(n);
    hc.m_selector_indices.resize(n);
    if (!kind)
    {
        hc.m_color_clusters.resize(n);
        for (uint32_t b = 0; b < n; b++)
        {
            hc.m_endpoint_indices[b].color = (uint16)b;
            for (uint s = 0; s < 4; s++)
                hc.m_color_clusters[b].color_values[s] = color_quad_u8(values[b * 16 + s * 4], values[b * 16 + s * 4 + 1], values[b * 16 + s * 4 + 2], values[b * 16 + s * 4 + 3]);
        }
        hc.m_color_selectors.resize(K);
        for (uint32_t i = 0; i < K; i++) hc.m_color_selectors[i] = (uint32)codebook[i];
        crnlib::vector<shim_color_selector_details> details(K);
        hc.create_color_selector_codebook_task(0, &details);
        for (uint32_t i = 0; i < K; i++) { memcpy(errors + (size_t)i * 64, details[i].error, 256); used[i] = details[i].used; }
        for (uint32_t b = 0; b < n; b++) best_index[b] = hc.m_selector_indices[b].color;
    }
    else
    {
        hc.m_num_alpha_blocks = 1;
        hc.m_params.m_alpha_component_indices[0] = comp;
        hc.m_alpha_clusters.resize(n);
        for (uint32_t b = 0; b < n; b++)
        {
            hc.m_endpoint_indices[b].component[1] = (uint16)b;
            hc.m_alpha_clusters[b].refined_alpha = values_accum != NULL;
            for (uint s = 0; s < 8; s++)
            {
                hc.m_alpha_clusters[b].alpha_values[s] = values[b * 8 + s];
                hc.m_alpha_clusters[b].refined_alpha_values[s] = values_accum ? values_accum[b * 8 + s] : 0;
            }
        }
        hc.m_alpha_selectors.resize(K);
        for (uint32_t i = 0; i < K; i++) hc.m_alpha_selectors[i] = codebook[i];
        crnlib::vector<shim_alpha_selector_details> details(K);
        hc.create_alpha_selector_codebook_task(0, &details);
        for (uint32_t i = 0; i < K; i++) { memcpy(errors + (size_t)i * 128, details[i].error, 512); used[i] = details[i].used; }
        for (uint32_t b = 0; b < n; b++) best_index[b] = hc.m_selector_indices[b].component[1];
    }
    hc.m_pTask_pool = NULL;
    return 1;
}

// ref_hc_nearest_codebook -> tree_clusterizer<vec6F / vec2F>::generate_codebook over the n training vectors (weights 1..),
// then dxt_hc::determine_color_endpoint_clusters_task (dims 6, crn_dxt_hc.cpp:836) / determine_alpha_endpoint_clusters_task
// (dims 2, :1132) with one tile per vector.  Returns the codebook size; codebook_out: up to max_size x dims floats.
SHIM_API uint32_t ref_hc_nearest_codebook(uint32_t dims, const float* vecs, const uint32_t* weights, uint32_t n, uint32_t max_size, float* codebook_out, uint32_t* index_out)
{
    task_pool tp;
    if (!tp.init(0)) return 0;
    dxt_hc hc;
    hc.m_pTask_pool = &tp;
    hc.m_tiles.resize(n);
    crnlib::vector<uint> w(n);
    for (uint32_t i = 0; i < n; i++) { w[i] = weights[i]; hc.m_tiles[i].pixels.resize(1); }
    uint32_t k = 0;
    if (dims == 6)
    {
        typedef vec<6, float> vec6F; crnlib::vector<vec6F> v(n);
        for (uint32_t i = 0; i < n; i++) { for (uint d = 0; d < 6; d++) v[i][d] = vecs[i * 6 + d]; hc.m_tiles[i].color_endpoint = v[i]; }
        tree_clusterizer<vec6F> vq;
        vq.generate_codebook(v.get_ptr(), w.get_ptr(), n, max_size, true, &tp);
        k = vq.get_codebook_size();
        for (uint32_t i = 0; i < k; i++) for (uint d = 0; d < 6; d++) codebook_out[i * 6 + d] = vq.get_codebook_entry(i)[d];
        hc.determine_color_endpoint_clusters_task(0, &vq);
        for (uint32_t i = 0; i < n; i++) index_out[i] = hc.m_tiles[i].cluster_indices[0];
    }
    else
    {
        hc.m_num_alpha_blocks = 1;
        typedef vec<2, float> vec2F; crnlib::vector<vec2F> v(n);
        for (uint32_t i = 0; i < n; i++) { for (uint d = 0; d < 2; d++) v[i][d] = vecs[i * 2 + d]; hc.m_tiles[i].alpha_endpoints[0] = v[i]; }
        tree_clusterizer<vec2F> vq;
        vq.generate_codebook(v.get_ptr(), w.get_ptr(), n, max_size, false, &tp);
        k = vq.get_codebook_size();
        for (uint32_t i = 0; i < k; i++) for (uint d = 0; d < 2; d++) codebook_out[i * 2 + d] = vq.get_codebook_entry(i)[d];
        hc.determine_alpha_endpoint_clusters_task(0, &vq);
        for (uint32_t i = 0; i < n; i++) index_out[i] = hc.m_tiles[i].cluster_indices[1];
    }
    hc.m_pTask_pool = NULL;
    return k;
}

// ref_hc_compress -> dxt_hc::compress (crn_dxt_hc.cpp:98-312) with `threads` helper threads (0 = the single-task tree
// quantiser: no "alternative" sub-trees, the deterministic configuration the device pipeline is compared against).
// levels: num_levels x {first_block, num_blocks, block_width, weight-as-float-bits}.  Outputs mirror crn_gpu_hc_*:
// endpoint_indices / selector_indices n x 4 uint16 (color, alpha0, alpha1, reference | 0); palettes up to 65536
// entries each; sizes[4] = color endpoints, alpha endpoints, color selectors, alpha selectors.  encodings / tile_indices
// (n each) optional.  blocks are copied (dxt_hc only reads them for the DXT formats).
SHIM_API int ref_hc_compress(int format, uint32_t n, uint32_t num_levels, uint32_t num_faces, const uint32_t* levels, int perceptual,
                             const uint32_t* codebook_sizes, const float* deratings, const uint32_t* alpha_comps, uint32_t threads,
                             const uint8_t* blocks, uint16_t* endpoint_indices, uint16_t* selector_indices,
                             uint32_t* color_endpoints, uint32_t* alpha_endpoints, uint32_t* color_selectors, uint64_t* alpha_selectors,
                             uint32_t* sizes, uint8_t* encodings, uint32_t* tile_indices)
{
    task_pool tp;
    if (!tp.init(threads)) return 0;
    dxt_hc::params p;
    p.m_num_blocks = n; p.m_num_levels = num_levels; p.m_num_faces = num_faces;
    for (uint32_t l = 0; l < num_levels; l++)
    {
        p.m_levels[l].m_first_block = levels[l * 4]; p.m_levels[l].m_num_blocks = levels[l * 4 + 1]; p.m_levels[l].m_block_width = levels[l * 4 + 2];
        memcpy(&p.m_levels[l].m_weight, &levels[l * 4 + 3], 4);
    }
    p.m_format = (dxt_format)format;
    p.m_perceptual = perceptual != 0;
    p.m_hierarchical = true;
    p.m_color_endpoint_codebook_size = codebook_sizes[0]; p.m_color_selector_codebook_size = codebook_sizes[1];
    p.m_alpha_endpoint_codebook_size = codebook_sizes[2]; p.m_alpha_selector_codebook_size = codebook_sizes[3];
    p.m_adaptive_tile_color_psnr_derating = deratings[0]; p.m_adaptive_tile_alpha_psnr_derating = deratings[1];
    p.m_adaptive_tile_color_alpha_weighting_ratio = deratings[2];
    p.m_alpha_component_indices[0] = alpha_comps[0]; p.m_alpha_component_indices[1] = alpha_comps[1];
    p.m_pTask_pool = &tp;
    crnlib::vector<color_quad_u8> copy(n * 16);
    memcpy(copy.get_ptr(), blocks, (size_t)n * 64);
    crnlib::vector<dxt_hc::endpoint_indices_details> ei;
    crnlib::vector<dxt_hc::selector_indices_details> si;
    crnlib::vector<uint32> ce, ae, cs;
    crnlib::vector<uint64> as;
    dxt_hc hc;
    if (!hc.compress((color_quad_u8(*)[16])copy.get_ptr(), ei, si, ce, ae, cs, as, p)) return 0;
    for (uint32_t b = 0; b < n; b++)
    {
        for (uint c = 0; c < 3; c++) { endpoint_indices[b * 4 + c] = ei[b].component[c]; selector_indices[b * 4 + c] = si[b].component[c]; }
        endpoint_indices[b * 4 + 3] = ei[b].reference; selector_indices[b * 4 + 3] = 0;
        if (encodings) encodings[b] = hc.m_block_encodings[b];
        if (tile_indices) tile_indices[b] = hc.m_tile_indices[b];
    }
    sizes[0] = ce.size(); sizes[1] = ae.size(); sizes[2] = cs.size(); sizes[3] = as.size();
    memcpy(color_endpoints, ce.get_ptr(), ce.size() * 4); memcpy(alpha_endpoints, ae.get_ptr(), ae.size() * 4);
    memcpy(color_selectors, cs.get_ptr(), cs.size() * 4); memcpy(alpha_selectors, as.get_ptr(), as.size() * 8);
    return 1;
}

// ref_unpack_image -> dxt_image::init(fmt, w, h) + the caller's elements + dxt_image::unpack (crn_dxt_image.cpp:495-567).
SHIM_API int ref_unpack_image(int fmt, const uint8_t* blocks, uint32_t width, uint32_t height, uint8_t* rgba_out)
{
    dxt_image di;
    if (!di.init((dxt_format)fmt, width, height, false)) return 0;
    memcpy(di.get_element_ptr(), blocks, (size_t)di.get_total_elements() * sizeof(dxt_image::element));
    image_u8 img;
    if (!di.unpack(img)) return 0;
    for (uint32_t y = 0; y < height; y++) memcpy(rgba_out + (size_t)y * width * 4, img.get_scanline(y), (size_t)width * 4);
    return 1;
}

// ref_resample -> image_utils::resample (crn_image_utils.cpp:876-886): multithreaded != 0 takes the task-pool form
// (threaded_resampler), 0 the single-thread Resampler.  RGBA8 in and out.
#include "crn_image_utils.h"
SHIM_API int ref_resample(const uint8_t* src, uint32_t sw, uint32_t sh, uint8_t* dst, uint32_t dw, uint32_t dh, const char* filter, float filter_scale,
                          int srgb, float gamma, int wrapping, uint32_t num_comps, int multithreaded)
{
    image_u8 s(sw, sh), d;
    for (uint32_t y = 0; y < sh; y++) memcpy(s.get_scanline(y), src + (size_t)y * sw * 4, (size_t)sw * 4);
    image_utils::resample_params rp;
    rp.m_dst_width = dw; rp.m_dst_height = dh; rp.m_pFilter = filter; rp.m_filter_scale = filter_scale; rp.m_srgb = srgb != 0;
    rp.m_wrapping = wrapping != 0; rp.m_first_comp = 0; rp.m_num_comps = num_comps; rp.m_source_gamma = gamma; rp.m_multithreaded = multithreaded != 0;
    if (!image_utils::resample(s, d, rp)) return 0;
    for (uint32_t y = 0; y < dh; y++) memcpy(dst + (size_t)y * dw * 4, d.get_scanline(y), (size_t)dw * 4);
    return 1;
}

// ref_dds_to_images -> what crn_decompress_dds_to_images (crnlib/crnlib.cpp:293-333) does in a build with asserts enabled: read_dds, then
// per level mip_level::get_unpacked_image(tmp, cUnpackFlagUncook) -- the call mip_level::unpack_from_dxt makes INSIDE CRNLIB_ASSERT
// (crn_mipmapped_texture.cpp:188-197), which a release (NDEBUG) build such as this one compiles out, so the public function itself returns
// NULL images here.  out: tight RGBA8 images, index level + levels * face, one after the other; desc: faces, width, height, levels, fourcc.
#include "crn_mipmapped_texture.h"
#include "crn_buffer_stream.h"
SHIM_API int ref_dds_to_images(const uint8_t* dds, uint32_t dds_size, uint8_t* out, uint64_t out_capacity, uint32_t* desc)
{
    mipmapped_texture tex;
    buffer_stream in_stream(dds, dds_size);
    data_stream_serializer in_serializer(in_stream);
    if (!tex.read_dds(in_serializer)) return 0;
    uint64_t ofs = 0;
    pixel_format fmt = tex.get_format();
    for (uint f = 0; f < tex.get_num_faces(); f++)
        for (uint l = 0; l < tex.get_num_levels(); l++) {
            mip_level* pLevel = tex.get_level(f, l);
            image_u8 tmp;
            image_u8* pImg = pLevel->get_unpacked_image(tmp, cUnpackFlagUncook);
            if (!pImg) return 0;
            const uint64_t bytes = (uint64_t)pImg->get_width() * pImg->get_height() * 4;
            if (ofs + bytes > out_capacity) return 0;
            for (uint y = 0; y < pImg->get_height(); y++) memcpy(out + ofs + (size_t)y * pImg->get_width() * 4, pImg->get_scanline(y), (size_t)pImg->get_width() * 4);
            ofs += bytes;
            if (!f && !l && tex.is_packed()) {                                  // mip_level::assign(image, PIXEL_FMT_INVALID) (:99-120)
                if (pImg->is_grayscale()) fmt = pImg->is_component_valid(3) ? PIXEL_FMT_A8L8 : PIXEL_FMT_L8;
                else fmt = pImg->is_component_valid(3) ? PIXEL_FMT_A8R8G8B8 : PIXEL_FMT_R8G8B8;
            }
        }
    desc[0] = tex.get_num_faces(); desc[1] = tex.get_width(); desc[2] = tex.get_height(); desc[3] = tex.get_num_levels(); desc[4] = (uint32_t)fmt;
    return 1;
}

// ref_convert_image -> image_utils::convert_image (crn_image_utils.cpp:1181-1380), conversion_type as the enum's integer value.
SHIM_API int ref_convert_image(uint8_t* rgba, uint32_t w, uint32_t h, int conv_type)
{
    image_u8 img(w, h);
    for (uint32_t y = 0; y < h; y++) memcpy(img.get_scanline(y), rgba + (size_t)y * w * 4, (size_t)w * 4);
    image_utils::convert_image(img, (image_utils::conversion_type)conv_type);
    for (uint32_t y = 0; y < h; y++) memcpy(rgba + (size_t)y * w * 4, img.get_scanline(y), (size_t)w * 4);
    return 1;
}
