// crnlib_dropin.cpp -- the reference's PUBLIC C++ API over the B200 library: libcrnlib_b200.so exports the same (mangled)
// symbols as the reference's libcrn, so a program built against inc/crnlib.h / inc/crn_defs.h links against it unchanged.
//
// Compiled with -I<reference>/inc: the reference's headers are included where they lie (never copied), which is what keeps
// crn_comp_params / crn_mipmap_params / crnd::crn_texture_info binary-identical.  Every call goes DOWN through the C ABI of
// include/crn_b200.h (plain C++ here, no CUDA): SURVEY.md section 8(b).
//
//   reference entry (file:line)                                     -> C ABI underneath
//   crn_compress (inc/crnlib.h:609, crnlib/crnlib.cpp:215-240)        crn_gpu_compress_crn / crn_gpu_compress_dds
//   crn_compress + crn_mipmap_params (:614, crnlib.cpp:242-267)       crn_gpu_compress_mip_chain (generate mode) / the call above
//   crn_decompress_crn_to_dds (:620, crnlib.cpp:269-291)              crn_gpu_crn_to_dds
//   crn_decompress_dds_to_images (:634, crnlib.cpp:293-333)           crn_gpu_dds_to_images
//   crn_free_block / crn_free_all_images / crn_set_memory_callbacks   the allocator below (crnlib/crn_mem.cpp:164-345 contract)
//   crn_create_block_compressor / crn_compress_block / ...            crn_gpu_pack_image_host / crn_gpu_unpack_image_host on one 4x4 block
//   crnd::crnd_unpack_begin / _level / _end (inc/crn_defs.h:139-221)  crn_gpu_crnd_unpack_begin / _unpack_level_host / _unpack_end
//   crnd::crnd_get_texture_info / _get_level_info / _validate_file /  host-only header arithmetic (inc/crn_decomp.h:2657-2830)
//         _get_level_data / _get_data / segmented-file helpers
//
// Contract kept: outputs zeroed first and NULL / false on any failure (crnlib.cpp:217-225, :272-278), crn_comp_params::check(),
// returned blocks come from the (replaceable) library allocator and are released with crn_free_block, the progress callback runs
// on the calling thread and cancels the call when it returns false, calls are re-entrant from different host threads (one GPU
// context per calling thread).  There is no CPU fallback: without the CUDA library / a device every call fails.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "crnlib.h"
#include "crn_defs.h"                        // the crnd:: API declarations + header structs (the bodies of inc/crn_decomp.h are what this file replaces)
#include "../../include/crn_b200.h"
#include <malloc.h>
#include <mutex>
#include <new>
#include <vector>

namespace {

// ---- allocator (crn_set_memory_callbacks; realloc-style + msize, crnlib/crn_mem.cpp:164-345) ---------------------------
void* default_realloc(void* p, size_t size, size_t* actual, bool movable, void*)
{
    void* r = nullptr;
    if (!p) { r = malloc(size); if (actual) *actual = r ? malloc_usable_size(r) : 0; }
    else if (!size) { free(p); if (actual) *actual = 0; }
    else {
        if (movable) r = realloc(p, size);
        else if (malloc_usable_size(p) >= size) r = p;          // in place only
        if (actual) *actual = malloc_usable_size(r ? r : p);
    }
    return r;
}
size_t default_msize(void* p, void*) { return p ? malloc_usable_size(p) : 0; }

std::mutex g_mem_mutex;
crn_realloc_func g_realloc = default_realloc;
crn_msize_func g_msize = default_msize;
void* g_mem_user = nullptr;

void* lib_alloc(size_t size)
{
    size_t actual = 0;
    return g_realloc(nullptr, size ? size : 1, &actual, true, g_mem_user);
}
void lib_free(void* p)
{
    if (p) { size_t actual = 0; g_realloc(p, 0, &actual, true, g_mem_user); }
}
// moves a malloc'ed result of the C ABI into a block of the library allocator (so crn_free_block / user allocators own it)
void* adopt(void* file, uint32_t size)
{
    void* p = lib_alloc(size);
    if (p) memcpy(p, file, size);
    crn_gpu_free_file(file);
    return p;
}

// ---- one GPU context per calling thread (the reference's calls are re-entrant across threads; a crn_gpu_ctx is not) -----
struct ThreadCtx {
    crn_gpu_ctx* ctx = nullptr;
    ~ThreadCtx() { if (ctx) crn_gpu_destroy(ctx); }
};
crn_gpu_ctx* gpu()
{
    static thread_local ThreadCtx t;
    if (!t.ctx) {
        const char* e = getenv("CRN_B200_DEVICE");
        if (crn_gpu_create(e && *e ? atoi(e) : 0, &t.ctx) != CRN_GPU_OK) t.ctx = nullptr;
    }
    return t.ctx;
}

struct ProgressScope {                       // installs crn_comp_params' callback for the duration of one call
    crn_gpu_ctx* ctx;
    const crn_comp_params* p;
    static int thunk(uint32_t a, uint32_t b, uint32_t c, uint32_t d, void* user)
    {
        const crn_comp_params* p = static_cast<const crn_comp_params*>(user);
        return p->m_pProgress_func(a, b, c, d, p->m_pProgress_func_data) ? 1 : 0;
    }
    ProgressScope(crn_gpu_ctx* c, const crn_comp_params& params) : ctx(c), p(&params)
    {
        if (p->m_pProgress_func) crn_gpu_set_progress(ctx, thunk, const_cast<crn_comp_params*>(p));
    }
    ~ProgressScope() { crn_gpu_set_progress(ctx, nullptr, nullptr); }
};

bool is_non_srgb(crn_format f)               // pixel_format_helpers::is_crn_format_non_srgb (crnlib/crn_pixel_format.h)
{
    switch (f) {
    case cCRNFmtDXN_XY: case cCRNFmtDXN_YX: case cCRNFmtDXT5A: case cCRNFmtDXT5_CCxY: case cCRNFmtDXT5_xGxR: case cCRNFmtDXT5_xGBR: case cCRNFmtDXT5_AGBR:
        return true;
    default: return false;
    }
}

bool images_present(const crn_comp_params& p, std::vector<const void*>& flat)
{
    flat.clear();
    for (crn_uint32 f = 0; f < p.m_faces; f++)
        for (crn_uint32 l = 0; l < p.m_levels; l++) {
            if (!p.m_pImages[f][l]) return false;       // create_dds_tex / crn_comp::compress_init refuse missing images
            flat.push_back(p.m_pImages[f][l]);
        }
    return true;
}

void fill_crn_params(const crn_comp_params& p, crn_gpu_crn_params& cp)
{
    crn_gpu_default_crn_params(&cp);
    cp.crn_format = (uint32_t)p.m_format; cp.width = p.m_width; cp.height = p.m_height; cp.levels = p.m_levels; cp.faces = p.m_faces;
    cp.quality_level = p.m_quality_level;
    cp.perceptual = (p.m_flags & cCRNCompFlagPerceptual) && !is_non_srgb(p.m_format);      // create_compressed_texture, crn_texture_comp.cpp:52-60
    cp.alpha_component = p.m_alpha_component;
    cp.userdata0 = p.m_userdata0; cp.userdata1 = p.m_userdata1;
    if (p.m_flags & cCRNCompFlagManualPaletteSizes) {
        cp.palette_sizes[0] = p.m_crn_color_endpoint_palette_size; cp.palette_sizes[1] = p.m_crn_color_selector_palette_size;
        cp.palette_sizes[2] = p.m_crn_alpha_endpoint_palette_size; cp.palette_sizes[3] = p.m_crn_alpha_selector_palette_size;
    }
    cp.adaptive_tile_color_psnr_derating = p.m_crn_adaptive_tile_color_psnr_derating;
    cp.adaptive_tile_alpha_psnr_derating = p.m_crn_adaptive_tile_alpha_psnr_derating;
    cp.target_bitrate = p.m_target_bitrate;
}

void fill_dds_params(const crn_comp_params& p, crn_gpu_dds_params& dp)
{
    crn_gpu_default_dds_params(&dp);
    dp.crn_format = (uint32_t)p.m_format; dp.width = p.m_width; dp.height = p.m_height; dp.levels = p.m_levels; dp.faces = p.m_faces;
    dp.quality_level = p.m_quality_level;
    dp.dxt1a_for_transparency = (p.m_flags & cCRNCompFlagDXT1AForTransparency) != 0;
    // dxt_image::pack_params::init(const crn_comp_params&) (crnlib/crn_dxt_image.h:192-203)
    dp.pack.dxt_quality = (uint32_t)p.m_dxt_quality;
    dp.pack.perceptual = (p.m_flags & cCRNCompFlagPerceptual) && !is_non_srgb(p.m_format);
    dp.pack.use_both_block_types = (p.m_flags & cCRNCompFlagUseBothBlockTypes) != 0;
    dp.pack.dxt1a_alpha_threshold = p.m_dxt1a_alpha_threshold;
    dp.pack.use_transparent_indices_for_black = (p.m_flags & cCRNCompFlagUseTransparentIndicesForBlack) != 0;
    dp.pack.grayscale_sampling = (p.m_flags & cCRNCompFlagGrayscaleSampling) != 0;
    dp.target_bitrate = p.m_target_bitrate;
    dp.hierarchical = (p.m_flags & cCRNCompFlagHierarchical) != 0;
}

bool supported(const crn_comp_params& p)
{
    // what this path implements; anything else fails like an invalid parameter would (NULL), never falls back to a CPU
    if (p.m_dxt_compressor_type != cCRNDXTCompressorCRN) return false;                       // CRNF / RYG block compressors: out of scope
    if (p.m_file_type == cCRNFileTypeCRN) {
        // (cCRNCompFlagHierarchical: dxt_hc never reads m_hierarchical in this revision, crnlib/crn_dxt_hc.cpp; nothing to switch)
        switch (p.m_format) {
        case cCRNFmtDXT1: case cCRNFmtDXT5: case cCRNFmtDXT5_CCxY: case cCRNFmtDXT5_xGxR: case cCRNFmtDXT5_xGBR: case cCRNFmtDXT5_AGBR:
        case cCRNFmtDXN_XY: case cCRNFmtDXN_YX: case cCRNFmtDXT5A: return true;
        default: return false;                                                               // DXT3 is refused by the reference too; ETC: not built
        }
    }
    switch (p.m_format) {
    case cCRNFmtDXT1: case cCRNFmtDXT3: case cCRNFmtDXT5: case cCRNFmtDXT5_CCxY: case cCRNFmtDXT5_xGxR: case cCRNFmtDXT5_xGBR: case cCRNFmtDXT5_AGBR:
    case cCRNFmtDXN_XY: case cCRNFmtDXN_YX: case cCRNFmtDXT5A: return true;
    default: return false;
    }
}

void* compress_common(const crn_comp_params& p, const crn_mipmap_params* mip, crn_uint32& compressed_size, crn_uint32* pq, float* pb)
{
    compressed_size = 0;
    if (pq) *pq = 0;
    if (pb) *pb = 0.0f;
    if (!p.check() || (mip && !mip->check()) || !supported(p)) return nullptr;
    crn_gpu_ctx* ctx = gpu();
    if (!ctx) return nullptr;
    ProgressScope scope(ctx, p);
    void* file = nullptr; uint32_t size = 0; float rate = 0.0f; uint32_t quality = p.m_quality_level;
    int rc;
    crn_gpu_crn_params cp; crn_gpu_dds_params dp;
    const bool crn = p.m_file_type == cCRNFileTypeCRN;
    if (crn) fill_crn_params(p, cp); else fill_dds_params(p, dp);
    // create_texture_mipmaps (crnlib/crn_texture_comp.cpp:352-575): which levels go in
    bool generate = false;
    const void* faces[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    struct Replaced { void* p[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr }; ~Replaced() { for (void* q : p) if (q) crn_gpu_free_file(q); } } replaced;
    crn_gpu_resample_params rp;
    crn_gpu_default_resample_params(&rp);
    bool source_changed = false;
    if (mip) {
        switch (mip->m_mode) {
        case cCRNMipModeUseSourceOrGenerateMips: generate = p.m_levels == 1; break;
        case cCRNMipModeUseSourceMips: break;
        case cCRNMipModeGenerateMips: generate = true; break;
        case cCRNMipModeNoMips: if (crn) cp.levels = 1; else dp.levels = 1; break;
        default: return nullptr;
        }
        rp.filter = (uint32_t)mip->m_filter; rp.filter_scale = mip->m_blurriness; rp.srgb = mip->m_gamma_filtering ? 1u : 0u;
        rp.source_gamma = mip->m_gamma; rp.wrapping = mip->m_tiled ? 1u : 0u; rp.num_comps = 0; rp.renormalize = mip->m_renormalize ? 1u : 0u;
        // crop / clamp / rescale / renormalise the top level (crnlib/crn_texture_comp.cpp:392-540); a crop or resize drops every other source level
        crn_gpu_mip_source_params sp;
        memset(&sp, 0, sizeof(sp));
        sp.struct_size = sizeof(sp);
        sp.window_left = mip->m_window_left; sp.window_top = mip->m_window_top; sp.window_right = mip->m_window_right; sp.window_bottom = mip->m_window_bottom;
        sp.clamp_width = mip->m_clamp_width; sp.clamp_height = mip->m_clamp_height; sp.clamp_scale = mip->m_clamp_scale ? 1u : 0u;
        sp.scale_mode = (uint32_t)mip->m_scale_mode; sp.scale_x = mip->m_scale_x; sp.scale_y = mip->m_scale_y; sp.rtopmip = mip->m_rtopmip ? 1u : 0u;
        const void* src[6];
        for (crn_uint32 f = 0; f < p.m_faces; f++) { if (!p.m_pImages[f][0]) return nullptr; src[f] = p.m_pImages[f][0]; }
        uint32_t nw = 0, nh = 0, changed = 0;
        if (crn_gpu_prepare_mip_source(ctx, &sp, &rp, p.m_faces, p.m_width, p.m_height, src, replaced.p, &nw, &nh, &changed) != CRN_GPU_OK) return nullptr;
        if (changed) {
            source_changed = true;
            for (crn_uint32 f = 0; f < p.m_faces; f++) faces[f] = replaced.p[f];
            if (crn) { cp.width = nw; cp.height = nh; cp.levels = 1; } else { dp.width = nw; dp.height = nh; dp.levels = 1; }
        }
    }
    std::vector<const void*> flat;
    if (generate) {
        for (crn_uint32 f = 0; f < p.m_faces; f++) if (!faces[f]) { if (!p.m_pImages[f][0]) return nullptr; faces[f] = p.m_pImages[f][0]; }
        if (crn && cp.target_bitrate > 0.0f) {}                                               // handled inside crn_gpu_compress_crn
        rc = crn_gpu_compress_mip_chain(ctx, crn ? 0u : 1u, crn ? &cp : nullptr, crn ? nullptr : &dp, &rp, mip->m_min_mip_size, mip->m_max_levels, faces, &file, &size);
        if (rc == CRN_GPU_OK && crn) {                                                        // the chain call has no rate outputs: file bits / texels (crn_comp.cpp:1640-1653)
            crn_gpu_texture_info ti; ti.struct_size = sizeof(ti);
            if (crn_gpu_crnd_get_texture_info(file, size, &ti) == CRN_GPU_OK) {
                uint64_t texels = 0;
                for (uint32_t l = 0; l < ti.levels; l++) texels += (uint64_t)(ti.width >> l ? ti.width >> l : 1) * (ti.height >> l ? ti.height >> l : 1);
                rate = texels ? size * 8.0f / (float)(texels * ti.faces) : 0.0f;
            }
        }
    } else {
        const crn_uint32 levels = crn ? cp.levels : dp.levels;
        for (crn_uint32 f = 0; f < p.m_faces; f++)
            for (crn_uint32 l = 0; l < levels; l++) {
                const void* img = (source_changed && l == 0) ? faces[f] : p.m_pImages[f][l];
                if (!img) return nullptr;
                flat.push_back(img);
            }
        if (crn) rc = crn_gpu_compress_crn(ctx, &cp, flat.data(), &file, &size, &rate, &quality);
        else rc = crn_gpu_compress_dds_ex(ctx, &dp, flat.data(), &file, &size, pb ? &rate : nullptr, &quality);
    }
    if (rc != CRN_GPU_OK || !file) return nullptr;
    void* out = adopt(file, size);
    if (!out) return nullptr;
    compressed_size = size;
    if (pq) *pq = quality;                      // target bitrate: the level the search picked; otherwise m_quality_level (crn_texture_comp.cpp:101-104, :253-256)
    if (pb) *pb = rate;
    return out;
}

}  // namespace

// ---- crnlib.h -------------------------------------------------------------------------------------------------------
void crn_set_memory_callbacks(crn_realloc_func pRealloc, crn_msize_func pMSize, void* pUser_data)
{
    std::lock_guard<std::mutex> lock(g_mem_mutex);
    if (!pRealloc || !pMSize) { g_realloc = default_realloc; g_msize = default_msize; g_mem_user = nullptr; }   // crn_mem.cpp:332-345
    else { g_realloc = pRealloc; g_msize = pMSize; g_mem_user = pUser_data; }
}

void crn_free_block(void* pBlock) { lib_free(pBlock); }

void* crn_compress(const crn_comp_params& comp_params, crn_uint32& compressed_size, crn_uint32* pActual_quality_level, float* pActual_bitrate)
{
    return compress_common(comp_params, nullptr, compressed_size, pActual_quality_level, pActual_bitrate);
}

void* crn_compress(const crn_comp_params& comp_params, const crn_mipmap_params& mip_params, crn_uint32& compressed_size, crn_uint32* pActual_quality_level, float* pActual_bitrate)
{
    return compress_common(comp_params, &mip_params, compressed_size, pActual_quality_level, pActual_bitrate);
}

void* crn_decompress_crn_to_dds(const void* pCRN_file_data, crn_uint32& file_size)
{
    const crn_uint32 in_size = file_size;
    file_size = 0;                                                                            // crnlib.cpp:272-278
    crn_gpu_ctx* ctx = gpu();
    if (!ctx || !pCRN_file_data) return nullptr;
    void* file = nullptr; uint32_t size = 0;
    if (crn_gpu_crn_to_dds(ctx, pCRN_file_data, in_size, &file, &size) != CRN_GPU_OK) return nullptr;
    void* out = adopt(file, size);
    if (out) file_size = size;
    return out;
}

bool crn_decompress_dds_to_images(const void* pDDS_file_data, crn_uint32 dds_file_size, crn_uint32** ppImages, crn_texture_desc& tex_desc)
{
    memset(&tex_desc, 0, sizeof(tex_desc));                                                   // crnlib.cpp:295
    crn_gpu_ctx* ctx = gpu();
    if (!ctx || !pDDS_file_data || !ppImages) return false;
    crn_gpu_dds_desc d; d.struct_size = sizeof(d);
    if (crn_gpu_dds_get_desc(pDDS_file_data, dds_file_size, &d) != CRN_GPU_OK) return false;
    const uint32_t count = d.faces * d.levels;
    std::vector<void*> imgs(count, nullptr);
    bool ok = true;
    for (uint32_t f = 0; f < d.faces && ok; f++)
        for (uint32_t l = 0; l < d.levels && ok; l++) {
            const size_t bytes = (size_t)(d.width >> l ? d.width >> l : 1) * (d.height >> l ? d.height >> l : 1) * 4;
            ok = (imgs[l + d.levels * f] = lib_alloc(bytes)) != nullptr;
        }
    if (ok) ok = crn_gpu_dds_to_images(ctx, pDDS_file_data, dds_file_size, imgs.data(), count, &d) == CRN_GPU_OK;    // d: final format (DXT1 -> DXT1A scan)
    if (!ok) { for (void* p : imgs) lib_free(p); return false; }
    tex_desc.m_faces = d.faces; tex_desc.m_width = d.width; tex_desc.m_height = d.height; tex_desc.m_levels = d.levels; tex_desc.m_fmt_fourcc = d.pixel_format;
    for (uint32_t i = 0; i < count; i++) ppImages[i] = static_cast<crn_uint32*>(imgs[i]);    // index l + levels * f (crnlib.cpp:327)
    return true;
}

void crn_free_all_images(crn_uint32** ppImages, const crn_texture_desc& desc)
{
    for (crn_uint32 f = 0; f < desc.m_faces; f++)
        for (crn_uint32 l = 0; l < desc.m_levels; l++) crn_free_block(ppImages[l + desc.m_levels * f]);
}

crn_uint32 crn_get_format_fourcc(crn_format fmt) { return crnd::crnd_crn_format_to_fourcc(fmt); }
crn_uint32 crn_get_format_bits_per_texel(crn_format fmt) { return crnd::crnd_get_crn_format_bits_per_texel(fmt); }
crn_uint32 crn_get_bytes_per_dxt_block(crn_format fmt) { return crnd::crnd_get_bytes_per_dxt_block(fmt); }
crn_format crn_get_fundamental_dxt_format(crn_format fmt) { return crnd::crnd_get_fundamental_dxt_format(fmt); }

const char* crn_get_file_type_ext(crn_file_type t) { return t == cCRNFileTypeDDS ? "dds" : (t == cCRNFileTypeCRN ? "crn" : "?"); }

const char* crn_get_format_string(crn_format fmt)
{   // pixel_format_helpers::get_crn_format_string (crnlib/crn_pixel_format.cpp)
    static const char* const names[] = { "DXT1", "DXT3", "DXT5", "DXT5_CCxY", "DXT5_xGxR", "DXT5_xGBR", "DXT5_AGBR", "DXN_XY", "DXN_YX", "DXT5A",
                                         "ETC1", "ETC2", "ETC2A", "ETC1S", "ETC2AS" };
    return ((int)fmt >= 0 && (int)fmt < (int)(sizeof(names) / sizeof(names[0]))) ? names[(int)fmt] : "?";
}

const char* crn_get_dxt_quality_string(crn_dxt_quality q)
{
    static const char* const names[] = { "SuperFast", "Fast", "Normal", "Better", "Uber" };
    return (uint32_t)q < 5 ? names[(uint32_t)q] : "?";
}

const char* crn_get_mip_mode_desc(crn_mip_mode m)
{
    static const char* const d[] = { "Use source/generate if none", "Only use source MIP maps (if any)", "Always generate new MIP maps", "No MIP maps" };
    return (uint32_t)m < 4 ? d[(uint32_t)m] : "?";
}
const char* crn_get_mip_mode_name(crn_mip_mode m)
{
    static const char* const d[] = { "UseSourceOrGenerate", "UseSource", "Generate", "None" };
    return (uint32_t)m < 4 ? d[(uint32_t)m] : "?";
}
const char* crn_get_mip_filter_name(crn_mip_filter f)
{
    static const char* const d[] = { "box", "tent", "lanczos4", "mitchell", "kaiser" };
    return (uint32_t)f < 5 ? d[(uint32_t)f] : "?";
}
const char* crn_get_scale_mode_desc(crn_scale_mode sm)
{
    static const char* const d[] = { "disabled", "absolute", "relative", "lowerpow2", "nearestpow2", "nextpow2" };
    return (uint32_t)sm < 6 ? d[(uint32_t)sm] : "?";
}

const char* crn_get_version() { return "1.2.0"; }
int crn_get_version_number() { return 120; }
int crn_get_version_major() { return 1; }
int crn_get_version_minor() { return 2; }
int crn_get_version_patch() { return 0; }

// ---- 4x4 block API (crnlib.cpp:345-420, :451-533): one block = one 4x4 image through the same kernels ---------------------
namespace {
struct BlockCompressor { uint32_t fmt; crn_gpu_pack_params pack; };
int gpu_block_format(crn_format f, bool dxt1a)
{
    switch (crnd::crnd_get_fundamental_dxt_format(f)) {
    case cCRNFmtDXT1: return dxt1a ? CRN_GPU_FMT_DXT1A : CRN_GPU_FMT_DXT1;
    case cCRNFmtDXT3: return CRN_GPU_FMT_DXT3;
    case cCRNFmtDXT5: return CRN_GPU_FMT_DXT5;
    case cCRNFmtDXN_XY: return CRN_GPU_FMT_DXN_XY;
    case cCRNFmtDXN_YX: return CRN_GPU_FMT_DXN_YX;
    case cCRNFmtDXT5A: return CRN_GPU_FMT_DXT5A;
    default: return -1;
    }
}
}  // namespace

crn_block_compressor_context_t crn_create_block_compressor(const crn_comp_params& params)
{
    const int fmt = gpu_block_format(params.m_format, (params.m_flags & cCRNCompFlagDXT1AForTransparency) != 0);
    if (fmt < 0 || !gpu()) return nullptr;
    BlockCompressor* b = static_cast<BlockCompressor*>(lib_alloc(sizeof(BlockCompressor)));
    if (!b) return nullptr;
    crn_gpu_dds_params dp;
    fill_dds_params(params, dp);
    b->fmt = (uint32_t)fmt; b->pack = dp.pack;
    return b;
}

void crn_compress_block(crn_block_compressor_context_t pContext, const crn_uint32* pPixels, void* pDst_block)
{
    BlockCompressor* b = static_cast<BlockCompressor*>(pContext);
    crn_gpu_ctx* ctx = gpu();
    if (b && ctx && pPixels && pDst_block) crn_gpu_pack_image_host(ctx, b->fmt, &b->pack, pPixels, 4, 4, 16, pDst_block);
}

void crn_free_block_compressor(crn_block_compressor_context_t pContext) { lib_free(pContext); }

bool crn_decompress_block(const void* pSrc_block, crn_uint32* pDst_pixels, crn_format crn_fmt)
{
    int fmt = gpu_block_format(crn_fmt, false);
    // The reference's switch has no `break` after its DXT5 case (crnlib.cpp:469-490): control falls into the DXN case, which
    // rewrites all 16 pixels as (values1[s1], values0[s0], 255, 255) -- the DXN_YX reading of the same 16 bytes.  Reproduced.
    if (fmt == CRN_GPU_FMT_DXT5) fmt = CRN_GPU_FMT_DXN_YX;
    crn_gpu_ctx* ctx = gpu();
    if (fmt < 0 || !ctx || !pSrc_block || !pDst_pixels) return false;
    if (crn_gpu_unpack_image_host(ctx, (uint32_t)fmt, pSrc_block, 4, 4, pDst_pixels, 16) != CRN_GPU_OK) return false;
    // dxt_image::unpack leaves the channels a format does not carry at (0, 0, 0, 255); this entry point fills them with 255
    // (set_noclamp_rgba(x, y, 255, 255) for DXN, (255, 255, 255, a) for DXT5A, crnlib.cpp:492-531)
    const crn_uint32 fill = (fmt == CRN_GPU_FMT_DXN_XY || fmt == CRN_GPU_FMT_DXN_YX) ? 0x00FF0000u : (fmt == CRN_GPU_FMT_DXT5A ? 0x00FFFFFFu : 0u);
    for (int i = 0; i < 16; i++) pDst_pixels[i] |= fill;
    return true;
}

// ---- crn_defs.h: namespace crnd ------------------------------------------------------------------------------------------------
namespace crnd {

namespace {
crnd_realloc_func g_crnd_realloc = nullptr;
crnd_msize_func g_crnd_msize = nullptr;
void* g_crnd_user = nullptr;
void* crnd_alloc(size_t n) { size_t a = 0; return g_crnd_realloc ? g_crnd_realloc(nullptr, n, &a, true, g_crnd_user) : malloc(n); }
void crnd_release(void* p) { size_t a = 0; if (g_crnd_realloc) g_crnd_realloc(p, 0, &a, true, g_crnd_user); else free(p); }

struct UnpackContext {                       // crn_unpacker (inc/crn_decomp.h:3519-3619): magic + borrowed file + our device object
    uint32 magic;
    const void* data; uint32 size;
    crn_gpu_texture* tex;
};
const uint32 kMagic = 0x1EF9CABD;            // crn_unpacker::cMagicValue

inline uint32 be16(const uint8* p) { return (uint32)p[0] << 8 | p[1]; }
inline uint32 be24(const uint8* p) { return (uint32)p[0] << 16 | (uint32)p[1] << 8 | p[2]; }
inline uint32 be32(const uint8* p) { return (uint32)p[0] << 24 | (uint32)p[1] << 16 | (uint32)p[2] << 8 | p[3]; }

// crnd_get_header (inc/crn_decomp.h:2657-2670): signature, header_size >= sizeof(crn_header), data_size <= buffer
const uint8* get_header(const void* pData, uint32 data_size)
{
    if (!pData || data_size < sizeof(crn_header)) return nullptr;
    const uint8* d = static_cast<const uint8*>(pData);
    if (be16(d) != crn_header::cCRNSigValue) return nullptr;
    if (be16(d + 2) < sizeof(crn_header) || data_size < be32(d + 6)) return nullptr;
    return d;
}
uint16 crc16(const uint8* p, size_t n, uint16 crc = 0)
{   // inc/crn_decomp.h:2372-2390
    crc = ~crc;
    while (n--) {
        const uint16 q = *p++ ^ (crc >> 8);
        uint16 k = (q >> 4) ^ q;
        crc = (((crc << 8) ^ k) ^ (k << 5)) ^ (k << 12);
    }
    return ~crc;
}
}  // namespace

void crnd_set_memory_callbacks(crnd_realloc_func pRealloc, crnd_msize_func pMSize, void* pUser_data)
{
    if (!pRealloc || !pMSize) { g_crnd_realloc = nullptr; g_crnd_msize = nullptr; g_crnd_user = nullptr; }
    else { g_crnd_realloc = pRealloc; g_crnd_msize = pMSize; g_crnd_user = pUser_data; }
}

uint32 crnd_crn_format_to_fourcc(crn_format fmt)
{
#define FCC(a, b, c, d) ((uint32)(a) | ((uint32)(b) << 8) | ((uint32)(c) << 16) | ((uint32)(d) << 24))
    switch (fmt) {
    case cCRNFmtDXT1: return FCC('D', 'X', 'T', '1');
    case cCRNFmtDXT3: return FCC('D', 'X', 'T', '3');
    case cCRNFmtDXT5: return FCC('D', 'X', 'T', '5');
    case cCRNFmtDXN_XY: return FCC('A', '2', 'X', 'Y');
    case cCRNFmtDXN_YX: return FCC('A', 'T', 'I', '2');
    case cCRNFmtDXT5A: return FCC('A', 'T', 'I', '1');
    case cCRNFmtDXT5_CCxY: return FCC('C', 'C', 'x', 'Y');
    case cCRNFmtDXT5_xGxR: return FCC('x', 'G', 'x', 'R');
    case cCRNFmtDXT5_xGBR: return FCC('x', 'G', 'B', 'R');
    case cCRNFmtDXT5_AGBR: return FCC('A', 'G', 'B', 'R');
    case cCRNFmtETC1: return FCC('E', 'T', 'C', '1');
    case cCRNFmtETC2: return FCC('E', 'T', 'C', '2');
    case cCRNFmtETC2A: return FCC('E', 'T', '2', 'A');
    case cCRNFmtETC1S: return FCC('E', 'T', '1', 'S');
    case cCRNFmtETC2AS: return FCC('E', '2', 'A', 'S');
    default: return 0;
    }
#undef FCC
}

crn_format crnd_get_fundamental_dxt_format(crn_format fmt)
{
    return (fmt == cCRNFmtDXT5_CCxY || fmt == cCRNFmtDXT5_xGxR || fmt == cCRNFmtDXT5_xGBR || fmt == cCRNFmtDXT5_AGBR) ? cCRNFmtDXT5 : fmt;
}

uint32 crnd_get_crn_format_bits_per_texel(crn_format fmt)
{
    switch (fmt) {
    case cCRNFmtDXT1: case cCRNFmtDXT5A: case cCRNFmtETC1: case cCRNFmtETC2: case cCRNFmtETC1S: return 4;
    case cCRNFmtDXT3: case cCRNFmtDXT5: case cCRNFmtDXN_XY: case cCRNFmtDXN_YX: case cCRNFmtDXT5_CCxY: case cCRNFmtDXT5_xGxR: case cCRNFmtDXT5_xGBR:
    case cCRNFmtDXT5_AGBR: case cCRNFmtETC2A: case cCRNFmtETC2AS: return 8;
    default: return 0;
    }
}

uint32 crnd_get_bytes_per_dxt_block(crn_format fmt) { return (crnd_get_crn_format_bits_per_texel(fmt) << 4) >> 3; }

bool crnd_validate_file(const void* pData, uint32 data_size, crn_file_info* pFile_info)
{   // inc/crn_decomp.h:2672-2735
    if (pFile_info) {
        if (pFile_info->m_struct_size != sizeof(crn_file_info)) return false;
        memset(&pFile_info->m_struct_size + 1, 0, sizeof(crn_file_info) - sizeof(pFile_info->m_struct_size));
    }
    if (!pData || data_size < cCRNHeaderMinSize) return false;
    const uint8* h = get_header(pData, data_size);
    if (!h) return false;
    const uint32 header_size = be16(h + 2), file_size = be32(h + 6);
    if (header_size > data_size || header_size > file_size) return false;
    if (crc16(h + 6, header_size - 6) != be16(h + 4)) return false;
    if (crc16(h + header_size, file_size - header_size) != be16(h + 10)) return false;
    const uint32 width = be16(h + 12), height = be16(h + 14), levels = h[16], faces = h[17], format = h[18];
    if (faces != 1 && faces != 6) return false;
    if (width < 1 || width > cCRNMaxLevelResolution || height < 1 || height > cCRNMaxLevelResolution) return false;
    uint32 max_mips = 1;
    for (uint32 s = width > height ? width : height; s > 1; s >>= 1) max_mips++;          // utils::compute_max_mips
    if (levels < 1 || levels > max_mips || format >= (uint32)cCRNFmtTotal) return false;
    if (pFile_info) {
        pFile_info->m_actual_data_size = file_size;
        pFile_info->m_header_size = header_size;
        pFile_info->m_total_palette_size = be24(h + 33 + 3) + be24(h + 41 + 3) + be24(h + 49 + 3) + be24(h + 57 + 3);
        pFile_info->m_tables_size = be16(h + 65);
        pFile_info->m_levels = levels;
        for (uint32 i = 0; i < levels; i++) {
            const uint32 next = i + 1 < levels ? be32(h + 70 + 4 * (i + 1)) : file_size;
            pFile_info->m_level_compressed_size[i] = next - be32(h + 70 + 4 * i);
        }
        pFile_info->m_color_endpoint_palette_entries = be16(h + 33 + 6);
        pFile_info->m_color_selector_palette_entries = be16(h + 41 + 6);
        pFile_info->m_alpha_endpoint_palette_entries = be16(h + 49 + 6);
        pFile_info->m_alpha_selector_palette_entries = be16(h + 57 + 6);
    }
    return true;
}

bool crnd_get_texture_info(const void* pData, uint32 data_size, crn_texture_info* pInfo)
{   // inc/crn_decomp.h:2737-2760
    if (!pData || data_size < sizeof(crn_header) || !pInfo || pInfo->m_struct_size != sizeof(crn_texture_info)) return false;
    const uint8* h = get_header(pData, data_size);
    if (!h) return false;
    pInfo->m_width = be16(h + 12); pInfo->m_height = be16(h + 14); pInfo->m_levels = h[16]; pInfo->m_faces = h[17];
    pInfo->m_format = static_cast<crn_format>((uint32)h[18]);
    const uint32 f = h[18];
    pInfo->m_bytes_per_block = (f == cCRNFmtDXT1 || f == cCRNFmtDXT5A || f == cCRNFmtETC1 || f == cCRNFmtETC2 || f == cCRNFmtETC1S) ? 8 : 16;
    pInfo->m_userdata0 = be32(h + 25); pInfo->m_userdata1 = be32(h + 29);
    return true;
}

bool crnd_get_level_info(const void* pData, uint32 data_size, uint32 level_index, crn_level_info* pLevel_info)
{   // inc/crn_decomp.h:2762-2790
    if (!pData || data_size < cCRNHeaderMinSize || !pLevel_info || pLevel_info->m_struct_size != sizeof(crn_level_info)) return false;
    const uint8* h = get_header(pData, data_size);
    if (!h || level_index >= h[16]) return false;
    const uint32 w = be16(h + 12) >> level_index, hh = be16(h + 14) >> level_index;
    pLevel_info->m_width = w ? w : 1; pLevel_info->m_height = hh ? hh : 1; pLevel_info->m_faces = h[17];
    pLevel_info->m_blocks_x = (pLevel_info->m_width + 3) >> 2; pLevel_info->m_blocks_y = (pLevel_info->m_height + 3) >> 2;
    pLevel_info->m_bytes_per_block = (h[18] == cCRNFmtDXT1 || h[18] == cCRNFmtDXT5A) ? 8 : 16;
    pLevel_info->m_format = static_cast<crn_format>((uint32)h[18]);
    return true;
}

const void* crnd_get_level_data(const void* pData, uint32 data_size, uint32 level_index, uint32* pSize)
{   // inc/crn_decomp.h:2792-2822
    if (pSize) *pSize = 0;
    if (!pData || data_size < cCRNHeaderMinSize) return nullptr;
    const uint8* h = get_header(pData, data_size);
    if (!h || level_index >= h[16]) return nullptr;
    const uint32 cur = be32(h + 70 + 4 * level_index);
    if (pSize) {
        const uint32 next = level_index + 1 < h[16] ? be32(h + 70 + 4 * (level_index + 1)) : be32(h + 6);
        *pSize = next - cur;
    }
    return h + cur;
}

uint32 crnd_get_segmented_file_size(const void* pData, uint32 data_size)
{   // inc/crn_decomp.h:2824-2843: everything before the first level
    if (!pData || data_size < cCRNHeaderMinSize) return 0;
    const uint8* h = get_header(pData, data_size);
    if (!h) return 0;
    uint32 size = be16(h + 2);
    const uint32 ends[5] = { be24(h + 33) + be24(h + 36), be24(h + 41) + be24(h + 44), be24(h + 49) + be24(h + 52), be24(h + 57) + be24(h + 60), be24(h + 67) + be16(h + 65) };
    for (int i = 0; i < 5; i++) if (ends[i] > size) size = ends[i];
    return size;
}

bool crnd_create_segmented_file(const void* pData, uint32 data_size, void* pBase_data, uint base_data_size)
{   // inc/crn_decomp.h:2839-2868: copy the base data, flag it segmented, set data_size, refresh both CRCs (level offsets stay as they are)
    if (!pData || data_size < cCRNHeaderMinSize || !pBase_data) return false;
    const uint8* h = get_header(pData, data_size);
    if (!h || (be16(h + 19) & cCRNHeaderFlagSegmented)) return false;
    const uint32 actual = crnd_get_segmented_file_size(pData, data_size);
    if (base_data_size < actual) return false;
    uint8* o = static_cast<uint8*>(pBase_data);
    memcpy(o, pData, actual);
    const uint32 flags = be16(o + 19) | cCRNHeaderFlagSegmented;
    o[19] = (uint8)(flags >> 8); o[20] = (uint8)flags;
    o[6] = (uint8)(actual >> 24); o[7] = (uint8)(actual >> 16); o[8] = (uint8)(actual >> 8); o[9] = (uint8)actual;
    const uint32 header_size = be16(o + 2);
    const uint16 dcrc = crc16(o + header_size, actual - header_size);
    o[10] = (uint8)(dcrc >> 8); o[11] = (uint8)dcrc;
    const uint16 hcrc = crc16(o + 6, header_size - 6);
    o[4] = (uint8)(hcrc >> 8); o[5] = (uint8)hcrc;
    return true;
}

crnd_unpack_context crnd_unpack_begin(const void* pData, uint32 data_size)
{   // inc/crn_decomp.h:4404-4420
    if (!pData || data_size < cCRNHeaderMinSize) return nullptr;
    crn_gpu_ctx* ctx = gpu();
    if (!ctx) return nullptr;
    UnpackContext* c = static_cast<UnpackContext*>(crnd_alloc(sizeof(UnpackContext)));
    if (!c) return nullptr;
    c->magic = kMagic; c->data = pData; c->size = data_size; c->tex = nullptr;
    if (crn_gpu_crnd_unpack_begin(ctx, pData, data_size, &c->tex) != CRN_GPU_OK) { crnd_release(c); return nullptr; }
    return c;
}

bool crnd_get_data(crnd_unpack_context pContext, const void** ppData, uint32* pData_size)
{
    UnpackContext* c = static_cast<UnpackContext*>(pContext);
    if (!c || c->magic != kMagic) return false;
    if (ppData) *ppData = c->data;
    if (pData_size) *pData_size = c->size;
    return true;
}

bool crnd_unpack_level(crnd_unpack_context pContext, void** ppDst, uint32 dst_size_in_bytes, uint32 row_pitch_in_bytes, uint32 level_index)
{   // inc/crn_decomp.h:4441-4458 -> crn_unpacker::unpack_level (:3552-3619); ppDst are HOST pointers, one per face
    UnpackContext* c = static_cast<UnpackContext*>(pContext);
    if (!c || !ppDst || dst_size_in_bytes < 8 || level_index >= cCRNMaxLevels || c->magic != kMagic) return false;
    return crn_gpu_crnd_unpack_level_host(c->tex, ppDst, dst_size_in_bytes, row_pitch_in_bytes, level_index) == CRN_GPU_OK;
}

bool crnd_unpack_level_segmented(crnd_unpack_context, const void*, uint32, void**, uint32, uint32, uint32)
{
    return false;                               // segmented level data living outside the file image: not built on the device path
}

bool crnd_unpack_end(crnd_unpack_context pContext)
{
    UnpackContext* c = static_cast<UnpackContext*>(pContext);
    if (!c || c->magic != kMagic) return false;
    crn_gpu_crnd_unpack_end(c->tex);
    c->magic = 0;
    crnd_release(c);
    return true;
}

}  // namespace crnd
