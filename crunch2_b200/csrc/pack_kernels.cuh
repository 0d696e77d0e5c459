// pack_kernels.cuh -- block-by-block packer kernels: one warp per 4x4 block.
//
// Replaces dxt_image::init_task + set_block_pixels (reference crnlib/crn_dxt_image.cpp:283-349,
// :1427-1541): clamped 4x4 gather, per-element optimiser, selector bit packing into 8-byte elements.
// Each kernel writes ONE 8-byte element of every block (element stride = bytes per block), so the
// alpha and colour elements of DXT5 / the two halves of DXN are independent launches that can
// overlap on the device.
#pragma once
#include "dxt5a_opt.cuh"
#include "dxt1_opt.cuh"

namespace crn {

struct ImageView {
    const uint8_t* rgba;   // device, RGBA8 row-major
    uint32_t width, height, pitch;
    uint32_t blocks_x, blocks_y;
};

// lanes 0..15 fetch pixel (x = lane&3, y = lane>>2) of block (bx,by) with edge clamping
// (crn_dxt_image.cpp:334-344).  Other lanes return 0.
__device__ __forceinline__ uint32_t fetch_block_pixel(const ImageView& img, uint32_t bx, uint32_t by)
{
    const unsigned lane = lane_id();
    uint32_t px = 0;
    if (lane < 16) {
        uint32_t x = min(bx * 4 + (lane & 3), img.width - 1);
        uint32_t y = min(by * 4 + (lane >> 2), img.height - 1);
        px = *reinterpret_cast<const uint32_t*>(img.rgba + (size_t)y * img.pitch + (size_t)x * 4);
    }
    return px;
}

constexpr int kPackWarpsPerCta = 8;
#ifndef CRN_COLOR_MIN_CTAS
#define CRN_COLOR_MIN_CTAS 2   // resident CTAs per SM the colour kernel is register-budgeted for
#endif

// DXT5A-type element (alpha of DXT5, DXT5A, both halves of DXN) ------------------------------------
__global__ void __launch_bounds__(kPackWarpsPerCta * 32)
pack_alpha_element_kernel(ImageView img, uint32_t comp, int quality, int both_types,
                          uint8_t* __restrict__ out, uint32_t bytes_per_block, uint32_t elem_ofs)
{
    __shared__ Dxt5aScratch scratch[kPackWarpsPerCta];
    const unsigned warp = threadIdx.x >> 5;
    Dxt5aScratch* sc = &scratch[warp];
    const uint32_t total = img.blocks_x * img.blocks_y;
    for (uint32_t b = blockIdx.x * kPackWarpsPerCta + warp; b < total; b += gridDim.x * kPackWarpsPerCta) {
        const uint32_t bx = b % img.blocks_x, by = b / img.blocks_x;
        const uint32_t px = fetch_block_pixel(img, bx, by);
        const unsigned value = (px >> (8 * comp)) & 0xffu;
        const unsigned long long elem = dxt5a_pack_block(sc, value, quality, both_types != 0);
        if (lane_id() == 0)
            *reinterpret_cast<unsigned long long*>(out + (size_t)b * bytes_per_block + elem_ofs) = elem;
    }
}

// DXT3 explicit 4-bit alpha element (crn_dxt_image.cpp:1524-1536, crn_dxt.h dxt3_block::set_alpha with
// scaled=true: (v*15+128)/255 ... see dxt3_quantize) -- one thread per block, trivially bandwidth bound.
__device__ __forceinline__ unsigned dxt3_quantize(unsigned v)
{   // reference crnlib/crn_dxt.cpp dxt3_block::set_alpha scaled path: value = (value * 15U + 128U) / 255U
    return (v * 15u + 128u) / 255u;
}
__global__ void pack_dxt3_alpha_kernel(ImageView img, uint32_t comp, uint8_t* __restrict__ out, uint32_t bytes_per_block, uint32_t elem_ofs)
{
    const uint32_t total = img.blocks_x * img.blocks_y;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < total; b += gridDim.x * blockDim.x) {
        const uint32_t bx = b % img.blocks_x, by = b / img.blocks_x;
        unsigned long long bits = 0;
        for (unsigned i = 0; i < 16; i++) {
            uint32_t x = min(bx * 4 + (i & 3), img.width - 1);
            uint32_t y = min(by * 4 + (i >> 2), img.height - 1);
            unsigned v = img.rgba[(size_t)y * img.pitch + (size_t)x * 4 + comp];
            bits |= (unsigned long long)dxt3_quantize(v) << (4 * i);
        }
        *reinterpret_cast<unsigned long long*>(out + (size_t)b * bytes_per_block + elem_ofs) = bits;
    }
}

// DXT1 colour element: five phase kernels over a chunk of blocks (see Dxt1BlockState) -----------------
__device__ __forceinline__ void state_load(Dxt1Scratch* sc, const Dxt1BlockState* g)
{
    const unsigned lane = lane_id();
    if (lane < (unsigned)kDxt1StateVec4) reinterpret_cast<int4*>(sc)[lane] = reinterpret_cast<const int4*>(g)[lane];
    __syncwarp();
}
__device__ __forceinline__ void state_store(Dxt1BlockState* g, const Dxt1Scratch* sc)
{
    const unsigned lane = lane_id();
    __syncwarp();
    if (lane < (unsigned)kDxt1StateVec4) reinterpret_cast<int4*>(g)[lane] = reinterpret_cast<const int4*>(sc)[lane];
}

// PHASE: 0 set-up, 1 LBG (try_median4), 2 sweep passes, 3 post passes, 4 finish (writes the element).
// Blocks [first, first + count) of the image; states[i] belongs to block first + i.
// black: try_alpha_as_black_optimization (crn_dxt1.cpp:2001-2079, :2241-2244; cCRNCompFlagUseTransparentIndicesForBlack) as a second run of
// the five phases.  1 = the normal run, which also records its error per block in err_buf; 2 = the run with every near-black pixel
// (r, g, b <= 4) made transparent, for the blocks the reference retries (no transparent pixel of their own, some but not all unique colours
// near black): its element replaces the first one when its error over ALL 16 pixels -- black standing for the transparent index -- is lower.
template <int PHASE>
__global__ void __launch_bounds__(kPackWarpsPerCta * 32, CRN_COLOR_MIN_CTAS)
pack_color_phase_kernel(ImageView img, Dxt1Params prm, int dxt1a, Dxt1BlockState* __restrict__ states, uint32_t first, uint32_t count,
                        uint8_t* __restrict__ out, uint32_t bytes_per_block, uint32_t elem_ofs, int black, unsigned long long* __restrict__ err_buf)
{
    __shared__ __align__(16) Dxt1Scratch scratch[kPackWarpsPerCta];
    const unsigned warp = threadIdx.x >> 5;
    Dxt1Scratch* sc = &scratch[warp];
    for (uint32_t i = blockIdx.x * kPackWarpsPerCta + warp; i < count; i += gridDim.x * kPackWarpsPerCta) {
        const uint32_t b = first + i;
        uint32_t px = 0;
        if (PHASE == 0 || PHASE == 4) px = fetch_block_pixel(img, b % img.blocks_x, b / img.blocks_x);      // lanes 0..15: pixel 4y + x (edge clamped)
        const bool is_dark = (px & 0xffu) <= 4u && ((px >> 8) & 0xffu) <= 4u && ((px >> 16) & 0xffu) <= 4u;
        if (PHASE == 0) {
            int pha = 0;
            if (dxt1a)   // crn_dxt_image.cpp:1440-1451
                pha = __ballot_sync(CRN_FULL_MASK, lane_id() < 16 && (px >> 24) < prm.alpha_threshold) != 0;
            if (black == 2) {
                // unique colours of the block and how many of them are near black (:2003-2016)
                const unsigned lanes16 = 0xffffu;
                unsigned peers = 0;
                if (lane_id() < 16) peers = __match_any_sync(lanes16, px | 0xFF000000u);
                const bool leader = lane_id() < 16 && (unsigned)(__ffs((int)peers) - 1) == lane_id();
                const unsigned uniq = __popc(__ballot_sync(CRN_FULL_MASK, leader)), uniq_dark = __popc(__ballot_sync(CRN_FULL_MASK, leader && is_dark));
                if (pha || !uniq_dark || uniq_dark == uniq) {
                    if (lane_id() == 0) sc->stage = 3;                   // not retried: the later phases and the comparison skip this block
                    __syncwarp();
                } else dxt1_phase_setup(sc, is_dark ? (px & 0x00ffffffu) : px, prm, 1);
            } else dxt1_phase_setup(sc, px, prm, pha);
            state_store(&states[i], sc);
        } else {
            state_load(sc, &states[i]);
            if (sc->stage == 3) { __syncwarp(); continue; }
            if (sc->stage == 0 || PHASE == 4) dxt1_build_eval_colours(sc, dxt1_make_cfg(prm, sc->pixels_have_alpha, sc->U));
            if (PHASE == 1) dxt1_phase_median4(sc, prm);
            if (PHASE == 2) dxt1_phase_passes(sc, prm);
            if (PHASE == 3) dxt1_phase_post(sc, prm);
            if (PHASE < 4) {
                if (sc->stage == 0) state_store(&states[i], sc);
            } else if (black != 2) {
                const unsigned long long elem = dxt1_phase_finish(sc, px, prm);
                if (lane_id() == 0) {
                    *reinterpret_cast<unsigned long long*>(out + (size_t)b * bytes_per_block + elem_ofs) = elem;
                    if (black == 1) err_buf[i] = sc->stage == 2 ? 0ull : sc->best.err;
                }
            } else if (sc->stage != 3) {
                const unsigned long long elem = dxt1_phase_finish(sc, is_dark ? (px & 0x00ffffffu) : px, prm);
                // error of the retried block over all 16 pixels against its 3-colour palette, index 3 = black (:2046-2068; get_block_colors3)
                const Dxt1Cfg cfg = dxt1_make_cfg(prm, 1, sc->U);
                int r0, g0, b0, r1, g1, b1;
                unpack565((unsigned)(elem & 0xffff), true, r0, g0, b0);
                unpack565((unsigned)((elem >> 16) & 0xffff), true, r1, g1, b1);
                unsigned long long te = 0;
                if (lane_id() < 16) {
                    const unsigned sel = (unsigned)(elem >> (32 + 2 * lane_id())) & 3u;
                    const int pr = sel == 0 ? r0 : (sel == 1 ? r1 : (sel == 2 ? (r0 + r1) >> 1 : 0));
                    const int pg = sel == 0 ? g0 : (sel == 1 ? g1 : (sel == 2 ? (g0 + g1) >> 1 : 0));
                    const int pb = sel == 0 ? b0 : (sel == 1 ? b1 : (sel == 2 ? (b0 + b1) >> 1 : 0));
                    te = dxt1_dist(cfg, (int)(px & 0xffu), (int)((px >> 8) & 0xffu), (int)((px >> 16) & 0xffu), pr, pg, pb);
                }
                te = warp_sum_u64(te);
                if (lane_id() == 0 && te < err_buf[i])
                    *reinterpret_cast<unsigned long long*>(out + (size_t)b * bytes_per_block + elem_ofs) = elem;
            }
        }
        __syncwarp();
    }
}

}  // namespace crn
