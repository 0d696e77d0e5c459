/* oracle/port/oracle_port.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded restatement of the reference's hot path (crnlib 1.2.0), written from the
 * algorithm, not copied: every function cites the reference file:line it follows.  Built into
 * oracle/liboracle_port.so by oracle/Makefile and pinned against the unmodified reference
 * (oracle/_ref/liboracle_ref.so) by tests/test_oracle_port.py and the committed fixtures in
 * tests/golden/.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it;
 * the product (crunch2_b200/, libcrn_b200.so) never does.
 */
#ifndef ORACLE_PORT_H
#define ORACLE_PORT_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct op_dxt1_params {
    uint32_t quality;              /* crn_dxt_quality 0..4 */
    uint32_t perceptual;
    uint32_t pixels_have_alpha;
    uint32_t use_alpha_blocks;
    uint32_t alpha_threshold;
    uint32_t grayscale_sampling;
    uint32_t transparent_for_black;
    uint32_t force_alpha_blocks;
} op_dxt1_params;

typedef struct op_dxt1_result {
    uint64_t error;
    uint16_t low, high;
    uint8_t alpha_block;
} op_dxt1_result;

/* crnlib::dxt1_endpoint_optimizer::compute with endpoint caching disabled (crn_dxt1.cpp:2234).
 * pixels: num_pixels RGBA8 (r first).  selectors: num_pixels bytes out.  Returns 1 on success. */
int op_dxt1_optimize(const uint8_t* pixels, uint32_t num_pixels, const op_dxt1_params* p,
                     op_dxt1_result* r, uint8_t* selectors);

/* crnlib::dxt5_endpoint_optimizer::compute (crn_dxt5a.cpp:40). */
int op_dxt5_optimize(const uint8_t* pixels, uint32_t num_pixels, uint32_t comp_index, uint32_t quality,
                     uint32_t use_both_block_types, uint8_t* first, uint8_t* second, uint8_t* selectors,
                     uint64_t* error, uint8_t* block_type);

/* dxt_image::init for the CRN compressor, endpoint caching disabled (crn_dxt_image.cpp:283-349,
 * :1427-1541).  fmt is crnlib::dxt_format (0 DXT1, 1 DXT1A, 2 DXT3, 3 DXT5, 4 DXT5A, 5 DXN_XY, 6 DXN_YX).
 * Returns bytes written. */
uint32_t op_pack_image(int fmt, const uint8_t* rgba, uint32_t width, uint32_t height, uint32_t pitch,
                       uint32_t quality, uint32_t perceptual, uint32_t use_both_block_types,
                       uint32_t alpha_threshold, uint32_t transparent_for_black, uint8_t* out);

/* dxt_fast primitives and the qdxt1/qdxt5 tile analysis (see dxt_fast_port.c). */
void op_fast_color_block(uint32_t n, const uint8_t* px, uint32_t* low16, uint32_t* high16, uint8_t* sel);
void op_fast_alpha_block(uint32_t n, const uint8_t* px, uint32_t comp, uint32_t* low8, uint32_t* high8, uint8_t* sel);
void op_find_representative_colors(uint32_t n, const uint8_t* px, uint8_t* lo, uint8_t* hi);
void op_qdxt_training(int kind, uint32_t comp, const uint8_t* blocks, uint32_t n_blocks, const uint32_t* mips, uint32_t num_mips,
                      int hierarchical, uint8_t* out_vecs, uint32_t* out_weights, uint8_t* out_encoding);

/* crnlib::dxt_endpoint_refiner::refine (crn_dxt_endpoint_refiner.cpp:36-301).  px: n RGBA8, sel: n selectors (DXT1 or
 * DXT5 block order).  Returns r.m_error < error_to_beat; low / high / error are always written. */
int op_refine(int dxt1_selectors, int perceptual, uint32_t comp, const uint8_t* px, uint32_t n, const uint8_t* sel,
              uint64_t error_to_beat, uint32_t* low, uint32_t* high, uint64_t* error);
/* nearest codebook entry, first minimum (crn_dxt_hc.cpp:836-886, :1132-1163) */
void op_nearest_codebook(uint32_t dims, const float* vecs, uint32_t n, const float* codebook, uint32_t k, uint32_t* out);

/* selector codebook assignment + re-vote (crn_dxt_hc.cpp:1306-1360, :1488-1503; alpha :1516-1586, :1702-1720) */
void op_assign_selectors(int kind, int perceptual, uint32_t comp, const uint8_t* blocks, uint32_t n, const uint8_t* values, const uint8_t* values_accum,
                         const uint64_t* codebook, uint32_t K, uint32_t* best_index, uint64_t* refined, uint8_t* used);

/* CRN -> DXTn transcoder (inc/crn_decomp.h): crnd_unpack_begin / crnd_get_texture_info / crnd_unpack_level /
 * crnd_unpack_end.  info out[0..7] = width,height,levels,faces,bytes_per_block,format,userdata0,userdata1. */
typedef struct op_crnd op_crnd;
op_crnd* op_crnd_begin(const uint8_t* data, uint32_t size);
int op_crnd_info(const uint8_t* data, uint32_t size, uint32_t* out);
int op_crnd_unpack_level(op_crnd* c, void** dst_faces, uint32_t dst_size, uint32_t row_pitch, uint32_t level);
void op_crnd_end(op_crnd* c);

#ifdef __cplusplus
}
#endif
#endif
