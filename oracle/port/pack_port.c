/* oracle/port/pack_port.c -- TEST INFRASTRUCTURE ONLY (see oracle_port.h).
 * Block-by-block packer: dxt_image::init_task + set_block_pixels for the CRN compressor with endpoint
 * caching disabled (reference crnlib/crn_dxt_image.cpp:283-349, :1427-1541) and the element bit
 * layouts of crnlib/crn_dxt.h:109-172 (DXT1), :266-361 (DXT5 alpha), crn_dxt.cpp:327-348 (DXT3). */
#include "oracle_port.h"
#include <string.h>

static void put_color(uint8_t* e, const uint8_t* px, uint32_t fmt, uint32_t quality, uint32_t perceptual,
                      uint32_t both, uint32_t thresh, uint32_t tfb)
{
    op_dxt1_params p;
    memset(&p, 0, sizeof(p));
    p.quality = quality; p.perceptual = perceptual; p.alpha_threshold = thresh; p.transparent_for_black = tfb;
    p.use_alpha_blocks = both;
    if (fmt == 1) for (int i = 0; i < 16; i++) if (px[4 * i + 3] < thresh) { p.pixels_have_alpha = 1; break; }
    if (fmt != 0 && fmt != 1) p.use_alpha_blocks = 0;
    op_dxt1_result r; uint8_t sel[16];
    op_dxt1_optimize(px, 16, &p, &r, sel);
    e[0] = (uint8_t)r.low; e[1] = (uint8_t)(r.low >> 8); e[2] = (uint8_t)r.high; e[3] = (uint8_t)(r.high >> 8);
    for (int y = 0; y < 4; y++) {
        uint8_t b = 0;
        for (int x = 0; x < 4; x++) b |= (uint8_t)(sel[y * 4 + x] << (2 * x));
        e[4 + y] = b;
    }
}
static void put_alpha(uint8_t* e, const uint8_t* px, uint32_t comp, uint32_t quality, uint32_t both)
{
    uint8_t f, s, bt, sel[16]; uint64_t err;
    op_dxt5_optimize(px, 16, comp, quality, both, &f, &s, sel, &err, &bt);
    uint64_t bits = 0;
    for (int i = 0; i < 16; i++) bits |= (uint64_t)sel[i] << (3 * i);
    e[0] = f; e[1] = s;
    for (int i = 0; i < 6; i++) e[2 + i] = (uint8_t)(bits >> (8 * i));
}
static void put_dxt3(uint8_t* e, const uint8_t* px, uint32_t comp)
{
    memset(e, 0, 8);
    for (int i = 0; i < 16; i++) {
        uint32_t v = (px[4 * i + comp] * 15u + 128u) / 255u;
        e[i >> 1] |= (uint8_t)(v << ((i & 1) * 4));
    }
}

uint32_t op_pack_image(int fmt, const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t pitch, uint32_t quality,
                       uint32_t perceptual, uint32_t both, uint32_t thresh, uint32_t tfb, uint8_t* out)
{
    const uint32_t bx = (w + 3) >> 2, by = (h + 3) >> 2;
    const uint32_t bpb = (fmt == 0 || fmt == 1 || fmt == 4) ? 8 : 16;
    for (uint32_t y = 0; y < by; y++)
        for (uint32_t x = 0; x < bx; x++) {
            uint8_t px[64];
            for (uint32_t j = 0; j < 4; j++) {
                uint32_t iy = y * 4 + j < h - 1 ? y * 4 + j : h - 1;
                for (uint32_t i = 0; i < 4; i++) {
                    uint32_t ix = x * 4 + i < w - 1 ? x * 4 + i : w - 1;
                    memcpy(px + 4 * (j * 4 + i), rgba + (size_t)iy * pitch + 4 * (size_t)ix, 4);
                }
            }
            uint8_t* e = out + ((size_t)y * bx + x) * bpb;
            switch (fmt) {
            case 0: case 1: put_color(e, px, (uint32_t)fmt, quality, perceptual, both, thresh, tfb); break;
            case 2: put_dxt3(e, px, 3); put_color(e + 8, px, 2, quality, perceptual, both, thresh, tfb); break;
            case 3: put_alpha(e, px, 3, quality, both); put_color(e + 8, px, 3, quality, perceptual, both, thresh, tfb); break;
            case 4: put_alpha(e, px, 3, quality, both); break;
            case 5: put_alpha(e, px, 0, quality, both); put_alpha(e + 8, px, 1, quality, both); break;
            case 6: put_alpha(e, px, 1, quality, both); put_alpha(e + 8, px, 0, quality, both); break;
            default: return 0;
            }
        }
    return bx * by * bpb;
}
