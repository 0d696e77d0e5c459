// dxt5a_opt.cuh -- one-warp DXT5A (8-bit single channel) endpoint optimiser for sm_100a.
//
// Replaces crnlib::dxt5_endpoint_optimizer::compute / evaluate_solution
// (reference crnlib/crn_dxt5a.cpp:40-196, :198-262).  The reference walks a candidate list serially
// and accepts a candidate when its error is strictly below the running best; parts of the list are
// generated from the *live* best (crn_dxt5a.cpp:114, :127).  Here one lane owns one candidate:
//   phase 1 (all pairs of unique values, :93-103) has no live dependence -> per-lane running minimum
//           and one (error, sequence-number) warp reduction;
//   phase 2 (the +-16 / +-8 window around the live best, :105-149) is evaluated 32 candidates at a
//           time speculatively; the first improving lane is committed and the sequence is replayed
//           from the candidate after it with the updated endpoints, which reproduces the serial
//           result exactly.  The reference's m_flags bitmap only skips re-evaluations that can never
//           win under strict '<' (the 8- and 6-value palettes are symmetric in l<->h), so it is not
//           needed for parity and is omitted.
#pragma once
#include "warp_util.cuh"

namespace crn {

struct Dxt5aScratch {       // per-warp shared memory
    uint32_t wgt[256];      // weight of unique value i (first-appearance order)
    uint8_t val[256];       // unique value i
    uint8_t sel[256];       // final selector of unique value i
};

struct Dxt5aBest {
    unsigned first, second, block_type;
    unsigned long long error;
};

__device__ __forceinline__ void dxt5a_values8(unsigned l, unsigned h, unsigned (&p)[8])
{   // crn_dxt.cpp:418-430
    p[0] = l; p[1] = h;
    p[2] = (l * 6 + h) / 7; p[3] = (l * 5 + h * 2) / 7; p[4] = (l * 4 + h * 3) / 7;
    p[5] = (l * 3 + h * 4) / 7; p[6] = (l * 2 + h * 5) / 7; p[7] = (l + h * 6) / 7;
}
__device__ __forceinline__ void dxt5a_values6(unsigned l, unsigned h, unsigned (&p)[8])
{   // crn_dxt.cpp:404-416
    p[0] = l; p[1] = h;
    p[2] = (l * 4 + h) / 5; p[3] = (l * 3 + h * 2) / 5; p[4] = (l * 2 + h * 3) / 5; p[5] = (l + h * 4) / 5;
    p[6] = 0; p[7] = 255;
}

// Error of candidate (l,h) for this lane.  WRAP reproduces the reference's 32-bit `int` product
// d*d*weight (crn_dxt5a.cpp:225-228) bit-for-bit when a weight is large enough to overflow; for
// <= 33025 pixels the product cannot overflow and the weight is factored out of the inner loop.
template <bool WRAP>
__device__ __forceinline__ void dxt5a_eval(const Dxt5aScratch* sc, int U, unsigned l, unsigned h, bool both,
                                           unsigned long long& err, unsigned& type)
{
    unsigned p8[8], p6[8];
    dxt5a_values8(l, h, p8);
    dxt5a_values6(l, h, p6);
    unsigned long long e8 = 0, e6 = 0;
    for (int i = 0; i < U; i++) {
        const int v = sc->val[i];
        const unsigned w = sc->wgt[i];
        unsigned b8 = 0xffffffffu, b6 = 0xffffffffu;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int d8 = v - (int)p8[j], d6 = v - (int)p6[j];
            unsigned x8 = (unsigned)(d8 * d8), x6 = (unsigned)(d6 * d6);
            if (WRAP) { x8 *= w; x6 *= w; }
            b8 = min(b8, x8);
            b6 = min(b6, x6);
        }
        if (WRAP) { e8 += b8; e6 += b6; }
        else { e8 += b8 * w; e6 += b6 * w; }
    }
    err = e8;
    type = 0;
    if (both && e6 < e8) { err = e6; type = 1; }
}

// Warp-collective search.  sc->val/wgt[0..U) must be visible to the whole warp.  U >= 2.
template <bool WRAP>
__device__ __forceinline__ Dxt5aBest dxt5a_search(const Dxt5aScratch* sc, int U, int quality, bool both)
{
    const unsigned lane = lane_id();
    Dxt5aBest best;
    // ---- phase 1: every pair (i<j) of unique values, i-major (crn_dxt5a.cpp:93-103)
    {
        unsigned long long my_err = ~0ull;
        unsigned my_k = 0xffffffffu, my_type = 0, my_l = 0, my_h = 0;
        const unsigned npairs = (unsigned)(U * (U - 1) / 2);
        int i = 0, j = 1;
        // advance lane to its first pair
        for (unsigned s = 0; s < lane; s++) { if (++j >= U) { i++; j = i + 1; } }
        for (unsigned k = lane; k < npairs; k += 32) {
            unsigned long long e; unsigned ty;
            const unsigned l = sc->val[i], h = sc->val[j];
            dxt5a_eval<WRAP>(sc, U, l, h, both, e, ty);
            if (e < my_err) { my_err = e; my_k = k; my_type = ty; my_l = l; my_h = h; }
            for (int s = 0; s < 32; s++) { if (++j >= U) { i++; j = i + 1; if (i >= U - 1) break; } }
        }
        unsigned long long key = my_err; unsigned idx = my_k;
        warp_argmin_u64(key, idx);
        const unsigned src = idx & 31u;  // pair k is owned by lane k%32
        best.error = key;
        best.first = __shfl_sync(CRN_FULL_MASK, my_l, src);
        best.second = __shfl_sync(CRN_FULL_MASK, my_h, src);
        best.block_type = __shfl_sync(CRN_FULL_MASK, my_type, src);
    }
    // ---- phase 2: probe window around the live best (crn_dxt5a.cpp:105-149)
    if (quality >= 3 && best.error) {
        const int P = (quality == 4) ? 16 : 8;
        const int W = 2 * P + 1;
        int k0 = 0, row_ld = -1000, row_l = 0;
        while (k0 < W * W && best.error) {
            const int ld0 = k0 / W - P;
            if (ld0 != row_ld) { row_ld = ld0; row_l = (int)best.first + ld0; }
            if (row_l > 255) break;                                   // :119-122
            const int k = k0 + (int)lane;
            const int ld = k / W - P, hd = k % W - P;
            const int l = (ld == row_ld) ? row_l : (int)best.first + ld;
            const int h = (int)best.second + hd;
            const bool valid = (k < W * W) && l >= 0 && l <= 255 && h >= 0 && h <= 255;
            unsigned long long e = ~0ull; unsigned ty = 0;
            if (valid) dxt5a_eval<WRAP>(sc, U, (unsigned)l, (unsigned)h, both, e, ty);
            const unsigned m = __ballot_sync(CRN_FULL_MASK, valid && e < best.error);
            if (!m) { k0 += 32; continue; }
            const int t = __ffs((int)m) - 1;
            best.error = __shfl_sync(CRN_FULL_MASK, e, t);
            best.block_type = __shfl_sync(CRN_FULL_MASK, ty, t);
            const int wl = __shfl_sync(CRN_FULL_MASK, l, t), wh = __shfl_sync(CRN_FULL_MASK, h, t);
            const int wld = __shfl_sync(CRN_FULL_MASK, ld, t);
            best.first = (unsigned)wl; best.second = (unsigned)wh;
            row_ld = wld; row_l = wl;                                  // the rest of this row keeps its l
            k0 = k0 + t + 1;
        }
    }
    return best;
}

// Final endpoint ordering + selector assignment for the unique values (crn_dxt5a.cpp:151-184 and the
// first-minimum selector rule of :223-237).  Writes sc->sel[0..U); returns ordered (first, second).
__device__ __forceinline__ void dxt5a_finish(Dxt5aScratch* sc, int U, const Dxt5aBest& best, unsigned& out_first, unsigned& out_second)
{
    const unsigned lane = lane_id();
    unsigned p[8];
    if (best.block_type) dxt5a_values6(best.first, best.second, p);
    else dxt5a_values8(best.first, best.second, p);
    unsigned first = best.first, second = best.second;
    int mode = 0;  // 0 none, 1 zero, 2 six-invert, 3 eight-invert
    if (first == second) mode = 1;
    else if (best.block_type) { if (first > second) mode = 2; }
    else if (first <= second) mode = 3;
    if (mode >= 2) { unsigned t = first; first = second; second = t; }
    for (int i = (int)lane; i < U; i += 32) {
        const int v = sc->val[i];
        unsigned bs = 0, be = 0xffffffffu;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int d = v - (int)p[j];
            unsigned e = (unsigned)(d * d);   // weight is a common positive factor: argmin unchanged (ties keep lowest j)
            if (e < be) { be = e; bs = j; }
        }
        // six: {1,0,5,4,3,2,6,7}  eight: {1,0,7,6,5,4,3,2}   (crn_dxt.cpp:41-42)
        if (mode == 1) bs = 0;
        else if (mode == 2) bs = (bs < 2) ? (bs ^ 1) : (bs < 6 ? 7 - bs : bs);
        else if (mode == 3) bs = (bs < 2) ? (bs ^ 1) : 9 - bs;
        sc->sel[i] = (uint8_t)bs;
    }
    out_first = first;
    out_second = second;
    __syncwarp();
}

// 4x4 block form: lanes 0..15 hold the 16 channel values (pixel index 4y+x).  Returns the packed
// 8-byte DXT5A element (crn_dxt.h:266-361: e0, e1, then 16 x 3-bit selectors LSB-first) on every lane.
__device__ __forceinline__ unsigned long long dxt5a_pack_block(Dxt5aScratch* sc, unsigned value, int quality, bool both)
{
    const unsigned lane = lane_id();
    const unsigned m16 = 0x0000ffffu;
    unsigned uidx = 0;
    int U = 0;
    {
        // unique values in first-appearance order (crn_dxt5a.cpp:58-75)
        unsigned peers = 0;
        if (lane < 16) peers = __match_any_sync(m16, value);
        const bool leader = lane < 16 && (unsigned)(__ffs((int)peers) - 1) == lane;
        const unsigned leaders = __ballot_sync(CRN_FULL_MASK, leader);
        U = __popc(leaders);
        const unsigned my_u = __popc(leaders & lanemask_lt());
        if (leader) { sc->val[my_u] = (uint8_t)value; sc->wgt[my_u] = (unsigned)__popc(peers); }
        const int lead_lane = lane < 16 ? __ffs((int)peers) - 1 : 0;
        uidx = __shfl_sync(CRN_FULL_MASK, my_u, lead_lane);
        __syncwarp();
    }
    unsigned first, second;
    if (U == 1) {   // crn_dxt5a.cpp:77-86
        first = second = sc->val[0];
        if (lane == 0) sc->sel[0] = 0;
        __syncwarp();
    } else {
        Dxt5aBest best = dxt5a_search<false>(sc, U, quality, both);
        dxt5a_finish(sc, U, best, first, second);
    }
    unsigned long long bits = 0;
    if (lane < 16) bits = (unsigned long long)sc->sel[uidx] << (3 * lane);
    // OR-reduce the 48 selector bits
#pragma unroll
    for (int ofs = 8; ofs > 0; ofs >>= 1) bits |= __shfl_xor_sync(CRN_FULL_MASK, bits, ofs);
    bits = __shfl_sync(CRN_FULL_MASK, bits, 0);
    __syncwarp();
    return (unsigned long long)first | ((unsigned long long)second << 8) | (bits << 16);
}

}  // namespace crn
