// warp_util.cuh -- warp-level helpers shared by the optimiser kernels (sm_100a).
#pragma once
#include "launch.h"

namespace crn {

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// Lexicographic (key, idx) minimum across the full warp; every lane receives the winner.
// Ties on key resolve to the lowest idx, which is how "first candidate in reference order wins"
// (strict '<' acceptance in crn_dxt1.cpp:1585 / crn_dxt5a.cpp:248) is reproduced after a parallel batch.
__device__ __forceinline__ void warp_argmin_u64(unsigned long long& key, unsigned& idx)
{
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) {
        unsigned long long k2 = __shfl_xor_sync(CRN_FULL_MASK, key, ofs);
        unsigned i2 = __shfl_xor_sync(CRN_FULL_MASK, idx, ofs);
        if (k2 < key || (k2 == key && i2 < idx)) {
            key = k2;
            idx = i2;
        }
    }
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v)
{
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) {
        unsigned long long o = __shfl_xor_sync(CRN_FULL_MASK, v, ofs);
        v = o < v ? o : v;
    }
    return v;
}

// butterfly sums: every lane ends with the same value (each step adds the same two numbers on both lanes of a pair)
__device__ __forceinline__ float warp_sum_f32(float v)
{
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) v += __shfl_xor_sync(CRN_FULL_MASK, v, ofs);
    return v;
}
__device__ __forceinline__ double warp_sum_f64(double v)
{
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) v += __shfl_xor_sync(CRN_FULL_MASK, v, ofs);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1)
        v += __shfl_xor_sync(CRN_FULL_MASK, v, ofs);
    return v;
}

__device__ __forceinline__ unsigned lanemask_lt() { return (1u << lane_id()) - 1u; }

}  // namespace crn
