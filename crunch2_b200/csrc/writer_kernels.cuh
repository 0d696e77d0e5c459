// writer_kernels.cuh -- the O(n^2) palette orderings of the .CRN writer on the device.
//
// crn_comp orders every palette so that neighbouring entries are similar and frequent index transitions become small deltas:
//   * sort_color_endpoints / optimize_color_selectors (crnlib/crn_comp.cpp:767-798, :935-1058): a greedy nearest-neighbour chain --
//     n steps, each a minimum over the entries still unplaced;
//   * optimize_color_endpoints_task (:800-933), three weights: a chain grown at both ends, each step the unplaced entry with the largest
//     max(similarity to front, similarity to back) * (transition frequency to the placed ones + normaliser), put at the end that "pulls" harder.
// Each step depends on the previous one, so the step count stays n; what the device removes is the O(n) scan inside a step: one CTA per
// ordering, the unplaced entries spread over its 1024 threads (<= 8 each, in registers), one block-wide (key, index) reduction per step.
// 8192 entries: ~3 ms per ordering against ~95 ms for the host loop, and the five orderings of a colour palette run as five CTAs of one launch.
// Results are identical to crn_writer.h's host loops (integer arithmetic, the same tie rules), which stay as the path of crn_gpu_crn_write
// (no context, no device) and as the cross-check of tests/test_writer_order_cpu.py.
#pragma once
#include "warp_util.cuh"

namespace crn {

constexpr int kOrderThreads = 1024;                     // default CTA size (CRN_B200_ORDER_THREADS = 256 / 512 picks another instantiation)
constexpr int kOrderMaxN = 8192;

struct OrderColorJob {
    const uint32_t* ep_lo; const uint32_t* ep_hi;        // n endpoints, components expanded to 8 bits: r | g << 8 | b << 16
    const uint32_t* row_start; const uint32_t* col; const uint32_t* cnt;   // transition counts as adjacency lists (crn_writer.h Transitions)
    const uint32_t* selectors;                           // n_sel colour selectors (2 bits per pixel)
    uint32_t n, n_sel, selected;
    uint32_t base[3];                                    // 4000 * (1 + weight) of the three weighted trials
    uint16_t* remap;                                     // [4][n]: trial 0 = greedy chain, 1..3 = weighted
    uint16_t* sel_remap;                                 // [n_sel]
};

struct OrderSmem {
    uint32_t freq[kOrderMaxN];
    uint32_t row_start[kOrderMaxN + 1];                  // weighted trials: the transition lists' row starts (one dependent global load less per step)
    uint32_t data_a[kOrderMaxN], data_b[kOrderMaxN];     // greedy chains: the items by id
    int16_t pos[kOrderMaxN];
    uint16_t chosen[2 * kOrderMaxN + 2];                 // weighted trials: the chain; greedy chains: id by slot
    unsigned long long red[32];
    uint32_t red_f[32], red_b[32];
    uint32_t win_fs, win_bs, win_lo, win_hi;
};

__device__ __forceinline__ uint32_t order_dist3(uint32_t a, uint32_t b)
{
    const int dr = (int)(a & 255u) - (int)(b & 255u), dg = (int)((a >> 8) & 255u) - (int)((b >> 8) & 255u), db = (int)((a >> 16) & 255u) - (int)((b >> 16) & 255u);
    return (uint32_t)(dr * dr + dg * dg + db * db);
}
__device__ __forceinline__ unsigned long long order_warp_max(unsigned long long v)
{
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) { const unsigned long long o = __shfl_xor_sync(CRN_FULL_MASK, v, ofs); v = o > v ? o : v; }
    return v;
}
// every thread gets the maximum of `v` over the CTA; one barrier (the caller's next barrier protects sm->red)
__device__ __forceinline__ unsigned long long order_block_max(OrderSmem* sm, unsigned long long v)
{
    v = order_warp_max(v);
    if (lane_id() == 0) sm->red[threadIdx.x >> 5] = v;
    __syncthreads();
    return order_warp_max(lane_id() < (blockDim.x >> 5) ? sm->red[lane_id()] : 0ull);
}

// remap_color_endpoints of crn_writer.h (optimize_color_endpoints_task, crn_comp.cpp:800-878)
template <int T>
__device__ __forceinline__ void order_weighted_chain(OrderSmem* sm, const OrderColorJob& J, uint32_t base, uint16_t* __restrict__ remap)
{
    constexpr int kOrderThreads = T, kOrderPer = kOrderMaxN / T;
    const unsigned tid = threadIdx.x;
    const uint32_t n = J.n;
    uint32_t lo[kOrderPer], hi[kOrderPer], fs[kOrderPer], bs[kOrderPer];
    unsigned alive = 0;
#pragma unroll
    for (int k = 0; k < kOrderPer; k++) {
        const uint32_t i = tid + k * kOrderThreads;
        lo[k] = hi[k] = fs[k] = bs[k] = 0;
        if (i < n) { lo[k] = J.ep_lo[i]; hi[k] = J.ep_hi[i]; alive |= 1u << k; sm->freq[i] = 0; sm->pos[i] = -1; }
    }
    for (uint32_t i = tid; i <= n; i += kOrderThreads) sm->row_start[i] = J.row_start[i];
    uint32_t selected = J.selected;
    int front = (int)n, back = (int)n;
    uint32_t front_lo = J.ep_lo[selected], front_hi = J.ep_hi[selected], back_lo = front_lo, back_hi = front_hi;
    bool front_updated = true, back_updated = true;
    uint32_t normalizer = 0;
    __syncthreads();
    if (tid == 0) { sm->chosen[front] = (uint16_t)selected; sm->pos[selected] = (int16_t)front; }
    if ((selected % kOrderThreads) == tid) alive &= ~(1u << (selected / kOrderThreads));
    __syncthreads();
    for (uint32_t k = sm->row_start[selected] + tid; k < sm->row_start[selected + 1]; k += kOrderThreads) sm->freq[J.col[k]] += J.cnt[k];
    __syncthreads();
    for (uint32_t left = n - 1; left; left--) {
        // the unplaced entry of largest value; ties: the lowest index (the reference's `value == best && index < selected`)
        unsigned long long key = 0;
#pragma unroll
        for (int k = 0; k < kOrderPer; k++) {
            if (!((alive >> k) & 1u)) continue;
            const uint32_t i = tid + k * kOrderThreads;
            if (front_updated) fs[k] = base - min(4000u, order_dist3(lo[k], front_lo) + order_dist3(hi[k], front_hi));
            if (back_updated) bs[k] = base - min(4000u, order_dist3(lo[k], back_lo) + order_dist3(hi[k], back_hi));
            const uint32_t value = max(fs[k], bs[k]) * (sm->freq[i] + normalizer) + 1u;         // 32-bit product, as the reference's
            const unsigned long long kk = ((unsigned long long)value << 32) | (0xFFFFFFFFu - i);
            key = kk > key ? kk : key;
        }
        key = order_block_max(sm, key);
        selected = 0xFFFFFFFFu - (uint32_t)key;
        if ((selected % kOrderThreads) == tid) {
            const int k = (int)(selected / kOrderThreads);
            uint32_t wfs = 0, wbs = 0, wlo = 0, whi = 0;
#pragma unroll
            for (int q = 0; q < kOrderPer; q++) if (q == k) { wfs = fs[q]; wbs = bs[q]; wlo = lo[q]; whi = hi[q]; }
            sm->win_fs = wfs; sm->win_bs = wbs; sm->win_lo = wlo; sm->win_hi = whi;
            alive &= ~(1u << k);
        }
        // one pass over the winner's transitions: its pull towards either end of the chain (chain_pull) and, for the next step, the
        // frequencies it adds to the entries still unplaced
        uint32_t pf = 0, pb = 0;
        const int L = back - front;
        for (uint32_t k = sm->row_start[selected] + tid; k < sm->row_start[selected + 1]; k += kOrderThreads) {
            const uint32_t c = J.col[k], w = J.cnt[k];
            const int at = sm->pos[c];
            if (at >= 0) {
                const int p = at - front, q = back - at;
                if (L - 2 * p > 0) pf += (uint32_t)(L - 2 * p) * w;
                if (L - 2 * q > 0) pb += (uint32_t)(L - 2 * q) * w;
            }
            sm->freq[c] += w;
        }
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) { pf += __shfl_xor_sync(CRN_FULL_MASK, pf, ofs); pb += __shfl_xor_sync(CRN_FULL_MASK, pb, ofs); }
        if (lane_id() == 0) { sm->red_f[tid >> 5] = pf; sm->red_b[tid >> 5] = pb; }
        __syncthreads();
        pf = lane_id() < (unsigned)(kOrderThreads >> 5) ? sm->red_f[lane_id()] : 0u; pb = lane_id() < (unsigned)(kOrderThreads >> 5) ? sm->red_b[lane_id()] : 0u;
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) { pf += __shfl_xor_sync(CRN_FULL_MASK, pf, ofs); pb += __shfl_xor_sync(CRN_FULL_MASK, pb, ofs); }
        const uint32_t wfs = sm->win_fs, wbs = sm->win_bs;
        normalizer = sm->freq[selected] << 3;
        front_updated = back_updated = false;
        if ((unsigned long long)wfs * pf > (unsigned long long)wbs * pb) { front--; front_lo = sm->win_lo; front_hi = sm->win_hi; front_updated = true; if (tid == 0) { sm->chosen[front] = (uint16_t)selected; sm->pos[selected] = (int16_t)front; } }
        else { back++; back_lo = sm->win_lo; back_hi = sm->win_hi; back_updated = true; if (tid == 0) { sm->chosen[back] = (uint16_t)selected; sm->pos[selected] = (int16_t)back; } }
        // (sm->red / win_* are rewritten only after the next step's first barrier; pos / chosen / freq are read after it)
    }
    __syncthreads();
    for (int i = front + (int)tid; i <= back; i += kOrderThreads) remap[sm->chosen[i]] = (uint16_t)(i - front);
}

// greedy_chain of crn_writer.h: nearest unplaced item to the last placed one, first minimum in the reference's array order (candidates in an
// array, the winner replaced by the last one).  DIST(a_item, b_item, a_cur, b_cur).
template <int T, typename Dist>
__device__ __forceinline__ void order_greedy_chain(OrderSmem* sm, const uint32_t* __restrict__ ga, const uint32_t* __restrict__ gb, uint32_t n, Dist dist, uint16_t* __restrict__ remap)
{
    constexpr int kOrderThreads = T;
    const unsigned tid = threadIdx.x;
    for (uint32_t i = tid; i < n; i += kOrderThreads) { sm->data_a[i] = ga[i]; sm->data_b[i] = gb ? gb[i] : 0u; sm->chosen[i] = (uint16_t)i; }
    uint32_t cur_a = 0, cur_b = 0;
    __syncthreads();
    for (uint32_t left = n; left; left--) {
        unsigned long long key = 0;
        for (uint32_t s = tid; s < left; s += kOrderThreads) {
            const uint32_t id = sm->chosen[s];
            const uint32_t e = dist(sm->data_a[id], sm->data_b[id], cur_a, cur_b);
            // smallest distance, then smallest slot: as a maximum of the complemented pair
            const unsigned long long kk = ((unsigned long long)(0xFFFFFFFFu - e) << 32) | (0xFFFFFFFFu - s);
            key = kk > key ? kk : key;
        }
        key = order_block_max(sm, key);
        const uint32_t best = 0xFFFFFFFFu - (uint32_t)key;
        const uint32_t id = sm->chosen[best];
        cur_a = sm->data_a[id]; cur_b = sm->data_b[id];
        __syncthreads();                                     // everyone has read chosen[best] before the slot is refilled
        if (tid == 0) { remap[id] = (uint16_t)(n - left); sm->chosen[best] = sm->chosen[left - 1]; }
        __syncthreads();
    }
}

// CTA 0: greedy chain of the endpoints (trial 0); CTAs 1..3: the weighted trials; CTA 4: greedy chain of the selectors
template <int T>
__global__ void __launch_bounds__(T) crn_order_color_kernel(OrderColorJob J)
{
    CRN_DYN_SMEM(OrderSmem, sm);
    if (blockIdx.x == 0)
        order_greedy_chain<T>(sm, J.ep_lo, J.ep_hi, J.n, [](uint32_t alo, uint32_t ahi, uint32_t clo, uint32_t chi) { return order_dist3(alo, clo) + order_dist3(ahi, chi); }, J.remap);
    else if (blockIdx.x < 4)
        order_weighted_chain<T>(sm, J, J.base[blockIdx.x - 1], J.remap + (size_t)blockIdx.x * J.n);
    else
        // per-pixel selector distance {0, 5, 14, 10} on the XOR of the 2-bit selectors (crn_comp.cpp:941-953) = 5 b0 + 14 b1 - 9 (b0 & b1)
        order_greedy_chain<T>(sm, J.selectors, nullptr, J.n_sel, [](uint32_t s, uint32_t, uint32_t ref, uint32_t) {
            const uint32_t x = s ^ ref;
            return 5u * (uint32_t)__popc(x & 0x55555555u) + 14u * (uint32_t)__popc(x & 0xAAAAAAAAu) - 9u * (uint32_t)__popc(x & (x >> 1) & 0x55555555u);
        }, J.sel_remap);
}

}  // namespace crn
