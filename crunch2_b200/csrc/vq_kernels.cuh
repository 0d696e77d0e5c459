// vq_kernels.cuh -- top-down binary-split vector quantiser (SURVEY 8(a) row a19, and a13's split step)
// for sm_100a: the clusterizer<VectorType> of the clustered-DDS path, frontier-batched.
//
// Replaces crnlib::clusterizer<V>::generate_codebook / split_node / compute_split_pca /
// compute_split_estimate (reference crnlib/crn_clusterizer.h:65-167, :740-873, :487-600, :449-485) for
// V = vec2F, vec6F, vec16F.  The reference pops the highest-variance leaf from a heap and splits it, one
// node at a time.  A split only looks at the node's own vectors, so here EVERY splittable leaf of the
// current frontier is split in the same round (one round = a fixed sequence of kernels over all vectors),
// and the reference's split ORDER -- which decides where the codebook budget is spent and what
// retrieve_clusters() prunes -- is replayed afterwards on the host from the recorded variances
// (vq_host.h).
//
// Arithmetic.  The reference accumulates its per-node sums in FLOAT, in member order, and every decision
// (split side, heap order) hangs on those roundings, so they are reproduced, not approximated:
//  * all vectors of the path have small integer components (endpoints 0..255, linear selectors 0..7) and
//    integer weights.  Sums of w*v are accumulated exactly in 64-bit integers with atomics (order free,
//    reproducible); while such a sum stays below 2^24 the reference's float accumulation is exact too and the
//    two agree bit for bit.  The few nodes (the top of the tree) whose sums pass 2^24 are re-accumulated
//    in float, sequentially in member order, one lane per accumulator (vq_float_sums_kernel);
//  * the covariance sums have fractional terms and always round: one warp per node accumulates them in
//    float in member order, one lane per matrix entry (vq_covariance_kernel);
//  * ttsum / weights are doubles / integers in the reference and exact here.
#pragma once
#include "warp_util.cuh"

namespace crn {

constexpr unsigned kVqNoSlot = 0xFFFFFFFFu;

template <int D> struct VqSlot {              // split state of one frontier node
    float covar[D * (D + 1) / 2];             // sum (v - c)[x] * ((v - c)[y] * w), x <= y, float in member order
    float fs1[2][D];                          // per side: sum w * v as the reference's float accumulation gives it
    unsigned long long s1[2][D];              // per side: sum w * v, exact
    unsigned long long wsum[2];
    unsigned long long tt[2];                 // per side: sum w * (v . v)
    unsigned long long far_key, opp_key;      // compute_split_estimate fallback: (float bits of dist << 32) | ~position
    float axis[D];
    float child[2][D];
    float centroid[D];
    float var[2];
    float prev_total;
    unsigned node, begin, count;
    int state;                                // 0 iterating, 1 split done, 2 unsplittable
    int loops;
    int mode;                                 // 0 PCA, 1 two vectors, 2 estimate fallback pending
    unsigned long long node_weight;
};

struct VqNodes {                              // SoA node table (capacity 2N + 1)
    unsigned* begin; unsigned* count; int* left; unsigned* flags; float* variance; unsigned long long* weight; float* centroid;   // centroid: [node][D]
};

template <int D> __device__ __forceinline__ void vq_load(const uint8_t* __restrict__ vecs, unsigned id, float (&v)[D])
{
#pragma unroll
    for (int d = 0; d < D; d++) v[d] = (float)vecs[(size_t)id * D + d];
}

// run-head segmented sum inside a warp; `take` bit k says lane + 2^k belongs to the same run
__device__ __forceinline__ unsigned vq_take_mask(unsigned key)
{
    unsigned take = 0;
    const unsigned lane = lane_id();
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const unsigned k2 = __shfl_down_sync(CRN_FULL_MASK, key, 1u << k);
        if (lane + (1u << k) < 32 && k2 == key) take |= 1u << k;
    }
    return take;
}
__device__ __forceinline__ int vq_seg_sum(int val, unsigned take)
{
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const int v2 = __shfl_down_sync(CRN_FULL_MASK, val, 1u << k);
        if (take >> k & 1) val += v2;
    }
    return val;
}
__device__ __forceinline__ bool vq_is_head(unsigned key)
{
    const unsigned kp = __shfl_up_sync(CRN_FULL_MASK, key, 1);
    return key != kVqNoSlot && (lane_id() == 0 || kp != key);
}
__device__ __forceinline__ void vq_add(unsigned long long* p, int v) { if (v) atomicAdd(p, (unsigned long long)(long long)v); }

// Variance as the reference rounds it (crn_clusterizer.h:91, :816-817): ttsum is a double, the centroid sum's
// dot product and the division by the weight are FLOAT operations.  With the integer sums of this path the
// fs1 is the float-accumulated centroid sum (see vq_float_sums_kernel).
template <int D> __device__ __forceinline__ float vq_variance(const float* fs1, unsigned long long w, unsigned long long tt)
{
    float dot = fs1[0] * fs1[0];
    for (int d = 1; d < D; d++) dot += fs1[d] * fs1[d];
    return (float)((double)tt - (double)(dot / (float)w));
}

constexpr int kVqSeqWarps = 4;                // warps per CTA of the warp-per-slot kernels
constexpr unsigned kVqCovBig = 256;           // slots with more members stream through vq_stream_kernel
constexpr int kVqStreamThreads = 256;         // = members per tile

__device__ __forceinline__ void vq_big_append(unsigned* __restrict__ big_count, unsigned* __restrict__ big_list, unsigned slot)
{
    big_list[atomicAdd(big_count, 1u)] = slot;
}

// K1: covariance sums in the reference's order and precision (compute_split_pca, crn_clusterizer.h:496-508;
// threaded_clusterizer::compute_pca, crn_threaded_clusterizer.h:252-266).  One warp per slot, one lane per matrix
// entry (several for D = 16); slots with more than kVqCovBig members are handed to vq_stream_kernel<D, 1>.
template <int D>
__global__ void __launch_bounds__(kVqSeqWarps * 32) vq_covariance_kernel(const uint8_t* __restrict__ vecs, const unsigned* __restrict__ wts, const unsigned* __restrict__ perm,
                                                                        VqSlot<D>* __restrict__ slots, unsigned nslots,
                                                                        unsigned* __restrict__ big_count, unsigned* __restrict__ big_list)
{
    constexpr int P = D * (D + 1) / 2, PER = (P + 31) / 32;
    __shared__ float dv[kVqSeqWarps][32][D + 1], dw[kVqSeqWarps][32][D + 1];
    const unsigned wi = threadIdx.x >> 5, lane = lane_id();
    const unsigned s = blockIdx.x * kVqSeqWarps + wi;
    if (s >= nslots) return;
    VqSlot<D>& sl = slots[s];
    if (sl.mode != 0) return;
    const unsigned begin = sl.begin, count = sl.count;
    if (count > kVqCovBig) { if (lane == 0) vq_big_append(big_count, big_list, s); return; }
    int xa[PER], ya[PER];
    float acc[PER];
#pragma unroll
    for (int k = 0; k < PER; k++) {
        int a = (int)lane + 32 * k, x = 0;
        xa[k] = 0; ya[k] = 0; acc[k] = 0.0f;
        if (a < P) { int rem = a; while (rem >= D - x) { rem -= D - x; x++; } xa[k] = x; ya[k] = x + rem; }
    }
    float c[D];
#pragma unroll
    for (int d = 0; d < D; d++) c[d] = sl.centroid[d];
    for (unsigned base = 0; base < count; base += 32) {
        const unsigned m = base + lane;
        if (m < count) {
            const unsigned id = perm[begin + m];
            const float w = (float)wts[id];
#pragma unroll
            for (int d = 0; d < D; d++) { const float v = (float)vecs[(size_t)id * D + d] - c[d]; dv[wi][lane][d] = v; dw[wi][lane][d] = v * w; }
        }
        __syncwarp();
        const unsigned cnt = count - base < 32u ? count - base : 32u;
        for (unsigned j = 0; j < cnt; j++) {
#pragma unroll
            for (int k = 0; k < PER; k++) acc[k] = acc[k] + dv[wi][j][xa[k]] * dw[wi][j][ya[k]];
        }
        __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < PER; k++) if ((int)lane + 32 * k < P) sl.covar[lane + 32 * k] = acc[k];
}

// Float centroid sums of every slot that is being split: (float) of the exact integer sum while that is below 2^24
// (the reference's running float sum is exact there); slots beyond that are handed to vq_stream_kernel<D, 0>.
// phase 0: after the projection (mode 0 slots), phase 1: after a Lloyd assignment (state 0 slots), 2: every slot.
template <int D>
__global__ void __launch_bounds__(256) vq_float_sums_kernel(VqSlot<D>* __restrict__ slots, unsigned nslots, int phase,
                                                           unsigned* __restrict__ big_count, unsigned* __restrict__ big_list)
{
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    VqSlot<D>& sl = slots[s];
    if (phase == 0 ? sl.mode != 0 : (phase == 1 ? sl.state != 0 : false)) return;
    bool big = false;
    for (int sd = 0; sd < 2; sd++)
        for (int d = 0; d < D; d++) {
            const unsigned long long e = sl.s1[sd][d];
            big |= e >= (1ull << 24);
            sl.fs1[sd][d] = (float)(long long)e;
        }
    if (big) vq_big_append(big_count, big_list, s);
}

// Member-order float accumulation for the few large slots: one CTA per slot streams the members through shared
// memory in tiles.  Per tile: (1) one thread per member stages its vector, (2) all threads expand the staged
// members into one addend per (member, accumulator), (3) one thread per accumulator adds its column up in member
// order -- the only serial part, a bare LDS + FADD chain.  The next tile's gathers are in flight meanwhile.
// MODE 0: fs1[side][d] += w * v[d]           (2D accumulators; the addend is an exact small integer, or 0 for the other side)
// MODE 1: covar[x][y] += (v-c)[x] * ((v-c)[y] * w)
template <int D, int MODE> struct VqStreamCfg {
    static constexpr int P = D * (D + 1) / 2;
    static constexpr int NACC = MODE == 0 ? 2 * D : P;
    static constexpr int T = D == 16 ? 128 : 256;            // members per tile
    static constexpr int ROW = NACC | 1;                     // odd row length: conflict-free column writes
    static constexpr size_t smem_bytes = sizeof(float) * (size_t)T * ROW;
};

template <int D, int MODE>
__global__ void __launch_bounds__(kVqStreamThreads) vq_stream_kernel(const uint8_t* __restrict__ vecs, const unsigned* __restrict__ wts, const unsigned* __restrict__ perm,
                                                                    const uint8_t* __restrict__ side, VqSlot<D>* __restrict__ slots,
                                                                    const unsigned* __restrict__ big_count, const unsigned* __restrict__ big_list)
{
    using Cfg = VqStreamCfg<D, MODE>;
    constexpr int T = Cfg::T, ROW = Cfg::ROW, NACC = Cfg::NACC;
    CRN_DYN_SMEM(float, smem);
    float (*add)[ROW] = reinterpret_cast<float (*)[ROW]>(smem);
    const unsigned tid = threadIdx.x;
    const unsigned nbig = *big_count;
    for (unsigned e = blockIdx.x; e < nbig; e += gridDim.x) {
        VqSlot<D>& sl = slots[big_list[e]];
        const unsigned begin = sl.begin, count = sl.count, tiles = (count + T - 1) / T;
        float c[D];
#pragma unroll
        for (int d = 0; d < D; d++) c[d] = MODE == 1 ? sl.centroid[d] : 0.0f;
        float acc = 0.0f;
        // software pipeline: id of tile t+2, data of tile t+1 in registers
        unsigned id1 = 0, id2 = 0;
        uint8_t vb[D]; unsigned wv = 0, sv = 0;
#pragma unroll
        for (int d = 0; d < D; d++) vb[d] = 0;
        auto load_id = [&](unsigned t) -> unsigned { const unsigned m = t * T + tid; return (tid < (unsigned)T && m < count) ? perm[begin + m] : 0u; };
        auto load_data = [&](unsigned t, unsigned id) {
            const unsigned m = t * T + tid;
            if (tid < (unsigned)T && m < count) {
                wv = wts[id];
                sv = (MODE == 0 && side) ? side[begin + m] : 0u;
#pragma unroll
                for (int d = 0; d < D; d++) vb[d] = vecs[(size_t)id * D + d];
            }
        };
        load_data(0, load_id(0));
        if (tiles > 1) id1 = load_id(1);
        for (unsigned t = 0; t < tiles; t++) {
            // expand this thread's member into its row of addends (indices are compile-time constants)
            if (tid < (unsigned)T) {
                const float w = (float)wv;
                if (MODE == 0) {
#pragma unroll
                    for (int d = 0; d < D; d++) {
                        const float p = (float)vb[d] * w;
                        add[tid][d] = sv == 0 ? p : 0.0f;
                        add[tid][D + d] = sv == 1 ? p : 0.0f;
                    }
                } else {
                    float dv[D];
#pragma unroll
                    for (int d = 0; d < D; d++) dv[d] = (float)vb[d] - c[d];
                    int a = 0;
#pragma unroll
                    for (int x = 0; x < D; x++)
#pragma unroll
                        for (int y = x; y < D; y++, a++) add[tid][a] = dv[x] * (dv[y] * w);
                }
            }
            if (t + 1 < tiles) load_data(t + 1, id1);
            if (t + 2 < tiles) id2 = load_id(t + 2);
            __syncthreads();
            const unsigned cnt = count - t * T < (unsigned)T ? count - t * T : (unsigned)T;
            if (tid < (unsigned)NACC) {
                // batches of 8: the loads of the next batch are issued before the ordered adds of this one
                unsigned j = 0;
                float a0[8];
                if (cnt >= 8) {
#pragma unroll
                    for (int u = 0; u < 8; u++) a0[u] = add[u][tid];
#pragma unroll 2
                    for (; j + 16 <= cnt; j += 8) {
                        float a1[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) a1[u] = add[j + 8 + u][tid];
#pragma unroll
                        for (int u = 0; u < 8; u++) acc = acc + a0[u];         // x + 0.0f == x (MODE 0, other side)
#pragma unroll
                        for (int u = 0; u < 8; u++) a0[u] = a1[u];
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++) acc = acc + a0[u];
                    j += 8;
                }
                for (; j < cnt; j++) acc = acc + add[j][tid];
            }
            __syncthreads();
            id1 = id2;
        }
        if (tid < (unsigned)NACC) {
            if (MODE == 0) sl.fs1[tid / D][tid % D] = acc; else sl.covar[tid] = acc;
        }
    }
}

// Covariance of the large slots, second generation (replaces vq_stream_kernel<D, 1> on the hot path).  The float
// accumulation of one matrix entry is a serial chain by definition of the reference (compute_split_pca adds in member
// order, crn_clusterizer.h:496-508), so the only things to win are (a) keeping that chain at the FADD latency and
// (b) running the independent chains side by side:
//   * warp-specialised CTA: 128 producer threads gather one member each and expand it into addends in shared memory,
//     ADD_WARPS warps own one accumulator per thread and add their column up in member order.  Two tile buffers: the
//     producers fill tile t while the adders consume tile t-1, one barrier per tile.
//   * for D = 16 the 136 chains of a slot are split over S = 4 CTAs (34 each); each computes only its own products.
template <int D> struct VqCovCfg {
    static constexpr int P = D * (D + 1) / 2;
    static constexpr int S = D == 16 ? 4 : 1;
    static constexpr int CP = (P + S - 1) / S;
    static constexpr int T = 128;
    static constexpr int ROW = CP | 1;                       // odd row length: conflict-free column reads and row writes
    static constexpr int ADD_WARPS = (CP + 31) / 32;
    static constexpr int THREADS = T + 32 * ADD_WARPS;
};

template <int D> struct VqPacked { unsigned w[(D + 3) / 4]; };
template <int D> __device__ __forceinline__ VqPacked<D> vq_load_packed(const uint8_t* __restrict__ vecs, unsigned id)
{
    VqPacked<D> r;
    if (D == 16) {
        const uint4 q = *reinterpret_cast<const uint4*>(vecs + (size_t)id * 16);
        r.w[0] = q.x; r.w[1] = q.y; r.w[2 % ((D + 3) / 4)] = q.z; r.w[3 % ((D + 3) / 4)] = q.w;
    } else {
        const unsigned short* h = reinterpret_cast<const unsigned short*>(vecs + (size_t)id * D);      // D even: 2-byte aligned
#pragma unroll
        for (int k = 0; k < (D + 3) / 4; k++) {
            unsigned v = h[2 * k];
            if (4 * k + 2 < D) v |= (unsigned)h[2 * k + 1] << 16;
            r.w[k] = v;
        }
    }
    return r;
}
template <int D> __device__ __forceinline__ float vq_packed_get(const VqPacked<D>& p, int d) { return (float)((p.w[d >> 2] >> (8 * (d & 3))) & 0xffu); }

template <int D, int PART>
__device__ __forceinline__ void vq_cov_expand(const VqPacked<D>& pk, float w, const float (&c)[D], float* __restrict__ row)
{
    constexpr int LO = PART * VqCovCfg<D>::CP, HI = LO + VqCovCfg<D>::CP;
    float dv[D];
#pragma unroll
    for (int d = 0; d < D; d++) dv[d] = vq_packed_get<D>(pk, d) - c[d];
    int a = 0;
#pragma unroll
    for (int x = 0; x < D; x++)
#pragma unroll
        for (int y = x; y < D; y++, a++)
            if (a >= LO && a < HI) row[a - LO] = dv[x] * (dv[y] * w);
}

template <int D>
__global__ void __launch_bounds__(VqCovCfg<D>::THREADS) vq_stream_cov_kernel(const uint8_t* __restrict__ vecs, const unsigned* __restrict__ wts, const unsigned* __restrict__ perm,
                                                                           VqSlot<D>* __restrict__ slots, const unsigned* __restrict__ big_count, const unsigned* __restrict__ big_list)
{
    using Cfg = VqCovCfg<D>;
    constexpr int T = Cfg::T, ROW = Cfg::ROW, S = Cfg::S, CP = Cfg::CP, P = Cfg::P;
    __shared__ float add[2][T][ROW];
    const unsigned tid = threadIdx.x;
    const bool producer = tid < (unsigned)T;
    const unsigned total = *big_count * (unsigned)S;
    for (unsigned e = blockIdx.x; e < total; e += gridDim.x) {
        const unsigned part = e % S;
        VqSlot<D>& sl = slots[big_list[e / S]];
        const unsigned begin = sl.begin, count = sl.count, tiles = (count + T - 1) / T;
        float c[D];
#pragma unroll
        for (int d = 0; d < D; d++) c[d] = sl.centroid[d];
        // producer pipeline (two dependent gathers per member): member data two tiles ahead, ids four tiles ahead
        unsigned id_c = 0xffffffffu, id_d = 0xffffffffu;
        VqPacked<D> v0 = {}, v1 = {};
        unsigned w0 = 0, w1 = 0;
        auto load_id = [&](unsigned t) -> unsigned { const unsigned m = t * T + tid; return (t < tiles && m < count) ? perm[begin + m] : 0xffffffffu; };
        if (producer) {
            const unsigned id_a = load_id(0), id_b = load_id(1);
            id_c = load_id(2); id_d = load_id(3);
            if (id_a != 0xffffffffu) { v0 = vq_load_packed<D>(vecs, id_a); w0 = wts[id_a]; }
            if (id_b != 0xffffffffu) { v1 = vq_load_packed<D>(vecs, id_b); w1 = wts[id_b]; }
        }
        const unsigned a_local = tid - (unsigned)T;                       // adders: accumulator within this part
        const bool adds = !producer && a_local < (unsigned)CP && part * CP + a_local < (unsigned)P;
        float acc = 0.0f;
        for (unsigned t = 0; t <= tiles; t++) {
            if (producer) {
                if (t < tiles) {
                    float* row = add[t & 1][tid];
                    const float w = (float)w0;
                    if (t * T + tid < count) {
                        switch (part) {
                        case 0: vq_cov_expand<D, 0>(v0, w, c, row); break;
                        case 1: if (S > 1) vq_cov_expand<D, (S > 1 ? 1 : 0)>(v0, w, c, row); break;
                        case 2: if (S > 2) vq_cov_expand<D, (S > 2 ? 2 : 0)>(v0, w, c, row); break;
                        default: if (S > 3) vq_cov_expand<D, (S > 3 ? 3 : 0)>(v0, w, c, row); break;
                        }
                    }
                    v0 = v1; w0 = w1;
                    if (id_c != 0xffffffffu) { v1 = vq_load_packed<D>(vecs, id_c); w1 = wts[id_c]; }
                    id_c = id_d;
                    id_d = load_id(t + 4);
                }
            } else if (adds && t > 0) {
                const unsigned tt = t - 1;
                const unsigned cnt = count - tt * T < (unsigned)T ? count - tt * T : (unsigned)T;
                const float (*buf)[ROW] = add[tt & 1];
                unsigned j = 0;
                if (cnt >= 16) {
                    float a0[16];
#pragma unroll
                    for (int u = 0; u < 16; u++) a0[u] = buf[u][a_local];
                    for (; j + 32 <= cnt; j += 16) {
                        float a1[16];
#pragma unroll
                        for (int u = 0; u < 16; u++) a1[u] = buf[j + 16 + u][a_local];
#pragma unroll
                        for (int u = 0; u < 16; u++) acc = acc + a0[u];
#pragma unroll
                        for (int u = 0; u < 16; u++) a0[u] = a1[u];
                    }
#pragma unroll
                    for (int u = 0; u < 16; u++) acc = acc + a0[u];
                    j += 16;
                }
                for (; j < cnt; j++) acc = acc + buf[j][a_local];
            }
            __syncthreads();
        }
        if (adds) sl.covar[part * CP + a_local] = acc;
    }
}

// ---- exact parallel evaluation of a member-order FLOAT sum of non-negative integers ---------------------
// acc <- RN(acc + x_i) for i = 0, 1, ... is sequential, but inside one binade [2^(23+e), 2^(24+e)) the running
// value is a multiple of ulp = 2^e and a step only depends on the PARITY of acc / ulp (ties round to even):
//     x = q * ulp + r;   acc/ulp <- acc/ulp + q + [r > ulp/2] + [r == ulp/2 and (acc/ulp + q) odd].
// So for a chunk of 256 members the map acc -> acc' is "add delta[e][parity] * 2^e" as long as the chunk stays
// inside binade e.  vq_chunk_sim_kernel tabulates delta for every plausible e and both parities, all chunks in
// parallel; vq_chunk_apply_kernel then walks the chunks in order (one table lookup per chunk instead of 256 adds)
// and re-runs the few chunks that cross a binade boundary step by step.  Everything is integer arithmetic, and
// the result is bit-identical to the sequential float accumulation.
constexpr int kVqChunk = 256;                 // members per chunk
constexpr int kVqEMax = 16;                   // exponents 0..16: sums below 2^40
constexpr unsigned kVqMaxBig = 1024;          // big slots per pass handled this way (the rest streams)

struct VqBigDir {                             // written by vq_big_dir_kernel
    unsigned nbig;                            // slots handled by the chunk kernels
    unsigned nchunks;
    unsigned noverflow;                       // slots left to vq_stream_kernel<D, 0>
    unsigned pad;
};

__device__ __forceinline__ int vq_binade(unsigned long long y) { return y < (1ull << 24) ? 0 : 40 - __clzll((long long)y); }

// RN(acc + x) for a float-representable integer acc
__device__ __forceinline__ unsigned long long vq_fl_add(unsigned long long acc, unsigned x)
{
    const unsigned long long y = acc + x;
    if (y < (1ull << 24)) return y;
    const int e = 40 - __clzll((long long)y);
    const unsigned long long ulp = 1ull << e, r = y & (ulp - 1), half = ulp >> 1;
    unsigned long long y0 = y - r;
    if (r > half || (r == half && ((y0 >> e) & 1))) y0 += ulp;
    return y0;
}

// chunk directory: chunk_start[i] = first chunk of big slot i; slots beyond kVqMaxBig / max_chunks overflow
template <int D>
__global__ void vq_big_dir_kernel(const unsigned* __restrict__ big_count, unsigned* __restrict__ big_list, const VqSlot<D>* __restrict__ slots,
                                  unsigned* __restrict__ chunk_start, VqBigDir* __restrict__ dir, unsigned* __restrict__ overflow_count,
                                  unsigned* __restrict__ overflow_list, unsigned max_chunks)
{
    if (threadIdx.x || blockIdx.x) return;
    const unsigned n = *big_count;
    unsigned nb = 0, nc = 0, nov = 0;
    for (unsigned i = 0; i < n; i++) {
        const unsigned s = big_list[i];
        const unsigned c = (slots[s].count + kVqChunk - 1) / kVqChunk;
        if (nb < kVqMaxBig && nc + c <= max_chunks) { big_list[nb] = s; chunk_start[nb] = nc; nb++; nc += c; }
        else overflow_list[nov++] = s;
    }
    chunk_start[nb] = nc;
    dir->nbig = nb; dir->nchunks = nc; dir->noverflow = nov;
    *overflow_count = nov;
}

// addends of 32 members: tile_x[j][d] = w * v[d] (exact), tile_s[j] = side
template <int D>
__device__ __forceinline__ void vq_load_subtile(const uint8_t* __restrict__ vecs, const unsigned* __restrict__ wts, const unsigned* __restrict__ perm,
                                                const uint8_t* __restrict__ side, unsigned pos, unsigned end, unsigned (*tile_x)[D + 1], uint8_t* tile_s)
{
    const unsigned lane = lane_id(), m = pos + lane;
    if (m < end) {
        const unsigned id = perm[m], w = wts[id];
        tile_s[lane] = side ? side[m] : (uint8_t)0;
#pragma unroll
        for (int d = 0; d < D; d++) tile_x[lane][d] = (unsigned)vecs[(size_t)id * D + d] * w;
    }
}

// table layout: [chunk][accumulator 0..2D-1][e 0..kVqEMax][parity]
template <int D> __device__ __forceinline__ size_t vq_tab_index(unsigned chunk, unsigned a, int e, unsigned p)
{
    return (((size_t)chunk * (2 * D) + a) * (kVqEMax + 1) + (unsigned)e) * 2 + p;
}

template <int D>
__global__ void __launch_bounds__(kVqSeqWarps * 32) vq_chunk_sim_kernel(const uint8_t* __restrict__ vecs, const unsigned* __restrict__ wts, const unsigned* __restrict__ perm,
                                                                       const uint8_t* __restrict__ side, const VqSlot<D>* __restrict__ slots,
                                                                       const unsigned* __restrict__ big_list, const unsigned* __restrict__ chunk_start,
                                                                       const VqBigDir* __restrict__ dir, unsigned* __restrict__ table)
{
    __shared__ unsigned tile_x[kVqSeqWarps][32][D + 1];
    __shared__ uint8_t tile_s[kVqSeqWarps][32];
    const unsigned wi = threadIdx.x >> 5, lane = lane_id();
    const unsigned chunk = blockIdx.x * kVqSeqWarps + wi;
    if (chunk >= dir->nchunks) return;
    unsigned lo = 0, hi = dir->nbig;                     // big slot that owns the chunk
    while (hi - lo > 1) { const unsigned mid = (lo + hi) >> 1; if (chunk_start[mid] <= chunk) lo = mid; else hi = mid; }
    const VqSlot<D>& sl = slots[big_list[lo]];
    const unsigned k = chunk - chunk_start[lo];
    const unsigned pos0 = sl.begin + k * kVqChunk, end = sl.begin + sl.count;
    const unsigned stop = pos0 + kVqChunk < end ? pos0 + kVqChunk : end;
    const bool mine = lane < 2 * D;
    const unsigned my_side = lane / D, my_d = lane % D;
    // exponents worth tabulating: up to one past the binade of the slot's largest exact sum
    unsigned long long mx = mine ? sl.s1[my_side][my_d] : 0ull;
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) { const unsigned long long o = __shfl_xor_sync(CRN_FULL_MASK, mx, ofs); mx = o > mx ? o : mx; }
    int emax = vq_binade(mx) + 1;
    emax = emax > kVqEMax ? kVqEMax : emax;
    unsigned A[kVqEMax + 1][2];
#pragma unroll
    for (int e = 0; e <= kVqEMax; e++) { A[e][0] = 0; A[e][1] = 1; }
    unsigned sum = 0;
    for (unsigned pos = pos0; pos < stop; pos += 32) {
        vq_load_subtile<D>(vecs, wts, perm, side, pos, stop, tile_x[wi], tile_s[wi]);
        __syncwarp();
        const unsigned cnt = stop - pos < 32u ? stop - pos : 32u;
        if (mine)
            for (unsigned j = 0; j < cnt; j++) {
                if (tile_s[wi][j] != my_side) continue;
                const unsigned x = tile_x[wi][j][my_d];
                sum += x;
#pragma unroll
                for (int e = 1; e <= kVqEMax; e++) {
                    if (e <= emax) {
                        const unsigned q = x >> e, r = x & ((1u << e) - 1), half = 1u << (e - 1);
#pragma unroll
                        for (int p = 0; p < 2; p++) {
                            const unsigned t = A[e][p] + q;
                            A[e][p] = t + ((r > half) | ((r == half) & (t & 1u)));
                        }
                    }
                }
            }
        __syncwarp();
    }
    if (mine) {
        table[vq_tab_index<D>(chunk, lane, 0, 0)] = sum;
        table[vq_tab_index<D>(chunk, lane, 0, 1)] = sum;
#pragma unroll
        for (int e = 1; e <= kVqEMax; e++)
            if (e <= emax) { table[vq_tab_index<D>(chunk, lane, e, 0)] = A[e][0]; table[vq_tab_index<D>(chunk, lane, e, 1)] = A[e][1] - 1u; }
    }
}

template <int D>
__global__ void __launch_bounds__(kVqSeqWarps * 32) vq_chunk_apply_kernel(const uint8_t* __restrict__ vecs, const unsigned* __restrict__ wts, const unsigned* __restrict__ perm,
                                                                         const uint8_t* __restrict__ side, VqSlot<D>* __restrict__ slots,
                                                                         const unsigned* __restrict__ big_list, const unsigned* __restrict__ chunk_start,
                                                                         const VqBigDir* __restrict__ dir, const unsigned* __restrict__ table)
{
    __shared__ unsigned tile_x[kVqSeqWarps][32][D + 1];
    __shared__ uint8_t tile_s[kVqSeqWarps][32];
    const unsigned wi = threadIdx.x >> 5, lane = lane_id();
    const unsigned b = blockIdx.x * kVqSeqWarps + wi;
    if (b >= dir->nbig) return;
    VqSlot<D>& sl = slots[big_list[b]];
    const unsigned c0 = chunk_start[b], nchunks = chunk_start[b + 1] - c0;
    const unsigned end = sl.begin + sl.count;
    const bool mine = lane < 2 * D;
    const unsigned my_side = lane / D, my_d = lane % D;
    unsigned long long mx = mine ? sl.s1[my_side][my_d] : 0ull;
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) { const unsigned long long o = __shfl_xor_sync(CRN_FULL_MASK, mx, ofs); mx = o > mx ? o : mx; }
    int emax = vq_binade(mx) + 1;
    emax = emax > kVqEMax ? kVqEMax : emax;
    unsigned long long acc = 0;
    for (unsigned k0 = 0; k0 < nchunks; k0 += 8) {
        // both parities of up to 8 chunks at the current binade, fetched before they are needed
        const int e0 = vq_binade(acc);
        unsigned d0[8], d1[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            d0[u] = d1[u] = 0;
            if (mine && k0 + u < nchunks && e0 <= emax) {
                const uint2 v = *reinterpret_cast<const uint2*>(&table[vq_tab_index<D>(c0 + k0 + u, lane, e0, 0)]);
                d0[u] = v.x; d1[u] = v.y;
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (k0 + u >= nchunks) break;                       // warp-uniform
            bool ok = true;
            unsigned long long next = acc;
            if (mine) {
                const int e = vq_binade(acc);
                ok = e <= emax;
                if (ok) {
                    unsigned lo = d0[u], hi = d1[u];
                    if (e != e0) { const uint2 v = *reinterpret_cast<const uint2*>(&table[vq_tab_index<D>(c0 + k0 + u, lane, e, 0)]); lo = v.x; hi = v.y; }
                    const unsigned dlt = ((acc >> e) & 1ull) ? hi : lo;
                    next = acc + ((unsigned long long)dlt << e);
                    ok = next < (1ull << (24 + e));               // the whole chunk stayed inside the binade
                }
            }
            if (__all_sync(CRN_FULL_MASK, ok)) { acc = next; continue; }
            // some accumulator crosses a binade boundary inside this chunk: every lane replays it step by step
            const unsigned pos0 = sl.begin + (k0 + u) * kVqChunk;
            const unsigned stop = pos0 + kVqChunk < end ? pos0 + kVqChunk : end;
            // (the float unit does the rounding here: acc is a float-representable integer < 2^63, the addends are < 2^24)
            float accf = (float)(long long)acc;
            for (unsigned pos = pos0; pos < stop; pos += 32) {
                vq_load_subtile<D>(vecs, wts, perm, side, pos, stop, tile_x[wi], tile_s[wi]);
                __syncwarp();
                const unsigned cnt = stop - pos < 32u ? stop - pos : 32u;
                if (mine)
                    for (unsigned j = 0; j < cnt; j++) accf += tile_s[wi][j] == my_side ? (float)tile_x[wi][j][my_d] : 0.0f;
                __syncwarp();
            }
            acc = (unsigned long long)accf;
        }
    }
    if (mine) sl.fs1[my_side][my_d] = (float)(long long)acc;       // exactly representable
}

// K2: covariance -> principal axis by power iteration (compute_split_pca, crn_clusterizer.h:510-579; presplit:
// threaded_clusterizer::compute_pca, crn_threaded_clusterizer.h:268-329, which scales by a double reciprocal)
template <int D>
__global__ void vq_axis_kernel(VqSlot<D>* __restrict__ slots, unsigned nslots, int presplit)
{
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    VqSlot<D>& sl = slots[s];
    if (sl.mode != 0) return;
    float covar[D][D];
    const float inv_w = 1.0f / (float)sl.node_weight;
    const double inv_wd = 1.0 / (double)sl.node_weight;
    int k = 0;
    for (int x = 0; x < D; x++)
        for (int y = x; y < D; y++, k++) {
            covar[x][y] = presplit ? (float)((double)sl.covar[k] * inv_wd) : sl.covar[k] * inv_w;
            covar[y][x] = covar[x][y];
        }
    float axis[D], prev[D];
    for (int i = 0; i < D; i++) {
        axis[i] = D == 1 ? 1.0f : .75f + (1.25f - .75f) * ((float)i * (1.0f / (float)(D - 1 > 1 ? D - 1 : 1)));
        prev[i] = axis[i];
    }
    for (int iter = 0; iter < 10; iter++) {
        float x[D];
        double max_sum = 0;
        for (int i = 0; i < D; i++) {
            double sum = 0;
            for (int j = 0; j < D; j++) sum += (double)(axis[j] * covar[i][j]);
            x[i] = (float)sum;
            max_sum = max_sum > fabs(sum) ? max_sum : fabs(sum);
        }
        if (max_sum != 0.0) { const float m = (float)(1.0 / max_sum); for (int i = 0; i < D; i++) x[i] *= m; }
        float nrm = 0;
        for (int i = 0; i < D; i++) { const float dlt = prev[i] - x[i]; nrm += dlt * dlt; }
        for (int i = 0; i < D; i++) { prev[i] = axis[i]; axis[i] = x[i]; }
        if (nrm < .0025f) break;
    }
    {   // vec::normalize (crn_vec.h:674-689)
        double nn = (double)(axis[0] * axis[0]);
        for (int i = 1; i < D; i++) nn += (double)(axis[i] * axis[i]);
        if (nn != 0) { const float sc = (float)(1.0 / sqrt(nn)); for (int i = 0; i < D; i++) axis[i] *= sc; }
    }
    for (int i = 0; i < D; i++) sl.axis[i] = axis[i];
}

// accumulate (w v, w, w v.v) of the vector at position i into side `side` of its slot
template <int D>
__device__ __forceinline__ void vq_accumulate_side(VqSlot<D>* __restrict__ slots, unsigned slot, unsigned take, bool head,
                                                   const int (&v)[D], int w, int side, bool with_tt)
{
#pragma unroll
    for (int sd = 0; sd < 2; sd++) {
        const int m = (slot != kVqNoSlot && side == sd) ? w : 0;
#pragma unroll
        for (int d = 0; d < D; d++) {
            const int s = vq_seg_sum(m * v[d], take);
            if (head) vq_add(&slots[slot].s1[sd][d], s);
        }
        const int sw = vq_seg_sum(m, take);
        if (head) vq_add(&slots[slot].wsum[sd], sw);
        if (with_tt) {
            int vv = 0;
#pragma unroll
            for (int d = 0; d < D; d++) vv += v[d] * v[d];
            const int st = vq_seg_sum(m * vv, take);
            if (head) vq_add(&slots[slot].tt[sd], st);
        }
    }
}

// K3: initial division by the sign of the projection onto the axis (crn_clusterizer.h:577-597)
template <int D>
__global__ void __launch_bounds__(256) vq_project_kernel(const uint8_t* __restrict__ vecs, const unsigned* __restrict__ wts, const unsigned* __restrict__ perm,
                                                         const unsigned* __restrict__ pos_slot, VqSlot<D>* __restrict__ slots, uint8_t* __restrict__ side_out,
                                                         unsigned n, int presplit)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned slot = i < n ? pos_slot[i] : kVqNoSlot;
    if (slot != kVqNoSlot && slots[slot].mode != 0) slot = kVqNoSlot;
    const unsigned take = vq_take_mask(slot);
    const bool head = vq_is_head(slot);
    if (!__any_sync(CRN_FULL_MASK, slot != kVqNoSlot)) return;
    int v[D]; int w = 0, side = 0;
#pragma unroll
    for (int d = 0; d < D; d++) v[d] = 0;
    if (slot != kVqNoSlot) {
        const unsigned id = perm[i];
        w = (int)wts[id];
        const VqSlot<D>& sl = slots[slot];
        float t = 0;
#pragma unroll
        for (int d = 0; d < D; d++) {
            v[d] = vecs[(size_t)id * D + d];
            const float df = (float)v[d] - sl.centroid[d];
            if (d == 0) t = df * sl.axis[0]; else t += df * sl.axis[d];
        }
        side = ((double)t < 0.0) ? 0 : 1;
        side_out[i] = (uint8_t)side;
    }
    vq_accumulate_side<D>(slots, slot, take, head, v, w, side, presplit != 0);
}

// threaded_clusterizer::compute_split (crn_threaded_clusterizer.h:335-369): the division itself is the split, no
// Lloyd iterations; the two sides become roots, with the statistics generate_codebook() gives a root (:77-93)
template <int D>
__global__ void vq_presplit_children_kernel(VqSlot<D>* __restrict__ slots, unsigned nslots, int children_are_pca_nodes)
{
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    VqSlot<D>& sl = slots[s];
    for (int sd = 0; sd < 2; sd++) {
        sl.var[sd] = 0;
        if (!sl.wsum[sd]) { for (int d = 0; d < D; d++) sl.child[sd][d] = 0; continue; }
        if (children_are_pca_nodes) {      // compute_pca's centroid (crn_threaded_clusterizer.h:243-247)
            const double inv = 1.0 / (double)sl.wsum[sd];
            for (int d = 0; d < D; d++) sl.child[sd][d] = (float)((double)sl.fs1[sd][d] * inv);
        } else {                           // a clusterizer root (crn_clusterizer.h:91-93)
            sl.var[sd] = vq_variance<D>(sl.fs1[sd], sl.wsum[sd], sl.tt[sd]);
            const float inv = 1.0f / (float)sl.wsum[sd];
            for (int d = 0; d < D; d++) sl.child[sd][d] = sl.fs1[sd][d] * inv;
        }
    }
    sl.state = 1;
}

// K4: children estimates from the division, or flag the furthest/opposite fallback
template <int D>
__global__ void vq_children_kernel(const uint8_t* __restrict__ vecs, const unsigned* __restrict__ perm, VqSlot<D>* __restrict__ slots, unsigned nslots)
{
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    VqSlot<D>& sl = slots[s];
    if (sl.mode == 1) {            // exactly two vectors (:489-494)
        for (int d = 0; d < D; d++) { sl.child[0][d] = (float)vecs[(size_t)perm[sl.begin] * D + d]; sl.child[1][d] = (float)vecs[(size_t)perm[sl.begin + 1] * D + d]; }
    } else if (sl.wsum[0] && sl.wsum[1]) {
        const float il = (float)(1.0 / (double)sl.wsum[0]), ir = (float)(1.0 / (double)sl.wsum[1]);
        for (int d = 0; d < D; d++) { sl.child[0][d] = sl.fs1[0][d] * il; sl.child[1][d] = sl.fs1[1][d] * ir; }
    } else sl.mode = 2;
    for (int d = 0; d < D; d++) { sl.s1[0][d] = 0; sl.s1[1][d] = 0; }
    sl.wsum[0] = sl.wsum[1] = 0; sl.tt[0] = sl.tt[1] = 0;
    sl.far_key = 0; sl.opp_key = 0;
}

// K4b/K4c: compute_split_estimate (:449-485): furthest vector from the centroid, then the vector furthest from it
template <int D, int PASS>
__global__ void __launch_bounds__(256) vq_estimate_kernel(const uint8_t* __restrict__ vecs, const unsigned* __restrict__ perm,
                                                          const unsigned* __restrict__ pos_slot, VqSlot<D>* __restrict__ slots, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned slot = pos_slot[i];
    if (slot == kVqNoSlot || slots[slot].mode != 2) return;
    VqSlot<D>& sl = slots[slot];
    float ref[D];
    if (PASS == 0) { for (int d = 0; d < D; d++) ref[d] = sl.centroid[d]; }
    else { const unsigned fp = ~(unsigned)(sl.far_key & 0xffffffffu); for (int d = 0; d < D; d++) ref[d] = (float)vecs[(size_t)perm[fp] * D + d]; }
    float dist = 0;
    for (int d = 0; d < D; d++) { const float df = (float)vecs[(size_t)perm[i] * D + d] - ref[d]; dist += df * df; }
    const unsigned long long key = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned long long)(~i);
    atomicMax(PASS == 0 ? &sl.far_key : &sl.opp_key, key);
}
template <int D>
__global__ void vq_estimate_finish_kernel(const uint8_t* __restrict__ vecs, const unsigned* __restrict__ perm, VqSlot<D>* __restrict__ slots, unsigned nslots)
{
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    VqSlot<D>& sl = slots[s];
    if (sl.mode != 2) return;
    const unsigned fp = ~(unsigned)(sl.far_key & 0xffffffffu), op = ~(unsigned)(sl.opp_key & 0xffffffffu);
    for (int d = 0; d < D; d++) {
        sl.child[0][d] = ((float)vecs[(size_t)perm[fp] * D + d] + sl.centroid[d]) * .5f;
        sl.child[1][d] = ((float)vecs[(size_t)perm[op] * D + d] + sl.centroid[d]) * .5f;
    }
    sl.mode = 0;
}

// K5: one Lloyd iteration, assignment half (split_node loop body, :780-812)
template <int D>
__global__ void __launch_bounds__(256) vq_assign_kernel(const uint8_t* __restrict__ vecs, const unsigned* __restrict__ wts, const unsigned* __restrict__ perm,
                                                        const unsigned* __restrict__ pos_slot, VqSlot<D>* __restrict__ slots, uint8_t* __restrict__ side_out, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned slot = i < n ? pos_slot[i] : kVqNoSlot;
    if (slot != kVqNoSlot && slots[slot].state != 0) slot = kVqNoSlot;
    const unsigned take = vq_take_mask(slot);
    const bool head = vq_is_head(slot);
    if (!__any_sync(CRN_FULL_MASK, slot != kVqNoSlot)) return;
    int v[D]; int w = 0, side = 0;
#pragma unroll
    for (int d = 0; d < D; d++) v[d] = 0;
    if (slot != kVqNoSlot) {
        const unsigned id = perm[i];
        w = (int)wts[id];
        const VqSlot<D>& sl = slots[slot];
        float dl = 0, dr = 0;
#pragma unroll
        for (int d = 0; d < D; d++) {
            v[d] = vecs[(size_t)id * D + d];
            const float a = sl.child[0][d] - (float)v[d], b = sl.child[1][d] - (float)v[d];
            dl += a * a; dr += b * b;
        }
        side = ((double)dl < (double)dr) ? 0 : 1;
        side_out[i] = (uint8_t)side;
    }
    vq_accumulate_side<D>(slots, slot, take, head, v, w, side, true);
}

// K6: Lloyd iteration, update half + convergence tests (:814-846)
template <int D>
__global__ void vq_update_kernel(VqSlot<D>* __restrict__ slots, unsigned nslots, unsigned* __restrict__ active_count)
{
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    VqSlot<D>& sl = slots[s];
    if (sl.state != 0) return;
    if (!sl.wsum[0] || !sl.wsum[1]) { sl.state = 2; return; }       // unsplittable (:814-818)
    float var[2];
    for (int sd = 0; sd < 2; sd++) {
        var[sd] = vq_variance<D>(sl.fs1[sd], sl.wsum[sd], sl.tt[sd]);
        const float inv = 1.0f / (float)sl.wsum[sd];
        for (int d = 0; d < D; d++) sl.child[sd][d] = sl.fs1[sd][d] * inv;
    }
    sl.var[0] = var[0]; sl.var[1] = var[1];
    const float total = var[0] + var[1];
    sl.loops++;
    bool done = false;
    if (total < .00001f) done = true;
    else if (((sl.prev_total - total) / total) < .00125f) done = true;
    else if (sl.loops >= 8) done = true;
    sl.prev_total = total;
    if (done) { sl.state = 1; return; }
    // another iteration: clear the accumulators (child weights are re-accumulated)
    for (int d = 0; d < D; d++) { sl.s1[0][d] = 0; sl.s1[1][d] = 0; }
    sl.wsum[0] = sl.wsum[1] = 0; sl.tt[0] = sl.tt[1] = 0;
    atomicAdd(active_count, 1u);
}

// K7: flags for the stable partition: 1 where the element moves to the LEFT child of a node that was split
__global__ void __launch_bounds__(256) vq_left_flags_kernel(const unsigned* __restrict__ pos_slot, const int* __restrict__ slot_state, const uint8_t* __restrict__ side,
                                                            unsigned* __restrict__ flags, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned slot = pos_slot[i];
    flags[i] = (slot != kVqNoSlot && slot_state[slot] == 1 && side[i] == 0) ? 1u : 0u;
}

// exclusive scan of 32-bit values: per-block (1024 elements) totals, scan of totals, final pass
__global__ void __launch_bounds__(256) vq_scan_block_kernel(const unsigned* __restrict__ in, unsigned* __restrict__ out, unsigned* __restrict__ block_sums, unsigned n)
{
    __shared__ unsigned warp_tot[8];
    const unsigned base = blockIdx.x * 1024u + threadIdx.x * 4u;
    unsigned v[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { v[k] = base + k < n ? in[base + k] : 0u; s += v[k]; }
    unsigned incl = s;
#pragma unroll
    for (int ofs = 1; ofs < 32; ofs <<= 1) { const unsigned t = __shfl_up_sync(CRN_FULL_MASK, incl, ofs); if ((int)lane_id() >= ofs) incl += t; }
    if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned woff = 0;
    for (unsigned w = 0; w < (threadIdx.x >> 5); w++) woff += warp_tot[w];
    unsigned run = woff + incl - s;
#pragma unroll
    for (int k = 0; k < 4; k++) { if (base + k < n) out[base + k] = run; run += v[k]; }
    if (threadIdx.x == 255) block_sums[blockIdx.x] = woff + incl;
}
__global__ void vq_scan_sums_kernel(unsigned* __restrict__ block_sums, unsigned nblocks)
{   // single thread block, serial over <= a few thousand entries per thread chunk
    __shared__ unsigned part[256];
    const unsigned per = (nblocks + 255) / 256, b = threadIdx.x * per, e = umin(b + per, nblocks);
    unsigned s = 0;
    for (unsigned i = b; i < e; i++) s += block_sums[i];
    part[threadIdx.x] = s;
    __syncthreads();
    unsigned off = 0;
    for (unsigned t = 0; t < threadIdx.x; t++) off += part[t];
    for (unsigned i = b; i < e; i++) { const unsigned v = block_sums[i]; block_sums[i] = off; off += v; }
}
__global__ void __launch_bounds__(256) vq_scan_add_kernel(unsigned* __restrict__ out, const unsigned* __restrict__ block_sums, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += block_sums[i >> 10];
}

// K8: create the two children of every split slot
template <int D>
__global__ void vq_finalize_kernel(VqSlot<D>* __restrict__ slots, unsigned nslots, const unsigned* __restrict__ left_scan, unsigned n, VqNodes nodes,
                                   unsigned* __restrict__ node_counter)
{
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    VqSlot<D>& sl = slots[s];
    if (sl.state == 2) { nodes.flags[sl.node] |= 1u; return; }
    if (sl.state != 1) return;
    const unsigned b = sl.begin, c = sl.count;
    const unsigned nleft = left_scan[b + c] - left_scan[b];                 // left_scan has n + 1 entries
    const unsigned child = atomicAdd(node_counter, 2u);
    nodes.left[sl.node] = (int)child;
    for (int sd = 0; sd < 2; sd++) {
        const unsigned id = child + sd;
        nodes.begin[id] = sd ? b + nleft : b;
        nodes.count[id] = sd ? c - nleft : nleft;
        nodes.left[id] = -1;
        nodes.flags[id] = 0;
        nodes.variance[id] = sl.var[sd];
        nodes.weight[id] = sl.wsum[sd];
        for (int d = 0; d < D; d++) nodes.centroid[(size_t)id * D + d] = sl.child[sd][d];
    }
}

// K9: scatter into the partitioned order; elements of nodes that were not split keep their place
template <int D>
__global__ void __launch_bounds__(256) vq_scatter_kernel(const unsigned* __restrict__ perm, unsigned* __restrict__ perm_out, const unsigned* __restrict__ pos_slot,
                                                         const VqSlot<D>* __restrict__ slots, const uint8_t* __restrict__ side, const unsigned* __restrict__ left_scan, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned slot = pos_slot[i];
    unsigned dst = i;
    if (slot != kVqNoSlot && slots[slot].state == 1) {
        const unsigned b = slots[slot].begin, c = slots[slot].count;
        const unsigned lb = left_scan[b], li = left_scan[i];
        const unsigned nleft = left_scan[b + c] - lb;
        dst = side[i] == 0 ? b + (li - lb) : b + nleft + ((i - b) - (li - lb));
    }
    perm_out[dst] = perm[i];
}

// K0: split state of the frontier nodes the host selected (slot order = ascending first position)
template <int D>
__global__ void vq_init_slots_kernel(const unsigned* __restrict__ slot_node, VqNodes nodes, VqSlot<D>* __restrict__ slots, unsigned* __restrict__ slot_starts,
                                     unsigned nslots, int presplit)
{
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    const unsigned id = slot_node[s];
    VqSlot<D>& sl = slots[s];
    for (int k = 0; k < D * (D + 1) / 2; k++) sl.covar[k] = 0;
    for (int d = 0; d < D; d++) { sl.s1[0][d] = 0; sl.s1[1][d] = 0; sl.fs1[0][d] = 0; sl.fs1[1][d] = 0; sl.centroid[d] = nodes.centroid[(size_t)id * D + d]; sl.axis[d] = 0; sl.child[0][d] = 0; sl.child[1][d] = 0; }
    sl.wsum[0] = sl.wsum[1] = 0; sl.tt[0] = sl.tt[1] = 0; sl.far_key = 0; sl.opp_key = 0;
    sl.var[0] = sl.var[1] = 0; sl.prev_total = 1e+10f;
    sl.node = id; sl.begin = nodes.begin[id]; sl.count = nodes.count[id];
    sl.state = 0; sl.loops = 0; sl.mode = (!presplit && nodes.count[id] == 2) ? 1 : 0;
    sl.node_weight = nodes.weight[id];
    slot_starts[s] = sl.begin;
}
// every position learns its slot: binary search over the slots' first positions
template <int D>
__global__ void __launch_bounds__(256) vq_pos_slot_kernel(const unsigned* __restrict__ slot_starts, const VqSlot<D>* __restrict__ slots, unsigned nslots,
                                                          unsigned* __restrict__ pos_slot, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned lo = 0, hi = nslots;           // last slot with start <= i
    while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (slot_starts[mid] <= i) lo = mid + 1; else hi = mid; }
    unsigned res = kVqNoSlot;
    if (lo > 0) { const unsigned s = lo - 1; if (i < slots[s].begin + slots[s].count) res = s; }
    pos_slot[i] = res;
}
template <int D>
__global__ void vq_slot_states_kernel(const VqSlot<D>* __restrict__ slots, unsigned nslots, int* __restrict__ states)
{
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < nslots) states[s] = slots[s].state;
}

struct VqSlotResult { int state; unsigned child; unsigned left_count, right_count; float var_left, var_right; };
template <int D>
__global__ void vq_export_kernel(const VqSlot<D>* __restrict__ slots, unsigned nslots, VqNodes nodes, VqSlotResult* __restrict__ out)
{
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    const VqSlot<D>& sl = slots[s];
    VqSlotResult r = {sl.state, 0u, 0u, 0u, 0.0f, 0.0f};
    if (sl.state == 1) {
        r.child = (unsigned)nodes.left[sl.node];
        r.left_count = nodes.count[r.child]; r.right_count = nodes.count[r.child + 1];
        r.var_left = sl.var[0]; r.var_right = sl.var[1];
    }
    out[s] = r;
}

// The root goes through the same float-sum machinery as a one-sided pseudo slot (slot 0, every member on side 0).
template <int D>
__global__ void vq_root_prepare_kernel(const unsigned long long* __restrict__ acc, VqSlot<D>* __restrict__ slots, unsigned n)
{
    if (threadIdx.x || blockIdx.x) return;
    VqSlot<D>& sl = slots[0];
    for (int d = 0; d < D; d++) { sl.s1[0][d] = acc[d]; sl.s1[1][d] = 0; sl.fs1[0][d] = 0; sl.fs1[1][d] = 0; }
    sl.begin = 0; sl.count = n; sl.mode = 0; sl.state = 0; sl.node = 0;
}
// the root node: centroid, weight and variance (generate_codebook, :77-93; with presplit the root is
// threaded_clusterizer::compute_pca's node, whose centroid is scaled by a double reciprocal)
template <int D>
__global__ void vq_root_finish_kernel(const unsigned long long* __restrict__ acc, const VqSlot<D>* __restrict__ slots, VqNodes nodes, unsigned n, int presplit)
{
    if (threadIdx.x || blockIdx.x) return;
    const float* fs = slots[0].fs1[0];
    const unsigned long long W = acc[D];
    nodes.begin[0] = 0; nodes.count[0] = n; nodes.left[0] = -1; nodes.flags[0] = 0; nodes.weight[0] = W;
    nodes.variance[0] = W ? vq_variance<D>(fs, W, acc[D + 1]) : 0.0f;
    if (presplit) { const double inv = W ? 1.0 / (double)W : 0.0; for (int d = 0; d < D; d++) nodes.centroid[d] = (float)((double)fs[d] * inv); }
    else { const float inv = W ? 1.0f / (float)W : 0.0f; for (int d = 0; d < D; d++) nodes.centroid[d] = fs[d] * inv; }
}

// root statistics: sum w v, sum w, sum w v.v over all vectors (generate_codebook, :77-92)
template <int D>
__global__ void __launch_bounds__(256) vq_root_kernel(const uint8_t* __restrict__ vecs, const unsigned* __restrict__ wts, const unsigned* __restrict__ ids, unsigned n,
                                                      unsigned long long* __restrict__ acc /* D + 2 */)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    int v[D]; int w = 0;
#pragma unroll
    for (int d = 0; d < D; d++) v[d] = 0;
    if (i < n) { const unsigned id = ids[i]; w = (int)wts[id]; for (int d = 0; d < D; d++) v[d] = vecs[(size_t)id * D + d]; }
    int vv = 0;
#pragma unroll
    for (int d = 0; d < D; d++) {
        vv += v[d] * v[d];
        const int s = __reduce_add_sync(CRN_FULL_MASK, w * v[d]);
        if (lane_id() == 0) vq_add(&acc[d], s);
    }
    const int sw = __reduce_add_sync(CRN_FULL_MASK, w), st = __reduce_add_sync(CRN_FULL_MASK, w * vv);
    if (lane_id() == 0) { vq_add(&acc[D], sw); vq_add(&acc[D + 1], st); }
}

}  // namespace crn
