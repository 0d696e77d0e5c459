// vq_fast.cuh -- crnlib::clusterizer<V>::split_node as ONE launch per frontier (tolerance-class form).
//
// Replaces clusterizer<V>::split_node + compute_split_pca (+ compute_split_estimate) (reference crnlib/crn_clusterizer.h:740-873, :478-616,
// :446-476) and threaded_clusterizer<V>::compute_split (crnlib/crn_threaded_clusterizer.h:222-366) for V = vec2F / vec6F / vec16F with
// byte components, as the clustered-DDS quantiser uses them (qdxt1 / qdxt5: endpoint trees in init, selector clustering in pack).
//
// vq_kernels.cuh reproduces the reference's member-order FLOAT accumulations bit for bit (so the cluster assignment is the reference's own),
// at the price of ~60 launches per round and serial accumulation chains.  The contract for the clustered path is a tolerance (BASELINE.json:
// PSNR within 0.05 dB, bitrate within 1 %), and the reference's own result already moves with its helper-thread count, so this file does the
// same algorithm with sums in a fixed PARALLEL order (per-thread partial sums in double, shuffle tree, warps in order, CTAs in rank order):
// deterministic, not bit-identical.  One group of threads owns one node for the whole of split_node, as in hc_tree_split_kernel:
// a warp below 1024 members, a 256-thread CTA below 8192, a thread-block cluster of 8 / 16 CTAs x 512 threads above (partials through DSMEM).
// Everything of one split -- covariance, power iteration (start vector lerp(.75, 1.25), early exit on |delta| < .0025), projection split,
// <= 8 Lloyd rounds (stop at relative gain < .00125), stable partition -- happens inside that one launch.
#pragma once
#include "hc_kernels.cuh"

namespace crn {

// The node table lives in HBM for the whole build: a split reads its node's range / centroid / weight there and writes the two children's
// records (the host hands out the child ids before the launch), so a round moves 8 bytes per node to the device and 16 bytes back.
struct VqFastNodes {
    uint32_t* begin; uint32_t* end;      // member range of the node in `perm`
    float* centroid;                     // D floats per node
    unsigned long long* weight;
};
struct VqFastResult { int state; uint32_t n_left; float lvar, rvar; };      // state: 1 split, 2 unsplittable

// A training vector as stored (D bytes), fetched first and unpacked to floats only when it is used -- so that a row can be kept in flight across an
// iteration in 4 registers instead of 16.
template <int D> struct VqfRaw { uint32_t w[D == 16 ? 4 : (D == 2 ? 1 : D / 2)]; };
template <int D> __device__ __forceinline__ VqfRaw<D> vqf_load_raw(const uint8_t* __restrict__ vecs, uint32_t id)
{
    VqfRaw<D> r;
    if (D == 16) {
        const uint4 q = *reinterpret_cast<const uint4*>(vecs + (size_t)id * 16);
        r.w[0] = q.x; r.w[1 % (D == 16 ? 4 : 1)] = q.y; r.w[2 % (D == 16 ? 4 : 1)] = q.z; r.w[3 % (D == 16 ? 4 : 1)] = q.w;
    } else if (D == 2) {
        r.w[0] = *reinterpret_cast<const uint16_t*>(vecs + (size_t)id * 2);
    } else {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(vecs + (size_t)id * D);      // D = 6: 2-byte aligned
#pragma unroll
        for (int k = 0; k < D / 2; k++) r.w[k % (D == 16 ? 4 : (D == 2 ? 1 : D / 2))] = p[k];
    }
    return r;
}
template <int D> __device__ __forceinline__ void vqf_unpack(const VqfRaw<D>& r, float (&v)[D])
{
    if (D == 16) {
#pragma unroll
        for (int k = 0; k < 16; k++) v[k % D] = (float)((r.w[(k >> 2) % (D == 16 ? 4 : 1)] >> (8 * (k & 3))) & 255u);
    } else if (D == 2) {
        v[0] = (float)(r.w[0] & 255u); v[1 % D] = (float)(r.w[0] >> 8);
    } else {
#pragma unroll
        for (int k = 0; k < D / 2; k++) { const uint32_t q = r.w[k % (D == 16 ? 4 : (D == 2 ? 1 : D / 2))]; v[2 * k] = (float)(q & 255u); v[(2 * k + 1) % D] = (float)(q >> 8); }
    }
}
template <int D> __device__ __forceinline__ void vqf_load(const uint8_t* __restrict__ vecs, uint32_t id, float (&v)[D])
{
    vqf_unpack<D>(vqf_load_raw<D>(vecs, id), v);
}
// Walks the members first, first + stride, ... below e, calling f(i, id, v, w) once per member in that order, with the next member's packed row and
// weight and the index after that already in flight (see hc_for_members in hc_kernels.cuh: the split kernels wait on the perm -> row gather chain).
template <int D, typename F>
__device__ __forceinline__ void vqf_for_members(const uint8_t* __restrict__ vecs, const uint32_t* __restrict__ wts, const uint32_t* __restrict__ perm,
                                                uint32_t first, uint32_t e, uint32_t stride, F&& f)
{
    if (first >= e) return;
    uint32_t i = first, id = perm[i];
    uint32_t id_next = (i + stride < e) ? perm[i + stride] : 0;
    VqfRaw<D> raw = vqf_load_raw<D>(vecs, id);
    uint32_t w = wts[id];
    for (;;) {
        const uint32_t in = i + stride;
        const bool more = in < e;
        VqfRaw<D> raw_next = raw; uint32_t wn = 0, id2 = 0;
        if (more) {
            raw_next = vqf_load_raw<D>(vecs, id_next); wn = wts[id_next];
            if (in + stride < e) id2 = perm[in + stride];
        }
        float v[D]; vqf_unpack<D>(raw, v);
        f(i, id, v, w);
        if (!more) break;
        i = in; id = id_next; id_next = id2; w = wn; raw = raw_next;
    }
}

// mode: 0 = clusterizer<V>::split_node; 1 = threaded_clusterizer<V>::compute_split (own centroid, PCA division only; the children come back
// with the root statistics generate_codebook would compute for them, crn_clusterizer.h:76-93)
// -DCRN_VQ_WARP_MINB=n: minimum CTAs per SM of the warp-per-node instantiation (T = 32), i.e. a register cap.  Measured at the end of round 2
// (tools/gpu_r2ai.sh, configs[1] step): 12 CTAs (168 registers, 40 B spilled) takes the <16, 32, 1> launches from 22.4 to 18.4 ms of kernel time
// per step, 16 CTAs (128 registers) to 20.0 ms; the step's wall time did not move beyond its noise, and the change came too late for a full
// validation run, so the default build leaves the bound off.
template <int D, int T, int G>
#ifdef CRN_VQ_WARP_MINB
__global__ void __launch_bounds__(T, T == 32 ? CRN_VQ_WARP_MINB : 1)
#else
__global__ void __launch_bounds__(T)
#endif
vq_fast_split_kernel(const uint8_t* __restrict__ vecs, const uint32_t* __restrict__ wts, uint32_t* __restrict__ perm, uint32_t* __restrict__ perm_tmp,
                     VqFastNodes N, const uint2* __restrict__ slots, const uint32_t* __restrict__ slot_list, uint32_t nslots, VqFastResult* __restrict__ results, int mode)
{
    constexpr int NC = D * (D + 1) / 2;
    constexpr int K = D == 16 ? 18 : NC + D + 2;
    constexpr int COVN = D == 16 ? T * 16 : D * D;
    __shared__ HcRed<T, K> red;
    __shared__ float s_cov[COVN];
    __shared__ double s_cpart[D == 16 ? 256 : 1];
    __shared__ float s_axis[D];
    __shared__ uint32_t s_warp_cnt[T / 32 + 1];
    const unsigned tid = threadIdx.x;
    const unsigned rank = hc_cta_rank<G>();
    unsigned parity = 0;
    for (uint32_t si = blockIdx.x / G; si < nslots; si += gridDim.x / G) {
        const uint32_t slot = slot_list[si], node = slots[slot].x, child = slots[slot].y;      // children: child, child + 1
        const uint32_t begin = N.begin[node], end = N.end[node];
        const uint32_t seg = (end - begin + G - 1) / G;
        const uint32_t sb = min(end, begin + rank * seg), se = min(end, sb + seg);          // this CTA's members
        // node totals: sum of weighted vectors, sum of weighted dot products, total weight (right-hand sums = total - left-hand sums)
        double tot[D + 2];
        {
#pragma unroll
            for (int d = 0; d < D + 2; d++) tot[d] = 0;
            vqf_for_members<D>(vecs, wts, perm, sb + tid, se, T, [&](uint32_t, uint32_t, const float (&v)[D], uint32_t wi) {
                const float w = (float)wi;
                float dot = v[0] * v[0];
#pragma unroll
                for (int d = 1; d < D; d++) dot += v[d] * v[d];
#pragma unroll
                for (int d = 0; d < D; d++) tot[d] += (double)(v[d] * w);
                tot[D] += (double)(dot * w); tot[D + 1] += (double)w;
            });
            hc_group_reduce<T, G, K, true>(red, parity, tot, D + 2, nullptr, 0, nullptr);
        }
        float centroid[D];
        unsigned long long total_weight;
        if (mode == 1) {                         // compute_pca's own centroid (crn_threaded_clusterizer.h:226-245)
            total_weight = (unsigned long long)tot[D + 1];
            const double inv = tot[D + 1] != 0.0 ? 1.0 / tot[D + 1] : 0.0;
#pragma unroll
            for (int d = 0; d < D; d++) centroid[d] = (float)((double)(float)tot[d] * inv);
        } else {
            total_weight = N.weight[node];
#pragma unroll
            for (int d = 0; d < D; d++) centroid[d] = N.centroid[(size_t)node * D + d];
        }
        float left[D], right[D];
        bool have_split = false;
        if (mode == 0 && end - begin == 2) {     // compute_split_pca :480-485
            vqf_load<D>(vecs, perm[begin], left); vqf_load<D>(vecs, perm[begin + 1], right);
            have_split = true;
        } else {
            // covariance of the centred members (:489-521)
            if (D == 16) {
                constexpr int LANES = T / 16;
                const int x = tid & 15, ml = tid >> 4;
                float acc[16];
#pragma unroll
                for (int y = 0; y < 16; y++) acc[y] = 0.0f;
                vqf_for_members<D>(vecs, wts, perm, sb + ml, se, LANES, [&](uint32_t, uint32_t, const float (&vr)[D], uint32_t wi) {
                    const float w = (float)wi;
                    float v[D], vx = 0;
#pragma unroll
                    for (int d = 0; d < D; d++) { v[d] = vr[d] - centroid[d]; if (d == x) vx = v[d]; }
#pragma unroll
                    for (int y = 0; y < 16; y++) acc[y % D] += vx * (v[y % D] * w);
                });
                __syncthreads();
#pragma unroll
                for (int y = 0; y < 16; y++) s_cov[(tid * 16 + y) % COVN] = acc[y];
                __syncthreads();
                for (int e = tid; e < 256; e += T) {
                    const int ex = e >> 4, ey = e & 15;
                    double s = 0;
                    for (int l = 0; l < LANES; l++) s += (double)s_cov[((l * 16 + ex) * 16 + ey) % COVN];
                    s_cpart[e % (D == 16 ? 256 : 1)] = s;
                }
                hc_group_sync<G>();
                for (int e = tid; e < 256; e += T) {
                    double s = s_cpart[e % (D == 16 ? 256 : 1)];
#ifdef __CUDACC__
                    if (G > 1) { s = 0; for (unsigned r = 0; r < (unsigned)G; r++) s += cg::this_cluster().map_shared_rank(&s_cpart[0], r)[e % (D == 16 ? 256 : 1)]; }
#endif
                    s_cov[e % COVN] = (float)s * (1.0f / (float)total_weight);
                }
                hc_group_sync<G>();
            } else {
                float acc[NC];
#pragma unroll
                for (int k = 0; k < NC; k++) acc[k] = 0.0f;
                vqf_for_members<D>(vecs, wts, perm, sb + tid, se, T, [&](uint32_t, uint32_t, const float (&vr)[D], uint32_t wi) {
                    const float w = (float)wi;
                    float v[D];
#pragma unroll
                    for (int d = 0; d < D; d++) v[d] = vr[d] - centroid[d];
                    int k = 0;
#pragma unroll
                    for (int x = 0; x < D; x++)
#pragma unroll
                        for (int y = x; y < D; y++) acc[k++] += v[x] * (v[y] * w);
                });
                double dacc[NC];
#pragma unroll
                for (int k = 0; k < NC; k++) dacc[k] = (double)acc[k];
                hc_group_reduce<T, G, K, true>(red, parity, dacc, NC, nullptr, 0, nullptr);
                if (tid == 0) {
                    int k = 0;
                    for (int x = 0; x < D; x++) for (int y = x; y < D; y++) { const float c = (float)dacc[k++] * (1.0f / (float)total_weight); s_cov[(x * D + y) % COVN] = c; s_cov[(y * D + x) % COVN] = c; }
                }
                __syncthreads();
            }
            // power iteration (:523-571): start lerp(.75, 1.25, i / (N - 1)), max-normalised, stops when |axis[k - 1] - axis[k + 1]| < .0025
            if (tid == 0) {
                float axis[D], prev[D];
#pragma unroll
                for (int d = 0; d < D; d++) { axis[d] = .75f + (1.25f - .75f) * ((float)d * (1.0f / (float)(D - 1))); prev[d] = axis[d]; }
                for (int iter = 0; iter < 10; iter++) {
                    float xv[D];
                    double max_sum = 0;
                    for (int i = 0; i < D; i++) {
                        double sum = 0;
                        for (int j = 0; j < D; j++) {
                            const float c = i <= j ? s_cov[(i * D + j) % COVN] : s_cov[(j * D + i) % COVN];
                            sum += (double)(axis[j] * c);
                        }
                        xv[i] = (float)sum;
                        const double a = sum < 0 ? -sum : sum;
                        max_sum = max_sum > a ? max_sum : a;
                    }
                    if (max_sum != 0.0) { const float sc = (float)(1.0f / max_sum); for (int i = 0; i < D; i++) xv[i] *= sc; }
                    float dn = 0;
                    for (int i = 0; i < D; i++) { const float dd = prev[i] - xv[i]; dn += dd * dd; prev[i] = axis[i]; axis[i] = xv[i]; }
                    if (sqrtf(dn) < .0025f) break;
                }
                double n = (double)(axis[0] * axis[0]);
                for (int i = 1; i < D; i++) n += (double)(axis[i] * axis[i]);
                if (n != 0) { const float sc = (float)(1.0f / sqrt(n)); for (int i = 0; i < D; i++) axis[i] *= sc; }
                for (int i = 0; i < D; i++) s_axis[i] = axis[i];
            }
            __syncthreads();
            float axis[D];
#pragma unroll
            for (int d = 0; d < D; d++) axis[d] = s_axis[d];
            if (mode == 1) {
                // compute_division (:335-358): members with a negative projection go left.  Then the root statistics of both sides.
#pragma unroll
                for (int d = 0; d < D; d++) { left[d] = axis[d]; right[d] = centroid[d]; }      // carried to the partition below as (axis, centroid)
            } else {
                // projection split (:573-606): means of the two sides, or compute_split_estimate when one side is empty
                double sl[D + 1];
#pragma unroll
                for (int d = 0; d <= D; d++) sl[d] = 0;
                vqf_for_members<D>(vecs, wts, perm, sb + tid, se, T, [&](uint32_t, uint32_t, const float (&v)[D], uint32_t wi) {
                    const float w = (float)wi;
                    float t = (v[0] - centroid[0]) * axis[0];
#pragma unroll
                    for (int d = 1; d < D; d++) t += (v[d] - centroid[d]) * axis[d];
                    if (t < 0.0f) {
#pragma unroll
                        for (int d = 0; d < D; d++) sl[d] += (double)(v[d] * w);
                        sl[D] += (double)w;
                    }
                });
                hc_group_reduce<T, G, K, true>(red, parity, sl, D + 1, nullptr, 0, nullptr);
                const double lwt = sl[D], rwt = tot[D + 1] - sl[D];
                if (lwt > 0.0 && rwt > 0.0) {
                    const float fl = (float)(1.0f / lwt), fr = (float)(1.0f / rwt);
#pragma unroll
                    for (int d = 0; d < D; d++) { left[d] = (float)sl[d] * fl; right[d] = (float)(tot[d] - sl[d]) * fr; }
                    have_split = true;
                }
            }
        }
        if (mode == 0 && !have_split) {
            // compute_split_estimate (:446-476): furthest from the centroid, then furthest from that one (first maximum in member order)
            float seed[2][D];
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                unsigned long long key = 0;
                vqf_for_members<D>(vecs, wts, perm, sb + tid, se, T, [&](uint32_t i, uint32_t, const float (&v)[D], uint32_t) {
                    const float d2 = pass ? hc_sqdist<D>(v, seed[0]) : hc_sqdist<D>(v, centroid);
                    const unsigned long long k = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)(~i);
                    key = k > key ? k : key;
                });
                hc_group_reduce<T, G, K, false>(red, parity, nullptr, 0, &key, 1, nullptr);
                const uint32_t pos = ~(unsigned)key;
                vqf_load<D>(vecs, perm[pos], seed[pass]);
            }
#pragma unroll
            for (int d = 0; d < D; d++) { left[d] = (seed[0][d] + centroid[d]) * .5f; right[d] = (seed[1][d] + centroid[d]) * .5f; }
        }
        // Lloyd rounds (:761-836), at most 8; a presplit runs the loop body once with the projection as the assignment
        float prev_total_variance = 1e+10f, lvar = 0, rvar = 0;
        unsigned long long lw = 0, rw = 0;
        uint32_t n_left = 0, left_before = 0;
        bool unsplittable = false;
        float used_left[D], used_right[D];
        const unsigned max_loops = mode == 1 ? 1u : 8u;
#pragma unroll 1
        for (unsigned loops = 0; loops < max_loops; loops++) {
            double sl[D + 1]; unsigned long long uu[2] = { 0, 0 }, upre[2];
#pragma unroll
            for (int d = 0; d <= D; d++) sl[d] = 0;
#pragma unroll
            for (int d = 0; d < D; d++) { used_left[d] = left[d]; used_right[d] = right[d]; }
            vqf_for_members<D>(vecs, wts, perm, sb + tid, se, T, [&](uint32_t, uint32_t, const float (&v)[D], uint32_t wi) {
                bool is_left;
                if (mode == 1) {
                    float t = (v[0] - right[0]) * left[0];
#pragma unroll
                    for (int d = 1; d < D; d++) t += (v[d] - right[d]) * left[d];
                    is_left = t < 0.0f;
                } else is_left = hc_sqdist<D>(left, v) < hc_sqdist<D>(right, v);
                if (is_left) {
                    const float w = (float)wi;
                    float dot = v[0] * v[0];
#pragma unroll
                    for (int d = 1; d < D; d++) dot += v[d] * v[d];
#pragma unroll
                    for (int d = 0; d < D; d++) sl[d] += (double)(v[d] * w);
                    sl[D] += (double)(dot * w); uu[0] += wi; uu[1]++;
                }
            });
            hc_group_reduce<T, G, K, true>(red, parity, sl, D + 1, uu, 2, upre);
            lw = uu[0]; n_left = (uint32_t)uu[1]; left_before = (uint32_t)upre[1];
            rw = (unsigned long long)tot[D + 1] - lw;
            if (mode == 1) {
                // the children of a division may be empty (their partition is then skipped, crn_threaded_clusterizer.h:101-118)
                float nl[D], nr[D];
#pragma unroll
                for (int d = 0; d < D; d++) { nl[d] = (float)sl[d]; nr[d] = (float)(tot[d] - sl[d]); }
                float ldot = nl[0] * nl[0], rdot = nr[0] * nr[0];
#pragma unroll
                for (int d = 1; d < D; d++) { ldot += nl[d] * nl[d]; rdot += nr[d] * nr[d]; }
                lvar = lw ? (float)(sl[D] - (double)(ldot / (float)lw)) : 0.0f;
                rvar = rw ? (float)((tot[D] - sl[D]) - (double)(rdot / (float)rw)) : 0.0f;
                const float fl = lw ? 1.0f / (float)lw : 0.0f, fr = rw ? 1.0f / (float)rw : 0.0f;
#pragma unroll
                for (int d = 0; d < D; d++) { left[d] = nl[d] * fl; right[d] = nr[d] * fr; }
                break;
            }
            if (!lw || !rw) { unsplittable = true; break; }
            float nl[D], nr[D];
#pragma unroll
            for (int d = 0; d < D; d++) { nl[d] = (float)sl[d]; nr[d] = (float)(tot[d] - sl[d]); }
            float ldot = nl[0] * nl[0], rdot = nr[0] * nr[0];
#pragma unroll
            for (int d = 1; d < D; d++) { ldot += nl[d] * nl[d]; rdot += nr[d] * nr[d]; }
            lvar = (float)(sl[D] - (double)(ldot / (float)lw)); rvar = (float)((tot[D] - sl[D]) - (double)(rdot / (float)rw));
            const float fl = 1.0f / (float)lw, fr = 1.0f / (float)rw;
#pragma unroll
            for (int d = 0; d < D; d++) { left[d] = nl[d] * fl; right[d] = nr[d] * fr; }
            const float total_variance = lvar + rvar;
            if (total_variance < .00001f) break;
            if (((prev_total_variance - total_variance) / total_variance) < .00125f) break;
            prev_total_variance = total_variance;
        }
        if (!unsplittable) {
            // stable partition by the last assignment (:838-871): lefts of lower-ranked CTAs come first
            uint32_t base_l = begin + left_before, base_r = begin + n_left + ((sb - begin) - left_before);
            for (uint32_t i0 = sb; i0 < se; i0 += T) {
                const uint32_t i = i0 + tid;
                uint32_t id = 0; bool valid = i < se, is_left = false;
                if (valid) {
                    id = perm[i];
                    float v[D]; vqf_load<D>(vecs, id, v);
                    if (mode == 1) {
                        float t = (v[0] - used_right[0]) * used_left[0];
#pragma unroll
                        for (int d = 1; d < D; d++) t += (v[d] - used_right[d]) * used_left[d];
                        is_left = t < 0.0f;
                    } else is_left = hc_sqdist<D>(used_left, v) < hc_sqdist<D>(used_right, v);
                }
                const unsigned bl = __ballot_sync(CRN_FULL_MASK, valid && is_left), br = __ballot_sync(CRN_FULL_MASK, valid && !is_left);
                uint32_t pre_l = 0, pre_r = 0, tot_l = __popc(bl), tot_r = __popc(br);
                if (T > 32) {
                    __syncthreads();
                    if ((tid & 31) == 0) { s_warp_cnt[tid >> 5] = (uint32_t)__popc(bl) | ((uint32_t)__popc(br) << 16); }
                    __syncthreads();
                    tot_l = tot_r = 0;
                    for (unsigned w = 0; w < T / 32; w++) {
                        const uint32_t c = s_warp_cnt[w];
                        if (w < (tid >> 5)) { pre_l += c & 0xffff; pre_r += c >> 16; }
                        tot_l += c & 0xffff; tot_r += c >> 16;
                    }
                }
                if (valid) {
                    const unsigned lt_mask = (1u << (tid & 31)) - 1u;
                    if (is_left) perm_tmp[base_l + pre_l + __popc(bl & lt_mask)] = id;
                    else perm_tmp[base_r + pre_r + __popc(br & lt_mask)] = id;
                }
                base_l += tot_l; base_r += tot_r;
            }
            hc_group_sync<G>();
            for (uint32_t i = sb + tid; i < se; i += T) perm[i] = perm_tmp[i];
        }
        if (tid == 0 && rank == 0) {
            VqFastResult r; r.state = unsplittable ? 2 : 1; r.n_left = n_left; r.lvar = lvar; r.rvar = rvar;
            results[slot] = r;
            if (!unsplittable) {
                N.begin[child] = begin; N.end[child] = begin + n_left; N.begin[child + 1] = begin + n_left; N.end[child + 1] = end;
                N.weight[child] = lw; N.weight[child + 1] = rw;
#pragma unroll
                for (int d = 0; d < D; d++) { N.centroid[(size_t)child * D + d] = left[d]; N.centroid[(size_t)(child + 1) * D + d] = right[d]; }
            }
        }
        hc_group_sync<G>();
    }
}

// Tiny nodes, one THREAD per node.  A 600 K-leaf selector tree ends in hundreds of thousands of nodes with two or three members; a warp (let
// alone its shuffle reductions and barriers) per such node costs ~100 us, a thread walking the members serially a few microseconds.  The
// thread follows split_node literally (member-order sums, as the reference); only the power iteration is applied matrix-free
// (covar * axis = sum_m w_m ((v_m - c) . axis)(v_m - c) / W) instead of through the N x N matrix, to stay in registers.
constexpr uint32_t kVqTinyNode = 12;
template <int D>
__global__ void __launch_bounds__(128)
vq_fast_tiny_kernel(const uint8_t* __restrict__ vecs, const uint32_t* __restrict__ wts, uint32_t* __restrict__ perm, VqFastNodes N, const uint2* __restrict__ slots,
                    const uint32_t* __restrict__ slot_list, uint32_t nslots, VqFastResult* __restrict__ results)
{
    const uint32_t si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= nslots) return;
    const uint32_t slot = slot_list[si], node = slots[slot].x, child = slots[slot].y;
    const uint32_t begin = N.begin[node], end = N.end[node], cnt = end - begin;
    uint32_t ids[kVqTinyNode];
#pragma unroll
    for (uint32_t i = 0; i < kVqTinyNode; i++) ids[i] = i < cnt ? perm[begin + i] : 0u;
    float centroid[D];
#pragma unroll
    for (int d = 0; d < D; d++) centroid[d] = N.centroid[(size_t)node * D + d];
    const unsigned long long total_weight = N.weight[node];
    float left[D], right[D];
    bool have = false;
    if (cnt == 2) { vqf_load<D>(vecs, ids[0], left); vqf_load<D>(vecs, ids[1], right); have = true; }
    else {
        float axis[D], prev[D];
#pragma unroll
        for (int d = 0; d < D; d++) { axis[d] = .75f + (1.25f - .75f) * ((float)d * (1.0f / (float)(D - 1))); prev[d] = axis[d]; }
        const float inv_w = 1.0f / (float)total_weight;
        for (int iter = 0; iter < 10; iter++) {
            float xv[D];
#pragma unroll
            for (int d = 0; d < D; d++) xv[d] = 0.0f;
            for (uint32_t i = 0; i < cnt; i++) {
                float v[D]; vqf_load<D>(vecs, ids[i], v);
                float t = 0.0f;
#pragma unroll
                for (int d = 0; d < D; d++) { v[d] -= centroid[d]; t += v[d] * axis[d]; }
                t *= (float)wts[ids[i]] * inv_w;
#pragma unroll
                for (int d = 0; d < D; d++) xv[d] += t * v[d];
            }
            float max_sum = 0.0f;
#pragma unroll
            for (int d = 0; d < D; d++) max_sum = fmaxf(max_sum, fabsf(xv[d]));
            if (max_sum != 0.0f) { const float sc = 1.0f / max_sum;
#pragma unroll
                for (int d = 0; d < D; d++) xv[d] *= sc; }
            float dn = 0.0f;
#pragma unroll
            for (int d = 0; d < D; d++) { const float dd = prev[d] - xv[d]; dn += dd * dd; prev[d] = axis[d]; axis[d] = xv[d]; }
            if (sqrtf(dn) < .0025f) break;
        }
        {
            double n2 = 0.0;
#pragma unroll
            for (int d = 0; d < D; d++) n2 += (double)(axis[d] * axis[d]);
            if (n2 != 0.0) { const float sc = (float)(1.0f / sqrt(n2));
#pragma unroll
                for (int d = 0; d < D; d++) axis[d] *= sc; }
        }
        float ls[D], rs[D]; double lw = 0.0, rw = 0.0;
#pragma unroll
        for (int d = 0; d < D; d++) { ls[d] = 0.0f; rs[d] = 0.0f; }
        for (uint32_t i = 0; i < cnt; i++) {
            float v[D]; vqf_load<D>(vecs, ids[i], v);
            const float w = (float)wts[ids[i]];
            float t = 0.0f;
#pragma unroll
            for (int d = 0; d < D; d++) t += (v[d] - centroid[d]) * axis[d];
            if (t < 0.0f) { lw += (double)w;
#pragma unroll
                for (int d = 0; d < D; d++) ls[d] += v[d] * w; }
            else { rw += (double)w;
#pragma unroll
                for (int d = 0; d < D; d++) rs[d] += v[d] * w; }
        }
        if (lw > 0.0 && rw > 0.0) {
            const float fl = (float)(1.0f / lw), fr = (float)(1.0f / rw);
#pragma unroll
            for (int d = 0; d < D; d++) { left[d] = ls[d] * fl; right[d] = rs[d] * fr; }
            have = true;
        }
    }
    if (!have) {                                 // compute_split_estimate (:446-476)
        float far_[D], opp[D]; float best = -1.0f;
        for (uint32_t i = 0; i < cnt; i++) { float v[D]; vqf_load<D>(vecs, ids[i], v); const float d2 = hc_sqdist<D>(v, centroid); if (d2 > best) { best = d2;
#pragma unroll
            for (int d = 0; d < D; d++) far_[d] = v[d]; } }
        best = -1.0f;
        for (uint32_t i = 0; i < cnt; i++) { float v[D]; vqf_load<D>(vecs, ids[i], v); const float d2 = hc_sqdist<D>(v, far_); if (d2 > best) { best = d2;
#pragma unroll
            for (int d = 0; d < D; d++) opp[d] = v[d]; } }
#pragma unroll
        for (int d = 0; d < D; d++) { left[d] = (far_[d] + centroid[d]) * .5f; right[d] = (opp[d] + centroid[d]) * .5f; }
    }
    // Lloyd rounds (:761-836)
    float prev_total_variance = 1e+10f, lvar = 0.0f, rvar = 0.0f;
    unsigned long long lwt = 0, rwt = 0;
    uint32_t side = 0, n_left = 0;
    bool unsplittable = false;
    for (unsigned loops = 0; loops < 8; loops++) {
        float nl[D], nr[D]; double ltt = 0.0, rtt = 0.0;
#pragma unroll
        for (int d = 0; d < D; d++) { nl[d] = 0.0f; nr[d] = 0.0f; }
        lwt = 0; rwt = 0; side = 0; n_left = 0;
        for (uint32_t i = 0; i < cnt; i++) {
            float v[D]; vqf_load<D>(vecs, ids[i], v);
            const unsigned wi = wts[ids[i]]; const float w = (float)wi;
            float dot = 0.0f;
#pragma unroll
            for (int d = 0; d < D; d++) dot += v[d] * v[d];
            if (hc_sqdist<D>(left, v) < hc_sqdist<D>(right, v)) { side |= 1u << i; n_left++; lwt += wi; ltt += (double)(dot * w);
#pragma unroll
                for (int d = 0; d < D; d++) nl[d] += v[d] * w; }
            else { rwt += wi; rtt += (double)(dot * w);
#pragma unroll
                for (int d = 0; d < D; d++) nr[d] += v[d] * w; }
        }
        if (!lwt || !rwt) { unsplittable = true; break; }
        float ldot = 0.0f, rdot = 0.0f;
#pragma unroll
        for (int d = 0; d < D; d++) { ldot += nl[d] * nl[d]; rdot += nr[d] * nr[d]; }
        lvar = (float)(ltt - (double)(ldot / (float)lwt)); rvar = (float)(rtt - (double)(rdot / (float)rwt));
        const float fl = 1.0f / (float)lwt, fr = 1.0f / (float)rwt;
#pragma unroll
        for (int d = 0; d < D; d++) { left[d] = nl[d] * fl; right[d] = nr[d] * fr; }
        const float total_variance = lvar + rvar;
        if (total_variance < .00001f) break;
        if (((prev_total_variance - total_variance) / total_variance) < .00125f) break;
        prev_total_variance = total_variance;
    }
    VqFastResult r; r.state = unsplittable ? 2 : 1; r.n_left = n_left; r.lvar = lvar; r.rvar = rvar;
    results[slot] = r;
    if (unsplittable) return;
    uint32_t pl = begin, pr = begin + n_left;                                 // stable partition
#pragma unroll
    for (uint32_t i = 0; i < kVqTinyNode; i++) if (i < cnt) { if ((side >> i) & 1u) perm[pl++] = ids[i]; else perm[pr++] = ids[i]; }
    N.begin[child] = begin; N.end[child] = begin + n_left; N.begin[child + 1] = begin + n_left; N.end[child + 1] = end;
    N.weight[child] = lwt; N.weight[child + 1] = rwt;
#pragma unroll
    for (int d = 0; d < D; d++) { N.centroid[(size_t)child * D + d] = left[d]; N.centroid[(size_t)(child + 1) * D + d] = right[d]; }
}

// root of a clusterizer (generate_codebook :76-93): statistics of the whole training set.  out: D weighted sums, the weighted dot-product sum, the weight
template <int D>
__global__ void __launch_bounds__(512)
vq_fast_root_kernel(const uint8_t* __restrict__ vecs, const uint32_t* __restrict__ wts, const uint32_t* __restrict__ ids, uint32_t n, uint32_t* __restrict__ perm, double* __restrict__ out)
{
    constexpr int T = 512;
    __shared__ HcRed<T, D + 2> red;
    unsigned parity = 0;
    double s[D + 2];
#pragma unroll
    for (int d = 0; d < D + 2; d++) s[d] = 0;
    for (uint32_t i = blockIdx.x * T + threadIdx.x; i < n; i += gridDim.x * T) {
        const uint32_t id = ids ? ids[i] : i;
        float v[D]; vqf_load<D>(vecs, id, v);
        const float w = (float)wts[id];
        float dot = v[0] * v[0];
#pragma unroll
        for (int d = 1; d < D; d++) dot += v[d] * v[d];
#pragma unroll
        for (int d = 0; d < D; d++) s[d] += (double)(v[d] * w);
        s[D] += (double)(dot * w); s[D + 1] += (double)w;
        perm[i] = id;
    }
    hc_group_reduce<T, 1, D + 2, true>(red, parity, s, D + 2, nullptr, 0, nullptr);
    if (threadIdx.x == 0) for (int d = 0; d < D + 2; d++) out[(size_t)blockIdx.x * (D + 2) + d] = s[d];
}

}  // namespace crn
