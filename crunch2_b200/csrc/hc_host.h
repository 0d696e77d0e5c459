// hc_host.h -- host driver of the dxt_hc pipeline (included by crn_b200.cu after its helpers).
//
// Replaces crnlib::dxt_hc::compress (reference crnlib/crn_dxt_hc.cpp:98-312) for the DXT formats: tile determination
// (a12) -> endpoint tree quantisers (a13) -> nearest-codebook assignment (a14) -> per-cluster endpoint codebooks (a15:
// cluster optimiser + per-block selectors + refiner) -> selector codebooks (a16: tree quantiser + exhaustive search +
// re-vote) -> palette dedup / index remap / reference flags (a17).  Everything per-pixel, per-block, per-cluster or
// per-tree-node runs on the device; the host keeps what is inherently a small serial structure: the tree quantiser's
// priority queue (HcTreeVq, replaying crn_tree_clusterizer.h:89-171 on split results the device produced a frontier
// at a time), the sort + dedup of training vectors, the CSR build, and the final remap.
#pragma once
#include <queue>
#include <thread>

namespace {

struct HcBuf {                                   // pooled device buffer
    crn_gpu_ctx* ctx = nullptr; void* p = nullptr; size_t cap = 0;
    ~HcBuf() { if (p) pool_free(ctx, p, cap); }
    cudaError_t alloc(crn_gpu_ctx* c, size_t bytes) { ctx = c; return pool_alloc(c, &p, bytes, &cap); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

template <typename T>
struct HcHost {                                  // host array for a transfer: pinned when possible, plain heap otherwise
    crn_gpu_ctx* ctx = nullptr; T* p = nullptr; size_t cap = 0, count = 0; bool pinned = false;
    HcHost(crn_gpu_ctx* c, size_t n) : ctx(c), count(n)
    {
        p = static_cast<T*>(pin_alloc(c, n * sizeof(T), &cap));
        pinned = p != nullptr;
        if (!p) p = static_cast<T*>(malloc(n ? n * sizeof(T) : 1));
        if (!p) throw std::bad_alloc();
    }
    ~HcHost() { if (pinned) pin_free(ctx, p, cap); else free(p); }
    HcHost(const HcHost&) = delete;
    HcHost& operator=(const HcHost&) = delete;
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
    T* data() { return p; }
    const T* data() const { return p; }
    size_t size() const { return count; }
};

#define HC_ALLOC(buf, bytes)                                                                         \
    do {                                                                                             \
        cudaError_t ce_ = (buf).alloc(ctx, (bytes));                                                 \
        if (ce_ != cudaSuccess) return set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "dxt_hc: cudaMalloc", ce_); \
    } while (0)
#define HC_RC(call) do { int rc_ = (call); if (rc_) return rc_; } while (0)

// ---- tree_clusterizer<V>::generate_codebook (crn_tree_clusterizer.h:89-171), single-task semantics ------------------
template <int D>
struct HcTreeVq {
    struct Node {
        float centroid[D]; unsigned long long total_weight; float variance; uint32_t begin, end; int left, right;
        bool have_result; crn::HcTreeSlot<D> res;
    };
    struct HeapEntry {                            // NodeInfo (:46-59): larger variance first, then lower index
        uint32_t index; float variance;
        bool operator<(const HeapEntry& o) const { return index < o.index ? variance < o.variance : o.variance >= variance; }
    };
    std::vector<float> codebook;                  // K x D
    static constexpr int kClusterCtas = 8, kWideClusterCtas = 16, kClusterThreads = 512;
    bool wide_ok = true;
#ifdef __CUDACC__
    template <int G>
    static bool launch_cluster(crn_gpu_ctx* ctx, const HcBuf& d_vecs, const HcBuf& d_wts, const HcBuf& d_perm, const HcBuf& d_tmp, const HcBuf& d_slots, const uint32_t* dl, uint32_t count)
    {
        auto kernel = crn::hc_tree_split_kernel<D, kClusterThreads, G>;
        if (G > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return false;
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        const unsigned nclusters = (unsigned)std::min<size_t>(count, (size_t)std::max(1, ctx->sm_count / G) * 2);
        cfg.gridDim = dim3(nclusters * G); cfg.blockDim = dim3(kClusterThreads); cfg.stream = ctx->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kernel, (const float*)d_vecs.as<float>(), (const uint32_t*)d_wts.as<uint32_t>(), d_perm.as<uint32_t>(), d_tmp.as<uint32_t>(),
                                  d_slots.as<crn::HcTreeSlot<D>>(), dl, count) == cudaSuccess;
    }
#endif
    static constexpr uint32_t kHugeNode = 8192;
    uint32_t rounds = 0, device_splits = 0;

    // vecs / wts: host arrays (on_device == false, uploaded here) or device arrays that stay valid during the call
    int build(crn_gpu_ctx* ctx, const float* vecs, const uint32_t* wts_in, uint32_t n, uint32_t max_splits, bool on_device = false)
    {
        codebook.clear();
        if (!n || !max_splits) return CRN_GPU_OK;
        HcBuf d_vecs, d_wts, d_perm, d_tmp, d_slots, d_list, d_root;
        HC_ALLOC(d_perm, (size_t)n * 4); HC_ALLOC(d_tmp, (size_t)n * 4);
        HC_ALLOC(d_root, (D + 2) * 8);
        if (on_device) {                           // borrowed: never returned to the pool by this object
            d_vecs.p = const_cast<float*>(vecs); d_wts.p = const_cast<uint32_t*>(wts_in);
        } else {
            HC_ALLOC(d_vecs, (size_t)n * D * 4); HC_ALLOC(d_wts, (size_t)n * 4);
            CRN_CUDA(ctx, cudaMemcpyAsync(d_vecs.p, vecs, (size_t)n * D * 4, cudaMemcpyHostToDevice, ctx->stream));
            CRN_CUDA(ctx, cudaMemcpyAsync(d_wts.p, wts_in, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
        }
        struct Borrow { HcBuf& a; HcBuf& b; bool on; ~Borrow() { if (on) { a.p = nullptr; b.p = nullptr; } } } borrow{ d_vecs, d_wts, on_device };
        CRN_LAUNCH(crn::hc_tree_root_kernel<D>, 1, 512, 0, ctx->stream, d_vecs.as<float>(), d_wts.as<uint32_t>(), n, d_perm.as<uint32_t>(), d_root.as<double>());
        ctx->launches++;
        double h_root[D + 2];
        CRN_CUDA(ctx, cudaMemcpyAsync(h_root, d_root.p, sizeof(h_root), cudaMemcpyDeviceToHost, ctx->stream));
        CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        std::vector<Node> nodes;
        nodes.reserve((size_t)max_splits * 2 + 2);
        {
            Node r; memset(&r, 0, sizeof(r));
            r.begin = 0; r.end = n; r.left = r.right = -1; r.have_result = false;
            r.total_weight = (unsigned long long)h_root[D + 1];
            float dot = 0;
            for (int d = 0; d < D; d++) { r.centroid[d] = (float)h_root[d]; dot = d ? dot + r.centroid[d] * r.centroid[d] : r.centroid[d] * r.centroid[d]; }
            r.variance = (float)(h_root[D] - (double)(dot / (float)r.total_weight));
            const float inv = 1.0f / (float)r.total_weight;
            for (int d = 0; d < D; d++) r.centroid[d] *= inv;
            nodes.push_back(r);
        }
        std::vector<HeapEntry> heap;               // std::priority_queue semantics via push_heap / pop_heap, iterable
        heap.push_back({ 0u, nodes[0].variance });
        uint32_t splits = 1;
        std::vector<crn::HcTreeSlot<D>> h_slots;
        std::vector<uint32_t> round_nodes, small_list, large_list, huge_list;
        size_t slot_cap = 0, list_cap = 0;
        bool done = false;
        while (!done) {
            // ---- replay the reference's loop as far as the device results reach
            bool need_round = false;
            while (splits < max_splits) {
                if (heap.empty()) { done = true; break; }
                const uint32_t ni = heap.front().index;
                if (nodes[ni].variance <= 0.0f || nodes[ni].begin + 1 == nodes[ni].end) { done = true; break; }
                if (!nodes[ni].have_result) { need_round = true; break; }
                std::pop_heap(heap.begin(), heap.end()); heap.pop_back();
                const crn::HcTreeSlot<D> r = nodes[ni].res;
                if (r.state == 1) {
                    Node l, rt; memset(&l, 0, sizeof(l)); memset(&rt, 0, sizeof(rt));
                    l.left = l.right = rt.left = rt.right = -1;
                    l.begin = nodes[ni].begin; l.end = rt.begin = nodes[ni].begin + r.n_left; rt.end = nodes[ni].end;
                    for (int d = 0; d < D; d++) { l.centroid[d] = r.lc[d]; rt.centroid[d] = r.rc[d]; }
                    l.total_weight = r.lw; rt.total_weight = r.rw; l.variance = r.lvar; rt.variance = r.rvar;
                    nodes[ni].left = (int)nodes.size(); nodes.push_back(l);
                    nodes[ni].right = (int)nodes.size(); nodes.push_back(rt);
                    heap.push_back({ (uint32_t)nodes[ni].left, l.variance }); std::push_heap(heap.begin(), heap.end());
                    heap.push_back({ (uint32_t)nodes[ni].right, rt.variance }); std::push_heap(heap.begin(), heap.end());
                }
                splits++;
            }
            if (splits >= max_splits) done = true;
            if (done || !need_round) break;
            // ---- one device round: every leaf the queue could still reach
            round_nodes.clear();
            for (const HeapEntry& e : heap) {
                const Node& nd = nodes[e.index];
                if (!nd.have_result && nd.variance > 0.0f && nd.begin + 1 != nd.end) round_nodes.push_back(e.index);
            }
            const uint32_t budget = max_splits - splits;
            if (round_nodes.size() > budget) {
                std::partial_sort(round_nodes.begin(), round_nodes.begin() + budget, round_nodes.end(), [&](uint32_t a, uint32_t b) {
                    return HeapEntry{ b, nodes[b].variance } < HeapEntry{ a, nodes[a].variance };
                });
                round_nodes.resize(budget);
            }
            const uint32_t nr = (uint32_t)round_nodes.size();
            h_slots.resize(nr);
            small_list.clear(); large_list.clear(); huge_list.clear();
            for (uint32_t i = 0; i < nr; i++) {
                const Node& nd = nodes[round_nodes[i]];
                crn::HcTreeSlot<D>& s = h_slots[i];
                memset(&s, 0, sizeof(s));
                s.begin = nd.begin; s.end = nd.end; s.total_weight = nd.total_weight;
                for (int d = 0; d < D; d++) s.centroid[d] = nd.centroid[d];
                const uint32_t sz = nd.end - nd.begin;
#ifdef __CUDACC__
                (sz >= kHugeNode ? huge_list : (sz >= 1024 ? large_list : small_list)).push_back(i);
#else
                (sz >= 1024 ? large_list : small_list).push_back(i);      // the emulator has no thread-block clusters
#endif
            }
            if (nr > slot_cap) {
                if (d_slots.p) { pool_free(ctx, d_slots.p, d_slots.cap); d_slots.p = nullptr; }
                if (d_list.p) { pool_free(ctx, d_list.p, d_list.cap); d_list.p = nullptr; }
                slot_cap = (size_t)nr * 2 + 64; list_cap = slot_cap;
                HC_ALLOC(d_slots, slot_cap * sizeof(crn::HcTreeSlot<D>)); HC_ALLOC(d_list, list_cap * 4);
            }
            CRN_CUDA(ctx, cudaMemcpyAsync(d_slots.p, h_slots.data(), (size_t)nr * sizeof(crn::HcTreeSlot<D>), cudaMemcpyHostToDevice, ctx->stream));
            uint32_t* dl = d_list.as<uint32_t>();
#ifdef __CUDACC__
            if (!huge_list.empty()) {          // a thread-block cluster per node, partial sums exchanged through distributed shared memory
                CRN_CUDA(ctx, cudaMemcpyAsync(dl, huge_list.data(), huge_list.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
                // the first rounds have a handful of very large nodes: 16-CTA clusters (non-portable size, one per GPC) put twice
                // the SMs on each of them; from 9 nodes on, 8-CTA clusters fill the machine anyway
                bool launched = false;
                if (huge_list.size() <= 8 && wide_ok) {
                    launched = launch_cluster<kWideClusterCtas>(ctx, d_vecs, d_wts, d_perm, d_tmp, d_slots, dl, (uint32_t)huge_list.size());
                    if (!launched) { wide_ok = false; (void)cudaGetLastError(); }
                }
                if (!launched && !launch_cluster<kClusterCtas>(ctx, d_vecs, d_wts, d_perm, d_tmp, d_slots, dl, (uint32_t)huge_list.size()))
                    return set_err(ctx, CRN_GPU_ERR_CUDA, "dxt_hc: cluster launch of hc_tree_split_kernel failed", cudaGetLastError());
                ctx->launches++;
                dl += huge_list.size();
            }
#endif
            if (!large_list.empty()) {
                CRN_CUDA(ctx, cudaMemcpyAsync(dl, large_list.data(), large_list.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
                CRN_LAUNCH((crn::hc_tree_split_kernel<D, 256, 1>), (unsigned)large_list.size(), 256, 0, ctx->stream, d_vecs.as<float>(), d_wts.as<uint32_t>(), d_perm.as<uint32_t>(),
                           d_tmp.as<uint32_t>(), d_slots.as<crn::HcTreeSlot<D>>(), dl, (uint32_t)large_list.size());
                ctx->launches++;
                dl += large_list.size();
            }
            if (!small_list.empty()) {
                CRN_CUDA(ctx, cudaMemcpyAsync(dl, small_list.data(), small_list.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
                const unsigned grid = (unsigned)std::min<size_t>(small_list.size(), (size_t)ctx->sm_count * 32);
                CRN_LAUNCH((crn::hc_tree_split_kernel<D, 32, 1>), grid, 32, 0, ctx->stream, d_vecs.as<float>(), d_wts.as<uint32_t>(), d_perm.as<uint32_t>(),
                           d_tmp.as<uint32_t>(), d_slots.as<crn::HcTreeSlot<D>>(), dl, (uint32_t)small_list.size());
                ctx->launches++;
            }
            CRN_CUDA(ctx, cudaGetLastError());
            CRN_CUDA(ctx, cudaMemcpyAsync(h_slots.data(), d_slots.p, (size_t)nr * sizeof(crn::HcTreeSlot<D>), cudaMemcpyDeviceToHost, ctx->stream));
            CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            for (uint32_t i = 0; i < nr; i++) { nodes[round_nodes[i]].res = h_slots[i]; nodes[round_nodes[i]].have_result = true; }
            rounds++; device_splits += nr;
        }
        for (size_t i = 0; i < nodes.size(); i++)
            if (nodes[i].left == -1) codebook.insert(codebook.end(), nodes[i].centroid, nodes[i].centroid + D);
        return CRN_GPU_OK;
    }
    uint32_t size() const { return (uint32_t)(codebook.size() / D); }
};

template <int D> struct HcVecKey {
    float v[D]; uint32_t w;
    bool operator<(const HcVecKey& o) const {      // std::pair<vec, uint> ordering (crn_vec.h:279-294, then the weight)
        for (int d = 0; d < D; d++) { if (v[d] < o.v[d]) return true; if (v[d] != o.v[d]) return false; }
        return w < o.w;
    }
    bool same_vec(const HcVecKey& o) const { for (int d = 0; d < D; d++) if (v[d] != o.v[d]) return false; return true; }
};

// sort + merge of equal vectors with saturating weights (determine_color_endpoints, crn_dxt_hc.cpp:888-968)
template <int D>
void hc_sort_dedup(std::vector<HcVecKey<D>>& keys, std::vector<float>& vecs, std::vector<uint32_t>& wts)
{
    std::sort(keys.begin(), keys.end());
    vecs.clear(); wts.clear();
    for (size_t i = 0; i < keys.size(); i++) {
        if (wts.empty() || !keys[i].same_vec(keys[i - 1])) { vecs.insert(vecs.end(), keys[i].v, keys[i].v + D); wts.push_back(keys[i].w); }
        else if (wts.back() > 0xffffffffu - keys[i].w) wts.back() = 0xffffffffu;
        else wts.back() += keys[i].w;
    }
}

}  // namespace

struct crn_gpu_hc {
    crn_gpu_hc_info info;
    std::vector<uint16_t> endpoint_indices, selector_indices;     // n x 4
    std::vector<uint32_t> color_endpoints, alpha_endpoints, color_selectors;
    std::vector<uint64_t> alpha_selectors;
    std::vector<uint8_t> block_encodings;
    std::vector<uint32_t> tile_indices;
};

namespace {

// Quality-independent products of one dxt_hc call: the tile pass (a12) and, per endpoint kind, the sorted unique weighted training vectors
// (a13 front half).  crn_comp restarts from pixels on every trial of a bitrate search (crn_texture_comp.cpp:120-262 -> crn_comp::compress_pass);
// none of this depends on the quality level, so crn_gpu_compress_crn keeps one HcPrepared across the trials.
template <int D> struct HcTrainingSet {
    bool valid = false;
    HcBuf d_tvec, d_uv, d_uw;                 // tile vectors (what the nearest-codebook search reads), unique vectors + weights
    uint32_t n_unique = 0;
    std::vector<float> uv; std::vector<uint32_t> uw;     // emulation build: the unique set on the host
};
struct HcPrepared {
    bool valid = false;
    uint32_t n = 0, num_tiles = 0;
    HcBuf d_enc, d_tile, d_npix, d_pixofs, d_vpix, d_cvec, d_avec, d_used;
    std::vector<uint8_t> h_npix, h_pixofs, h_enc;
    std::vector<uint32_t> h_tile, used_slots, slot_rank;
    HcTrainingSet<6> ts6;
    HcTrainingSet<2> ts2;
};

// a13 front half: the (component, tile) training vectors, compacted on the device into ts.d_tvec (what the nearest-codebook
// search reads afterwards), sorted and merged (once per HcPrepared); then the tree quantiser at this call's codebook size.  nvcc build:
// six / two stable LSD radix passes over the float bit patterns (cub) + run heads + scan, never leaving the device; emulation build: the
// same on the host.
template <int D>
int hc_endpoint_codebook(crn_gpu_ctx* ctx, int kind, const float* d_src, const uint32_t* d_used, uint32_t num_tiles, const uint8_t* d_npix,
                         uint32_t n, int ncp, const crn::HcLevelWeights& LW,
                         uint32_t max_size, HcTrainingSet<D>& ts, std::vector<float>& codebook, uint32_t& rounds, uint32_t& n_unique)
{
    cudaStream_t st = ctx->stream;
    const uint32_t NT = (uint32_t)ncp * num_tiles;
    HcTreeVq<D> vq;
    if (!ts.valid) {
    HcBuf& d_tvec = ts.d_tvec;
    HcBuf d_w;
    HC_ALLOC(d_tvec, (size_t)NT * D * 4); HC_ALLOC(d_w, (size_t)NT * 4);
    CRN_LAUNCH(crn::hc_compact_tiles_kernel<D>, (NT + 255) / 256, 256, 0, st, d_src, d_used, d_npix, n, num_tiles, NT, kind, LW, d_tvec.as<float>(), d_w.as<uint32_t>());
    ctx->launches++;
#ifdef __CUDACC__
    {
        HcBuf d_perm[2], d_keys[2], d_temp, d_head, d_rank, d_bsums;
        HcBuf& d_uv = ts.d_uv; HcBuf& d_uw = ts.d_uw;
        for (int k = 0; k < 2; k++) { HC_ALLOC(d_perm[k], (size_t)NT * 4); HC_ALLOC(d_keys[k], (size_t)NT * 4); }
        size_t temp_bytes = 0;
        CRN_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, d_keys[0].as<uint32_t>(), d_keys[1].as<uint32_t>(), d_perm[0].as<uint32_t>(), d_perm[1].as<uint32_t>(), (int)NT, 0, 32, st));
        HC_ALLOC(d_temp, temp_bytes);
        CRN_LAUNCH(crn::hc_iota_kernel, (NT + 255) / 256, 256, 0, st, d_perm[0].as<uint32_t>(), NT);
        int cur = 0;
        for (int comp = D - 1; comp >= 0; comp--) {       // least significant component first; every pass is stable
            CRN_LAUNCH(crn::hc_gather_key_kernel, (NT + 255) / 256, 256, 0, st, d_tvec.as<float>(), d_perm[cur].as<uint32_t>(), D, comp, NT, d_keys[0].as<uint32_t>());
            CRN_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_temp.p, temp_bytes, d_keys[0].as<uint32_t>(), d_keys[1].as<uint32_t>(), d_perm[cur].as<uint32_t>(), d_perm[cur ^ 1].as<uint32_t>(),
                                                          (int)NT, 0, 32, st));
            cur ^= 1;
            ctx->launches += 3;
        }
        const uint32_t m = NT + 1, nb = (m + 1023) / 1024;
        HC_ALLOC(d_head, (size_t)m * 4); HC_ALLOC(d_rank, (size_t)m * 4); HC_ALLOC(d_bsums, (size_t)(nb + 2) * 4);
        CRN_LAUNCH(crn::hc_vec_heads_kernel<D>, (m + 255) / 256, 256, 0, st, d_tvec.as<float>(), d_perm[cur].as<uint32_t>(), NT, d_head.as<uint32_t>());
        CRN_LAUNCH(crn::vq_scan_block_kernel, nb, 256, 0, st, d_head.as<uint32_t>(), d_rank.as<uint32_t>(), d_bsums.as<uint32_t>(), m);
        if (nb > 1) {
            CRN_LAUNCH(crn::vq_scan_sums_kernel, 1, 256, 0, st, d_bsums.as<uint32_t>(), nb);
            CRN_LAUNCH(crn::vq_scan_add_kernel, (m + 255) / 256, 256, 0, st, d_rank.as<uint32_t>(), d_bsums.as<uint32_t>(), m);
        }
        ctx->launches += 5;
        CRN_CUDA(ctx, cudaMemcpyAsync(&ts.n_unique, d_rank.as<uint32_t>() + NT, 4, cudaMemcpyDeviceToHost, st));
        CRN_CUDA(ctx, cudaStreamSynchronize(st));
        HC_ALLOC(d_uv, (size_t)ts.n_unique * D * 4); HC_ALLOC(d_uw, (size_t)ts.n_unique * 4);
        CRN_LAUNCH(crn::hc_vec_unique_kernel<D>, (NT + 255) / 256, 256, 0, st, d_tvec.as<float>(), d_w.as<uint32_t>(), d_perm[cur].as<uint32_t>(), d_head.as<uint32_t>(),
                   d_rank.as<uint32_t>(), NT, d_uv.as<float>(), d_uw.as<uint32_t>());
        ctx->launches++;
    }
#else
    {
        std::vector<HcVecKey<D>> keys(NT);
        std::vector<float> tv((size_t)NT * D); std::vector<uint32_t> tw(NT);
        CRN_CUDA(ctx, cudaMemcpyAsync(tv.data(), d_tvec.p, (size_t)NT * D * 4, cudaMemcpyDeviceToHost, st));
        CRN_CUDA(ctx, cudaMemcpyAsync(tw.data(), d_w.p, (size_t)NT * 4, cudaMemcpyDeviceToHost, st));
        CRN_CUDA(ctx, cudaStreamSynchronize(st));
        for (uint32_t i = 0; i < NT; i++) { memcpy(keys[i].v, &tv[(size_t)i * D], D * 4); keys[i].w = tw[i]; }
        hc_sort_dedup<D>(keys, ts.uv, ts.uw);
        ts.n_unique = (uint32_t)ts.uw.size();
    }
#endif
    ts.valid = true;
    }
    n_unique = ts.n_unique;
#ifdef __CUDACC__
    HC_RC(vq.build(ctx, ts.d_uv.template as<float>(), ts.d_uw.template as<uint32_t>(), ts.n_unique, max_size, true));
#else
    HC_RC(vq.build(ctx, ts.uv.data(), ts.uw.data(), ts.n_unique, max_size));
#endif
    codebook.swap(vq.codebook);
    rounds = vq.rounds;
    return CRN_GPU_OK;
}

int hc_compress_impl(crn_gpu_ctx* ctx, const crn_gpu_hc_params* prm, const void* blocks_rgba, int on_host, crn_gpu_hc* H, HcPrepared* prep = nullptr)
{
    const uint32_t n = prm->num_blocks;
    const uint32_t fmt = prm->format;
    const bool has_color = fmt == CRN_GPU_FMT_DXT1 || fmt == CRN_GPU_FMT_DXT5;
    const int na = fmt == CRN_GPU_FMT_DXT5 || fmt == CRN_GPU_FMT_DXT5A ? 1 : ((fmt == CRN_GPU_FMT_DXN_XY || fmt == CRN_GPU_FMT_DXN_YX) ? 2 : 0);
    if (!has_color && !na) return set_err(ctx, CRN_GPU_ERR_UNSUPPORTED, "crn_gpu_hc_compress: format must be DXT1, DXT5, DXT5A or DXN");
    const uint32_t comp0 = prm->alpha_component_indices[0], comp1 = prm->alpha_component_indices[1];
    const int perceptual = prm->perceptual ? 1 : 0;
    // ---- parameters of the tile pass (crn_dxt_hc.cpp:126-146)
    crn::HcTileParams TP; memset(&TP, 0, sizeof(TP));
    TP.num_levels = prm->num_levels; TP.num_faces = prm->num_faces; TP.has_color = has_color; TP.num_alpha = na;
    TP.alpha_comp[0] = comp0; TP.alpha_comp[1] = comp1; TP.color_alpha_ratio = prm->adaptive_tile_color_alpha_weighting_ratio;
    static const unsigned tile_derating[8] = { 0, 1, 1, 2, 2, 2, 2, 3 };
    uint32_t chunks = 0, expect = 0;
    for (uint32_t l = 0; l < prm->num_levels; l++) {
        const crn_gpu_hc_level& L = prm->levels[l];
        if (!L.block_width || (L.block_width & 1) || L.first_block != expect || !L.num_blocks || L.num_blocks % (2 * L.block_width * prm->num_faces))
            return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_hc_compress: levels must be contiguous, with even block widths and an even number of block rows per face");
        TP.levels[l].first_block = L.first_block; TP.levels[l].num_blocks = L.num_blocks; TP.levels[l].block_width = L.block_width;
        TP.levels[l].weight = L.weight; TP.levels[l].first_chunk = chunks;
        chunks += L.num_blocks / 4; expect += L.num_blocks;
        float der = prm->adaptive_tile_color_psnr_derating;
        if (l && der > .25f) { const float d = der / powf(3.0f, (float)l); der = d > .25f ? d : .25f; }
        for (int e = 0; e < 8; e++) TP.color_derating[l][e] = 0.0f + (der - 0.0f) * ((float)tile_derating[e] / 3.0f);
    }
    if (expect != n) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_hc_compress: levels do not cover num_blocks");
    for (int e = 0; e < 8; e++) TP.alpha_derating[e] = 0.0f + (prm->adaptive_tile_alpha_psnr_derating - 0.0f) * ((float)tile_derating[e] / 3.0f);
    TP.total_chunks = chunks;
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;

    QdxtTrace tr(ctx);
    // ---- a12: tiles (once per HcPrepared: the tile pass does not depend on the codebook sizes)
    HcPrepared local_prep;
    HcPrepared& PR = prep ? *prep : local_prep;
    if (PR.valid && PR.n != n) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_hc_compress: prepared state belongs to another block array");
    HcBuf d_blocks_own;
    HcBuf &d_enc = PR.d_enc, &d_tile = PR.d_tile, &d_npix = PR.d_npix, &d_pixofs = PR.d_pixofs, &d_vpix = PR.d_vpix, &d_cvec = PR.d_cvec, &d_avec = PR.d_avec, &d_used = PR.d_used;
    std::vector<uint8_t>& h_npix = PR.h_npix; std::vector<uint8_t>& h_pixofs = PR.h_pixofs;
    std::vector<uint32_t>& used_slots = PR.used_slots; std::vector<uint32_t>& slot_rank = PR.slot_rank;
    const uint32_t* d_blocks = static_cast<const uint32_t*>(blocks_rgba);
    if (!PR.valid) {
        if (on_host) {
            HC_ALLOC(d_blocks_own, (size_t)n * 64);
            CRN_CUDA(ctx, cudaMemcpyAsync(d_blocks_own.p, blocks_rgba, (size_t)n * 64, cudaMemcpyHostToDevice, st));
            d_blocks = d_blocks_own.as<uint32_t>();
        }
        HC_ALLOC(d_enc, n); HC_ALLOC(d_tile, (size_t)n * 4); HC_ALLOC(d_npix, n); HC_ALLOC(d_pixofs, n); HC_ALLOC(d_vpix, (size_t)n * 64);
        if (has_color) HC_ALLOC(d_cvec, (size_t)n * 24);
        if (na) HC_ALLOC(d_avec, (size_t)na * n * 8);
        CRN_LAUNCH(crn::hc_tiles_kernel, grid_for(ctx, chunks, crn::kHcTileWarps, 8), crn::kHcTileWarps * 32, 0, st, d_blocks, TP, d_enc.as<uint8_t>(), d_tile.as<uint32_t>(),
                   d_npix.as<uint8_t>(), d_pixofs.as<uint8_t>(), d_vpix.as<uint32_t>());
        const int ncomp = (has_color ? 1 : 0) + na;
        CRN_LAUNCH(crn::hc_palettize_kernel, (n * ncomp + 127) / 128, 128, 0, st, d_vpix.as<uint32_t>(), d_npix.as<uint8_t>(), d_pixofs.as<uint8_t>(), n, (int)has_color, na, comp0, comp1,
                   perceptual, d_cvec.as<float>(), d_avec.as<float>());
        ctx->launches += 2;
        CRN_CUDA(ctx, cudaGetLastError());
        HcHost<uint8_t> s_npix(ctx, n), s_pixofs(ctx, n), s_enc(ctx, n);
        HcHost<uint32_t> s_tile(ctx, n);
        CRN_CUDA(ctx, cudaMemcpyAsync(s_npix.data(), d_npix.p, n, cudaMemcpyDeviceToHost, st));
        CRN_CUDA(ctx, cudaMemcpyAsync(s_pixofs.data(), d_pixofs.p, n, cudaMemcpyDeviceToHost, st));
        CRN_CUDA(ctx, cudaMemcpyAsync(s_enc.data(), d_enc.p, n, cudaMemcpyDeviceToHost, st));
        CRN_CUDA(ctx, cudaMemcpyAsync(s_tile.data(), d_tile.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        CRN_CUDA(ctx, cudaStreamSynchronize(st));
        h_npix.assign(s_npix.data(), s_npix.data() + n); h_pixofs.assign(s_pixofs.data(), s_pixofs.data() + n);
        PR.h_enc.assign(s_enc.data(), s_enc.data() + n); PR.h_tile.assign(s_tile.data(), s_tile.data() + n);
        uint32_t nt = 0;
        for (uint32_t s = 0; s < n; s++) nt += h_npix[s] != 0;
        used_slots.resize(nt);                       // tile slots in order (m_tiles[t].pixels.size() != 0)
        for (uint32_t s = 0, i = 0; s < n; s++) if (h_npix[s]) used_slots[i++] = s;
        slot_rank.assign(n, 0xffffffffu);
        for (uint32_t i = 0; i < nt; i++) slot_rank[used_slots[i]] = i;
        PR.num_tiles = nt; PR.n = n;
        if (!nt) return set_err(ctx, CRN_GPU_ERR_BAD_DATA, "dxt_hc: no tiles");
        HC_ALLOC(d_used, (size_t)nt * 4);
        CRN_CUDA(ctx, cudaMemcpyAsync(d_used.p, used_slots.data(), (size_t)nt * 4, cudaMemcpyHostToDevice, st));
        CRN_CUDA(ctx, cudaStreamSynchronize(st));
        PR.valid = true;
        tr.mark("hc tiles + palettize + D2H", 0);
    }
    H->block_encodings = PR.h_enc; H->tile_indices = PR.h_tile;
    const uint32_t num_tiles = PR.num_tiles;
    H->info.num_tiles = num_tiles;
    crn::HcLevelWeights LW; memset(&LW, 0, sizeof(LW));
    LW.num_levels = prm->num_levels;
    for (uint32_t l = 0; l < prm->num_levels; l++) { LW.first_block[l] = prm->levels[l].first_block; LW.weight[l] = prm->levels[l].weight; }
    LW.first_block[prm->num_levels] = n;
    H->endpoint_indices.assign((size_t)n * 4, 0); H->selector_indices.assign((size_t)n * 4, 0);
    std::vector<uint16_t> raw_endpoint((size_t)n * 3, 0), raw_selector((size_t)n * 3, 0);
    std::vector<uint32_t> color_cluster_ep, alpha_cluster_ep;
    std::vector<uint8_t> color_cluster_used, alpha_cluster_used;
    std::vector<uint32_t> color_sel_cb; std::vector<uint64_t> alpha_sel_cb; std::vector<uint8_t> color_sel_used, alpha_sel_used;

    crn_gpu_pack_params pp; crn_gpu_default_pack_params(&pp);
    pp.dxt_quality = 4; pp.perceptual = (uint32_t)perceptual; pp.use_both_block_types = 0;

    // ---- one pass per endpoint kind: 0 colour, 1 alpha (all alpha channels share one codebook).  The two passes are
    // independent; for DXT5 they run concurrently, the alpha pass on a child context (own stream, scratch and buffer pool).
    CRN_CUDA(ctx, cudaStreamSynchronize(st));                          // d_used and everything the tile pass produced is complete
    crn_gpu_ctx* const parent = ctx;
    // Sharded calls: before the payload all-gather every rank exchanges one 16-byte status record (status, cluster count).  A rank that
    // leaves run_kind early -- error return or exception -- sends status 1 from this guard's destructor, so the other ranks see a failure
    // instead of waiting in the collective; a cluster-count mismatch (the ranks' tree quantisers diverged) is caught the same way.
    struct ShardAgreement {
        const crn_gpu_hc_params* prm; bool done;
        explicit ShardAgreement(const crn_gpu_hc_params* p) : prm(p), done(p->shard_count <= 1) {}
        int agree(uint32_t status, uint32_t K)
        {
            done = true;
            const uint32_t SC = prm->shard_count, SR = prm->shard_rank;
            std::vector<uint32_t> w((size_t)SC * 4, 0u);
            w[(size_t)SR * 4] = status; w[(size_t)SR * 4 + 1] = K; w[(size_t)SR * 4 + 2] = 0x43524E53u;
            if (prm->exchange(prm->exchange_user, w.data(), 16, SC) != 0) return -1;
            for (uint32_t r = 0; r < SC; r++) {
                if (w[(size_t)r * 4 + 2] != 0x43524E53u || w[(size_t)r * 4] != 0u) return 1;
                if (w[(size_t)r * 4 + 1] != K) return 2;
            }
            return 0;
        }
        ~ShardAgreement() { if (!done) { try { agree(1u, 0u); } catch (...) {} } }
    };
    auto run_kind = [&](crn_gpu_ctx* ctx, int kind) -> int {
        cudaStream_t st = ctx->stream;
        QdxtTrace tr(ctx);
        (void)parent;
        ShardAgreement agreement(prm);
        if (prm->shard_count > 1) {                                      // test hook: tests/test_shard_gloo.py makes one rank fail before the exchange
            const char* f = getenv("CRN_B200_TEST_FAIL_RANK");
            if (f && (uint32_t)atoi(f) == prm->shard_rank) return set_err(ctx, CRN_GPU_ERR_BAD_DATA, "crn_gpu_hc_compress: failure injected by CRN_B200_TEST_FAIL_RANK");
        }
        const int ncp = kind ? na : 1;                                   // components handled together
        const uint32_t NV = (uint32_t)ncp * n;                           // virtual blocks (= member blocks in the CSR)
        // a13: training vectors -> sorted unique weighted vectors -> tree quantiser
        std::vector<float> codebook; uint32_t K = 0;
        HcBuf& d_tvec = kind == 0 ? PR.ts6.d_tvec : PR.ts2.d_tvec;
        if (kind == 0) {
            HC_RC(hc_endpoint_codebook<6>(ctx, 0, d_cvec.as<float>(), d_used.as<uint32_t>(), num_tiles, d_npix.as<uint8_t>(), n, 1, LW,
                                          std::min(num_tiles, prm->color_endpoint_codebook_size), PR.ts6, codebook, H->info.vq_rounds[0], H->info.unique_vectors[0]));
            K = (uint32_t)(codebook.size() / 6);
        } else {
            HC_RC(hc_endpoint_codebook<2>(ctx, 1, d_avec.as<float>(), d_used.as<uint32_t>(), num_tiles, d_npix.as<uint8_t>(), n, na, LW,
                                          std::min(num_tiles, prm->alpha_endpoint_codebook_size), PR.ts2, codebook, H->info.vq_rounds[1], H->info.unique_vectors[1]));
            K = (uint32_t)(codebook.size() / 2);
        }
        tr.mark("hc endpoint sort + tree VQ", kind);
        if (!K) return set_err(ctx, CRN_GPU_ERR_BAD_DATA, "dxt_hc: empty endpoint codebook");
        // a14: nearest codebook entry per tile (per component)
        const uint32_t dims = kind ? 2 : 6, NT = (uint32_t)ncp * num_tiles;
        HcBuf d_cb, d_tcl;
        HC_ALLOC(d_cb, (size_t)K * dims * 4); HC_ALLOC(d_tcl, (size_t)NT * 4);
        CRN_CUDA(ctx, cudaMemcpyAsync(d_cb.p, codebook.data(), (size_t)K * dims * 4, cudaMemcpyHostToDevice, st));
        HC_RC(crn_gpu_nearest_codebook(ctx, dims, d_tvec.as<float>(), NT, d_cb.as<float>(), K, d_tcl.as<uint32_t>()));
        HcHost<uint32_t> tcl(ctx, NT);
        CRN_CUDA(ctx, cudaMemcpyAsync(tcl.data(), d_tcl.p, (size_t)NT * 4, cudaMemcpyDeviceToHost, st));
        CRN_CUDA(ctx, cudaStreamSynchronize(st));
        tr.mark("hc nearest codebook", kind);
        // cluster member lists: the tiles' 16-pixel virtual blocks, component-major then tile-slot order (:970-977, :1246-1260)
        std::vector<uint32_t> offs(K + 1, 0);
        HcHost<uint32_t> members(ctx, NV), block_cluster(ctx, NV);
        {
            // a stable counting sort of the (component, tile) sequence by cluster, in chunks: every chunk counts its clusters, the counts are
            // prefixed chunk-major inside each cluster, every chunk then writes its own slices -- the order inside a cluster stays the sequence's
#ifdef __CUDACC__
            const uint32_t nchunk = NT >= 65536 ? 8u : 1u;
#else
            const uint32_t nchunk = 1u;
#endif
            std::vector<std::vector<uint32_t>> cnt(nchunk, std::vector<uint32_t>(K, 0u));
            auto chunk_range = [&](uint32_t t, uint32_t& j0, uint32_t& j1) { j0 = (uint32_t)((uint64_t)NT * t / nchunk); j1 = (uint32_t)((uint64_t)NT * (t + 1) / nchunk); };
            auto count_chunk = [&](uint32_t t) {
                uint32_t j0, j1; chunk_range(t, j0, j1);
                std::vector<uint32_t>& c = cnt[t];
                for (uint32_t j = j0; j < j1; j++) c[tcl[j]] += h_npix[used_slots[j % num_tiles]] / 16;
            };
            auto fill_chunk = [&](uint32_t t) {
                uint32_t j0, j1; chunk_range(t, j0, j1);
                std::vector<uint32_t>& cur = cnt[t];              // now: this chunk's first slot in every cluster
                for (uint32_t j = j0; j < j1; j++) {
                    const uint32_t a = j / num_tiles, s = used_slots[j % num_tiles], c = tcl[j];
                    const uint32_t vb0 = a * n + (s & ~3u) + h_pixofs[s] / 16;
                    for (uint32_t k = 0; k < h_npix[s] / 16u; k++) members[cur[c]++] = vb0 + k;
                }
            };
            auto run_chunks = [&](auto fn) {
                std::vector<std::thread> th;
                for (uint32_t t = 1; t < nchunk; t++) th.emplace_back(fn, t);
                fn(0u);
                for (auto& x : th) x.join();
            };
            run_chunks(count_chunk);
            for (uint32_t c = 0; c < K; c++) {
                uint32_t run = offs[c];
                for (uint32_t t = 0; t < nchunk; t++) { const uint32_t v = cnt[t][c]; cnt[t][c] = run; run += v; }
                offs[c + 1] = run;
            }
            run_chunks(fill_chunk);
        }
        {
            auto per_block = [&](uint32_t b0, uint32_t b1) {
                for (int a = 0; a < ncp; a++)
                    for (uint32_t b = b0; b < b1; b++) {
                        const uint32_t c = tcl[(size_t)a * num_tiles + slot_rank[H->tile_indices[b]]];
                        block_cluster[(size_t)a * n + b] = c;
                        raw_endpoint[(size_t)b * 3 + (kind ? 1 + a : 0)] = (uint16_t)c;
                    }
            };
#ifdef __CUDACC__
            if (n >= 262144) {
                const uint32_t nt = 6;
                std::vector<std::thread> th;
                for (uint32_t t = 1; t < nt; t++) th.emplace_back(per_block, (uint32_t)((uint64_t)n * t / nt), (uint32_t)((uint64_t)n * (t + 1) / nt));
                per_block(0, n / nt);
                for (auto& x : th) x.join();
            } else
#endif
            per_block(0, n);
        }
        std::vector<uint32_t> poffs(K + 1);
        for (uint32_t c = 0; c <= K; c++) poffs[c] = offs[c] * 16;
        tr.mark("hc CSR build (host)", kind);
        // a15: per-cluster optimiser over the virtual blocks
        HcBuf d_vsrc, d_offs, d_mem, d_elem, d_ep, d_err, d_flags, d_bcl, d_bsel, d_bval, d_cpix, d_csel, d_poffs, d_rep, d_rerr, d_rok, d_bacc, d_grey;
        const uint32_t* d_vblocks = d_vpix.as<uint32_t>();
        if (kind == 1) {                                                 // grey copies of the tile-ordered pixels, one plane per channel
            HC_ALLOC(d_vsrc, (size_t)NV * 64);
            for (int a = 0; a < na; a++)
                CRN_LAUNCH(crn::hc_grey_kernel, (unsigned)(((size_t)n * 16 + 255) / 256), 256, 0, st, d_vpix.as<uint32_t>(), (size_t)n * 16, a ? comp1 : comp0, d_vsrc.as<uint32_t>() + (size_t)a * n * 16);
            ctx->launches += na;
            d_vblocks = d_vsrc.as<uint32_t>();
        }
        std::vector<uint32_t> order(K);
        for (uint32_t c = 0; c < K; c++) order[c] = c;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return offs[a + 1] - offs[a] > offs[b + 1] - offs[b]; });
        if (tr.on) fprintf(stderr, "[crn_b200] kind %d: %u clusters, %u member blocks, largest %u, median %u\n", kind, K, NV, offs[order[0] + 1] - offs[order[0]], offs[order[K / 2] + 1] - offs[order[K / 2]]);
        // The clusters this rank optimises: all of them, or every shard_count-th of the size-ordered list (crn_gpu_hc_params).
        const uint32_t SC = prm->shard_count, SR = prm->shard_rank;
        const uint32_t Kopt = SC > 1 ? (K > SR ? (K - SR + SC - 1) / SC : 0) : K;
        std::vector<uint32_t> opt_offs(Kopt + 1, 0), opt_poffs(Kopt + 1, 0), opt_order(Kopt);
        uint32_t NVopt = NV;
        if (SC > 1) {
            for (uint32_t j = 0; j < Kopt; j++) { const uint32_t c = order[SR + j * SC]; opt_offs[j + 1] = opt_offs[j] + (offs[c + 1] - offs[c]); }
            NVopt = opt_offs[Kopt];
        } else opt_offs = offs;
        HcHost<uint32_t> opt_members(ctx, SC > 1 ? NVopt : 0);
        if (SC > 1)
            for (uint32_t j = 0; j < Kopt; j++) { const uint32_t c = order[SR + j * SC]; memcpy(opt_members.data() + opt_offs[j], members.data() + offs[c], (size_t)(offs[c + 1] - offs[c]) * 4); }
        for (uint32_t j = 0; j <= Kopt; j++) opt_poffs[j] = opt_offs[j] * 16;
        for (uint32_t j = 0; j < Kopt; j++) opt_order[j] = SC > 1 ? j : order[j];               // a rank's own list is already largest-first
        const uint32_t* h_opt_members = SC > 1 ? opt_members.data() : members.data();
        HC_ALLOC(d_offs, (size_t)(Kopt + 1) * 4); HC_ALLOC(d_poffs, (size_t)(Kopt + 1) * 4); HC_ALLOC(d_mem, (size_t)NVopt * 4 + 4); HC_ALLOC(d_elem, (size_t)NV * 8);
        HC_ALLOC(d_ep, (size_t)K * 4); HC_ALLOC(d_err, (size_t)K * 8); HC_ALLOC(d_flags, (size_t)K * 4); HC_ALLOC(d_bcl, (size_t)NV * 4);
        CRN_CUDA(ctx, cudaMemcpyAsync(d_offs.p, opt_offs.data(), (size_t)(Kopt + 1) * 4, cudaMemcpyHostToDevice, st));
        CRN_CUDA(ctx, cudaMemcpyAsync(d_poffs.p, opt_poffs.data(), (size_t)(Kopt + 1) * 4, cudaMemcpyHostToDevice, st));
        if (NVopt) CRN_CUDA(ctx, cudaMemcpyAsync(d_mem.p, h_opt_members, (size_t)NVopt * 4, cudaMemcpyHostToDevice, st));
        CRN_CUDA(ctx, cudaMemcpyAsync(d_bcl.p, block_cluster.data(), (size_t)NV * 4, cudaMemcpyHostToDevice, st));
        CRN_CUDA(ctx, cudaMemsetAsync(d_ep.p, 0, (size_t)K * 4, st));
        CRN_CUDA(ctx, cudaMemsetAsync(d_err.p, 0, (size_t)K * 8, st));
        CRN_CUDA(ctx, cudaMemsetAsync(d_flags.p, 0, (size_t)K * 4, st));
        HcBuf d_order;
        HC_ALLOC(d_order, (size_t)Kopt * 4 + 4);
        if (Kopt) CRN_CUDA(ctx, cudaMemcpyAsync(d_order.p, opt_order.data(), (size_t)Kopt * 4, cudaMemcpyHostToDevice, st));
        ctx->d_cluster_flags = d_flags.as<uint32_t>(); ctx->d_cluster_order = d_order.as<uint32_t>();
        ctx->cluster_big_count = 0;
        for (uint32_t j = 0; j < Kopt; j++) { const uint32_t c = opt_order[j]; if (opt_offs[c + 1] - opt_offs[c] < crn::kClusterCoopMinBlocks) break; ctx->cluster_big_count++; }
        int rc = kind == 0 ? crn_gpu_dxt1_optimize_clusters(ctx, &pp, 0, d_vblocks, NV, d_offs.as<uint32_t>(), d_mem.as<uint32_t>(), Kopt, NVopt, d_elem.p, 8, 0, d_ep.as<uint32_t>(), d_err.as<uint64_t>())
                           : crn_gpu_dxt5_optimize_clusters(ctx, &pp, 0, d_vblocks, NV, d_offs.as<uint32_t>(), d_mem.as<uint32_t>(), Kopt, NVopt, d_elem.p, 8, 0, d_ep.as<uint32_t>(), d_err.as<uint64_t>());
        ctx->d_cluster_flags = nullptr; ctx->d_cluster_order = nullptr;
        if (rc) return rc;
        tr.mark("hc cluster optimiser", kind);
        // refiner over the cluster pixel lists with the optimiser's selectors (a10)
        HC_ALLOC(d_cpix, (size_t)NVopt * 64 + 64); HC_ALLOC(d_csel, (size_t)NVopt * 16 + 16); HC_ALLOC(d_rep, (size_t)K * 4); HC_ALLOC(d_rerr, (size_t)K * 8); HC_ALLOC(d_rok, K);
        if (NVopt) {
            CRN_LAUNCH(crn::hc_gather_cluster_pixels_kernel, (unsigned)(((size_t)NVopt * 16 + 255) / 256), 256, 0, st, d_vblocks, d_mem.as<uint32_t>(), NVopt * 16, kind,
                       d_elem.as<unsigned long long>(), d_cpix.as<uint32_t>(), d_csel.as<uint8_t>());
            ctx->launches++;
        }
        ctx->refine_parallel = (parent->vq_exact || getenv("CRN_B200_ORDERED_SUMS")) ? 0 : 1;
        {
            const int rrc = crn_gpu_refine_endpoints(ctx, kind == 0 ? 1 : 0, perceptual, 0, d_cpix.p, d_csel.as<uint8_t>(), d_poffs.as<uint32_t>(), Kopt, d_err.as<uint64_t>(),
                                                     d_rep.as<uint32_t>(), d_rerr.as<uint64_t>(), d_rok.as<uint8_t>());
            ctx->refine_parallel = 0;
            if (rrc) return rrc;
        }
        std::vector<uint32_t> ep(K, 0), rep(K, 0); std::vector<uint8_t> rok(K, 0);
        if (SC > 1) {
            // this rank's results -> all ranks: one 16-byte record per cluster through the caller's all-gather
            std::vector<uint32_t> s_ep(Kopt), s_fl(Kopt), s_rep(Kopt); std::vector<uint8_t> s_rok(Kopt);
            if (Kopt) {
                CRN_CUDA(ctx, cudaMemcpyAsync(s_ep.data(), d_ep.p, (size_t)Kopt * 4, cudaMemcpyDeviceToHost, st));
                CRN_CUDA(ctx, cudaMemcpyAsync(s_fl.data(), d_flags.p, (size_t)Kopt * 4, cudaMemcpyDeviceToHost, st));
                CRN_CUDA(ctx, cudaMemcpyAsync(s_rep.data(), d_rep.p, (size_t)Kopt * 4, cudaMemcpyDeviceToHost, st));
                CRN_CUDA(ctx, cudaMemcpyAsync(s_rok.data(), d_rok.p, Kopt, cudaMemcpyDeviceToHost, st));
            }
            CRN_CUDA(ctx, cudaStreamSynchronize(st));
            {
                const int ag = agreement.agree(0u, K);
                if (ag < 0) return set_err(ctx, CRN_GPU_ERR_CUDA, "crn_gpu_hc_compress: the exchange callback failed");
                if (ag == 1) return set_err(ctx, CRN_GPU_ERR_BAD_DATA, "crn_gpu_hc_compress (sharded): another rank failed before the exchange");
                if (ag == 2) return set_err(ctx, CRN_GPU_ERR_BAD_DATA, "crn_gpu_hc_compress (sharded): the ranks disagree on the number of clusters");
            }
            const uint32_t per = (K + SC - 1) / SC;
            std::vector<uint32_t> xbuf((size_t)SC * per * 4, 0);
            for (uint32_t j = 0; j < Kopt; j++) { uint32_t* r = &xbuf[((size_t)SR * per + j) * 4]; r[0] = s_ep[j]; r[1] = s_fl[j]; r[2] = s_rep[j]; r[3] = s_rok[j]; }
            if (prm->exchange(prm->exchange_user, xbuf.data(), (uint64_t)per * 16, SC) != 0) return set_err(ctx, CRN_GPU_ERR_CUDA, "crn_gpu_hc_compress: the exchange callback failed");
            std::vector<uint32_t> fl(K, 0);
            for (uint32_t r = 0; r < SC; r++)
                for (uint32_t j = 0; r + j * SC < K; j++) {
                    const uint32_t c = order[r + j * SC]; const uint32_t* q = &xbuf[((size_t)r * per + j) * 4];
                    ep[c] = q[0]; fl[c] = q[1]; rep[c] = q[2]; rok[c] = (uint8_t)q[3];
                }
            CRN_CUDA(ctx, cudaMemcpyAsync(d_ep.p, ep.data(), (size_t)K * 4, cudaMemcpyHostToDevice, st));
            CRN_CUDA(ctx, cudaMemcpyAsync(d_flags.p, fl.data(), (size_t)K * 4, cudaMemcpyHostToDevice, st));
            CRN_CUDA(ctx, cudaMemcpyAsync(d_rep.p, rep.data(), (size_t)K * 4, cudaMemcpyHostToDevice, st));
            CRN_CUDA(ctx, cudaMemcpyAsync(d_rok.p, rok.data(), K, cudaMemcpyHostToDevice, st));
            CRN_CUDA(ctx, cudaStreamSynchronize(st));              // fl lives on this frame
        } else {
            CRN_CUDA(ctx, cudaMemcpyAsync(ep.data(), d_ep.p, (size_t)K * 4, cudaMemcpyDeviceToHost, st));
            CRN_CUDA(ctx, cudaMemcpyAsync(rep.data(), d_rep.p, (size_t)K * 4, cudaMemcpyDeviceToHost, st));
            CRN_CUDA(ctx, cudaMemcpyAsync(rok.data(), d_rok.p, K, cudaMemcpyDeviceToHost, st));
        }
        // per-block selectors + weights against the cluster palette (all clusters, every rank)
        HC_ALLOC(d_bsel, (size_t)NV * 8); HC_ALLOC(d_bval, (size_t)NV * (kind ? 8 : 16));
        if (kind == 0) {
            CRN_LAUNCH(crn::hc_color_blocks_kernel, (n + 255) / 256, 256, 0, st, d_blocks, n, d_bcl.as<uint32_t>(), d_ep.as<uint32_t>(), d_flags.as<uint32_t>(), LW,
                       d_enc.as<uint8_t>(), perceptual, d_bsel.as<unsigned long long>(), d_bval.as<uint32_t>());
        } else {
            CRN_LAUNCH(crn::hc_alpha_blocks_kernel, (NV + 255) / 256, 256, 0, st, d_blocks, n, na, comp0, comp1, d_bcl.as<uint32_t>(), d_ep.as<uint32_t>(), d_flags.as<uint32_t>(),
                       d_enc.as<uint8_t>(), d_bsel.as<unsigned long long>(), d_bval.as<uint8_t>());
        }
        ctx->launches++;
        // a16: selector codebook.  Training set = the distinct block selectors with summed weights (:1379-1444, :1588-1660)
        std::vector<uint32_t>& cl_ep = kind ? alpha_cluster_ep : color_cluster_ep;
        std::vector<uint8_t>& cl_used = kind ? alpha_cluster_used : color_cluster_used;
        cl_ep.resize(K); cl_used.resize(K);
        HcTreeVq<16> svq;
        const uint32_t max_sel = kind ? prm->alpha_selector_codebook_size : prm->color_selector_codebook_size;
        uint32_t n_unique_sel = 0;
#ifdef __CUDACC__
        {   // sort, run heads, ranks, vectors + summed weights: all on the device
            HcBuf d_sorted, d_temp, d_head, d_rank, d_bsums, d_sv, d_sw;
            const int wshift = kind ? 16 : 32;
            HC_ALLOC(d_sorted, (size_t)NV * 8);
            size_t temp_bytes = 0;
            CRN_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, d_bsel.as<unsigned long long>(), d_sorted.as<unsigned long long>(), (int)NV, 0, 64, st));
            HC_ALLOC(d_temp, temp_bytes);
            CRN_CUDA(ctx, cub::DeviceRadixSort::SortKeys(d_temp.p, temp_bytes, d_bsel.as<unsigned long long>(), d_sorted.as<unsigned long long>(), (int)NV, 0, 64, st));
            const uint32_t m = NV + 1, nb = (m + 1023) / 1024;
            HC_ALLOC(d_head, (size_t)m * 4); HC_ALLOC(d_rank, (size_t)m * 4); HC_ALLOC(d_bsums, (size_t)(nb + 2) * 4);
            CRN_LAUNCH(crn::hc_sel_heads_kernel, (m + 255) / 256, 256, 0, st, d_sorted.as<unsigned long long>(), NV, wshift, d_head.as<uint32_t>());
            CRN_LAUNCH(crn::vq_scan_block_kernel, nb, 256, 0, st, d_head.as<uint32_t>(), d_rank.as<uint32_t>(), d_bsums.as<uint32_t>(), m);
            if (nb > 1) {
                CRN_LAUNCH(crn::vq_scan_sums_kernel, 1, 256, 0, st, d_bsums.as<uint32_t>(), nb);
                CRN_LAUNCH(crn::vq_scan_add_kernel, (m + 255) / 256, 256, 0, st, d_rank.as<uint32_t>(), d_bsums.as<uint32_t>(), m);
            }
            ctx->launches += 5;
            CRN_CUDA(ctx, cudaMemcpyAsync(&n_unique_sel, d_rank.as<uint32_t>() + NV, 4, cudaMemcpyDeviceToHost, st));
            CRN_CUDA(ctx, cudaStreamSynchronize(st));
            for (uint32_t c = 0; c < K; c++) {
                cl_used[c] = offs[c + 1] != offs[c];
                if (kind == 0) cl_ep[c] = rok[c] ? rep[c] : ep[c];                                           // dxt1_block::pack_endpoints(first, second) = low | high << 16
                else cl_ep[c] = rok[c] ? ((rep[c] & 0xffff) | ((rep[c] >> 16) << 8)) : (ep[c] & 0xffff);     // dxt5_block::pack_endpoints = first | second << 8
            }
            HC_ALLOC(d_sv, (size_t)n_unique_sel * 64); HC_ALLOC(d_sw, (size_t)n_unique_sel * 4);
            CRN_LAUNCH(crn::hc_sel_vectors_kernel, (NV + 255) / 256, 256, 0, st, d_sorted.as<unsigned long long>(), d_head.as<uint32_t>(), d_rank.as<uint32_t>(), NV, kind,
                       d_sv.as<float>(), d_sw.as<uint32_t>());
            ctx->launches++;
            tr.mark("hc selector sort + dedup (device)", kind);
            HC_RC(svq.build(ctx, d_sv.as<float>(), d_sw.as<uint32_t>(), n_unique_sel, max_sel, true));
        }
#else
        {   // emulation build (no cub): the same steps on the host
            std::vector<unsigned long long> bsel(NV);
            CRN_CUDA(ctx, cudaMemcpyAsync(bsel.data(), d_bsel.p, (size_t)NV * 8, cudaMemcpyDeviceToHost, st));
            CRN_CUDA(ctx, cudaStreamSynchronize(st));
            for (uint32_t c = 0; c < K; c++) {
                cl_used[c] = offs[c + 1] != offs[c];
                if (kind == 0) cl_ep[c] = rok[c] ? rep[c] : ep[c];
                else cl_ep[c] = rok[c] ? ((rep[c] & 0xffff) | ((rep[c] >> 16) << 8)) : (ep[c] & 0xffff);
            }
            std::sort(bsel.begin(), bsel.end());
            std::vector<float> sv; std::vector<uint32_t> sw;
            const int bits = kind ? 3 : 2, wshift = kind ? 16 : 32;
            unsigned long long prev = 0;
            float lut[8];
            for (int s = 0; s < 8; s++) lut[s] = kind ? ((float)s + 0.5f) * 0.125f : ((float)s + 0.5f) * 0.25f;
            for (size_t i = 0; i < bsel.size(); i++) {
                const uint32_t weight = kind ? (uint32_t)(bsel[i] & 0xffff) : (uint32_t)bsel[i];
                unsigned long long sel = bsel[i] >> wshift;
                if (sw.empty() || sel != prev) {
                    prev = sel;
                    float v[16];
                    for (int p = 0; p < 16; p++, sel >>= bits) v[15 - p] = lut[sel & ((1u << bits) - 1)];
                    sv.insert(sv.end(), v, v + 16); sw.push_back(weight);
                } else if (sw.back() > 0xffffffffu - weight) sw.back() = 0xffffffffu;
                else sw.back() += weight;
            }
            n_unique_sel = (uint32_t)sw.size();
            tr.mark("hc selector sort + dedup (host)", kind);
            HC_RC(svq.build(ctx, sv.data(), sw.data(), n_unique_sel, max_sel));
        }
#endif
        const uint32_t KS = svq.size();
        H->info.vq_rounds[2 + kind] = svq.rounds; H->info.unique_vectors[2 + kind] = n_unique_sel;
        tr.mark("hc selector tree VQ", kind);
        if (!KS) return set_err(ctx, CRN_GPU_ERR_BAD_DATA, "dxt_hc: empty selector codebook");
        std::vector<uint64_t> scb(KS);
        for (uint32_t i = 0; i < KS; i++) {
            uint64_t s = 0;
            for (int j = 0; j < 16; j++) s |= (uint64_t)(uint32_t)(svq.codebook[(size_t)i * 16 + j] * (kind ? 8.0f : 4.0f)) << ((kind ? 3 : 2) * j);
            scb[i] = kind ? s : (uint64_t)(uint32_t)s;
        }
        // exhaustive search + re-vote
        HcBuf d_scb, d_best, d_refined, d_used;
        HC_ALLOC(d_scb, (size_t)KS * 8); HC_ALLOC(d_best, (size_t)NV * 4); HC_ALLOC(d_refined, (size_t)KS * 8); HC_ALLOC(d_used, KS);
        CRN_CUDA(ctx, cudaMemcpyAsync(d_scb.p, scb.data(), (size_t)KS * 8, cudaMemcpyHostToDevice, st));
        const void* search_blocks = d_blocks;
        const void* accum = nullptr;
        if (kind == 1) {
            HC_ALLOC(d_grey, (size_t)NV * 64); HC_ALLOC(d_bacc, (size_t)NV * 8);
            for (int a = 0; a < na; a++)
                CRN_LAUNCH(crn::hc_grey_kernel, (unsigned)(((size_t)n * 16 + 255) / 256), 256, 0, st, d_blocks, (size_t)n * 16, a ? comp1 : comp0, d_grey.as<uint32_t>() + (size_t)a * n * 16);
            CRN_LAUNCH(crn::hc_alpha_refined_values_kernel, (NV + 255) / 256, 256, 0, st, NV, d_bcl.as<uint32_t>(), d_rep.as<uint32_t>(), d_rok.as<uint8_t>(), d_bval.as<uint8_t>(), d_bacc.as<uint8_t>());
            ctx->launches += na + 1;
            search_blocks = d_grey.p; accum = d_bacc.p;
        }
        HC_RC(crn_gpu_assign_selectors(ctx, (uint32_t)kind, perceptual, 0, search_blocks, NV, d_bval.p, accum, d_scb.as<uint64_t>(), KS, d_best.as<uint32_t>(), d_refined.as<uint64_t>(), d_used.as<uint8_t>()));
        HcHost<uint32_t> best(ctx, NV); std::vector<uint64_t> refined(KS); std::vector<uint8_t> used(KS);
        CRN_CUDA(ctx, cudaMemcpyAsync(best.data(), d_best.p, (size_t)NV * 4, cudaMemcpyDeviceToHost, st));
        CRN_CUDA(ctx, cudaMemcpyAsync(refined.data(), d_refined.p, (size_t)KS * 8, cudaMemcpyDeviceToHost, st));
        CRN_CUDA(ctx, cudaMemcpyAsync(used.data(), d_used.p, KS, cudaMemcpyDeviceToHost, st));
        CRN_CUDA(ctx, cudaStreamSynchronize(st));
        tr.mark("hc selector search + re-vote", kind);
        for (int a = 0; a < ncp; a++)
            for (uint32_t b = 0; b < n; b++) raw_selector[(size_t)b * 3 + (kind ? 1 + a : 0)] = (uint16_t)best[(size_t)a * n + b];
        if (kind == 0) { color_sel_cb.resize(KS); for (uint32_t i = 0; i < KS; i++) color_sel_cb[i] = (uint32_t)refined[i]; color_sel_used = used; }
        else { alpha_sel_cb = refined; alpha_sel_used = used; }
        return CRN_GPU_OK;
    };
    // (sharded calls stay sequential: the two passes' exchanges must reach every rank in the same order)
    if (has_color && na && prm->shard_count > 1) {
        HC_RC(run_kind(ctx, 0));
        HC_RC(run_kind(ctx, 1));
    } else if (has_color && na) {
#ifdef __CUDACC__
        if (!ctx->child[0] && crn_gpu_create(ctx->device, &ctx->child[0]) != CRN_GPU_OK) return set_err(ctx, CRN_GPU_ERR_CUDA, "crn_gpu_hc_compress: element stream");
        crn_gpu_ctx* child = ctx->child[0];
        const uint64_t l0 = child->launches;
        int rc1 = CRN_GPU_OK;
        std::thread alpha_thread([&] {
            cudaSetDevice(child->device);
            try { rc1 = run_kind(child, 1); }
            catch (const std::bad_alloc&) { rc1 = set_err(child, CRN_GPU_ERR_NO_MEMORY, "crn_gpu_hc_compress: out of host memory"); }
            catch (...) { rc1 = set_err(child, CRN_GPU_ERR_BAD_DATA, "crn_gpu_hc_compress: exception in the alpha pass"); }   // nothing escapes a thread
            child->d_cluster_flags = nullptr; child->d_cluster_order = nullptr;
        });
        int rc0;
        try { rc0 = run_kind(ctx, 0); }
        catch (const std::bad_alloc&) { rc0 = set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "crn_gpu_hc_compress: out of host memory"); }
        catch (...) { rc0 = set_err(ctx, CRN_GPU_ERR_BAD_DATA, "crn_gpu_hc_compress: exception in the colour pass"); }
        alpha_thread.join();                                        // never leave the frame with the thread running
        ctx->launches += child->launches - l0;
        if (rc0) return rc0;
        if (rc1) return set_err(ctx, rc1, child->err);
#else
        HC_RC(run_kind(ctx, 0));                                        // the emulator is single-threaded
        HC_RC(run_kind(ctx, 1));
#endif
    } else HC_RC(run_kind(ctx, has_color ? 0 : 1));
    tr.mark("hc endpoint + selector passes (total)", 0);

    // ---- a17: palette dedup, index remap, reference flags (crn_dxt_hc.cpp:200-310)
    auto dedup32 = [](const std::vector<uint32_t>& in, const std::vector<uint8_t>& used, std::vector<uint32_t>& out, std::vector<uint16_t>& remap) {
        remap.assign(in.size(), 0);
        std::vector<int64_t> table(1u << 17, -1);
        for (size_t i = 0; i < in.size(); i++) {
            if (!used[i]) continue;
            uint32_t h = (in[i] * 2654435761u) >> 15;
            for (;;) {
                if (table[h] < 0) { table[h] = (int64_t)out.size(); remap[i] = (uint16_t)out.size(); out.push_back(in[i]); break; }
                if (out[(size_t)table[h]] == in[i]) { remap[i] = (uint16_t)table[h]; break; }
                h = (h + 1) & ((1u << 17) - 1);
            }
        }
    };
    auto dedup64 = [](const std::vector<uint64_t>& in, const std::vector<uint8_t>& used, std::vector<uint64_t>& out, std::vector<uint16_t>& remap) {
        remap.assign(in.size(), 0);
        std::vector<int64_t> table(1u << 17, -1);
        for (size_t i = 0; i < in.size(); i++) {
            if (!used[i]) continue;
            uint32_t h = (uint32_t)((in[i] * 0x9E3779B97F4A7C15ull) >> 47);
            for (;;) {
                if (table[h] < 0) { table[h] = (int64_t)out.size(); remap[i] = (uint16_t)out.size(); out.push_back(in[i]); break; }
                if (out[(size_t)table[h]] == in[i]) { remap[i] = (uint16_t)table[h]; break; }
                h = (h + 1) & ((1u << 17) - 1);
            }
        }
    };
    std::vector<uint16_t> ce_remap, ae_remap, cs_remap, as_remap;
    dedup32(color_cluster_ep, color_cluster_used, H->color_endpoints, ce_remap);
    dedup32(alpha_cluster_ep, alpha_cluster_used, H->alpha_endpoints, ae_remap);
    dedup32(color_sel_cb, color_sel_used, H->color_selectors, cs_remap);
    dedup64(alpha_sel_cb, alpha_sel_used, H->alpha_selectors, as_remap);
    // every block's record depends only on the raw indices of itself and its left / upper neighbour, so the rows go to host threads
    for (uint32_t l = 0; l < prm->num_levels; l++) {
        const uint32_t first = prm->levels[l].first_block, nblk = prm->levels[l].num_blocks, bw = prm->levels[l].block_width;
        const uint32_t rows = bw ? nblk / bw : 0;
        auto do_rows = [&](uint32_t y0, uint32_t y1) {
            for (uint32_t by = y0; by < y1; by++)
                for (uint32_t bx = 0; bx < bw; bx++) {
                    const uint32_t b = first + by * bw + bx;
                    bool top = by != 0, left = top || bx;
                    for (int c = has_color ? 0 : 1; c < 1 + na; c++) {
                        const std::vector<uint16_t>& er = c ? ae_remap : ce_remap;
                        const uint16_t e = er[raw_endpoint[(size_t)b * 3 + c]];
                        left = left && e == er[raw_endpoint[(size_t)(b - 1) * 3 + c]];      // (b - 1 / b - bw are only read when the flag is still set)
                        top = top && e == er[raw_endpoint[(size_t)(b - bw) * 3 + c]];
                        H->endpoint_indices[(size_t)b * 4 + c] = e;
                        H->selector_indices[(size_t)b * 4 + c] = (c ? as_remap : cs_remap)[raw_selector[(size_t)b * 3 + c]];
                    }
                    H->endpoint_indices[(size_t)b * 4 + 3] = left ? 1 : (top ? 2 : 0);
                }
        };
#ifdef __CUDACC__
        if (nblk >= 65536) {
            const uint32_t nt = 8;
            std::vector<std::thread> th;
            for (uint32_t t = 1; t < nt; t++) th.emplace_back(do_rows, (uint32_t)((uint64_t)rows * t / nt), (uint32_t)((uint64_t)rows * (t + 1) / nt));
            do_rows(0, rows / nt);
            for (auto& x : th) x.join();
            continue;
        }
#endif
        do_rows(0, rows);
    }
    tr.mark("hc tail (host)", 0);
    H->info.num_blocks = n;
    H->info.n_color_endpoints = (uint32_t)H->color_endpoints.size(); H->info.n_alpha_endpoints = (uint32_t)H->alpha_endpoints.size();
    H->info.n_color_selectors = (uint32_t)H->color_selectors.size(); H->info.n_alpha_selectors = (uint32_t)H->alpha_selectors.size();
    return CRN_GPU_OK;
}

}  // namespace
