// vq_fast_host.h -- host driver of the single-launch-per-frontier vector quantiser (vq_fast.cuh), tolerance-class twin of VqBuilder.
//
// Same division of labour as vq_host.h: the device splits whole frontiers, VqTreeSim replays crnlib::clusterizer<V>::generate_codebook's
// heap loop (crnlib/crn_clusterizer.h:100-139) on the results to recover the split order and the stop point, and VqResult carries the tree
// for retrieve_clusters().  What changed is the device round: one launch per node-size class (warp / CTA / thread-block cluster per node)
// instead of ~60 launches, a node table that stays in HBM (8 bytes per node up, 16 bytes back per round), no member-order emulation.
#pragma once
#include "vq_fast.cuh"
#include "vq_host.h"

namespace crn {

// The reference pops the leaf of largest variance, splits it, and counts one leaf more (even when the split fails) until max_size leaves
// exist (generate_codebook, crn_clusterizer.h:100-139).  A node's variance here is its within-node squared error, and a split never increases
// the total, so a child's key is <= its parent's: the pops are then simply ALL candidate nodes in descending key order, parents before
// children, and the popped set is the (max_size - 1) largest keys of the tree.  That turns the reference's heap loop into a threshold:
// VqOrderSim keeps a histogram of the keys of every candidate seen so far (a candidate = a node with more than one member and a positive
// variance), sends to the device every unsplit candidate whose key could still be among the (max_size - 1) largest, and at the end ranks the
// winners by (key descending, node id ascending).  A child's key is clamped to its parent's (float rounding can put it a hair above), so the
// order stays parent-first.  O(1) per node and no pointer chasing, against ~20 dependent cache misses per pop for the heap replay of
// VqTreeSim -- which matters when a selector tree has 600 K leaves.  (Ties between equal keys may be popped in a different order than the
// reference's heap would: tolerance class, like the rest of this builder.)
struct VqOrderSim {
    uint32_t root = 0, max_size = 0;
    static constexpr uint32_t kBins = 1u << 16;
    std::vector<uint32_t> hist;                         // keys of all candidates, by the top 16 bits of the float (monotonic for keys >= 0)
    struct Cand { float key; uint32_t id; };
    std::vector<Cand> pending;                          // candidates not yet split on the device
    std::vector<Cand> done;                             // candidates the device has processed (split or found unsplittable)
    uint32_t total = 0;
    static uint32_t bin_of(float v) { uint32_t b; memcpy(&b, &v, 4); return b >> 16; }
    static float bin_floor(uint32_t bin) { const uint32_t b = bin << 16; float v; memcpy(&v, &b, 4); return v; }
    void reset(const VqNodeVec& nodes)
    {
        hist.assign(kBins, 0u); pending.clear(); done.clear(); total = 0;
        add(root, nodes[root].variance, nodes[root].count, 3.0e38f, true);
    }
    void add(uint32_t id, float var, uint32_t count, float parent_key, bool is_root = false)
    {
        // the root enters unconditionally (:100-102); children only with more than one member and a positive variance (:857-871)
        if (!is_root && !(count > 1 && var > 0.0f)) return;
        const float key = var < parent_key ? var : parent_key;
        hist[bin_of(key < 0.0f ? 0.0f : key)]++;
        pending.push_back(Cand{ key, id });
        total++;
    }
    uint32_t budget() const { return max_size ? max_size - 1 : 0; }
    // candidates that may still be among the budget() largest keys and have no device result yet; they leave `pending`
    void wanted(std::vector<uint32_t>& out)
    {
        const uint32_t b = budget();
        float thresh = -1.0f;
        if (total > b) {
            uint32_t seen = 0, bin = kBins;
            while (bin > 0 && seen < b) seen += hist[--bin];
            thresh = b ? bin_floor(bin) : 3.4e38f;
        }
        size_t keep = 0;
        for (size_t i = 0; i < pending.size(); i++) {
            if (pending[i].key >= thresh) { out.push_back(pending[i].id); done.push_back(pending[i]); }
            else pending[keep++] = pending[i];
        }
        pending.resize(keep);
    }
    // `v` by key descending, then id ascending -- the order `before` in finish() gives -- as an LSD radix sort of 64-bit (inverted key bits, id)
    // words: std::sort with the comparator took 4.7 ms of a 14 ms colour endpoint tree (150 K candidates).  Digits every word agrees on are skipped.
    static void sort_cands(std::vector<Cand>& v)
    {
        const size_t n = v.size();
        static const size_t radix_min = [] { const char* e = getenv("CRN_B200_RANK_RADIX_MIN"); return e ? (size_t)atoi(e) : (size_t)4096; }();      // (tests lower it)
        if (n < 2 || n < radix_min) { std::sort(v.begin(), v.end(), [](const Cand& x, const Cand& y) { return x.key != y.key ? x.key > y.key : x.id < y.id; }); return; }
        std::vector<uint64_t> a(n), b(n);
        for (size_t i = 0; i < n; i++) {
            const float k = v[i].key == 0.0f ? 0.0f : v[i].key;                        // -0 and +0 compare equal
            uint32_t u; memcpy(&u, &k, 4);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);                            // ascending in the float's order
            a[i] = ((uint64_t)(~u) << 32) | v[i].id;
        }
        constexpr unsigned kDigit = 11, kBuckets = 1u << kDigit;
        for (unsigned shift = 0; shift < 64; shift += kDigit) {
            uint32_t cnt[kBuckets] = { 0 };
            for (size_t i = 0; i < n; i++) cnt[(a[i] >> shift) & (kBuckets - 1)]++;
            if (cnt[(a[0] >> shift) & (kBuckets - 1)] == n) continue;
            uint32_t at = 0;
            for (unsigned k = 0; k < kBuckets; k++) { const uint32_t c = cnt[k]; cnt[k] = at; at += c; }
            for (size_t i = 0; i < n; i++) b[cnt[(a[i] >> shift) & (kBuckets - 1)]++] = a[i];
            a.swap(b);
        }
        for (size_t i = 0; i < n; i++) {
            uint32_t u = ~(uint32_t)(a[i] >> 32);
            u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
            float k; memcpy(&k, &u, 4);
            v[i] = Cand{ k, (uint32_t)a[i] };
        }
    }
    // after the last round: the budget() winners (the nodes the reference would have popped).  Everything above the histogram bin that holds
    // the budget-th key wins outright; only that bin's entries need a selection.  need_ranks: also order them as the reference pops them
    // (m_codebook_index, what retrieve_clusters(max) prunes by); a caller that keeps every leaf only needs to know WHICH nodes were split.
    void finish(VqNodeVec& nodes, uint32_t& split_index, bool need_ranks)
    {
        const uint32_t b = budget();
        auto before = [](const Cand& x, const Cand& y) { return x.key != y.key ? x.key > y.key : x.id < y.id; };
        if (done.size() > b) {
            uint32_t seen = 0, bin = kBins;
            while (bin > 0 && seen < b) seen += hist[--bin];          // keys in bins above `bin` all win; `bin` is split
            const uint32_t above = seen - hist[bin];
            size_t keep = 0;
            std::vector<Cand> edge;
            for (const Cand& c : done) {
                const uint32_t cb = bin_of(c.key < 0.0f ? 0.0f : c.key);
                if (cb > bin) done[keep++] = c;
                else if (cb == bin) edge.push_back(c);
            }
            done.resize(keep);
            const size_t want = b > above ? b - above : 0;
            if (edge.size() > want) { std::nth_element(edge.begin(), edge.begin() + want, edge.end(), before); edge.resize(want); }
            done.insert(done.end(), edge.begin(), edge.end());
        }
        if (need_ranks) sort_cands(done);
        uint32_t rank = 0;
        for (const Cand& c : done) {
            VqHostNode& nd = nodes[c.id];
            if (nd.count != 1 && !nd.unsplittable && nd.left >= 0) nd.split_rank = (int32_t)rank++;      // a failed split still used up its pop
        }
        split_index = rank;
    }
};

// Host-side arrays of a build, kept by the caller between builds: a 600 K-leaf tree needs ~100 MB of them, and fresh allocations of that size
// come back from the OS page by page (first-touch faults cost more than the work done on the data).
// page-locked host array that only grows: cudaMemcpyAsync to or from pageable memory blocks the caller until the copy is done, which
// would serialise the piecewise rounds (the host could not record piece c while the device works on piece c + 1)
template <typename T> struct VqPinned {
    T* p = nullptr; size_t cap = 0;
    VqPinned() = default;
    VqPinned(const VqPinned&) = delete;
    VqPinned& operator=(const VqPinned&) = delete;
    ~VqPinned() { if (p) cudaFreeHost(p); }
    bool ensure(size_t n)
    {
        if (n <= cap) return true;
        size_t c = cap ? cap : 4096;
        while (c < n) c *= 2;
        void* q = nullptr;
        if (cudaHostAlloc(&q, c * sizeof(T), cudaHostAllocDefault) != cudaSuccess) { (void)cudaGetLastError(); return false; }
        if (p) cudaFreeHost(p);                         // contents are rebuilt every round: nothing to keep
        p = static_cast<T*>(q); cap = c;
        return true;
    }
};

struct VqFastScratch {
    VqPinned<uint2> slots;
    VqPinned<VqFastResult> results;
    VqPinned<uint32_t> all;
    std::vector<uint32_t> lists[4];                     // node-size classes: thread-block cluster, CTA, warp, single thread
    std::vector<uint32_t> slot_node, node_sim, todo;        // node_sim: which sim (partition) a node belongs to
    std::vector<VqOrderSim> sims;
    std::vector<uint32_t> parts[4];
};

template <int D> class VqFastBuilder {
public:
    VqFastBuilder(cudaStream_t stream, uint64_t* launch_counter, VqWorkspace* ws, int sm_count, VqFastScratch* scratch)
        : stream_(stream), launches_(launch_counter), ws_(ws), sm_count_(sm_count), sc_(*scratch), lists_(scratch->lists),
          sims_(scratch->sims), node_sim_(scratch->node_sim) {}
    ~VqFastBuilder() { if (events_ready_) for (cudaEvent_t e : events_) cudaEventDestroy(e); }
    VqFastBuilder(const VqFastBuilder&) = delete;
    VqFastBuilder& operator=(const VqFastBuilder&) = delete;

    // Same contract as VqBuilder<D>::build (vq_host.h): d_vecs u8[][D], d_wts u32[], d_ids ascending ids (nullptr = 0..n-1);
    // threaded = threaded_clusterizer<V>::create_clusters (three PCA divisions, then one clusterizer per non-empty partition).
    // need_ranks = false: the caller keeps every leaf (retrieve_clusters(0) / vq_leaf_offsets) and does not need the split order.
    cudaError_t build(const uint8_t* d_vecs, const uint32_t* d_wts, const uint32_t* d_ids, uint32_t n, uint32_t max_size, bool threaded, VqResult& res,
                      uint32_t* d_perm_out = nullptr, bool need_ranks = true)
    {
        res.nodes.clear(); res.trees.clear(); res.perm.clear(); res.rounds = 0; res.device_splits = 0;      // capacity kept (see VqFastScratch)
        if (!n) return cudaSuccess;
        const double t_build0 = now_ms();
        n_ = n; vecs_ = d_vecs; wts_ = d_wts;
        cudaError_t ce = allocate(n, max_size);
        if (ce != cudaSuccess) return ce;
        VqNodeVec& nodes = res.nodes;
        // root statistics (generate_codebook :76-93)
        const unsigned rg = std::max(1u, std::min<unsigned>((n + 8191) / 8192, 64u));
        CRN_LAUNCH(vq_fast_root_kernel<D>, rg, 512, 0, stream_, vecs_, wts_, d_ids, n, d_perm_, d_root_); count();
        std::vector<double> part((size_t)rg * (D + 2));
        cudaMemcpyAsync(part.data(), d_root_, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, stream_);
        ce = cudaStreamSynchronize(stream_);
        if (ce != cudaSuccess) return ce;
        double tot[D + 2];
        for (int d = 0; d < D + 2; d++) { tot[d] = 0; for (unsigned g = 0; g < rg; g++) tot[d] += part[(size_t)g * (D + 2) + d]; }
        nodes.reserve(std::min<size_t>((size_t)2 * n + 16, (size_t)4 * max_size + 64));
        nodes.resize(1);
        nodes[0] = VqHostNode();
        node_count_ = 1;
        nodes[0].begin = 0; nodes[0].count = n;
        {
            float c[D], dot = 0;
            for (int d = 0; d < D; d++) { c[d] = (float)tot[d]; dot = d ? dot + c[d] * c[d] : c[d] * c[d]; }
            const unsigned long long tw = (unsigned long long)tot[D + 1];
            nodes[0].variance = (float)(tot[D] - (double)(dot / (float)tw));
            const float inv = 1.0f / (float)tw;
            for (int d = 0; d < D; d++) c[d] *= inv;
            const uint32_t be[2] = { 0u, n };
            cudaMemcpyAsync(nodes_.begin, &be[0], 4, cudaMemcpyHostToDevice, stream_);
            cudaMemcpyAsync(nodes_.end, &be[1], 4, cudaMemcpyHostToDevice, stream_);
            cudaMemcpyAsync(nodes_.centroid, c, sizeof(c), cudaMemcpyHostToDevice, stream_);
            cudaMemcpyAsync(nodes_.weight, &tw, 8, cudaMemcpyHostToDevice, stream_);
            ce = cudaStreamSynchronize(stream_);                      // the sources are locals
            if (ce != cudaSuccess) return ce;
        }
        std::vector<uint32_t> frontier;
        if (threaded && max_size >= 128) {
            // compute_split x3 (crn_threaded_clusterizer.h:93-95)
            frontier.assign(1, 0u);
            ce = round(frontier, nodes, 1);
            if (ce != cudaSuccess) return ce;
            frontier.clear();
            const uint32_t a = (uint32_t)nodes[0].left;
            for (uint32_t c = 0; c < 2; c++) if (nodes[a + c].count) frontier.push_back(a + c);
            ce = round(frontier, nodes, 1);
            if (ce != cudaSuccess) return ce;
            std::vector<uint32_t> parts;
            for (uint32_t c = 0; c < 2; c++) {
                if (!nodes[a + c].count) continue;
                const uint32_t b = (uint32_t)nodes[a + c].left;
                for (uint32_t k = 0; k < 2; k++) if (nodes[b + k].count) parts.push_back(b + k);
            }
            const uint32_t total = (uint32_t)parts.size();
            for (uint32_t p : parts) {
                VqTreeSim t;
                t.root = p; t.max_size = (max_size + total / 2) / total;
                res.trees.push_back(t);
            }
            for (uint32_t i = 0; i < node_count_; i++) nodes[i].processed = 0;           // the divisions are not clusterizer splits
            for (uint32_t p : parts) nodes[p].left = -1;
        } else {
            VqTreeSim t;
            t.root = 0; t.max_size = max_size;
            res.trees.push_back(t);
        }
        sims_.resize(res.trees.size());
        for (size_t i = 0; i < sims_.size(); i++) { sims_[i].root = res.trees[i].root; sims_[i].max_size = res.trees[i].max_size; sims_[i].reset(nodes); }
        node_sim_.assign(node_count_, 0u);
        for (size_t i = 0; i < sims_.size(); i++) node_sim_[sims_[i].root] = (uint32_t)i;
        for (;;) {
            const double th = now_ms();
            frontier.clear();
            bool par_wanted = false;
#ifdef __CUDACC__
            size_t live = 0;
            for (const VqOrderSim& t : sims_) live += t.pending.size();
            if (sims_.size() > 1 && sims_.size() <= 4 && live > 32768) {
                std::vector<uint32_t> (&parts)[4] = sc_.parts;
                for (int i = 0; i < 4; i++) parts[i].clear();
                if (!pool_) pool_.reset(new VqPool());
                pool_->run4([&](int i) { if ((size_t)i < sims_.size()) sims_[i].wanted(parts[i]); });
                for (size_t i = 0; i < sims_.size(); i++) frontier.insert(frontier.end(), parts[i].begin(), parts[i].end());
                par_wanted = true;
            }
#endif
            if (!par_wanted) for (VqOrderSim& t : sims_) t.wanted(frontier);
            t_wanted_ += now_ms() - th;
            if (frontier.empty()) break;
            ce = round(frontier, nodes, 0);
            if (ce != cudaSuccess) return ce;
            res.rounds++;
            res.device_splits += (uint32_t)frontier.size();
        }
        {
            const double th = now_ms();
            bool par = false;
#ifdef __CUDACC__
            if (sims_.size() > 1 && sims_.size() <= 4 && node_count_ > 65536) {
                if (!pool_) pool_.reset(new VqPool());
                pool_->run4([&](int i) { if ((size_t)i < sims_.size()) sims_[i].finish(nodes, res.trees[i].split_index, need_ranks); });
                par = true;
            }
#endif
            if (!par) for (size_t i = 0; i < sims_.size(); i++) sims_[i].finish(nodes, res.trees[i].split_index, need_ranks);      // m_codebook_index of the interior nodes
            t_finish_ += now_ms() - th;
        }
        if (d_perm_out) cudaMemcpyAsync(d_perm_out, d_perm_, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, stream_);
        else {
            res.perm.resize(n);
            cudaMemcpyAsync(res.perm.data(), d_perm_, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, stream_);
        }
        ce = cudaStreamSynchronize(stream_);
        if (getenv("CRN_B200_TRACE"))
            fprintf(stderr, "[crn_b200] vq_fast<%d> n=%u max=%u: %u rounds, %u device splits, enqueue %.1f ms (of which slot prep %.1f), waiting for the device %.1f ms, recording %.1f ms, wanted %.1f ms, finish %.1f ms; whole build %.1f ms\n", D, n,
                    max_size, res.rounds, res.device_splits, t_enqueue_, t_prep_, t_sync_, t_host_, t_wanted_, t_finish_, now_ms() - t_build0);
        return ce;
    }

    const uint32_t* device_perm() const { return d_perm_; }

private:
    cudaStream_t stream_;
    uint64_t* launches_;
    VqWorkspace* ws_;
    int sm_count_;
    uint32_t n_ = 0, node_cap_ = 0, slot_cap_ = 0, node_count_ = 0;
    const uint8_t* vecs_ = nullptr;
    const uint32_t* wts_ = nullptr;
    uint32_t *d_perm_ = nullptr, *d_tmp_ = nullptr, *d_list_ = nullptr;
    uint2* d_slots_ = nullptr;
    VqFastResult* d_results_ = nullptr;
    double* d_root_ = nullptr;
    VqFastNodes nodes_ = {};
    bool wide_ok_ = true;
    VqFastScratch& sc_;
    std::vector<uint32_t> (&lists_)[4];
#ifdef __CUDACC__
    std::unique_ptr<VqPool> pool_;
#endif
    std::vector<VqOrderSim>& sims_;
    std::vector<uint32_t>& node_sim_;
    double t_enqueue_ = 0, t_sync_ = 0, t_host_ = 0, t_wanted_ = 0, t_finish_ = 0, t_prep_ = 0;
    cudaEvent_t events_[4] = {};
    bool events_ready_ = false;
    static constexpr uint32_t kHugeNode = 8192, kLargeNode = 1024;
    static constexpr int kClusterCtas = 8, kWideClusterCtas = 16, kClusterThreads = 512;
    static double now_ms() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    void count() { if (launches_) ++*launches_; }

    // The node table: two children per device split.  A frontier may hold nodes the replay never pops, so the number of device splits of a
    // build is not known in advance: the table starts at 4 x max_size nodes and doubles when a round would overflow it (contents copied).
    static constexpr size_t node_bytes() { return 4 + 4 + 4 * (size_t)D + 8 + 1; }   // + slack for the 256-byte alignment of the four arrays at small capacities
    cudaError_t ensure_nodes(size_t need, size_t live)
    {
        if (need <= node_cap_) return cudaSuccess;
        size_t cap = std::max<size_t>(node_cap_, 1024);
        while (cap < need) cap *= 2;
        const size_t bytes = cap * node_bytes() + 1024;
        void* buf = ws_->nodes;
        const bool fresh = live || !ws_->nodes || ws_->nodes_cap < bytes;
        if (fresh) {
            const cudaError_t ce = cudaMalloc(&buf, bytes);
            if (ce != cudaSuccess) return ce;
        }
        VqFastNodes nn;
        uint8_t* b = static_cast<uint8_t*>(buf);
        nn.weight = reinterpret_cast<unsigned long long*>(b); b += (cap * 8 + 255) & ~(size_t)255;
        nn.begin = reinterpret_cast<uint32_t*>(b); b += (cap * 4 + 255) & ~(size_t)255;
        nn.end = reinterpret_cast<uint32_t*>(b); b += (cap * 4 + 255) & ~(size_t)255;
        nn.centroid = reinterpret_cast<float*>(b);
        if (live) {
            cudaMemcpyAsync(nn.weight, nodes_.weight, live * 8, cudaMemcpyDeviceToDevice, stream_);
            cudaMemcpyAsync(nn.begin, nodes_.begin, live * 4, cudaMemcpyDeviceToDevice, stream_);
            cudaMemcpyAsync(nn.end, nodes_.end, live * 4, cudaMemcpyDeviceToDevice, stream_);
            cudaMemcpyAsync(nn.centroid, nodes_.centroid, live * 4 * D, cudaMemcpyDeviceToDevice, stream_);
            const cudaError_t ce = cudaStreamSynchronize(stream_);
            if (ce != cudaSuccess) { cudaFree(buf); return ce; }
        }
        if (fresh) { if (ws_->nodes) cudaFree(ws_->nodes); ws_->nodes = buf; ws_->nodes_cap = bytes; }
        nodes_ = nn; node_cap_ = (uint32_t)cap;
        return cudaSuccess;
    }

    // one slab the caller keeps between builds
    cudaError_t allocate(uint32_t n, uint32_t max_size)
    {
        const size_t slots = std::max<size_t>(8, std::min<size_t>((size_t)n / 2 + 8, (size_t)max_size + 8));
        node_cap_ = 0;
        cudaError_t ne = ensure_nodes(std::min<size_t>((size_t)2 * n + 16, (size_t)4 * max_size + 64), 0);
        if (ne != cudaSuccess) return ne;
        for (int pass = 0; pass < 2; pass++) {
            size_t off = 0;
            uint8_t* base = pass ? static_cast<uint8_t*>(ws_->base) : nullptr;
            auto carve = [&](auto*& p, size_t cnt) {
                using T = std::remove_pointer_t<std::remove_reference_t<decltype(p)>>;
                if (pass) p = reinterpret_cast<T*>(base + off);
                off += (cnt * sizeof(T) + 255) & ~(size_t)255;
            };
            carve(d_perm_, n); carve(d_tmp_, n); carve(d_list_, slots); carve(d_slots_, slots); carve(d_results_, slots); carve(d_root_, (size_t)64 * (D + 2));
            if (!pass && off > ws_->cap) {
                if (ws_->base) cudaFree(ws_->base);
                ws_->base = nullptr; ws_->cap = 0;
                const cudaError_t ce = cudaMalloc(&ws_->base, off);
                if (ce != cudaSuccess) return ce;
                ws_->cap = off;
            }
        }
        slot_cap_ = (uint32_t)slots;
        return cudaSuccess;
    }

#ifdef __CUDACC__
    template <int G> bool launch_cluster(const uint32_t* dl, uint32_t cnt, int mode)
    {
        auto kernel = vq_fast_split_kernel<D, kClusterThreads, G>;
        if (G > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return false;
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        const unsigned nclusters = (unsigned)std::min<size_t>(cnt, (size_t)std::max(1, sm_count_ / G) * 2);
        cfg.gridDim = dim3(nclusters * G); cfg.blockDim = dim3(kClusterThreads); cfg.stream = stream_;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kernel, vecs_, wts_, d_perm_, d_tmp_, nodes_, (const uint2*)d_slots_, dl, cnt, d_results_, mode) == cudaSuccess;
    }
#endif

    // Split every node of `frontier` on the device and record the results; mode 1 = threaded_clusterizer's PCA division.
    // Large frontiers go in kChunks pieces, each with its own launches and its own copy back, so that the host records piece c (the node
    // table, the order simulations) while the device is splitting piece c + 1.  Pieces are cut across the partitions: piece c holds the c-th
    // quarter of EVERY partition's nodes, so recording a piece still spreads over the partitions' host threads.
    static constexpr uint32_t kChunks = 4, kPipelineMin = 16384;
    cudaError_t round(const std::vector<uint32_t>& frontier, VqNodeVec& nodes, int mode)
    {
        const double t0 = now_ms();
        std::vector<uint32_t>& slot_node = sc_.slot_node;
        slot_node.clear();
        std::vector<uint32_t>& todo = sc_.todo;
        todo.clear();
        for (uint32_t id : frontier) {
            VqHostNode& nd = nodes[id];
            if (nd.count < 2 && mode == 0) { nd.processed = 1; nd.unsplittable = 1; continue; }
            todo.push_back(id);
        }
        const uint32_t F = (uint32_t)todo.size();
        if (!F) return cudaSuccess;
        if (F > slot_cap_) return cudaErrorMemoryAllocation;
        if (!sc_.slots.ensure(F) || !sc_.results.ensure(F) || !sc_.all.ensure(F)) return cudaErrorMemoryAllocation;
        uint2* const h_slots = sc_.slots.p;
        VqFastResult* const h_results = sc_.results.p;
        uint32_t* const all = sc_.all.p;
        uint32_t nslots = 0, nall = 0;
        // partitions: contiguous runs of one sim in the frontier (wanted() is called sim by sim)
        uint32_t pb[5] = { 0, F, F, F, F };
        uint32_t np = 1;
        if (mode == 0 && sims_.size() > 1 && sims_.size() <= 4) {
            for (uint32_t s = 1; s < F && np < 4; s++) if (node_sim_[todo[s]] != node_sim_[todo[s - 1]]) pb[np++] = s;
            for (uint32_t k = np; k <= 4; k++) pb[k] = F;
        }
        uint32_t nchunks = 1;
        {
            const char* pm = getenv("CRN_B200_VQ_PIPELINE_MIN");          // tests lower it so that small inputs take the piecewise path too
            const uint32_t min_f = pm ? (uint32_t)atoi(pm) : kPipelineMin;
            if (mode == 0 && F >= min_f && F >= 2 * kChunks && !getenv("CRN_B200_VQ_NO_PIPELINE")) nchunks = kChunks;
        }
        // slot order: piece-major, partition inside; sub[c][p] .. sub[c][p + 1] = the slots of partition p in piece c
        uint32_t sub[kChunks][5];
        uint32_t next_child = node_count_;
        for (uint32_t c = 0; c < nchunks; c++)
            for (uint32_t p = 0; p < 4; p++) {
                sub[c][p] = nslots;
                if (p < np) {
                    const uint32_t len = pb[p + 1] - pb[p];
                    const uint32_t a = pb[p] + (uint32_t)((uint64_t)len * c / nchunks), b = pb[p] + (uint32_t)((uint64_t)len * (c + 1) / nchunks);
                    for (uint32_t k = a; k < b; k++) { h_slots[nslots++] = make_uint2(todo[k], next_child); slot_node.push_back(todo[k]); next_child += 2; }
                }
                sub[c][4] = nslots;
            }
        // size classes per piece, concatenated: [piece 0: cluster | cta | warp | thread][piece 1: ...]
        uint32_t cls_cnt[kChunks][4];
        for (uint32_t c = 0; c < nchunks; c++) {
            for (int k = 0; k < 4; k++) lists_[k].clear();
            for (uint32_t s = sub[c][0]; s < sub[c][4]; s++) {
                const uint32_t cnt = nodes[slot_node[s]].count;
                const bool tiny = mode == 0 && cnt <= kVqTinyNode;
#ifdef __CUDACC__
                lists_[tiny ? 3 : (cnt >= kHugeNode ? 0 : (cnt >= kLargeNode ? 1 : 2))].push_back(s);
#else
                lists_[tiny ? 3 : (cnt >= kLargeNode ? 1 : 2)].push_back(s);      // the emulator has no thread-block clusters
#endif
            }
            for (int k = 0; k < 4; k++) {
                cls_cnt[c][k] = (uint32_t)lists_[k].size();
                if (cls_cnt[c][k]) memcpy(all + nall, lists_[k].data(), sizeof(uint32_t) * cls_cnt[c][k]);
                nall += cls_cnt[c][k];
            }
        }
        const double t_prep = now_ms() - t0; t_prep_ += t_prep;
        if (next_child > node_cap_) { const cudaError_t ge = ensure_nodes(next_child, node_count_); if (ge != cudaSuccess) return ge; }
        cudaMemcpyAsync(d_slots_, h_slots, sizeof(uint2) * F, cudaMemcpyHostToDevice, stream_);
        cudaMemcpyAsync(d_list_, all, sizeof(uint32_t) * F, cudaMemcpyHostToDevice, stream_);
        if (nchunks > 1 && !events_ready_) {
            for (uint32_t c = 0; c < kChunks; c++) if (cudaEventCreateWithFlags(&events_[c], crn::event_flags()) != cudaSuccess) return cudaGetLastError();
            events_ready_ = true;
        }
        const uint32_t* dl = d_list_;
        size_t n_cluster = 0, n_cta = 0, n_warp = 0, n_thread = 0;
        for (uint32_t c = 0; c < nchunks; c++) {
#ifdef __CUDACC__
            if (cls_cnt[c][0]) {
                const uint32_t cnt = cls_cnt[c][0];
                bool launched = false;
                if (cnt <= 8 && wide_ok_) {
                    launched = launch_cluster<kWideClusterCtas>(dl, cnt, mode);
                    if (!launched) { wide_ok_ = false; (void)cudaGetLastError(); }
                }
                if (!launched && !launch_cluster<kClusterCtas>(dl, cnt, mode)) return cudaGetLastError();
                count();
                dl += cnt; n_cluster += cnt;
            }
#endif
            if (cls_cnt[c][1]) {
                const uint32_t cnt = cls_cnt[c][1];
                CRN_LAUNCH((vq_fast_split_kernel<D, 256, 1>), cnt, 256, 0, stream_, vecs_, wts_, d_perm_, d_tmp_, nodes_, (const uint2*)d_slots_, dl, cnt, d_results_, mode); count();
                dl += cnt; n_cta += cnt;
            }
            if (cls_cnt[c][2]) {
                const uint32_t cnt = cls_cnt[c][2];
                const unsigned grid = (unsigned)std::min<size_t>(cnt, (size_t)sm_count_ * 32);
                CRN_LAUNCH((vq_fast_split_kernel<D, 32, 1>), grid, 32, 0, stream_, vecs_, wts_, d_perm_, d_tmp_, nodes_, (const uint2*)d_slots_, dl, cnt, d_results_, mode); count();
                dl += cnt; n_warp += cnt;
            }
            if (cls_cnt[c][3]) {
                const uint32_t cnt = cls_cnt[c][3];
                CRN_LAUNCH(vq_fast_tiny_kernel<D>, (cnt + 127) / 128, 128, 0, stream_, vecs_, wts_, d_perm_, nodes_, (const uint2*)d_slots_, dl, cnt, d_results_); count();
                dl += cnt; n_thread += cnt;
            }
            const uint32_t s0 = sub[c][0], s1 = sub[c][4];
            if (s1 > s0) cudaMemcpyAsync(h_results + s0, d_results_ + s0, sizeof(VqFastResult) * (s1 - s0), cudaMemcpyDeviceToHost, stream_);
            if (nchunks > 1) cudaEventRecord(events_[c], stream_);
        }
        const double t1 = now_ms();
        nodes.resize(next_child);                                // (no construction: VqDefaultInitAlloc; the records are assigned below)
        node_count_ = next_child;
        if (mode == 0) node_sim_.resize(next_child, 0u);
        auto scatter = [&](uint32_t s0, uint32_t s1) {
            for (uint32_t s = s0; s < s1; s++) {
                const VqFastResult& r = h_results[s];
                VqHostNode& par = nodes[slot_node[s]];
                par.processed = 1;
                if (r.state != 1) { par.unsplittable = 1; continue; }
                const uint32_t child = h_slots[s].y;
                par.left = (int32_t)child;
                VqHostNode& l = nodes[child];
                VqHostNode& rr = nodes[child + 1];
                l = VqHostNode(); rr = VqHostNode();                 // (the table is reused between builds: stale records)
                l.begin = par.begin; l.count = r.n_left; l.variance = r.lvar;
                rr.begin = par.begin + r.n_left; rr.count = par.count - r.n_left; rr.variance = r.rvar;
                par.child_count[0] = l.count; par.child_count[1] = rr.count;
                par.child_var[0] = r.lvar; par.child_var[1] = r.rvar;
                if (mode == 0) {
                    const uint32_t si = node_sim_[slot_node[s]];
                    node_sim_[child] = si; node_sim_[child + 1] = si;
                    const float pk = par.variance;
                    l.variance = r.lvar < pk ? r.lvar : pk; rr.variance = r.rvar < pk ? r.rvar : pk;      // keys never exceed the parent's (see VqOrderSim)
                    sims_[si].add(child, l.variance, l.count, pk);
                    sims_[si].add(child + 1, rr.variance, rr.count, pk);
                }
            }
        };
        double t_wait = 0, t_rec = 0;
        for (uint32_t c = 0; c < nchunks; c++) {
            const double ta = now_ms();
            cudaError_t ce = nchunks > 1 ? cudaEventSynchronize(events_[c]) : cudaStreamSynchronize(stream_);
            if (ce != cudaSuccess) return ce;
            const double tb = now_ms();
            // every record written belongs to one partition only, so each partition's share of the piece goes to its own host thread
            bool done_parallel = false;
#ifdef __CUDACC__
            if (np > 1 && sub[c][4] - sub[c][0] > 4096) {
                if (!pool_) pool_.reset(new VqPool());
                pool_->run4([&](int i) { if (sub[c][i] < sub[c][i + 1]) scatter(sub[c][i], sub[c][i + 1]); });
                done_parallel = true;
            }
#endif
            if (!done_parallel) scatter(sub[c][0], sub[c][4]);
            t_wait += tb - ta; t_rec += now_ms() - tb;
        }
        cudaError_t ce = cudaGetLastError();
        if (ce != cudaSuccess) return ce;
        t_enqueue_ += t1 - t0; t_sync_ += t_wait; t_host_ += t_rec;
        if (getenv("CRN_B200_TRACE_ROUNDS"))
            fprintf(stderr, "[crn_b200]   vq_fast<%d> round F=%u in %u piece(s) (cluster %zu, cta %zu, warp %zu, thread %zu): enqueue %.2f ms, waiting for the device %.2f ms, recording %.2f ms\n",
                    D, F, nchunks, n_cluster, n_cta, n_warp, n_thread, t1 - t0, t_wait, t_rec);
        return cudaSuccess;
    }
};

}  // namespace crn
