// vq_host.h -- host driver of the frontier-batched vector quantiser (vq_kernels.cuh).
//
// The device splits whole frontiers; this file decides WHICH leaves are worth splitting and replays the
// reference's split order.  crnlib::clusterizer<V>::generate_codebook (crnlib/crn_clusterizer.h:100-139)
// keeps the leaves in a binary max-heap keyed by variance, pops the worst one, splits it, counts
// total_leaves up (even when the node turned out unsplittable) and stops at max_size.  VqTreeSim below is
// that loop, fed with the split results the device produced one round earlier; it stalls when the heap top
// has not been split on the device yet, which is what triggers the next round.  A leaf outside the
// `max_size - total_leaves` highest variances of the heap can never be popped, so it is not sent to the
// device.  retrieve_clusters() (:301-332) is the pruning walk over the recorded split ranks.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <algorithm>
#ifdef __CUDACC__
#include <atomic>
#include <memory>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#endif
#include <type_traits>
#include <vector>
#include "vq_kernels.cuh"

namespace crn {

struct VqHostNode {
    uint32_t begin = 0, count = 0;
    int32_t left = -1;               // device child id (right = left + 1), -1 while not split on the device
    float variance = 0;
    int32_t split_rank = -1;         // m_codebook_index of an interior node: order in which the reference split it
    uint8_t processed = 0, unsplittable = 0;
    // copies of the children's count / variance: the heap replay decides about both children from the parent's record
    // alone (one cache line per pop instead of three; the node table of a large build is tens of MB)
    uint32_t child_count[2] = {0, 0};
    float child_var[2] = {0, 0};
};

// Node tables grow to a million records per build; a plain vector would write every new record twice (its constructor, then the split
// result).  With this allocator resize() leaves new records uninitialised: whoever appends nodes assigns VqHostNode() or every field.
template <typename T> struct VqDefaultInitAlloc : std::allocator<T> {
    template <typename U> struct rebind { using other = VqDefaultInitAlloc<U>; };
    VqDefaultInitAlloc() = default;
    template <typename U> VqDefaultInitAlloc(const VqDefaultInitAlloc<U>&) {}
    template <typename U> void construct(U*) noexcept {}
    template <typename U, typename A0, typename... A> void construct(U* p, A0&& a0, A&&... a) { ::new ((void*)p) U(std::forward<A0>(a0), std::forward<A>(a)...); }
};
typedef std::vector<VqHostNode, VqDefaultInitAlloc<VqHostNode>> VqNodeVec;

struct VqHeapEntry { float variance; uint32_t id; };   // the key travels with the id: sift loops stay inside one array

struct VqTreeSim {                   // one clusterizer<V> instance
    uint32_t root = 0, max_size = 0, total_leaves = 1, split_index = 0;
    std::vector<VqHeapEntry> heap;   // 1-based, as the reference's
    uint32_t heap_size = 0;
    // Bookkeeping for wanted(): a histogram of the keys in the heap (8192 bins over the float's exponent + 4 mantissa
    // bits; variances are >= 0 so the bit pattern is monotonic) and the heap members not yet sent to the device.
    static constexpr uint32_t kBins = 8192;
    // Nodes split by the replay, in split order (rank = index).  Written to VqHostNode::split_rank once the build is over:
    // the node table is shared by the trees' threads and its records are interleaved, so writing ranks during the replay
    // would bounce cache lines between cores.
    std::vector<uint32_t> split_log;
    std::vector<uint32_t> hist;
    std::vector<VqHeapEntry> pending;
    static uint32_t bin_of(float v) { uint32_t b; memcpy(&b, &v, 4); return (b >> 19) & (kBins - 1); }
    static float bin_floor(uint32_t bin) { const uint32_t b = bin << 19; float v; memcpy(&v, &b, 4); return v; }

    void reset(const VqNodeVec& nodes)
    {
        heap.assign((size_t)max_size + 2, VqHeapEntry{0.0f, 0u});
        hist.assign(kBins, 0u);
        pending.clear();
        heap_size = 0;
        insert(root, nodes[root].variance);                    // the root enters the heap unconditionally (:100-102)
    }
    void insert(uint32_t id, float v)
    {   // insert_heap (:384-414): sift up while the parent is not strictly greater
        uint32_t pos = ++heap_size;
        if (heap_size >= heap.size()) heap.resize(heap_size + 1);
        for (;;) {
            const uint32_t parent = pos >> 1;
            if (!parent || heap[parent].variance > v) break;
            heap[pos] = heap[parent];
            pos = parent;
        }
        heap[pos] = VqHeapEntry{v, id};
        hist[bin_of(v)]++;
        pending.push_back(VqHeapEntry{v, id});
    }
    uint32_t pop()
    {   // generate_codebook :114-121 + down_heap (:416-444)
        const uint32_t top = heap[1].id;
        hist[bin_of(heap[1].variance)]--;
        heap[1] = heap[heap_size--];
        if (heap_size) {
            uint32_t pos = 1, child;
            const VqHeapEntry orig = heap[1];
            while ((child = pos << 1) <= heap_size) {
                if (child < heap_size && heap[child].variance < heap[child + 1].variance) child++;
                if (orig.variance > heap[child].variance) break;
                heap[pos] = heap[child];
                pos = child;
            }
            heap[pos] = orig;
        }
        return top;
    }
    bool finished() const { return !(total_leaves < max_size && heap_size); }
    // advance as far as the device results allow
    void run(VqNodeVec& nodes)
    {
        while (!finished()) {
            VqHostNode& nd = nodes[heap[1].id];
            if (nd.count != 1 && !nd.processed) return;
            // the next top is one of the root's children: have their records on the way
            if (heap_size >= 3) { __builtin_prefetch(&nodes[heap[2].id]); __builtin_prefetch(&nodes[heap[3].id]); }
            const uint32_t id = pop();
            const VqHostNode& node = nodes[id];
            if (node.count != 1 && !node.unsplittable) {          // split_node (:740-873)
                split_log.push_back(id);
                split_index++;
                for (uint32_t c = 0; c < 2; c++)
                    if (node.child_count[c] > 1 && node.child_var[c] > 0.0f) insert((uint32_t)node.left + c, node.child_var[c]);
            }
            total_leaves++;
        }
    }
    // Leaves of the heap that can still be popped and have not been split on the device: at most `budget` more pops
    // can happen, so a key below the budget-th largest of the heap is out of reach.  The cut is taken at the lower
    // edge of the histogram bin holding that key -- a few more nodes than strictly needed go to the device, never fewer.
    // Every node returned is split by the round that follows (VqBuilder::round is synchronous), so it leaves `pending`.
    void wanted(std::vector<uint32_t>& out)
    {
        if (finished()) { pending.clear(); return; }
        const uint32_t budget = max_size - total_leaves;
        float thresh = -1.0f;
        if (heap_size > budget) {
            uint32_t seen = 0, bin = kBins;
            while (bin > 0 && seen < budget) seen += hist[--bin];
            thresh = bin_floor(bin);
        }
        size_t keep = 0;
        for (size_t i = 0; i < pending.size(); i++) {
            if (pending[i].variance >= thresh) out.push_back(pending[i].id);
            else pending[keep++] = pending[i];
        }
        pending.resize(keep);
    }
};

struct VqResult {                    // host-side outcome of one build
    VqNodeVec nodes;
    std::vector<VqTreeSim> trees;
    std::vector<uint32_t> perm;      // vector indices, grouped by node range
    uint32_t rounds = 0, device_splits = 0;

    uint32_t codebook_size() const
    {
        uint32_t s = 0;
        for (const VqTreeSim& t : trees) s += 1 + t.split_index;
        return s;
    }
    // retrieve_clusters(max_clusters) over every tree, in the reference's order; cluster_of[vector] = cluster index.
    // max_clusters == 0 keeps every leaf.  Members of a cluster are, as in the reference, in ascending vector order
    // when they are listed by scanning cluster_of.
    uint32_t retrieve(uint32_t max_clusters, uint32_t* cluster_of) const
    {
        uint32_t nclusters = 0;
        std::vector<uint32_t> stack;
        for (const VqTreeSim& t : trees) {
            stack.clear();
            uint32_t cur = t.root;
            for (;;) {
                const VqHostNode& nd = nodes[cur];
                const bool leaf = nd.split_rank < 0;
                if (leaf || (max_clusters && (uint32_t)nd.split_rank + 2 > max_clusters)) {
                    for (uint32_t i = nd.begin; i < nd.begin + nd.count; i++) cluster_of[perm[i]] = nclusters;
                    nclusters++;
                    if (stack.empty()) break;
                    cur = stack.back();
                    stack.pop_back();
                    continue;
                }
                stack.push_back((uint32_t)nd.left + 1);
                cur = (uint32_t)nd.left;
            }
        }
        return nclusters;
    }
};

// Every node covers a contiguous range of the final permutation and the splits are stable partitions of an ascending id
// list, so with every leaf kept (retrieve_clusters(0)) cluster k is simply the k-th leaf in position order and its members
// are perm[begin .. begin + count) -- already ascending, already on the device.  Appends the leaves' first positions
// (+ base) to `offsets` (which ends with the running total) in the reference's retrieval order; returns the leaf count.
// The same for retrieve_clusters(max_clusters): a node whose split came too late for the budget is kept whole -- still one range.
// `offsets` must hold its leading 0; returns the number of clusters (numbered like VqResult::retrieve).
inline uint32_t vq_range_offsets(const VqResult& res, uint32_t max_clusters, std::vector<uint32_t>& offsets)
{
    uint32_t nclusters = 0, total = 0;
    std::vector<uint32_t> stack;
    offsets.pop_back();
    for (const VqTreeSim& t : res.trees) {
        stack.clear();
        uint32_t cur = t.root;
        for (;;) {
            const VqHostNode& nd = res.nodes[cur];
            const bool leaf = nd.split_rank < 0;
            if (leaf || (max_clusters && (uint32_t)nd.split_rank + 2 > max_clusters)) {
                offsets.push_back(nd.begin);
                total = nd.begin + nd.count;
                nclusters++;
                if (stack.empty()) break;
                cur = stack.back();
                stack.pop_back();
                continue;
            }
            stack.push_back((uint32_t)nd.left + 1);
            cur = (uint32_t)nd.left;
        }
    }
    offsets.push_back(total);
    return nclusters;
}

inline uint32_t vq_leaf_offsets(const VqResult& res, uint32_t base, std::vector<uint32_t>& offsets)
{
    // one depth-first walk per tree (threaded_clusterizer: four); the trees cover disjoint, ascending position ranges, so they are walked
    // on their own host threads and their leaf lists concatenated
    const size_t nt = res.trees.size();
    std::vector<std::vector<uint32_t>> part(nt);
    std::vector<uint32_t> last(nt, 0u);
    auto walk = [&](size_t ti) {
        const VqTreeSim& t = res.trees[ti];
        std::vector<uint32_t>& out = part[ti];
        std::vector<uint32_t> stack;
        uint32_t cur = t.root;
        for (;;) {
            const VqHostNode& nd = res.nodes[cur];
            if (nd.split_rank < 0) {
                out.push_back(base + nd.begin);
                last[ti] = nd.begin + nd.count;
                if (stack.empty()) break;
                cur = stack.back();
                stack.pop_back();
                continue;
            }
            stack.push_back((uint32_t)nd.left + 1);
            cur = (uint32_t)nd.left;
        }
    };
#ifdef __CUDACC__
    if (nt > 1 && res.nodes.size() > 65536) {
        std::vector<std::thread> th;
        for (size_t ti = 1; ti < nt; ti++) th.emplace_back(walk, ti);
        walk(0);
        for (std::thread& x : th) x.join();
    } else
#endif
    for (size_t ti = 0; ti < nt; ti++) walk(ti);
    uint32_t leaves = 0, total = 0;
    offsets.pop_back();                                            // the running total comes back at the end
    for (size_t ti = 0; ti < nt; ti++) {
        offsets.insert(offsets.end(), part[ti].begin(), part[ti].end());
        leaves += (uint32_t)part[ti].size();
        if (!part[ti].empty()) total = last[ti];
    }
    offsets.push_back(base + total);
    return leaves;
}

struct VqWorkspace {                 // device slab reused across builds (+ the growable node table of the single-launch builder, vq_fast_host.h)
    void* base = nullptr; size_t cap = 0;
    void* nodes = nullptr; size_t nodes_cap = 0;
};

#ifdef __CUDACC__
// Three helper threads that live as long as one build: the host replay and the result scatter of a round are split four
// ways (spawning threads per round cost more than the work they did).
class VqPool {
public:
    // How long an idle worker spins before it parks on the condition variable.  With cores to spare (>= 8 hardware threads per visible GPU)
    // 4 ms: rounds follow each other within a millisecond or two and a parked thread takes tens of microseconds to get going.  With ranks
    // sharing a few cores (8 GPUs on a 32-core host) the spinners compete with the threads doing the work: 0.05 ms.  CRN_B200_POOL_SPIN_MS overrides.
    static double spin_ms()
    {
        static const double v = [] {
            if (const char* e = getenv("CRN_B200_POOL_SPIN_MS")) { if (*e) return atof(e); }
            int ndev = 1;
            if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) ndev = 1;
            const unsigned hc = std::thread::hardware_concurrency();
            return (hc && hc / (unsigned)ndev < 8u) ? 0.05 : 4.0;
        }();
        return v;
    }
    VqPool() { for (int i = 0; i < 3; i++) workers_[i] = std::thread([this, i]() { loop(i + 1); }); }
    ~VqPool()
    {
        stop_.store(true);
        gen_.fetch_add(1);
        { std::lock_guard<std::mutex> l(m_); }
        cv_.notify_all();
        for (std::thread& t : workers_) t.join();
    }
    // fn(part) for part = 0..3, part 0 on the calling thread; returns when all four are done
    void run4(const std::function<void(int)>& fn)
    {
        fn_ = &fn;
        pending_.store(3);
        gen_.fetch_add(1);
        { std::lock_guard<std::mutex> l(m_); }
        cv_.notify_all();
        fn(0);
        while (pending_.load(std::memory_order_acquire) != 0) std::this_thread::yield();
    }
private:
    // Rounds follow each other within a millisecond or two, and a sleeping thread takes about that long to get going
    // again, so a worker first spins on the generation counter and only then parks on the condition variable.
    void loop(int part)
    {
        unsigned seen = 0;
        for (;;) {
            const double t0 = now();
            while (gen_.load(std::memory_order_acquire) == seen) {
                std::this_thread::yield();                  // several ranks may share a few cores: a spinning worker must not keep one to itself
                if (now() - t0 > spin_ms()) {
                    std::unique_lock<std::mutex> l(m_);
                    cv_.wait(l, [&]() { return gen_.load() != seen; });
                    break;
                }
            }
            seen = gen_.load();
            if (stop_.load()) return;
            (*fn_)(part);
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    static double now() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    std::thread workers_[3];
    std::mutex m_;
    std::condition_variable cv_;
    const std::function<void(int)>* fn_ = nullptr;
    std::atomic<unsigned> gen_{0};
    std::atomic<int> pending_{0};
    std::atomic<bool> stop_{false};
};
#endif

template <int D> class VqBuilder {
public:
    VqBuilder(cudaStream_t stream, uint64_t* launch_counter, VqWorkspace* ws) : stream_(stream), launches_(launch_counter), ws_(ws) {}

    // d_vecs: u8[][D], d_wts: u32[] (device), indexed by vector id.  d_ids: the n ids to quantise in ascending order
    // (nullptr = 0..n-1).  threaded: crnlib::threaded_clusterizer<V>::create_clusters (crn_threaded_clusterizer.h:70-174):
    // three PCA divisions into <= 4 partitions, each its own clusterizer.
    // d_perm_out (optional, device, n entries): receives the final permutation; the host copy res.perm is then skipped.
    cudaError_t build(const uint8_t* d_vecs, const uint32_t* d_wts, const uint32_t* d_ids, uint32_t n, uint32_t max_size, bool threaded, VqResult& res,
                      uint32_t* d_perm_out = nullptr)
    {
        res = VqResult();
        if (!n) return cudaSuccess;
        cudaError_t ce = allocate(n, max_size);
        if (ce != cudaSuccess) return ce;
        n_ = n; vecs_ = d_vecs; wts_ = d_wts;
        if (kSmemCov > 48 * 1024) cudaFuncSetAttribute(vq_stream_kernel<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCov);
        VqNodeVec& nodes = res.nodes;

        // root: identity order + statistics
        launch_fill_identity(d_ids, n);
        cudaMemsetAsync(d_acc_, 0, sizeof(unsigned long long) * (D + 2), stream_);
        CRN_LAUNCH(vq_root_kernel<D>, grid(n), 256, 0, stream_, vecs_, wts_, d_perm_[0], n, d_acc_); count();
        CRN_LAUNCH(vq_root_prepare_kernel<D>, 1, 32, 0, stream_, d_acc_, d_slots_, n); count();
        cudaMemsetAsync(d_side_, 0, n, stream_);
        cudaMemsetAsync(d_big_count_, 0, sizeof(unsigned) * kBigLists, stream_);
        float_sums(d_perm_[0], 1, 2, 0);
        CRN_LAUNCH(vq_root_finish_kernel<D>, 1, 32, 0, stream_, d_acc_, d_slots_, nodes_, n, (threaded && max_size >= 128) ? 1 : 0); count();
        const unsigned first_free_node = 1;
        cudaMemcpyAsync(d_node_counter_, &first_free_node, sizeof(unsigned), cudaMemcpyHostToDevice, stream_);
        float root_var = 0;
        cudaMemcpyAsync(&root_var, nodes_.variance, sizeof(float), cudaMemcpyDeviceToHost, stream_);
        ce = cudaStreamSynchronize(stream_);
        if (ce != cudaSuccess) return ce;
        nodes.reserve(std::min<size_t>((size_t)2 * n + 16, (size_t)4 * max_size + 64));
        nodes.resize(1);
        nodes[0] = VqHostNode();
        nodes[0].begin = 0; nodes[0].count = n; nodes[0].variance = root_var;

        std::vector<uint32_t> frontier;
        if (threaded && max_size >= 128) {
            // compute_split x3 (:93-95)
            frontier.assign(1, 0u);
            ce = round(frontier, nodes, 1);
            if (ce != cudaSuccess) return ce;
            frontier.clear();
            const uint32_t a = (uint32_t)nodes[0].left;
            for (uint32_t c = 0; c < 2; c++) if (nodes[a + c].count) frontier.push_back(a + c);
            ce = round(frontier, nodes, 2);
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
            for (VqHostNode& nd : nodes) nd.processed = 0;      // presplit results are not clusterizer splits
            for (uint32_t p : parts) nodes[p].left = -1;
        } else {
            VqTreeSim t;
            t.root = 0; t.max_size = max_size;
            res.trees.push_back(t);
        }
        for (VqTreeSim& t : res.trees) t.reset(nodes);

        for (;;) {
            const double th = now_ms();
            frontier.clear();
            advance_trees(res.trees, nodes, frontier);
            if (frontier.empty()) break;
            t_heap_ += now_ms() - th;
            ce = round(frontier, nodes, 0);
            if (ce != cudaSuccess) return ce;
            res.rounds++;
            res.device_splits += (uint32_t)frontier.size();
        }
        for (VqTreeSim& t : res.trees) {                          // m_codebook_index of the interior nodes
            for (size_t k = 0; k < t.split_log.size(); k++) nodes[t.split_log[k]].split_rank = (int32_t)k;
            std::vector<uint32_t>().swap(t.split_log);
        }
        if (d_perm_out) cudaMemcpyAsync(d_perm_out, d_perm_[cur_], sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, stream_);
        else {
            res.perm.resize(n);
            cudaMemcpyAsync(res.perm.data(), d_perm_[cur_], sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, stream_);
        }
        ce = cudaStreamSynchronize(stream_);
        if (getenv("CRN_B200_TRACE"))
            fprintf(stderr, "[crn_b200] vq<%d> n=%u max=%u: %u rounds, host enqueue %.1f ms, sync wait %.1f ms, host results %.1f ms, host heap %.1f ms\n", D, n, max_size,
                    res.rounds, t_enqueue_, t_sync_, t_results_, t_heap_);
        if (getenv("CRN_B200_TRACE"))
            for (int i = 0; i < 4; i++) fprintf(stderr, "[crn_b200]   tree %d: run %.1f ms, wanted %.1f ms, sort %.1f ms\n", i, tm_[i][0], tm_[i][1], tm_[i][2]);
        return ce;
    }

    const uint32_t* device_perm() const { return d_perm_[cur_]; }

private:
    cudaStream_t stream_;
    uint64_t* launches_;
    VqWorkspace* ws_;
    uint32_t n_ = 0, cap_n_ = 0, cap_slots_ = 0;
    const uint8_t* vecs_ = nullptr;
    const uint32_t* wts_ = nullptr;
    int cur_ = 0;
    unsigned* d_perm_[2] = {nullptr, nullptr};
    unsigned *d_pos_slot_ = nullptr, *d_flags_ = nullptr, *d_scan_ = nullptr, *d_block_sums_ = nullptr, *d_slot_node_ = nullptr, *d_slot_starts_ = nullptr;
    unsigned *d_node_counter_ = nullptr, *d_active_ = nullptr, *d_big_count_ = nullptr, *d_big_list_ = nullptr;
    unsigned *d_chunk_start_ = nullptr, *d_overflow_count_ = nullptr, *d_overflow_list_ = nullptr, *d_table_ = nullptr, max_chunks_ = 0;
    VqBigDir* d_dir_ = nullptr;
    static constexpr size_t kSmemSum = VqStreamCfg<D, 0>::smem_bytes, kSmemCov = VqStreamCfg<D, 1>::smem_bytes;
    static constexpr int kBigLists = 10;     // one per pass of a round: 8 Lloyd iterations, the projection, the covariance
    int* d_slot_states_ = nullptr;
    uint8_t* d_side_ = nullptr;
    unsigned long long* d_acc_ = nullptr;
    VqSlot<D>* d_slots_ = nullptr;
    VqSlotResult* d_results_ = nullptr;
    VqNodes nodes_ = {};
    std::vector<VqSlotResult> h_results_;
    double t_enqueue_ = 0, t_sync_ = 0, t_results_ = 0, t_heap_ = 0;      // CRN_B200_TRACE attribution of the host loop
    static double now_ms() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

    static unsigned grid(unsigned n) { return (n + 255) / 256; }
    void count() { if (launches_) ++*launches_; }

    // Replays the reference's heap loop for every tree and collects the next frontier.  threaded_clusterizer's partitions
    // are independent clusterizers over disjoint node sets, and with hundreds of thousands of leaves the replay is bound
    // by cache misses, so each tree gets its own host thread once the heaps are large (product build only).
    // one tree: replay, collect, and order its part of the frontier by first position (what the device kernels expect);
    // the keys are gathered first so that the sort itself runs on a contiguous array
    static void advance_tree(VqTreeSim& t, VqNodeVec& nodes, std::vector<uint32_t>& part, double* tm)
    {
        const double a = now_ms();
        t.run(nodes);
        const double b = now_ms();
        t.wanted(part);
        const double c = now_ms();
        tm[0] += b - a; tm[1] += c - b;
        if (part.size() < 2) return;
        std::vector<unsigned long long> keyed(part.size());
        for (size_t i = 0; i < part.size(); i++) keyed[i] = ((unsigned long long)nodes[part[i]].begin << 32) | part[i];
        std::sort(keyed.begin(), keyed.end());
        for (size_t i = 0; i < part.size(); i++) part[i] = (uint32_t)keyed[i];
        tm[2] += now_ms() - c;
    }
    double tm_[4][4] = {};           // CRN_B200_TRACE: per tree, ms in run / wanted / sort
#ifdef __CUDACC__
    std::unique_ptr<VqPool> pool_;
#endif
    void advance_trees(std::vector<VqTreeSim>& trees, VqNodeVec& nodes, std::vector<uint32_t>& frontier)
    {
        // threaded_clusterizer's partitions cover ascending, disjoint position ranges, so the per-tree parts concatenate
        // into a sorted frontier
        std::vector<std::vector<uint32_t>> parts(trees.size());
#ifdef __CUDACC__
        size_t live = 0;
        for (const VqTreeSim& t : trees) live += t.pending.size();
        if (trees.size() > 1 && trees.size() <= 4 && live > 4096) {
            if (!pool_) pool_.reset(new VqPool());
            pool_->run4([&](int i) { if ((size_t)i < trees.size()) advance_tree(trees[i], nodes, parts[i], tm_[i]); });
        } else
#endif
        for (size_t i = 0; i < trees.size(); i++) advance_tree(trees[i], nodes, parts[i], tm_[i & 3]);
        for (const std::vector<uint32_t>& p : parts) frontier.insert(frontier.end(), p.begin(), p.end());
    }

    // every device array is carved from one slab the caller keeps between builds (no cudaMalloc / cudaFree per build)
    cudaError_t allocate(uint32_t n, uint32_t max_size)
    {
        const uint32_t slots = std::max<uint32_t>(4u, std::min<uint32_t>(n / 2 + 4, max_size + 4));
        const size_t node_cap = (size_t)2 * n + 16;
        for (int pass = 0; pass < 2; pass++) {
            size_t off = 0;
            uint8_t* base = pass ? static_cast<uint8_t*>(ws_->base) : nullptr;
            auto carve = [&](auto*& p, size_t cnt) {
                using T = std::remove_pointer_t<std::remove_reference_t<decltype(p)>>;
                if (pass) p = reinterpret_cast<T*>(base + off);
                off += (cnt * sizeof(T) + 255) & ~(size_t)255;
            };
            carve(d_perm_[0], n); carve(d_perm_[1], n); carve(d_pos_slot_, n); carve(d_flags_, (size_t)n + 1); carve(d_scan_, (size_t)n + 1);
            carve(d_block_sums_, (size_t)n / 1024 + 2); carve(d_slot_node_, slots); carve(d_slot_starts_, slots); carve(d_node_counter_, 1); carve(d_active_, 1);
            carve(d_big_count_, kBigLists); carve(d_big_list_, (size_t)kBigLists * slots);
            max_chunks_ = n / kVqChunk + kVqMaxBig;
            carve(d_chunk_start_, kVqMaxBig + 1); carve(d_overflow_count_, kBigLists); carve(d_overflow_list_, slots); carve(d_dir_, kBigLists);
            carve(d_table_, (size_t)max_chunks_ * (2 * D) * (kVqEMax + 1) * 2);
            carve(d_slot_states_, slots); carve(d_side_, n); carve(d_acc_, D + 2); carve(d_slots_, slots); carve(d_results_, slots);
            carve(nodes_.begin, node_cap); carve(nodes_.count, node_cap); carve(nodes_.left, node_cap); carve(nodes_.flags, node_cap);
            carve(nodes_.variance, node_cap); carve(nodes_.weight, node_cap); carve(nodes_.centroid, node_cap * D);
            if (!pass && off > ws_->cap) {
                if (ws_->base) cudaFree(ws_->base);
                ws_->base = nullptr; ws_->cap = 0;
                const cudaError_t ce = cudaMalloc(&ws_->base, off);
                if (ce != cudaSuccess) return ce;
                ws_->cap = off;
            }
        }
        cap_n_ = n; cap_slots_ = slots;
        return cudaSuccess;
    }
    void launch_fill_identity(const uint32_t* d_ids, uint32_t n);

    // float centroid sums of the slots of one pass (see vq_float_sums_kernel); `list` selects this pass's big-slot list
    void float_sums(const unsigned* perm, unsigned F, int phase, int list)
    {
        unsigned* cnt = d_big_count_ + list;
        unsigned* lst = d_big_list_ + (size_t)list * cap_slots_;
        CRN_LAUNCH(vq_float_sums_kernel<D>, (F + 255) / 256, 256, 0, stream_, d_slots_, F, phase, cnt, lst); count();
        // slots whose sums passed 2^24: chunk tables + ordered walk (a few slots that do not fit stream instead)
        VqBigDir* dir = d_dir_ + list;
        unsigned* ovc = d_overflow_count_ + list;
        CRN_LAUNCH(vq_big_dir_kernel<D>, 1, 32, 0, stream_, cnt, lst, d_slots_, d_chunk_start_, dir, ovc, d_overflow_list_, max_chunks_); count();
        CRN_LAUNCH(vq_chunk_sim_kernel<D>, (max_chunks_ + kVqSeqWarps - 1) / kVqSeqWarps, kVqSeqWarps * 32, 0, stream_, vecs_, wts_, perm, d_side_, d_slots_, lst, d_chunk_start_, dir, d_table_); count();
        CRN_LAUNCH(vq_chunk_apply_kernel<D>, (std::min<unsigned>(F, kVqMaxBig) + kVqSeqWarps - 1) / kVqSeqWarps, kVqSeqWarps * 32, 0, stream_, vecs_, wts_, perm, d_side_, d_slots_, lst, d_chunk_start_, dir, d_table_); count();
        CRN_LAUNCH((vq_stream_kernel<D, 0>), stream_grid(F), kVqStreamThreads, kSmemSum, stream_, vecs_, wts_, perm, d_side_, d_slots_, ovc, d_overflow_list_); count();
    }
    static unsigned stream_grid(unsigned F) { return F < 128u ? F : 128u; }

    // exclusive scan of d_flags_[0..m) into d_scan_
    void scan(uint32_t m)
    {
        const unsigned nb = (m + 1023) / 1024;
        CRN_LAUNCH(vq_scan_block_kernel, nb, 256, 0, stream_, d_flags_, d_scan_, d_block_sums_, m); count();
        if (nb > 1) {
            CRN_LAUNCH(vq_scan_sums_kernel, 1, 256, 0, stream_, d_block_sums_, nb); count();
            CRN_LAUNCH(vq_scan_add_kernel, grid(m), 256, 0, stream_, d_scan_, d_block_sums_, m); count();
        }
    }

    // split every node of `frontier` (sorted by first position) on the device and record the results
    // presplit: 0 = clusterizer split; 1 / 2 = first / second level of threaded_clusterizer's PCA divisions
    cudaError_t round(const std::vector<uint32_t>& frontier, VqNodeVec& nodes, int presplit)
    {
        const unsigned F = (unsigned)frontier.size(), n = n_;
        if (!F) return cudaSuccess;
        if (F > cap_slots_) return cudaErrorInvalidValue;
        const unsigned gs = (F + 127) / 128;
        const double t0 = now_ms();
        cudaMemcpyAsync(d_slot_node_, frontier.data(), sizeof(unsigned) * F, cudaMemcpyHostToDevice, stream_);
        unsigned* perm = d_perm_[cur_];
        unsigned* perm_out = d_perm_[cur_ ^ 1];
        CRN_LAUNCH(vq_init_slots_kernel<D>, gs, 128, 0, stream_, d_slot_node_, nodes_, d_slots_, d_slot_starts_, F, presplit ? 1 : 0); count();
        CRN_LAUNCH(vq_pos_slot_kernel<D>, grid(n), 256, 0, stream_, d_slot_starts_, d_slots_, F, d_pos_slot_, n); count();
        const unsigned gw = (F + kVqSeqWarps - 1) / kVqSeqWarps;
        cudaMemsetAsync(d_big_count_, 0, sizeof(unsigned) * kBigLists, stream_);
        CRN_LAUNCH(vq_covariance_kernel<D>, gw, kVqSeqWarps * 32, 0, stream_, vecs_, wts_, perm, d_slots_, F, d_big_count_ + 9, d_big_list_ + (size_t)9 * cap_slots_); count();
        CRN_LAUNCH(vq_stream_cov_kernel<D>, std::min<unsigned>(F * VqCovCfg<D>::S, 592u), VqCovCfg<D>::THREADS, 0, stream_, vecs_, wts_, perm, d_slots_, d_big_count_ + 9, d_big_list_ + (size_t)9 * cap_slots_); count();
        CRN_LAUNCH(vq_axis_kernel<D>, gs, 128, 0, stream_, d_slots_, F, presplit ? 1 : 0); count();
        CRN_LAUNCH(vq_project_kernel<D>, grid(n), 256, 0, stream_, vecs_, wts_, perm, d_pos_slot_, d_slots_, d_side_, n, presplit ? 1 : 0); count();
        float_sums(perm, F, 0, 8);
        if (presplit) {
            CRN_LAUNCH(vq_presplit_children_kernel<D>, gs, 128, 0, stream_, d_slots_, F, presplit == 1 ? 1 : 0); count();
        } else {
            CRN_LAUNCH(vq_children_kernel<D>, gs, 128, 0, stream_, vecs_, perm, d_slots_, F); count();
            CRN_LAUNCH((vq_estimate_kernel<D, 0>), grid(n), 256, 0, stream_, vecs_, perm, d_pos_slot_, d_slots_, n); count();
            CRN_LAUNCH((vq_estimate_kernel<D, 1>), grid(n), 256, 0, stream_, vecs_, perm, d_pos_slot_, d_slots_, n); count();
            CRN_LAUNCH(vq_estimate_finish_kernel<D>, gs, 128, 0, stream_, vecs_, perm, d_slots_, F); count();
            for (int it = 0; it < 8; it++) {
                CRN_LAUNCH(vq_assign_kernel<D>, grid(n), 256, 0, stream_, vecs_, wts_, perm, d_pos_slot_, d_slots_, d_side_, n); count();
                float_sums(perm, F, 1, it);
                CRN_LAUNCH(vq_update_kernel<D>, gs, 128, 0, stream_, d_slots_, F, d_active_); count();
            }
        }
        CRN_LAUNCH(vq_slot_states_kernel<D>, gs, 128, 0, stream_, d_slots_, F, d_slot_states_); count();
        cudaMemsetAsync(d_flags_ + n, 0, sizeof(unsigned), stream_);
        CRN_LAUNCH(vq_left_flags_kernel, grid(n), 256, 0, stream_, d_pos_slot_, d_slot_states_, d_side_, d_flags_, n); count();
        scan(n + 1);
        CRN_LAUNCH(vq_finalize_kernel<D>, gs, 128, 0, stream_, d_slots_, F, d_scan_, n, nodes_, d_node_counter_); count();
        CRN_LAUNCH(vq_scatter_kernel<D>, grid(n), 256, 0, stream_, perm, perm_out, d_pos_slot_, d_slots_, d_side_, d_scan_, n); count();
        CRN_LAUNCH(vq_export_kernel<D>, gs, 128, 0, stream_, d_slots_, F, nodes_, d_results_); count();
        cur_ ^= 1;
        h_results_.resize(F);
        cudaMemcpyAsync(h_results_.data(), d_results_, sizeof(VqSlotResult) * F, cudaMemcpyDeviceToHost, stream_);
        const double t1 = now_ms();
        cudaError_t ce = cudaStreamSynchronize(stream_);
        const double t2 = now_ms();
        t_enqueue_ += t1 - t0; t_sync_ += t2 - t1;
        if (getenv("CRN_B200_TRACE_ROUNDS")) fprintf(stderr, "[crn_b200]   vq<%d> round F=%u enqueue %.2f ms wait %.2f ms\n", D, F, t1 - t0, t2 - t1);
        if (ce != cudaSuccess) return ce;
        ce = cudaGetLastError();
        if (ce != cudaSuccess) return ce;
        // children ids are handed out by a device counter: size the node table once, then scatter the results (disjoint
        // parents and children, so the four quarters of the frontier can go in parallel)
        size_t need = nodes.size();
        for (unsigned s = 0; s < F; s++) if (h_results_[s].state == 1) need = std::max(need, (size_t)h_results_[s].child + 2);
        if (nodes.size() < need) { const size_t old = nodes.size(); nodes.resize(need); for (size_t i = old; i < need; i++) nodes[i] = VqHostNode(); }
        auto scatter = [&](unsigned s0, unsigned s1) {
            for (unsigned s = s0; s < s1; s++) {
                const VqSlotResult& r = h_results_[s];
                VqHostNode& par = nodes[frontier[s]];
                par.processed = 1;
                if (r.state != 1) { par.unsplittable = 1; continue; }
                par.left = (int32_t)r.child;
                VqHostNode& l = nodes[r.child];
                VqHostNode& rr = nodes[r.child + 1];
                l = VqHostNode(); rr = VqHostNode();
                l.begin = par.begin; l.count = r.left_count; l.variance = r.var_left;
                rr.begin = par.begin + r.left_count; rr.count = r.right_count; rr.variance = r.var_right;
                par.child_count[0] = r.left_count; par.child_count[1] = r.right_count;
                par.child_var[0] = r.var_left; par.child_var[1] = r.var_right;
            }
        };
#ifdef __CUDACC__
        if (F > 8192) {
            if (!pool_) pool_.reset(new VqPool());
            pool_->run4([&](int i) { scatter((unsigned)((size_t)F * i / 4), (unsigned)((size_t)F * (i + 1) / 4)); });
        } else
#endif
        scatter(0, F);
        t_results_ += now_ms() - t2;
        return cudaSuccess;
    }
};

__global__ void __launch_bounds__(256) vq_identity_kernel(unsigned* __restrict__ perm, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[i] = i;
}
template <int D> void VqBuilder<D>::launch_fill_identity(const uint32_t* d_ids, uint32_t n)
{
    cur_ = 0;
    if (d_ids) { cudaMemcpyAsync(d_perm_[0], d_ids, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, stream_); return; }
    CRN_LAUNCH(vq_identity_kernel, grid(n), 256, 0, stream_, d_perm_[0], n); count();
}

}  // namespace crn
