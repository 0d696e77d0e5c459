// crn_b200.cu -- C-ABI entry points of libcrn_b200.so (see include/crn_b200.h).
//
// One translation unit: the kernels live in the .cuh files included below.  Built by nvcc for
// sm_100a (the product) and, for the CPU-side tests only, by g++ against tests/cusim (SIMT emulator).
#include "../../include/crn_b200.h"
#include "launch.h"
#include "pack_kernels.cuh"
#include "transcode_host.h"
#include "transcode_wide.cuh"
#include "transcode_streams.cuh"
#include "cluster_kernels.cuh"
#include "qdxt_kernels.cuh"
#include "vq_host.h"
#include "vq_fast_host.h"
#include "refiner_kernels.cuh"
#include "hc_kernels.cuh"
#include "unpack_kernels.cuh"
#include "dds_kernels.cuh"
#include "mip_kernels.cuh"
#include "mip_host.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <dlfcn.h>
#include <new>
#include <vector>
#include <algorithm>
#ifdef __CUDACC__
#include <thread>
#include <cub/device/device_radix_sort.cuh>
#endif

struct crn_gpu_ctx {
    int device;
    cudaStream_t stream;
    int sm_count;
    uint64_t launches;
    char err[256];
    crn_gpu_progress_fn progress; void* progress_user;   // crn_gpu_set_progress
    // reusable device staging for the *_host entry points
    void* d_in; size_t d_in_cap;
    void* d_out; size_t d_out_cap;
    void* d_state; size_t d_state_cap;   // Dxt1BlockState scratch of the colour phase kernels
    void* d_cluster_ws; size_t d_cluster_ws_cap;   // hash / colour workspace of the cluster optimiser
    const uint32_t* d_cluster_order;     // set by the dxt_hc pipeline: clusters in descending size, the order the work-stealing loop takes them
    int refine_parallel;                 // set by the dxt_hc pipeline outside exact mode: the refiner's least-squares sums lane-parallel (refiner_kernels.cuh)
    uint32_t cluster_big_count;          // how many leading entries of d_cluster_order have >= kClusterCoopMinBlocks member blocks (a CTA each)
    uint32_t* d_cluster_flags;           // set by the dxt_hc pipeline around a cluster-optimiser call: per-cluster m_reordered / m_alternate_rounding out
    void* d_files; size_t d_files_cap;   // TranscodeFile array of a batched transcode launch
    void* d_wide; size_t d_wide_cap;     // transition tables + pair offsets of the wide transcoder
    int wide_smem_set, streams_smem_set;
    crn::VqWorkspace vq_ws;              // slab of the vector quantiser
    crn::VqFastScratch* vq_scratch;      // host arrays of the single-launch builder, kept between builds
    int vq_exact;                        // crn_gpu_set_vq_mode: 1 = member-order float emulation (vq_kernels.cuh), 0 = single-launch frontier splits (vq_fast.cuh)
    int transcode_smem_set;
    // clustered path: per-element child contexts (own stream + scratch) and a cache of released device buffers, both
    // kept for the life of this context so that compressing texture after texture does not pay cudaMalloc / cudaFree
    crn_gpu_ctx* child[3];
    struct PoolBlock { void* p; size_t cap; };
    std::vector<PoolBlock>* pool;
    std::vector<PoolBlock>* pin_pool;    // pinned host staging blocks of the dxt_hc pipeline, cached like the device pool
    uint64_t pool_mallocs;               // cudaMalloc calls made by pool_alloc since creation (crn_gpu_pool_mallocs): 0 growth in steady state
};

namespace {

int set_err(crn_gpu_ctx* ctx, int code, const char* what, cudaError_t ce = cudaSuccess)
{
    if (ctx) {
        if (ce != cudaSuccess) snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what, cudaGetErrorString(ce));
        else snprintf(ctx->err, sizeof(ctx->err), "%s", what);
    }
    return code;
}

// crn_progress_callback_func semantics: called on the calling thread between phases; 0 from the callback cancels the call.
int progress_tick(crn_gpu_ctx* ctx, uint32_t phase, uint32_t total, uint32_t sub, uint32_t subtotal)
{
    if (!ctx || !ctx->progress) return CRN_GPU_OK;
    return ctx->progress(phase, total, sub, subtotal, ctx->progress_user) ? CRN_GPU_OK : set_err(ctx, CRN_GPU_ERR_CANCELLED, "cancelled by the progress callback");
}

// Every extern "C" body runs inside this: nothing may unwind across the C boundary (include/crn_b200.h: "never throws").
template <typename F>
int crn_guard(crn_gpu_ctx* ctx, F&& body) noexcept
{
    try { return body(); }
    catch (const std::bad_alloc&) { return set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "out of host memory"); }
    catch (const std::exception& e) { return set_err(ctx, CRN_GPU_ERR_BAD_DATA, e.what()); }
    catch (...) { return set_err(ctx, CRN_GPU_ERR_BAD_DATA, "unexpected exception"); }
}

#define CRN_CUDA(ctx, call)                                                          \
    do {                                                                             \
        cudaError_t ce_ = (call);                                                    \
        if (ce_ != cudaSuccess) return set_err((ctx), CRN_GPU_ERR_CUDA, #call, ce_); \
    } while (0)

int ensure(crn_gpu_ctx* ctx, void** p, size_t* cap, size_t need)
{
    if (*cap >= need) return CRN_GPU_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    // Grow with 1/8 headroom, in whole 2 MiB pages: a bitrate search asks for a few KB more every time the cluster count reaches a new
    // maximum, and re-allocating a 3 GB workspace (cudaFree + cudaMalloc: several hundred ms) for that was the slowest step of a trial.
    size_t want = need > (64u << 20) ? need + need / 8 : need;
    want = (want + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    cudaError_t ce = cudaMalloc(p, want);
    if (ce != cudaSuccess) { (void)cudaGetLastError(); want = need; ce = cudaMalloc(p, want); }
    if (ce != cudaSuccess) return set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "cudaMalloc", ce);
    *cap = want;
    return CRN_GPU_OK;
}

// device buffers of the clustered path: best fit from the context's cache (at most 2x oversize), else cudaMalloc
// Requests are rounded up to size classes (eight per power of two above 64 KiB, i.e. at most 12.5 % slack): the trials of a bitrate search ask
// for slightly different sizes every time (codebook sizes change with the quality level), and exact-size blocks would never be reused.
size_t pool_size_class(size_t bytes)
{
    if (bytes <= (64u << 10)) return (bytes + 4095) & ~(size_t)4095;
    size_t p2 = 64u << 10;
    while (p2 * 2 <= bytes) p2 *= 2;
    const size_t step = p2 / 8;
    return ((bytes + step - 1) / step) * step;
}

cudaError_t pool_alloc(crn_gpu_ctx* ctx, void** out, size_t bytes, size_t* cap_out)
{
    if (!bytes) bytes = 256;
    bytes = pool_size_class(bytes);
    if (ctx->pool) {
        int best = -1;
        for (size_t i = 0; i < ctx->pool->size(); i++) {
            const size_t c = (*ctx->pool)[i].cap;
            if (c >= bytes && c <= 2 * bytes + (1u << 20) && (best < 0 || c < (*ctx->pool)[best].cap)) best = (int)i;
        }
        if (best >= 0) {
            *out = (*ctx->pool)[best].p; *cap_out = (*ctx->pool)[best].cap;
            ctx->pool->erase(ctx->pool->begin() + best);
            return cudaSuccess;
        }
    }
    ctx->pool_mallocs++;
    cudaError_t ce = cudaMalloc(out, bytes);
    if (ce != cudaSuccess && ctx->pool && !ctx->pool->empty()) {       // give the cache back and retry once
        for (auto& b : *ctx->pool) cudaFree(b.p);
        ctx->pool->clear();
        (void)cudaGetLastError();
        ce = cudaMalloc(out, bytes);
    }
    *cap_out = bytes;
    return ce;
}

// pinned host memory for the transfers of the dxt_hc pipeline: DMA at link speed instead of the driver's pageable bounce
// copies, and no dependence on how busy the host's memory system is.  Cached per context; nullptr when pinning fails.
void* pin_alloc(crn_gpu_ctx* ctx, size_t bytes, size_t* cap_out)
{
    if (!bytes) bytes = 256;
    bytes = pool_size_class(bytes);
    if (ctx->pin_pool) {
        int best = -1;
        for (size_t i = 0; i < ctx->pin_pool->size(); i++) {
            const size_t c = (*ctx->pin_pool)[i].cap;
            if (c >= bytes && c <= 2 * bytes + (1u << 20) && (best < 0 || c < (*ctx->pin_pool)[best].cap)) best = (int)i;
        }
        if (best >= 0) {
            void* p = (*ctx->pin_pool)[best].p; *cap_out = (*ctx->pin_pool)[best].cap;
            ctx->pin_pool->erase(ctx->pin_pool->begin() + best);
            return p;
        }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    *cap_out = bytes;
    return p;
}
void pin_free(crn_gpu_ctx* ctx, void* p, size_t cap)
{
    if (!p) return;
    if (!ctx->pin_pool) ctx->pin_pool = new (std::nothrow) std::vector<crn_gpu_ctx::PoolBlock>();
    if (!ctx->pin_pool || ctx->pin_pool->size() >= 32) { cudaFreeHost(p); return; }
    ctx->pin_pool->push_back({p, cap});
}

void pool_free(crn_gpu_ctx* ctx, void* p, size_t cap)
{
    if (!p) return;
    if (!ctx->pool) ctx->pool = new (std::nothrow) std::vector<crn_gpu_ctx::PoolBlock>();
    if (!ctx->pool || ctx->pool->size() >= 160) { cudaFree(p); return; }   // one dxt_hc call holds ~70 buffers
    ctx->pool->push_back({p, cap});
}

int grid_for(const crn_gpu_ctx* ctx, uint32_t total_blocks, int warps_per_cta, int ctas_per_sm)
{
    const uint32_t need = (total_blocks + warps_per_cta - 1) / warps_per_cta;
    const uint32_t cap = (uint32_t)(ctx->sm_count * ctas_per_sm);
    uint32_t g = need < cap ? need : cap;
    return (int)(g ? g : 1);
}

}  // namespace

namespace {
template <int D>
int vq_clusterize(crn_gpu_ctx* ctx, const uint8_t* d_vecs, const uint32_t* d_wts, uint32_t n, uint32_t max_size, uint32_t retrieve, int threaded,
                  uint32_t* h_cluster_of, uint32_t* num_clusters, uint32_t* codebook_size)
{
    crn::VqResult res;
    cudaError_t ce;
    if (ctx->vq_exact) { crn::VqBuilder<D> builder(ctx->stream, &ctx->launches, &ctx->vq_ws); ce = builder.build(d_vecs, d_wts, nullptr, n, max_size, threaded != 0, res); }
    else {
        if (!ctx->vq_scratch) ctx->vq_scratch = new crn::VqFastScratch();
        crn::VqFastBuilder<D> builder(ctx->stream, &ctx->launches, &ctx->vq_ws, ctx->sm_count, ctx->vq_scratch);
        ce = builder.build(d_vecs, d_wts, nullptr, n, max_size, threaded != 0, res);
    }
    if (ce != cudaSuccess) return set_err(ctx, ce == cudaErrorMemoryAllocation ? CRN_GPU_ERR_NO_MEMORY : CRN_GPU_ERR_CUDA, "crn_gpu_vq_clusterize", ce);
    if (codebook_size) *codebook_size = res.codebook_size();
    const uint32_t k = res.retrieve(retrieve, h_cluster_of);
    if (num_clusters) *num_clusters = k;
    return CRN_GPU_OK;
}
}  // namespace

static bool build_mip_table(const crn_gpu_mip_desc* mips, uint32_t num_mips, uint32_t n_blocks, crn::QdxtMipTable& mt)
{
    if (!mips || !num_mips || num_mips > (uint32_t)crn::kQdxtMaxMips) return false;
    uint32_t chunks = 0;
    for (uint32_t i = 0; i < num_mips; i++) {
        if (!mips[i].block_width || !mips[i].block_height ||
            (uint64_t)mips[i].first_block + (uint64_t)mips[i].block_width * mips[i].block_height > n_blocks) return false;
        mt.m[i].first_block = mips[i].first_block; mt.m[i].block_width = mips[i].block_width; mt.m[i].block_height = mips[i].block_height;
        mt.m[i].first_chunk = chunks;
        chunks += ((mips[i].block_width + 1) / 2) * ((mips[i].block_height + 1) / 2);
    }
    mt.num_mips = num_mips; mt.total_chunks = chunks;
    return true;
}

struct crn_qdxt_element {
    int kind;                         // 0 colour (qdxt1), 1 alpha (qdxt5)
    uint32_t comp;                    // source channel of an alpha element
    uint32_t offset;                  // byte offset of the element inside a block
    int use_alpha_blocks;             // qdxt1_params::m_use_alpha_blocks
    crn::VqResult endpoint_tree;
    crn::VqResult sel_tree;           // selector tree of the last pack(): kept so that its host arrays are reused by the next one
    uint32_t max_selector_clusters;
    uint32_t endpoint_clusters, selector_clusters;
    // every element is an independent chain of short, latency-bound kernels with host decisions in between, so each
    // one gets its own stream + scratch (a private crn_gpu_ctx) and its own host thread
    crn_gpu_ctx* ctx;
    uint8_t* d_vecs;                  // n_blocks x 16: training / selector vectors
    uint32_t* d_wts;
    uint8_t* d_cat;
    uint32_t *d_offsets, *d_members, *d_ids;
    uint32_t* d_ep_perm;              // final permutation of the endpoint tree: every cluster pack() retrieves is a range of it
    unsigned long long* d_keys;       // per-block dxt_fast selector keys + the distinct-count table
    size_t caps[8];                   // pool capacities of d_vecs, d_wts, d_cat, d_offsets, d_members, d_ids, d_keys, d_ep_perm
    std::vector<uint32_t> cluster_of, offsets, members;
    std::vector<uint8_t> cat;
    cudaEvent_t ev_opt[2];            // brackets the endpoint optimisation of the last pack()
    float endpoint_opt_ms;
    uint64_t opt_stats[3];            // of the last pack(): candidates evaluated, unique colours they ranged over (sum of U), palette entries per evaluation
    int vq_exact;                     // copied from the parent context at init
    int rc;
};

struct crn_gpu_qdxt {
    crn_gpu_ctx* ctx;
    uint32_t format, n_blocks, num_levels, bytes_per_block, num_elements;
    float pow_mul;
    int flat;                         // qdxt1_params::m_hierarchical == false: every block is its own tile (crn_qdxt1.cpp:370-403)
    crn_gpu_pack_params params;
    std::vector<crn_gpu_mip_desc> mips;
    crn_qdxt_element el[3];
    uint32_t* d_blocks;               // n_blocks x 16 RGBA8
    uint8_t* d_out;                   // n_blocks x bytes_per_block
    size_t d_blocks_cap, d_out_cap;
};

namespace {

struct HcBufLite {                                // pooled device scratch, returned on scope exit
    crn_gpu_ctx* ctx; void* p = nullptr; size_t cap = 0;
    explicit HcBufLite(crn_gpu_ctx* c) : ctx(c) {}
    ~HcBufLite() { if (p) pool_free(ctx, p, cap); }
    cudaError_t alloc(size_t bytes) { return pool_alloc(ctx, &p, bytes ? bytes : 1, &cap); }
};

void qdxt_release(crn_gpu_qdxt* q)
{
    for (uint32_t i = 0; i < q->num_elements; i++) {
        crn_qdxt_element& e = q->el[i];
        void* ptrs[] = {e.d_vecs, e.d_wts, e.d_cat, e.d_offsets, e.d_members, e.d_ids, e.d_keys, e.d_ep_perm};
        for (int k = 0; k < 8; k++) pool_free(q->ctx, ptrs[k], e.caps[k]);
        for (cudaEvent_t ev : e.ev_opt) if (ev) cudaEventDestroy(ev);
        // e.ctx is q->ctx->child[i]: it stays with the parent context
    }
    pool_free(q->ctx, q->d_blocks, q->d_blocks_cap);
    pool_free(q->ctx, q->d_out, q->d_out_cap);
    delete q;
}

// cluster_of over `ids` (nullptr = all blocks) -> CSR appended to e.offsets / e.members; members ascending
void qdxt_append_csr(crn_qdxt_element& e, const uint32_t* ids, uint32_t n, uint32_t n_clusters)
{
    const size_t first_cluster = e.offsets.size() - 1, base = e.members.size();
    std::vector<uint32_t> cnt(n_clusters + 1, 0u);
    for (uint32_t i = 0; i < n; i++) cnt[e.cluster_of[ids ? ids[i] : i] + 1]++;
    for (uint32_t k = 0; k < n_clusters; k++) cnt[k + 1] += cnt[k];
    e.members.resize(base + n);
    e.offsets.resize(first_cluster + n_clusters + 1);
    for (uint32_t k = 0; k < n_clusters; k++) e.offsets[first_cluster + k + 1] = (uint32_t)(base + cnt[k + 1]);
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t b = ids ? ids[i] : i;
        e.members[base + cnt[e.cluster_of[b]]++] = b;
    }
}

int qdxt_upload_csr(crn_qdxt_element& e)
{
    crn_gpu_ctx* ctx = e.ctx;
    CRN_CUDA(ctx, cudaMemcpyAsync(e.d_offsets, e.offsets.data(), e.offsets.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (!e.members.empty())
        CRN_CUDA(ctx, cudaMemcpyAsync(e.d_members, e.members.data(), e.members.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));       // the host vectors are reused right away
    return CRN_GPU_OK;
}

template <int D>
int qdxt_vq(crn_qdxt_element& e, const uint32_t* d_ids, uint32_t n, uint32_t max_size, bool threaded, crn::VqResult& res, uint32_t* d_perm_out = nullptr, bool need_ranks = true)
{
    cudaError_t ce;
    if (e.vq_exact) { crn::VqBuilder<D> builder(e.ctx->stream, &e.ctx->launches, &e.ctx->vq_ws); ce = builder.build(e.d_vecs, e.d_wts, d_ids, n, max_size, threaded, res, d_perm_out); }
    else {
        if (!e.ctx->vq_scratch) e.ctx->vq_scratch = new crn::VqFastScratch();
        crn::VqFastBuilder<D> builder(e.ctx->stream, &e.ctx->launches, &e.ctx->vq_ws, e.ctx->sm_count, e.ctx->vq_scratch);
        ce = builder.build(e.d_vecs, e.d_wts, d_ids, n, max_size, threaded, res, d_perm_out, need_ranks);
    }
    if (ce != cudaSuccess) return set_err(e.ctx, ce == cudaErrorMemoryAllocation ? CRN_GPU_ERR_NO_MEMORY : CRN_GPU_ERR_CUDA, "clustered DDS: vector quantiser", ce);
    return CRN_GPU_OK;
}

// runs fn(element) for every element, one host thread each (the SIMT-emulation test build is single threaded)
template <typename F>
int qdxt_for_each_element(crn_gpu_qdxt* q, F fn)
{
#ifdef __CUDACC__
    std::vector<std::thread> threads;
    for (uint32_t i = 1; i < q->num_elements; i++)
        threads.emplace_back([q, i, &fn]() { cudaSetDevice(q->ctx->device); q->el[i].rc = fn(q->el[i]); });
    q->el[0].rc = fn(q->el[0]);
    for (std::thread& t : threads) t.join();
#else
    for (uint32_t i = 0; i < q->num_elements; i++) q->el[i].rc = fn(q->el[i]);
#endif
    for (uint32_t i = 0; i < q->num_elements; i++) {
        crn_qdxt_element& e = q->el[i];
        q->ctx->launches += e.ctx->launches; e.ctx->launches = 0;
        if (e.rc) { snprintf(q->ctx->err, sizeof(q->ctx->err), "%s", e.ctx->err); return e.rc; }
    }
    return CRN_GPU_OK;
}

// CRN_B200_TRACE=1: wall-clock per phase (synchronising), for tuning only
struct QdxtTrace {
    bool on; crn_gpu_ctx* ctx; double t0;
    static double now() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    explicit QdxtTrace(crn_gpu_ctx* c) : on(getenv("CRN_B200_TRACE") != nullptr), ctx(c), t0(0) { if (on) { cudaStreamSynchronize(ctx->stream); t0 = now(); } }
    void mark(const char* what, int el)
    {
        if (!on) return;
        cudaStreamSynchronize(ctx->stream);
        const double t = now();
        fprintf(stderr, "[crn_b200] element %d %-28s %8.2f ms  (launches so far %llu)\n", el, what, t - t0, (unsigned long long)ctx->launches);
        t0 = t;
    }
};

// qdxt1::init / qdxt5::init for one element
int qdxt_init_element(crn_gpu_qdxt* q, crn_qdxt_element& e)
{
    crn_gpu_ctx* ctx = e.ctx;
    const uint32_t n = q->n_blocks;
    crn::QdxtMipTable mt;
    if (!build_mip_table(q->mips.data(), (uint32_t)q->mips.size(), n, mt)) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "clustered DDS: bad level table");
    QdxtTrace tr(ctx);
    const int eli = (int)(&e - q->el);
    const int threads = crn::kQdxtWarpsPerCta * 32;
    const int grid = grid_for(ctx, mt.total_chunks, crn::kQdxtWarpsPerCta, 8);
    if (e.kind == 0)
        CRN_LAUNCH(crn::qdxt_training_kernel<0>, grid, threads, 0, ctx->stream, q->d_blocks, mt, 3u, e.d_vecs, e.d_wts, (uint8_t*)nullptr, e.d_keys, q->flat);
    else
        CRN_LAUNCH(crn::qdxt_training_kernel<1>, grid, threads, 0, ctx->stream, q->d_blocks, mt, e.comp, e.d_vecs, e.d_wts, (uint8_t*)nullptr, e.d_keys, q->flat);
    ctx->launches++;
    CRN_CUDA(ctx, cudaGetLastError());
    // distinct dxt_fast selector patterns (crn_qdxt1.cpp:415-438, crn_qdxt5.cpp:395-421)
    uint32_t cap = 1024;
    while (cap < 2 * n) cap <<= 1;
    unsigned long long* table = e.d_keys + n;
    unsigned* counter = reinterpret_cast<unsigned*>(table + cap);
    CRN_CUDA(ctx, cudaMemsetAsync(table, 0xff, sizeof(unsigned long long) * cap, ctx->stream));
    CRN_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned), ctx->stream));
    CRN_LAUNCH(crn::count_distinct_kernel, (n + 255) / 256, 256, 0, ctx->stream, e.d_keys, n, table, cap - 1, counter);
    ctx->launches++;
    unsigned distinct = 0;
    CRN_CUDA(ctx, cudaMemcpyAsync(&distinct, counter, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    e.max_selector_clusters = distinct + 128;
    tr.mark("init: tile analysis + distinct", eli);
    // endpoint codebook: generate_codebook(65535) (crn_qdxt1.cpp:405-413, crn_qdxt5.cpp:386-393)
    // the tree's final permutation stays on the device: retrieve_clusters() at any size is a set of ranges of it (vq_range_offsets)
    const int rc = e.kind == 0 ? qdxt_vq<6>(e, nullptr, n, 65535u, false, e.endpoint_tree, e.d_ep_perm) : qdxt_vq<2>(e, nullptr, n, 65535u, false, e.endpoint_tree, e.d_ep_perm);
    if (tr.on) fprintf(stderr, "[crn_b200] endpoint tree: %u rounds, %u device splits\n", e.endpoint_tree.rounds, e.endpoint_tree.device_splits);
    tr.mark("init: endpoint tree", eli);
    return rc;
}

uint32_t clampu(uint32_t v, uint32_t lo, uint32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }   // math::clamp

// qdxt1::pack (crn_qdxt1.cpp:910-1030) / qdxt5::pack (crn_qdxt5.cpp:838-960) for one element
int qdxt_pack_element(crn_gpu_qdxt* q, crn_qdxt_element& e, uint32_t quality_level)
{
    crn_gpu_ctx* ctx = e.ctx;
    const uint32_t n = q->n_blocks, stride = q->bytes_per_block;
    const float quality = quality_level / 255.0f;
    const uint32_t codebook = e.endpoint_tree.codebook_size();
    uint32_t max_endpoint_clusters, max_selector_clusters;
    if (e.kind == 0) {
        const float eq = powf(quality, 1.8f * q->pow_mul), sq = powf(quality, 1.65f * q->pow_mul);
        max_endpoint_clusters = clampu((uint32_t)(codebook * eq), 96u, codebook);
        max_selector_clusters = clampu((uint32_t)(e.max_selector_clusters * sq), 128u, e.max_selector_clusters);
    } else {
        const float eq = powf(quality, 2.1f), sq = powf(quality, 1.65f);
        max_endpoint_clusters = clampu((uint32_t)(codebook * eq), 16u, codebook);
        max_selector_clusters = clampu((uint32_t)(e.max_selector_clusters * sq), 32u, e.max_selector_clusters);
    }
    QdxtTrace tr(ctx);
    const int eli = (int)(&e - q->el);
    crn_gpu_pack_params pp = q->params;
    if (e.kind == 0) pp.use_both_block_types = e.use_alpha_blocks ? 1u : 0u;      // qdxt5 keeps pack_params' flag (crn_qdxt5.cpp:468)
    // endpoint clusters
    uint32_t k_end;
    e.offsets.assign(1, 0u); e.members.clear();
    int rc;
    if (quality >= 1.0f) {
        e.cluster_of.resize(n);
        for (uint32_t i = 0; i < n; i++) e.cluster_of[i] = i;
        k_end = n;
        qdxt_append_csr(e, nullptr, n, k_end);
        rc = qdxt_upload_csr(e);
    } else {
        // retrieve_clusters(max_endpoint_clusters): every cluster is a contiguous, ascending range of the tree's permutation, so only the
        // (<= 65535 + 1) offsets are built here and the member list is a device-to-device copy
        k_end = crn::vq_range_offsets(e.endpoint_tree, max_endpoint_clusters, e.offsets);
        rc = qdxt_upload_csr(e);                              // offsets only
#ifdef __CUDACC__
        if (rc == CRN_GPU_OK) {
            // d_vecs (16 bytes per block) is free between the endpoint tree and the selector vectors: cluster_of | sorted keys | ids
            uint32_t* d_cl = reinterpret_cast<uint32_t*>(e.d_vecs);
            uint32_t *d_cl_sorted = d_cl + n, *d_id = d_cl + 2 * (size_t)n;
            CRN_LAUNCH(crn::range_cluster_of_kernel, (n + 255) / 256, 256, 0, ctx->stream, e.d_ep_perm, e.d_offsets, k_end, n, d_cl, d_id);
            int bits = 1;
            while ((1u << bits) < k_end && bits < 32) bits++;
            size_t temp_bytes = 0;
            CRN_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, d_cl, d_cl_sorted, d_id, e.d_members, (int)n, 0, bits, ctx->stream));
            HcBufLite tmp(ctx);
            if (tmp.alloc(temp_bytes) != cudaSuccess) return set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "clustered DDS: sort scratch");
            CRN_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, temp_bytes, d_cl, d_cl_sorted, d_id, e.d_members, (int)n, 0, bits, ctx->stream));
            ctx->launches += 3;
            CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));       // tmp goes back to the pool
        }
#else
        if (rc == CRN_GPU_OK) {                               // emulation build: the same on the host
            std::vector<uint32_t> perm(n);
            CRN_CUDA(ctx, cudaMemcpyAsync(perm.data(), e.d_ep_perm, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            e.cluster_of.resize(n);
            for (uint32_t c = 0; c < k_end; c++) for (uint32_t i = e.offsets[c]; i < e.offsets[c + 1]; i++) e.cluster_of[perm[i]] = c;
            e.offsets.assign(1, 0u); e.members.clear();
            qdxt_append_csr(e, nullptr, n, k_end);
            rc = qdxt_upload_csr(e);
        }
#endif
    }
    e.endpoint_clusters = k_end;
    if (rc) return rc;
    if (tr.on) {
        std::vector<uint32_t> sz(k_end);
        for (uint32_t k = 0; k < k_end; k++) sz[k] = e.offsets[k + 1] - e.offsets[k];
        std::sort(sz.begin(), sz.end(), [](uint32_t a, uint32_t b) { return a > b; });
        fprintf(stderr, "[crn_b200] element %d endpoint clusters %u, largest (blocks):", eli, k_end);
        for (uint32_t k = 0; k < 8 && k < k_end; k++) fprintf(stderr, " %u", sz[k]);
        fprintf(stderr, "  median %u\n", sz[k_end / 2]);
    }
    tr.mark("pack: retrieve + CSR", eli);
    cudaEventRecord(e.ev_opt[0], ctx->stream);
    if (e.kind == 0)
        rc = crn_gpu_dxt1_optimize_clusters(ctx, &pp, e.use_alpha_blocks, q->d_blocks, n, e.d_offsets, e.d_members, k_end, n, q->d_out, stride, e.offset, nullptr, nullptr);
    else
        rc = crn_gpu_dxt5_optimize_clusters(ctx, &pp, e.comp, q->d_blocks, n, e.d_offsets, e.d_members, k_end, n, q->d_out, stride, e.offset, nullptr, nullptr);
    cudaEventRecord(e.ev_opt[1], ctx->stream);
    if (rc) return rc;
    tr.mark("pack: endpoint optimisation", eli);
    e.selector_clusters = 0;
    if (quality >= 1.0f) return CRN_GPU_OK;
    // selector training vectors
    if (e.kind == 0)
        CRN_LAUNCH(crn::selector_vectors_kernel<0>, (n + 255) / 256, 256, 0, ctx->stream, q->d_out, stride, e.offset, n, pp.perceptual ? 1 : 0, e.d_vecs, e.d_wts, (uint8_t*)nullptr);
    else
        CRN_LAUNCH(crn::selector_vectors_kernel<1>, (n + 255) / 256, 256, 0, ctx->stream, q->d_out, stride, e.offset, n, 0, e.d_vecs, e.d_wts, e.d_cat);
    ctx->launches++;
    CRN_CUDA(ctx, cudaGetLastError());
    e.offsets.assign(1, 0u); e.members.clear();
    crn::VqResult& sel_tree = e.sel_tree;
    // selector clusters keep every leaf, so the CSR lists are the leaves' position ranges over the final permutation,
    // which goes device-to-device into d_members (see vq_leaf_offsets); only the offsets travel
    uint32_t sel_members = 0;
    if (e.kind == 0) {
        tr.mark("pack: selector vectors", eli);
        rc = qdxt_vq<16>(e, nullptr, n, max_selector_clusters, true, sel_tree, e.d_members, false);
        if (rc) return rc;
        tr.mark("pack: selector VQ build", eli);
        crn::vq_leaf_offsets(sel_tree, 0u, e.offsets);
        sel_members = n;
    } else {
        e.cat.resize(n);
        CRN_CUDA(ctx, cudaMemcpyAsync(e.cat.data(), e.d_cat, n, cudaMemcpyDeviceToHost, ctx->stream));
        CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        std::vector<uint32_t> ids;
        for (uint32_t type = 0; type < 2; type++) {                 // crn_qdxt5.cpp:762-833
            ids.clear();
            for (uint32_t b = 0; b < n; b++) if (e.cat[b] == type) ids.push_back(b);
            const uint32_t m = (uint32_t)ids.size();
            if (!m) continue;
            if ((m / (float)n) < .01f) continue;
            uint32_t max_clusters = (uint32_t)(((uint64_t)m * max_selector_clusters + (n - 1)) / n);
            max_clusters = std::min(std::max(64u, max_clusters), m);
            if (max_clusters >= m) continue;
            CRN_CUDA(ctx, cudaMemcpyAsync(e.d_ids, ids.data(), (size_t)m * 4, cudaMemcpyHostToDevice, ctx->stream));
            rc = qdxt_vq<16>(e, e.d_ids, m, max_clusters, true, sel_tree, e.d_members + sel_members, false);
            if (rc) return rc;
            crn::vq_leaf_offsets(sel_tree, sel_members, e.offsets);
            sel_members += m;
        }
    }
    const uint32_t k_sel = (uint32_t)e.offsets.size() - 1;
    e.selector_clusters = k_sel;
    tr.mark("pack: selector VQ", eli);
    if (!k_sel) return CRN_GPU_OK;
    rc = qdxt_upload_csr(e);
    if (rc) return rc;
    rc = crn_gpu_optimize_selectors(ctx, (uint32_t)e.kind, &pp, e.comp, q->d_blocks, n, e.d_offsets, e.d_members, k_sel, q->d_out, stride, e.offset);
    tr.mark("pack: selector re-vote", eli);
    return rc;
}

}  // namespace


#include "hc_host.h"
#include "crn_writer.h"
#include "writer_kernels.cuh"

extern "C" {

uint32_t crn_gpu_abi_version(void) { return CRN_B200_ABI_VERSION; }

int crn_gpu_is_native(void)
{ return crn_guard(nullptr, [&]() -> int {
#ifdef __CUDACC__
    return 1;
#else
    return 0;
#endif
}); }

int crn_gpu_device_count(void)
{ return crn_guard(nullptr, [&]() -> int {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}); }

int crn_gpu_create(int device, crn_gpu_ctx** out_ctx)
{ return crn_guard(nullptr, [&]() -> int {
    if (!out_ctx) return CRN_GPU_ERR_BAD_PARAM;
    *out_ctx = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return CRN_GPU_ERR_NO_DEVICE;
    crn_gpu_ctx* ctx = new (std::nothrow) crn_gpu_ctx();
    if (!ctx) return CRN_GPU_ERR_NO_MEMORY;
    memset(static_cast<void*>(ctx), 0, sizeof(*ctx));
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return CRN_GPU_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return CRN_GPU_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    { const char* e = getenv("CRN_B200_VQ_EXACT"); ctx->vq_exact = (e && *e && *e != '0') ? 1 : 0; }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return CRN_GPU_ERR_CUDA; }
    *out_ctx = ctx;
    return CRN_GPU_OK;
}); }

void crn_gpu_destroy(crn_gpu_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->d_in) cudaFree(ctx->d_in);
    if (ctx->d_out) cudaFree(ctx->d_out);
    if (ctx->d_state) cudaFree(ctx->d_state);
    if (ctx->d_files) cudaFree(ctx->d_files);
    if (ctx->d_wide) cudaFree(ctx->d_wide);
    if (ctx->vq_ws.base) cudaFree(ctx->vq_ws.base);
    if (ctx->vq_ws.nodes) cudaFree(ctx->vq_ws.nodes);
    delete ctx->vq_scratch;
    if (ctx->d_cluster_ws) cudaFree(ctx->d_cluster_ws);
    for (crn_gpu_ctx* c : ctx->child) if (c) crn_gpu_destroy(c);
    if (ctx->pool) {
        for (auto& b : *ctx->pool) cudaFree(b.p);
        delete ctx->pool;
    }
    if (ctx->pin_pool) {
        for (auto& b : *ctx->pin_pool) cudaFreeHost(b.p);
        delete ctx->pin_pool;
    }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* crn_gpu_last_error(const crn_gpu_ctx* ctx) { return ctx ? ctx->err : "null context"; }
void* crn_gpu_stream(crn_gpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t crn_gpu_launch_count(const crn_gpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

int crn_gpu_synchronize(crn_gpu_ctx* ctx)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CRN_GPU_OK;
}); }

uint64_t crn_gpu_pool_mallocs(const crn_gpu_ctx* ctx)
{
    if (!ctx) return 0;
    uint64_t n = ctx->pool_mallocs;
    for (int i = 0; i < 3; i++) if (ctx->child[i]) n += ctx->child[i]->pool_mallocs;
    return n;
}

void crn_gpu_set_vq_mode(crn_gpu_ctx* ctx, int exact_member_order)
{
    if (ctx) ctx->vq_exact = exact_member_order ? 1 : 0;
}

void crn_gpu_set_progress(crn_gpu_ctx* ctx, crn_gpu_progress_fn fn, void* user)
{
    if (ctx) { ctx->progress = fn; ctx->progress_user = user; }
}

void crn_gpu_default_pack_params(crn_gpu_pack_params* p)
{   // defaults of crn_comp_params::clear() (inc/crnlib.h:239-273) as seen by the block packer
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->struct_size = sizeof(*p);
    p->dxt_quality = 4;
    p->perceptual = 1;
    p->use_both_block_types = 1;
    p->dxt1a_alpha_threshold = 128;
}

uint32_t crn_gpu_bytes_per_block(uint32_t format)
{
    switch (format) {
    case CRN_GPU_FMT_DXT1: case CRN_GPU_FMT_DXT1A: case CRN_GPU_FMT_DXT5A: return 8;
    case CRN_GPU_FMT_DXT3: case CRN_GPU_FMT_DXT5: case CRN_GPU_FMT_DXN_XY: case CRN_GPU_FMT_DXN_YX: return 16;
    default: return 0;
    }
}

int crn_gpu_pack_image(crn_gpu_ctx* ctx, uint32_t format, const crn_gpu_pack_params* params,
                       const void* d_rgba, uint32_t width, uint32_t height, uint32_t pitch_bytes, void* d_out)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (!params || params->struct_size != sizeof(crn_gpu_pack_params) || !d_rgba || !d_out || !width || !height ||
        pitch_bytes < width * 4u || (pitch_bytes & 3u))
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_pack_image: bad argument");
    const uint32_t bpb = crn_gpu_bytes_per_block(format);
    if (!bpb) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_pack_image: unknown format");
    if (params->dxt_quality > 4 || params->dxt1a_alpha_threshold > 255)
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_pack_image: parameter out of range");
    const bool has_color = format == CRN_GPU_FMT_DXT1 || format == CRN_GPU_FMT_DXT1A || format == CRN_GPU_FMT_DXT3 || format == CRN_GPU_FMT_DXT5;
    (void)has_color;
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));

    crn::ImageView img;
    img.rgba = static_cast<const uint8_t*>(d_rgba);
    img.width = width; img.height = height; img.pitch = pitch_bytes;
    img.blocks_x = (width + 3) >> 2; img.blocks_y = (height + 3) >> 2;
    const uint32_t total = img.blocks_x * img.blocks_y;
    uint8_t* out = static_cast<uint8_t*>(d_out);
    const int q = (int)params->dxt_quality;
    const int both = params->use_both_block_types ? 1 : 0;
    const int threads = crn::kPackWarpsPerCta * 32;
    const int grid = grid_for(ctx, total, crn::kPackWarpsPerCta, 8);

    auto launch_alpha = [&](uint32_t comp, uint32_t ofs) {
        CRN_LAUNCH(crn::pack_alpha_element_kernel, grid, threads, 0, ctx->stream, img, comp, q, both, out, bpb, ofs);
        ctx->launches++;
    };
    int color_rc = CRN_GPU_OK;
    auto launch_color = [&](uint32_t ofs) {
        crn::Dxt1Params dp;
        dp.quality = q;
        dp.perceptual = params->perceptual ? 1 : 0;
        dp.pixels_have_alpha = 0;
        // crn_dxt_image.cpp:1463-1473: 3-colour blocks only for DXT1 / DXT1A
        dp.use_alpha_blocks = (format == CRN_GPU_FMT_DXT1 || format == CRN_GPU_FMT_DXT1A) ? both : 0;
        dp.force_alpha_blocks = 0;
        dp.grayscale_sampling = params->grayscale_sampling ? 1 : 0;
        dp.parallel_sums = 0;
        dp.alpha_threshold = params->dxt1a_alpha_threshold;
        const int dxt1a = format == CRN_GPU_FMT_DXT1A;
        // five phase kernels per chunk of blocks; the per-block state lives in ctx->d_state between them
        const uint32_t chunk_cap = 1u << 18;
        const uint32_t chunk = total < chunk_cap ? total : chunk_cap;
        // try_alpha_as_black_optimization (crn_dxt1.cpp:2241-2244): only where 3-colour blocks are allowed at all
        const bool black = params->use_transparent_indices_for_black && dp.use_alpha_blocks;
        const size_t state_bytes = (((size_t)chunk * sizeof(crn::Dxt1BlockState)) + 255) & ~(size_t)255;
        color_rc = ensure(ctx, &ctx->d_state, &ctx->d_state_cap, state_bytes + (black ? (size_t)chunk * 8 : 0));
        if (color_rc) return;
        crn::Dxt1BlockState* st = static_cast<crn::Dxt1BlockState*>(ctx->d_state);
        unsigned long long* errs = black ? reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(ctx->d_state) + state_bytes) : nullptr;
        for (uint32_t first = 0; first < total; first += chunk) {
            const uint32_t count = total - first < chunk ? total - first : chunk;
            const int g = grid_for(ctx, count, crn::kPackWarpsPerCta, 8);
            for (int run = black ? 1 : 0; run <= (black ? 2 : 0); run++) {
                CRN_LAUNCH(crn::pack_color_phase_kernel<0>, g, threads, 0, ctx->stream, img, dp, dxt1a, st, first, count, out, bpb, ofs, run, errs);
                CRN_LAUNCH(crn::pack_color_phase_kernel<1>, g, threads, 0, ctx->stream, img, dp, dxt1a, st, first, count, out, bpb, ofs, run, errs);
                CRN_LAUNCH(crn::pack_color_phase_kernel<2>, g, threads, 0, ctx->stream, img, dp, dxt1a, st, first, count, out, bpb, ofs, run, errs);
                CRN_LAUNCH(crn::pack_color_phase_kernel<3>, g, threads, 0, ctx->stream, img, dp, dxt1a, st, first, count, out, bpb, ofs, run, errs);
                CRN_LAUNCH(crn::pack_color_phase_kernel<4>, g, threads, 0, ctx->stream, img, dp, dxt1a, st, first, count, out, bpb, ofs, run, errs);
                ctx->launches += 5;
            }
        }
    };

    switch (format) {
    case CRN_GPU_FMT_DXT1: case CRN_GPU_FMT_DXT1A: launch_color(0); break;
    case CRN_GPU_FMT_DXT3: {
        const int g3 = (int)((total + 255) / 256);
        CRN_LAUNCH(crn::pack_dxt3_alpha_kernel, g3 ? g3 : 1, 256, 0, ctx->stream, img, 3u, out, bpb, 0u);
        ctx->launches++;
        launch_color(8);
        break;
    }
    case CRN_GPU_FMT_DXT5: launch_alpha(3, 0); launch_color(8); break;
    case CRN_GPU_FMT_DXT5A: launch_alpha(3, 0); break;
    case CRN_GPU_FMT_DXN_XY: launch_alpha(0, 0); launch_alpha(1, 8); break;
    case CRN_GPU_FMT_DXN_YX: launch_alpha(1, 0); launch_alpha(0, 8); break;
    default: break;
    }
    if (color_rc) return color_rc;
    CRN_CUDA(ctx, cudaGetLastError());
    return CRN_GPU_OK;
}); }

int crn_gpu_pack_image_host(crn_gpu_ctx* ctx, uint32_t format, const crn_gpu_pack_params* params,
                            const void* h_rgba, uint32_t width, uint32_t height, uint32_t pitch_bytes, void* h_out)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (!h_rgba || !h_out || !width || !height) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_pack_image_host: bad argument");
    const uint32_t bpb = crn_gpu_bytes_per_block(format);
    if (!bpb) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_pack_image_host: unknown format");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t in_bytes = (size_t)pitch_bytes * height;
    const size_t out_bytes = (size_t)((width + 3) >> 2) * ((height + 3) >> 2) * bpb;
    int rc = ensure(ctx, &ctx->d_in, &ctx->d_in_cap, in_bytes);
    if (rc) return rc;
    rc = ensure(ctx, &ctx->d_out, &ctx->d_out_cap, out_bytes);
    if (rc) return rc;
    CRN_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, h_rgba, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = crn_gpu_pack_image(ctx, format, params, ctx->d_in, width, height, pitch_bytes, ctx->d_out);
    if (rc) return rc;
    CRN_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CRN_GPU_OK;
}); }


/* ---- cluster optimisers ---------------------------------------------------------------------------- */

int crn_gpu_dxt1_optimize_clusters(crn_gpu_ctx* ctx, const crn_gpu_pack_params* params, int dxt1a,
                                   const void* d_blocks_rgba, uint32_t n_blocks,
                                   const uint32_t* d_cluster_offsets, const uint32_t* d_cluster_blocks,
                                   uint32_t n_clusters, uint32_t total_member_blocks,
                                   void* d_out, uint32_t out_stride_bytes, uint32_t out_offset_bytes,
                                   uint32_t* d_cluster_endpoints, uint64_t* d_cluster_error)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (!params || params->struct_size != sizeof(crn_gpu_pack_params) || !d_blocks_rgba || !n_blocks || !d_cluster_offsets ||
        !d_cluster_blocks || !d_out || out_stride_bytes < 8 || (out_stride_bytes & 7) || (out_offset_bytes & 7))
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_dxt1_optimize_clusters: bad argument");
    if (params->dxt_quality > 4)
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_dxt1_optimize_clusters: dxt_quality out of range");
    if (!n_clusters || !total_member_blocks) return CRN_GPU_OK;
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t P = (size_t)total_member_blocks * 16;
    const size_t nb_scan = (P + 1 + 1023) / 1024;
    const size_t need = P * crn::kClusterWorkspaceBytesPerPixel + (size_t)n_clusters * 12 + nb_scan * 4 + 4096;
    int rc = ensure(ctx, &ctx->d_cluster_ws, &ctx->d_cluster_ws_cap, need);
    if (rc) return rc;
    uint8_t* base = static_cast<uint8_t*>(ctx->d_cluster_ws);
    crn::ClusterWorkspace ws;
    size_t o = 256;                                     // first 256 bytes: the work-stealing counter
    ws.hash = reinterpret_cast<crn::ClusterHashEntry*>(base + o); o += P * 2 * sizeof(crn::ClusterHashEntry);
    ws.cw = reinterpret_cast<int4*>(base + o); o += P * 16;
    ws.ce = reinterpret_cast<int4*>(base + o); o += P * 16;
    ws.ce2 = reinterpret_cast<int4*>(base + o); o += P * 16;
    ws.mark = reinterpret_cast<uint32_t*>(base + o); o += P * 4;
    uint32_t* flags = reinterpret_cast<uint32_t*>(base + o); o += (P + 2) * 4;
    uint32_t* rank = reinterpret_cast<uint32_t*>(base + o); o += (P + 2) * 4;
    uint32_t* block_sums = reinterpret_cast<uint32_t*>(base + o); o += (nb_scan + 2) * 4;
    uint32_t* transparent = reinterpret_cast<uint32_t*>(base + o); o += (size_t)n_clusters * 4;
    crn::ClusterResult* results = reinterpret_cast<crn::ClusterResult*>(base + o); o += (size_t)n_clusters * 8;
    ws.sel = base + o;
    // hash keys / first / count and the first-appearance marks must start at zero
    CRN_CUDA(ctx, cudaMemsetAsync(base, 0, 256 + P * 2 * sizeof(crn::ClusterHashEntry), ctx->stream));
    CRN_CUDA(ctx, cudaMemsetAsync(ws.mark, 0, P * 4, ctx->stream));
    CRN_CUDA(ctx, cudaMemsetAsync(transparent, 0, (size_t)n_clusters * 4, ctx->stream));
    crn::Dxt1Params dp;
    dp.quality = (int)params->dxt_quality;
    dp.perceptual = params->perceptual ? 1 : 0;
    dp.pixels_have_alpha = 0;
    dp.use_alpha_blocks = params->use_both_block_types ? 1 : 0;
    dp.force_alpha_blocks = 0;
    dp.grayscale_sampling = params->grayscale_sampling ? 1 : 0;
    dp.parallel_sums = (ctx->vq_exact || getenv("CRN_B200_ORDERED_SUMS")) ? 0 : 1;     // crn_gpu_set_vq_mode: exact keeps the reference's member order
    // qdxt1::pack (crn_qdxt1.cpp:920-923): without 3-colour blocks the alpha threshold is forced to 0
    dp.alpha_threshold = dp.use_alpha_blocks ? params->dxt1a_alpha_threshold : 0;
    const int scan_alpha = (dxt1a && dp.use_alpha_blocks) ? 1 : 0;
    const uint32_t* blocks = static_cast<const uint32_t*>(d_blocks_rgba);
    const uint32_t TP = (uint32_t)P;
    const unsigned gp = (TP + 255) / 256;
    CRN_LAUNCH(crn::cluster_hash_insert_kernel, gp, 256, 0, ctx->stream, blocks, d_cluster_offsets, d_cluster_blocks, n_clusters, TP, scan_alpha, dp.alpha_threshold, ws, transparent);
    CRN_LAUNCH(crn::cluster_mark_kernel, (2 * TP + 255) / 256, 256, 0, ctx->stream, d_cluster_offsets, n_clusters, TP, ws);
    CRN_LAUNCH(crn::cluster_first_flags_kernel, (TP + 1 + 255) / 256, 256, 0, ctx->stream, ws.mark, flags, TP);
    {   // exclusive scan of the TP + 1 flags
        const uint32_t m = TP + 1, nb = (m + 1023) / 1024;
        CRN_LAUNCH(crn::vq_scan_block_kernel, nb, 256, 0, ctx->stream, flags, rank, block_sums, m);
        ctx->launches++;
        if (nb > 1) {
            CRN_LAUNCH(crn::vq_scan_sums_kernel, 1, 256, 0, ctx->stream, block_sums, nb);
            CRN_LAUNCH(crn::vq_scan_add_kernel, (m + 255) / 256, 256, 0, ctx->stream, rank, block_sums, m);
            ctx->launches += 2;
        }
    }
    CRN_LAUNCH(crn::cluster_compact_kernel, gp, 256, 0, ctx->stream, d_cluster_offsets, n_clusters, TP, ws, rank);
    const int threads = crn::kClusterWarpsPerCta * 32;
    const int grid = grid_for(ctx, n_clusters, crn::kClusterWarpsPerCta, getenv("CRN_B200_CLUSTER_CTAS") ? atoi(getenv("CRN_B200_CLUSTER_CTAS")) : 5);
    const uint32_t n_big = (ctx->d_cluster_order && dp.quality >= 3 && !getenv("CRN_B200_NO_COOP")) ? std::min(ctx->cluster_big_count, n_clusters) : 0;
    if (n_big) {
        // large clusters: a CTA each, then the CTAs finish the small ones warp by warp (cluster_kernels.cuh)
        const unsigned want = n_big + (n_clusters - n_big + crn::kClusterCoopWarps - 1) / crn::kClusterCoopWarps;
        const int cgrid = (int)std::min<unsigned>(want, (unsigned)ctx->sm_count * (unsigned)CRN_COOP_OCC);
        CRN_LAUNCH(crn::dxt1_optimize_clusters_cta_kernel, cgrid, crn::kClusterCoopWarps * 32, 0, ctx->stream, d_cluster_offsets, n_clusters, n_big, dp, scan_alpha, ws, rank,
                   transparent, reinterpret_cast<unsigned int*>(base), results, d_cluster_endpoints, reinterpret_cast<unsigned long long*>(d_cluster_error),
                   ctx->d_cluster_flags, ctx->d_cluster_order);
    } else
    CRN_LAUNCH(crn::dxt1_optimize_clusters_kernel, grid, threads, 0, ctx->stream, d_cluster_offsets, n_clusters, dp, scan_alpha, ws, rank, transparent,
               reinterpret_cast<unsigned int*>(base), results, d_cluster_endpoints, reinterpret_cast<unsigned long long*>(d_cluster_error), ctx->d_cluster_flags, ctx->d_cluster_order);
#ifdef CRN_B200_PHASE_CLOCKS
    {
        unsigned long long clk[12];
        CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        CRN_CUDA(ctx, cudaMemcpy(clk, base + 80, sizeof(clk), cudaMemcpyDeviceToHost));
        static const char* names[12] = { "eval colours + order", "setup_common", "median4", "passes (incl. eval)", "post (incl. eval)", "combinatorial (incl. eval)", "best_selectors", "-",
                                         "coop eval cycles (owner warp) | warp mode: valid candidates", "coop batches", "warp mode: candidates alive after the first 32 colours", "warp mode: batches" };
        fprintf(stderr, "[crn_b200] cluster optimiser phase clocks (%u clusters, %u a CTA each), owner-warp SM cycles summed over clusters:\n", n_clusters, n_big);
        for (int k = 0; k < 12; k++) if (k != 7) fprintf(stderr, "[crn_b200]   %-52s %14llu\n", names[k], clk[k]);
    }
#endif
    CRN_LAUNCH(crn::cluster_write_kernel, gp, 256, 0, ctx->stream, blocks, d_cluster_offsets, d_cluster_blocks, n_clusters, TP, scan_alpha, dp.alpha_threshold, ws, transparent,
               results, static_cast<uint8_t*>(d_out), out_stride_bytes, out_offset_bytes);
    ctx->launches += 6;
    CRN_CUDA(ctx, cudaGetLastError());
    return CRN_GPU_OK;
}); }

int crn_gpu_dxt5_optimize_clusters(crn_gpu_ctx* ctx, const crn_gpu_pack_params* params, uint32_t component,
                                   const void* d_blocks_rgba, uint32_t n_blocks,
                                   const uint32_t* d_cluster_offsets, const uint32_t* d_cluster_blocks,
                                   uint32_t n_clusters, uint32_t total_member_blocks,
                                   void* d_out, uint32_t out_stride_bytes, uint32_t out_offset_bytes,
                                   uint32_t* d_cluster_endpoints, uint64_t* d_cluster_error)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (!params || params->struct_size != sizeof(crn_gpu_pack_params) || !d_blocks_rgba || !n_blocks || !d_cluster_offsets ||
        !d_cluster_blocks || !d_out || component > 3 || out_stride_bytes < 8 || (out_stride_bytes & 7) || (out_offset_bytes & 7) ||
        params->dxt_quality > 4)
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_dxt5_optimize_clusters: bad argument");
    if (!n_clusters || !total_member_blocks) return CRN_GPU_OK;
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ensure(ctx, &ctx->d_cluster_ws, &ctx->d_cluster_ws_cap, 1024);
    if (rc) return rc;
    CRN_CUDA(ctx, cudaMemsetAsync(ctx->d_cluster_ws, 0, 256, ctx->stream));
    const int threads = crn::kAlphaClusterWarps * 32;                    // a CTA per cluster, work-stealing over the size-ordered list
    const int grid = (int)std::min<uint32_t>(n_clusters, (uint32_t)ctx->sm_count * 8u);
    CRN_LAUNCH(crn::dxt5_optimize_clusters_kernel, grid, threads, 0, ctx->stream, static_cast<const uint32_t*>(d_blocks_rgba), d_cluster_offsets,
               d_cluster_blocks, n_clusters, component, (int)params->dxt_quality, params->use_both_block_types ? 1 : 0,
               reinterpret_cast<unsigned int*>(ctx->d_cluster_ws), static_cast<uint8_t*>(d_out), out_stride_bytes, out_offset_bytes,
               d_cluster_endpoints, reinterpret_cast<unsigned long long*>(d_cluster_error), ctx->d_cluster_flags, ctx->d_cluster_order);
    ctx->launches++;
    CRN_CUDA(ctx, cudaGetLastError());
    return CRN_GPU_OK;
}); }

int crn_gpu_qdxt_training(crn_gpu_ctx* ctx, uint32_t kind, uint32_t component, const void* d_blocks_rgba, uint32_t n_blocks,
                          const crn_gpu_mip_desc* mips, uint32_t num_mips, void* d_vectors, uint32_t* d_weights, uint8_t* d_chunk_encoding)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    crn::QdxtMipTable mt;
    if (kind > 1 || component > 3 || !d_blocks_rgba || !n_blocks || !d_vectors || !d_weights || !build_mip_table(mips, num_mips, n_blocks, mt))
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_qdxt_training: bad argument");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    const int threads = crn::kQdxtWarpsPerCta * 32;
    const int grid = grid_for(ctx, mt.total_chunks, crn::kQdxtWarpsPerCta, 8);
    const uint32_t* blocks = static_cast<const uint32_t*>(d_blocks_rgba);
    if (kind == 0)
        CRN_LAUNCH(crn::qdxt_training_kernel<0>, grid, threads, 0, ctx->stream, blocks, mt, component, static_cast<uint8_t*>(d_vectors), d_weights, d_chunk_encoding, (unsigned long long*)nullptr, 0);
    else
        CRN_LAUNCH(crn::qdxt_training_kernel<1>, grid, threads, 0, ctx->stream, blocks, mt, component, static_cast<uint8_t*>(d_vectors), d_weights, d_chunk_encoding, (unsigned long long*)nullptr, 0);
    ctx->launches++;
    CRN_CUDA(ctx, cudaGetLastError());
    return CRN_GPU_OK;
}); }

int crn_gpu_optimize_selectors(crn_gpu_ctx* ctx, uint32_t kind, const crn_gpu_pack_params* params, uint32_t component,
                               const void* d_blocks_rgba, uint32_t n_blocks,
                               const uint32_t* d_cluster_offsets, const uint32_t* d_cluster_blocks, uint32_t n_clusters,
                               void* d_elements, uint32_t stride_bytes, uint32_t offset_bytes)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (kind > 1 || !params || params->struct_size != sizeof(crn_gpu_pack_params) || component > 3 || !d_blocks_rgba || !n_blocks ||
        !d_cluster_offsets || !d_cluster_blocks || !d_elements || stride_bytes < 8 || (stride_bytes & 7) || (offset_bytes & 7))
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_optimize_selectors: bad argument");
    if (!n_clusters) return CRN_GPU_OK;
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    const int threads = crn::kClusterWarpsPerCta * 32;
    const int grid = grid_for(ctx, n_clusters, crn::kClusterWarpsPerCta, 8);
    const uint32_t* blocks = static_cast<const uint32_t*>(d_blocks_rgba);
    uint8_t* el = static_cast<uint8_t*>(d_elements);
    // qdxt1::pack (crn_qdxt1.cpp:920-923): without 3-colour blocks the alpha threshold is 0
    const uint32_t thr = params->use_both_block_types ? params->dxt1a_alpha_threshold : 0;
    if (kind == 0)
        CRN_LAUNCH(crn::optimize_selectors_kernel<false>, grid, threads, 0, ctx->stream, blocks, d_cluster_offsets, d_cluster_blocks, n_clusters,
                   component, params->perceptual ? 1 : 0, thr, el, stride_bytes, offset_bytes);
    else
        CRN_LAUNCH(crn::optimize_selectors_kernel<true>, grid, threads, 0, ctx->stream, blocks, d_cluster_offsets, d_cluster_blocks, n_clusters,
                   component, 0, 0u, el, stride_bytes, offset_bytes);
    ctx->launches++;
    CRN_CUDA(ctx, cudaGetLastError());
    return CRN_GPU_OK;
}); }

/* ---- vector quantiser ----------------------------------------------------------------------------- */

int crn_gpu_vq_clusterize(crn_gpu_ctx* ctx, uint32_t dims, const void* d_vectors, const uint32_t* d_weights, uint32_t n,
                          uint32_t max_codebook_size, uint32_t retrieve_max_clusters, int threaded,
                          uint32_t* h_cluster_of, uint32_t* num_clusters, uint32_t* codebook_size)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if ((dims != 2 && dims != 6 && dims != 16) || !d_vectors || !d_weights || !n || !max_codebook_size || !h_cluster_of)
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_vq_clusterize: bad argument");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint8_t* v = static_cast<const uint8_t*>(d_vectors);
    switch (dims) {
    case 2: return vq_clusterize<2>(ctx, v, d_weights, n, max_codebook_size, retrieve_max_clusters, threaded, h_cluster_of, num_clusters, codebook_size);
    case 6: return vq_clusterize<6>(ctx, v, d_weights, n, max_codebook_size, retrieve_max_clusters, threaded, h_cluster_of, num_clusters, codebook_size);
    default: return vq_clusterize<16>(ctx, v, d_weights, n, max_codebook_size, retrieve_max_clusters, threaded, h_cluster_of, num_clusters, codebook_size);
    }
}); }

/* ---- clustered DDS compression ------------------------------------------------------------------- */

int crn_gpu_qdxt_init(crn_gpu_ctx* ctx, uint32_t format, const crn_gpu_pack_params* params,
                      const crn_gpu_level_desc* levels, uint32_t num_levels, int pixels_on_host, crn_gpu_qdxt** out)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (!out || !params || params->struct_size != sizeof(crn_gpu_pack_params) || !levels || !num_levels || num_levels > (uint32_t)crn::kQdxtMaxMips)
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_qdxt_init: bad argument");
    *out = nullptr;
    if (format > CRN_GPU_FMT_DXN_YX || format == CRN_GPU_FMT_DXT3)
        return set_err(ctx, CRN_GPU_ERR_UNSUPPORTED, "crn_gpu_qdxt_init: format is not clustered (DXT3) or unknown");
    if (params->dxt_quality > 4)
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_qdxt_init: dxt_quality out of range");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    crn_gpu_qdxt* q = new (std::nothrow) crn_gpu_qdxt();
    if (!q) return set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "crn_gpu_qdxt_init: out of host memory");
    q->ctx = ctx; q->format = format; q->params = *params; q->num_levels = num_levels;
    q->bytes_per_block = crn_gpu_bytes_per_block(format);
    q->flat = params->non_hierarchical ? 1 : 0;
    q->pow_mul = format == CRN_GPU_FMT_DXT5 ? .75f : 1.0f;           // crn_mipmapped_texture.cpp:2529-2539
    q->d_blocks = nullptr; q->d_out = nullptr; q->num_elements = 0; q->d_blocks_cap = q->d_out_cap = 0;
    // element table (crn_mipmapped_texture.cpp:2319-2366)
    uint32_t ne = 0;
    auto add = [&](int kind, uint32_t comp, uint32_t offset, int use_alpha) {
        crn_qdxt_element& e = q->el[ne++];
        e.kind = kind; e.comp = comp; e.offset = offset; e.use_alpha_blocks = use_alpha;
        e.max_selector_clusters = 0; e.endpoint_clusters = e.selector_clusters = 0;
        e.ctx = nullptr; e.d_vecs = nullptr; e.d_wts = nullptr; e.d_cat = nullptr; e.d_offsets = e.d_members = e.d_ids = e.d_ep_perm = nullptr; e.d_keys = nullptr; e.rc = 0; e.ev_opt[0] = e.ev_opt[1] = nullptr; e.endpoint_opt_ms = 0;
        memset(e.caps, 0, sizeof(e.caps));
        q->num_elements = ne;
    };
    switch (format) {
    case CRN_GPU_FMT_DXT1: add(0, 3, 0, params->use_both_block_types ? 1 : 0); break;
    case CRN_GPU_FMT_DXT1A: add(0, 3, 0, 1); break;
    case CRN_GPU_FMT_DXT5: add(0, 3, 8, 0); add(1, 3, 0, 0); break;
    case CRN_GPU_FMT_DXT5A: add(1, 3, 0, 0); break;
    case CRN_GPU_FMT_DXN_XY: add(1, 0, 0, 0); add(1, 1, 8, 0); break;
    default: add(1, 1, 0, 0); add(1, 0, 8, 0); break;               // DXN_YX
    }
    q->num_elements = ne;
    uint64_t total_blocks = 0;
    for (uint32_t l = 0; l < num_levels; l++) {
        if (!levels[l].rgba || !levels[l].width || !levels[l].height || levels[l].pitch_bytes < levels[l].width * 4) {
            qdxt_release(q);
            return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_qdxt_init: bad level");
        }
        crn_gpu_mip_desc m;
        m.first_block = (uint32_t)total_blocks; m.block_width = (levels[l].width + 3) / 4; m.block_height = (levels[l].height + 3) / 4;
        q->mips.push_back(m);
        total_blocks += (uint64_t)m.block_width * m.block_height;
    }
    if (total_blocks > 0x7fffffffu / 16) { qdxt_release(q); return set_err(ctx, CRN_GPU_ERR_UNSUPPORTED, "crn_gpu_qdxt_init: too many blocks"); }
    const uint32_t n = q->n_blocks = (uint32_t)total_blocks;
    uint32_t cap = 1024;
    while (cap < 2 * n) cap <<= 1;
#define QDXT_ALLOC(ptr, bytes, cap)                                                                                    \
    do {                                                                                                               \
        cudaError_t ce_ = pool_alloc(ctx, (void**)&(ptr), (bytes), &(cap));                                            \
        if (ce_ != cudaSuccess) { (ptr) = nullptr; qdxt_release(q); return set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "crn_gpu_qdxt_init: cudaMalloc", ce_); } \
    } while (0)
    QDXT_ALLOC(q->d_blocks, (size_t)n * 64, q->d_blocks_cap);
    QDXT_ALLOC(q->d_out, (size_t)n * q->bytes_per_block, q->d_out_cap);
    for (uint32_t i = 0; i < ne; i++) {
        crn_qdxt_element& e = q->el[i];
        if (!ctx->child[i] && crn_gpu_create(ctx->device, &ctx->child[i]) != CRN_GPU_OK) { qdxt_release(q); return set_err(ctx, CRN_GPU_ERR_CUDA, "crn_gpu_qdxt_init: element stream"); }
        e.ctx = ctx->child[i];
        e.vq_exact = ctx->vq_exact;
        e.ctx->launches = 0;
        if (cudaEventCreate(&e.ev_opt[0]) != cudaSuccess || cudaEventCreate(&e.ev_opt[1]) != cudaSuccess) { qdxt_release(q); return set_err(ctx, CRN_GPU_ERR_CUDA, "crn_gpu_qdxt_init: events"); }
        QDXT_ALLOC(e.d_vecs, (size_t)n * 16, e.caps[0]);
        QDXT_ALLOC(e.d_wts, (size_t)n * 4, e.caps[1]);
        QDXT_ALLOC(e.d_cat, (size_t)n, e.caps[2]);
        QDXT_ALLOC(e.d_offsets, ((size_t)n + 1) * 4, e.caps[3]);
        QDXT_ALLOC(e.d_members, (size_t)n * 4, e.caps[4]);
        QDXT_ALLOC(e.d_ids, (size_t)n * 4, e.caps[5]);
        QDXT_ALLOC(e.d_keys, ((size_t)n + cap + 2) * 8, e.caps[6]);
        QDXT_ALLOC(e.d_ep_perm, (size_t)n * 4, e.caps[7]);
    }
#undef QDXT_ALLOC
    // pixel blocks of every level (crn_mipmapped_texture.cpp:2419-2472)
    int rc = CRN_GPU_OK;
    if (pixels_on_host) {              // one staging buffer for the whole chain: copies and gathers queue back to back
        size_t total = 0;
        for (uint32_t l = 0; l < num_levels; l++) total += (((size_t)levels[l].width * 4 * levels[l].height) + 255) & ~(size_t)255;
        rc = ensure(ctx, &ctx->d_in, &ctx->d_in_cap, total);
    }
    size_t stage = 0;
    for (uint32_t l = 0; l < num_levels && !rc; l++) {
        const crn_gpu_level_desc& lv = levels[l];
        const uint8_t* src = static_cast<const uint8_t*>(lv.rgba);
        uint32_t pitch = lv.pitch_bytes;
        if (pixels_on_host) {
            uint8_t* d = static_cast<uint8_t*>(ctx->d_in) + stage;
            stage += (((size_t)lv.width * 4 * lv.height) + 255) & ~(size_t)255;
            cudaError_t ce = cudaMemcpy2DAsync(d, (size_t)lv.width * 4, lv.rgba, lv.pitch_bytes, (size_t)lv.width * 4, lv.height, cudaMemcpyHostToDevice, ctx->stream);
            if (ce != cudaSuccess) { rc = set_err(ctx, CRN_GPU_ERR_CUDA, "crn_gpu_qdxt_init: H2D", ce); break; }
            src = d; pitch = lv.width * 4;
        }
        const uint32_t nb = q->mips[l].block_width * q->mips[l].block_height;
        CRN_LAUNCH(crn::blockify_kernel, (nb * 16 + 255) / 256, 256, 0, ctx->stream, src, lv.width, lv.height, pitch, q->d_blocks + (size_t)q->mips[l].first_block * 16);
        ctx->launches++;
    }
    if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = set_err(ctx, CRN_GPU_ERR_CUDA, "crn_gpu_qdxt_init: blockify");
    if (!rc) rc = qdxt_for_each_element(q, [q](crn_qdxt_element& e) { return qdxt_init_element(q, e); });
    if (rc) { qdxt_release(q); return rc; }
    *out = q;
    return CRN_GPU_OK;
}); }

uint64_t crn_gpu_qdxt_output_size(const crn_gpu_qdxt* q) { return q ? (uint64_t)q->n_blocks * q->bytes_per_block : 0; }

uint64_t crn_gpu_qdxt_level_offset(const crn_gpu_qdxt* q, uint32_t level)
{
    return (q && level < q->num_levels) ? (uint64_t)q->mips[level].first_block * q->bytes_per_block : 0;
}

int crn_gpu_qdxt_pack(crn_gpu_qdxt* q, uint32_t quality_level, void* dst, int dst_on_host)
{ return crn_guard(nullptr, [&]() -> int {
    if (!q) return CRN_GPU_ERR_BAD_PARAM;
    crn_gpu_ctx* ctx = q->ctx;
    if (quality_level > 255 || !dst) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_qdxt_pack: bad argument");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    const int rc = qdxt_for_each_element(q, [q, quality_level](crn_qdxt_element& e) {
        const int r = qdxt_pack_element(q, e, quality_level);
        if (!r && cudaStreamSynchronize(e.ctx->stream) != cudaSuccess) return set_err(e.ctx, CRN_GPU_ERR_CUDA, "crn_gpu_qdxt_pack: element stream");
        if (!r) cudaEventElapsedTime(&e.endpoint_opt_ms, e.ev_opt[0], e.ev_opt[1]);
        e.opt_stats[0] = e.opt_stats[1] = e.opt_stats[2] = 0;
        if (!r && e.kind == 0 && e.ctx->d_cluster_ws) {      // work counters the colour optimiser left behind its work-stealing counter
            uint64_t st[2] = { 0, 0 };
            if (cudaMemcpy(st, static_cast<uint8_t*>(e.ctx->d_cluster_ws) + 64, 16, cudaMemcpyDeviceToHost) == cudaSuccess) {
                e.opt_stats[0] = st[0]; e.opt_stats[1] = st[1];
                e.opt_stats[2] = e.use_alpha_blocks ? 7 : 4;     // 4-colour palette, plus the 3-colour one where both block types are tried
            }
        }
        return r;
    });
    if (rc) return rc;
    CRN_CUDA(ctx, cudaMemcpyAsync(dst, q->d_out, (size_t)q->n_blocks * q->bytes_per_block, dst_on_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
    CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CRN_GPU_OK;
}); }

int crn_gpu_qdxt_get_info(const crn_gpu_qdxt* q, crn_gpu_qdxt_info* info)
{ return crn_guard(nullptr, [&]() -> int {
    if (!q || !info || info->struct_size != sizeof(crn_gpu_qdxt_info)) return CRN_GPU_ERR_BAD_PARAM;
    memset(info, 0, sizeof(*info));
    info->struct_size = sizeof(*info);
    info->n_blocks = q->n_blocks; info->num_elements = q->num_elements;
    for (uint32_t i = 0; i < q->num_elements; i++) {
        info->endpoint_codebook_size[i] = q->el[i].endpoint_tree.codebook_size();
        info->max_selector_clusters[i] = q->el[i].max_selector_clusters;
        info->endpoint_clusters[i] = q->el[i].endpoint_clusters;
        info->selector_clusters[i] = q->el[i].selector_clusters;
        info->endpoint_opt_ms[i] = q->el[i].endpoint_opt_ms;
        info->opt_candidates[i] = q->el[i].opt_stats[0]; info->opt_colour_evals[i] = q->el[i].opt_stats[1]; info->opt_palette_entries[i] = (uint32_t)q->el[i].opt_stats[2];
    }
    return CRN_GPU_OK;
}); }

void crn_gpu_qdxt_free(crn_gpu_qdxt* q)
{
    if (!q) return;
    cudaSetDevice(q->ctx->device);
    cudaStreamSynchronize(q->ctx->stream);
    qdxt_release(q);
}

/* ---- CRN -> DXTn transcoding --------------------------------------------------------------------- */

struct crn_gpu_texture {
    crn_gpu_ctx* ctx;
    crn::CrnHeaderInfo hdr;
    uint32_t bytes_per_block;
    void* slab;                       // one device allocation holding everything below
    crn::TranscodeFile host_file;     // host mirror; device copy at d_file
    crn::TranscodeFile* d_file;
    uint32_t rowbuf_ofs[16];
    uint64_t level_ofs[16];           // tight layout offsets (face 0)
    uint64_t level_face_size[16];
    uint64_t total_size;
};

int crn_gpu_refine_endpoints(crn_gpu_ctx* ctx, int dxt1_selectors, int perceptual, uint32_t component,
                             const void* d_pixels_rgba, const uint8_t* d_selectors, const uint32_t* d_offsets, uint32_t n_clusters,
                             const uint64_t* d_error_to_beat, uint32_t* d_endpoints, uint64_t* d_error, uint8_t* d_ok)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (!d_pixels_rgba || !d_selectors || !d_offsets || !d_endpoints || !d_error || !d_ok || component > 3)
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_refine_endpoints: bad argument");
    if (!n_clusters) return CRN_GPU_OK;
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    const int grid = grid_for(ctx, n_clusters, crn::kRefineWarpsPerCta, 8);
    CRN_LAUNCH(crn::refine_endpoints_kernel, grid, crn::kRefineWarpsPerCta * 32, 0, ctx->stream, static_cast<const uint32_t*>(d_pixels_rgba), d_selectors, d_offsets,
               n_clusters, dxt1_selectors ? 1 : 0, perceptual ? 1 : 0, component, reinterpret_cast<const unsigned long long*>(d_error_to_beat),
               d_endpoints, reinterpret_cast<unsigned long long*>(d_error), d_ok, ctx->refine_parallel ? 1 : 0);
    ctx->launches++;
    CRN_CUDA(ctx, cudaGetLastError());
    return CRN_GPU_OK;
}); }

int crn_gpu_nearest_codebook(crn_gpu_ctx* ctx, uint32_t dims, const float* d_vectors, uint32_t n, const float* d_codebook, uint32_t codebook_size, uint32_t* d_out)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if ((dims != 2 && dims != 6) || !d_vectors || !d_codebook || !d_out || !codebook_size)
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_nearest_codebook: bad argument");
    if (!n) return CRN_GPU_OK;
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    const unsigned grid = (n + crn::kNearestThreads - 1) / crn::kNearestThreads;
    if (dims == 6) CRN_LAUNCH(crn::nearest_codebook_kernel<6>, grid, crn::kNearestThreads, 0, ctx->stream, d_vectors, n, d_codebook, codebook_size, d_out);
    else CRN_LAUNCH(crn::nearest_codebook_kernel<2>, grid, crn::kNearestThreads, 0, ctx->stream, d_vectors, n, d_codebook, codebook_size, d_out);
    ctx->launches++;
    CRN_CUDA(ctx, cudaGetLastError());
    return CRN_GPU_OK;
}); }

int crn_gpu_assign_selectors(crn_gpu_ctx* ctx, uint32_t kind, int perceptual, uint32_t component,
                             const void* d_blocks_rgba, uint32_t n_blocks, const void* d_block_values, const void* d_block_values_accum,
                             const uint64_t* d_codebook, uint32_t codebook_size,
                             uint32_t* d_best_index, uint64_t* d_refined_codebook, uint8_t* d_used)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (kind > 1 || component > 3 || !d_blocks_rgba || !d_block_values || !d_codebook || !codebook_size || !d_best_index || !d_refined_codebook || !d_used)
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_assign_selectors: bad argument");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t tab_bytes = (size_t)codebook_size * 16 * (kind ? 8 : 4) * sizeof(uint32_t);
    int rc = ensure(ctx, &ctx->d_cluster_ws, &ctx->d_cluster_ws_cap, tab_bytes);
    if (rc) return rc;
    uint32_t* tot = static_cast<uint32_t*>(ctx->d_cluster_ws);
    CRN_CUDA(ctx, cudaMemsetAsync(tot, 0, tab_bytes, ctx->stream));
    CRN_CUDA(ctx, cudaMemsetAsync(d_used, 0, codebook_size, ctx->stream));
    const unsigned long long* cb = reinterpret_cast<const unsigned long long*>(d_codebook);
    unsigned long long* refined = reinterpret_cast<unsigned long long*>(d_refined_codebook);
    if (n_blocks) {
        const int grid = grid_for(ctx, n_blocks, crn::kAssignWarpsPerCta, 8);
        if (kind == 0)
            CRN_LAUNCH(crn::assign_selectors_kernel<0>, grid, crn::kAssignWarpsPerCta * 32, 0, ctx->stream, static_cast<const uint32_t*>(d_blocks_rgba), n_blocks,
                       static_cast<const uint8_t*>(d_block_values), static_cast<const uint8_t*>(nullptr), cb, codebook_size, perceptual ? 1 : 0, component, d_best_index, tot, d_used);
        else
            CRN_LAUNCH(crn::assign_selectors_kernel<1>, grid, crn::kAssignWarpsPerCta * 32, 0, ctx->stream, static_cast<const uint32_t*>(d_blocks_rgba), n_blocks,
                       static_cast<const uint8_t*>(d_block_values), static_cast<const uint8_t*>(d_block_values_accum), cb, codebook_size, 0, component, d_best_index, tot, d_used);
        ctx->launches++;
    }
    if (kind == 0) CRN_LAUNCH(crn::revote_selectors_kernel<0>, (codebook_size + 255) / 256, 256, 0, ctx->stream, tot, codebook_size, refined);
    else CRN_LAUNCH(crn::revote_selectors_kernel<1>, (codebook_size + 255) / 256, 256, 0, ctx->stream, tot, codebook_size, refined);
    ctx->launches++;
    CRN_CUDA(ctx, cudaGetLastError());
    return CRN_GPU_OK;
}); }

int crn_gpu_unpack_image(crn_gpu_ctx* ctx, uint32_t format, const void* d_blocks, uint32_t width, uint32_t height, void* d_rgba, uint32_t pitch_bytes)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (!d_blocks || !d_rgba || !width || !height || pitch_bytes < width * 4u || (pitch_bytes & 3u) || !crn_gpu_bytes_per_block(format))
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_unpack_image: bad argument");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t bxn = (width + 3) >> 2, byn = (height + 3) >> 2, total = bxn * byn;
    CRN_LAUNCH(crn::unpack_blocks_kernel, (total + 255) / 256, 256, 0, ctx->stream, static_cast<const unsigned long long*>(d_blocks), format, width, height, bxn, total,
               static_cast<uint8_t*>(d_rgba), pitch_bytes);
    ctx->launches++;
    CRN_CUDA(ctx, cudaGetLastError());
    return CRN_GPU_OK;
}); }

int crn_gpu_unpack_image_host(crn_gpu_ctx* ctx, uint32_t format, const void* h_blocks, uint32_t width, uint32_t height, void* h_rgba, uint32_t pitch_bytes)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    const uint32_t bpb = crn_gpu_bytes_per_block(format);
    if (!h_blocks || !h_rgba || !width || !height || pitch_bytes < width * 4u || (pitch_bytes & 3u) || !bpb)
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_unpack_image_host: bad argument");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t in_bytes = (size_t)((width + 3) >> 2) * ((height + 3) >> 2) * bpb, out_bytes = (size_t)pitch_bytes * height;
    int rc = ensure(ctx, &ctx->d_in, &ctx->d_in_cap, in_bytes);
    if (rc) return rc;
    rc = ensure(ctx, &ctx->d_out, &ctx->d_out_cap, out_bytes);
    if (rc) return rc;
    CRN_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, h_blocks, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = crn_gpu_unpack_image(ctx, format, ctx->d_in, width, height, ctx->d_out, pitch_bytes);
    if (rc) return rc;
    CRN_CUDA(ctx, cudaMemcpyAsync(h_rgba, ctx->d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CRN_GPU_OK;
}); }

void crn_gpu_default_resample_params(crn_gpu_resample_params* p)
{   // what crn_mipmap_params (inc/crnlib.h:471-574) + create_texture_mipmaps (crnlib/crn_texture_comp.cpp:552-566) hand to generate_mipmaps
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->struct_size = sizeof(*p);
    p->filter = 4; p->filter_scale = 0.9f; p->srgb = 1; p->source_gamma = 2.2f; p->wrapping = 0; p->num_comps = 4;
}

namespace {
struct MipPlan {                 // one (src, dst) size pair: host contributor lists, then their device copies + the intermediate image
    uint32_t dw = 0, dh = 0;
    bool ok = false, same_xy = false;
    MipContribs cx, cy;
    HcBuf off_x, pix_x, wgt_x, off_y, pix_y, wgt_y, tmp;
};
void mip_plan_build(MipPlan* P, const crn_gpu_resample_params* prm, uint32_t sw, uint32_t sh)
{
    const MipFilter& F = g_mip_filters[prm->filter];
    P->same_xy = sw == sh && P->dw == P->dh;                // square level of a square image: one list serves both axes
    P->ok = mip_make_clist((int)sw, (int)P->dw, prm->wrapping != 0, F, prm->filter_scale, P->cx) &&
            (P->same_xy || mip_make_clist((int)sh, (int)P->dh, prm->wrapping != 0, F, prm->filter_scale, P->cy));
}
struct MipTablesDev { HcBuf to_linear, to_srgb; };
int mip_tables_upload(crn_gpu_ctx* ctx, const crn_gpu_resample_params* prm, MipTablesDev& T)
{
    HC_ALLOC(T.to_linear, 256 * 4); HC_ALLOC(T.to_srgb, 8192);
    if (prm->srgb) {   // crn_image_utils.cpp:686-712
        float to_linear[256]; uint8_t to_srgb[8192];
        const float source_gamma = prm->source_gamma;
        for (int i = 0; i < 256; ++i) to_linear[i] = (float)pow(i * 1.0f / 255.0f, source_gamma);
        const float inv_size = 1.0f / 8192, inv_gamma = 1.0f / source_gamma;
        for (int i = 0; i < 8192; ++i) {
            int k = (int)(255.0f * pow(i * inv_size, inv_gamma) + .5f);
            to_srgb[i] = (uint8_t)(k < 0 ? 0 : (k > 255 ? 255 : k));
        }
        CRN_CUDA(ctx, cudaMemcpyAsync(T.to_linear.p, to_linear, sizeof(to_linear), cudaMemcpyHostToDevice, ctx->stream));
        CRN_CUDA(ctx, cudaMemcpyAsync(T.to_srgb.p, to_srgb, sizeof(to_srgb), cudaMemcpyHostToDevice, ctx->stream));
        CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // the tables live on this stack frame
    }
    return CRN_GPU_OK;
}
// uploads the plan and enqueues the two passes; the caller keeps the plan alive until the stream is drained
int mip_plan_launch(crn_gpu_ctx* ctx, const crn_gpu_resample_params* prm, MipPlan& P, const MipTablesDev& T, const void* d_src, uint32_t sw, uint32_t sh, uint32_t spitch,
                    void* d_dst, uint32_t dpitch)
{
    cudaStream_t st = ctx->stream;
    if (!P.ok) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_resample: could not build the contributor lists");
    const uint32_t dw = P.dw, dh = P.dh;
    const MipContribs& cy = P.same_xy ? P.cx : P.cy;
    HC_ALLOC(P.off_x, P.cx.off.size() * 4); HC_ALLOC(P.pix_x, P.cx.pix.size() * 4); HC_ALLOC(P.wgt_x, P.cx.wgt.size() * 4);
    HC_ALLOC(P.tmp, (size_t)sh * dw * 16);
    CRN_CUDA(ctx, cudaMemcpyAsync(P.off_x.p, P.cx.off.data(), P.cx.off.size() * 4, cudaMemcpyHostToDevice, st));
    CRN_CUDA(ctx, cudaMemcpyAsync(P.pix_x.p, P.cx.pix.data(), P.cx.pix.size() * 4, cudaMemcpyHostToDevice, st));
    CRN_CUDA(ctx, cudaMemcpyAsync(P.wgt_x.p, P.cx.wgt.data(), P.cx.wgt.size() * 4, cudaMemcpyHostToDevice, st));
    const uint32_t* d_off_y = P.off_x.as<uint32_t>(); const uint32_t* d_pix_y = P.pix_x.as<uint32_t>(); const float* d_wgt_y = P.wgt_x.as<float>();
    if (!P.same_xy) {
        HC_ALLOC(P.off_y, cy.off.size() * 4); HC_ALLOC(P.pix_y, cy.pix.size() * 4); HC_ALLOC(P.wgt_y, cy.wgt.size() * 4);
        CRN_CUDA(ctx, cudaMemcpyAsync(P.off_y.p, cy.off.data(), cy.off.size() * 4, cudaMemcpyHostToDevice, st));
        CRN_CUDA(ctx, cudaMemcpyAsync(P.pix_y.p, cy.pix.data(), cy.pix.size() * 4, cudaMemcpyHostToDevice, st));
        CRN_CUDA(ctx, cudaMemcpyAsync(P.wgt_y.p, cy.wgt.data(), cy.wgt.size() * 4, cudaMemcpyHostToDevice, st));
        d_off_y = P.off_y.as<uint32_t>(); d_pix_y = P.pix_y.as<uint32_t>(); d_wgt_y = P.wgt_y.as<float>();
    }
    const int nc = (int)prm->num_comps, srgb = prm->srgb ? 1 : 0;
    const size_t tot_x = (size_t)sh * dw * 4, tot_y = (size_t)dh * dw * 4;
    CRN_LAUNCH(crn::mip_resample_x_kernel, (unsigned)((tot_x + 255) / 256), 256, 0, st, static_cast<const uint8_t*>(d_src), spitch, sh, dw, nc, srgb,
               P.off_x.as<uint32_t>(), P.pix_x.as<uint32_t>(), P.wgt_x.as<float>(), T.to_linear.as<float>(), P.tmp.as<float>());
    CRN_LAUNCH(crn::mip_resample_y_kernel, (unsigned)((tot_y + 255) / 256), 256, 0, st, P.tmp.as<float>(), dw, dh, nc, srgb,
               d_off_y, d_pix_y, d_wgt_y, T.to_srgb.as<uint8_t>(), static_cast<uint8_t*>(d_dst), dpitch);
    ctx->launches += 2;
    CRN_CUDA(ctx, cudaGetLastError());
    return CRN_GPU_OK;
}
int mip_resample(crn_gpu_ctx* ctx, const crn_gpu_resample_params* prm, const void* d_src, uint32_t sw, uint32_t sh, uint32_t spitch,
                 void* d_dst, uint32_t dw, uint32_t dh, uint32_t dpitch)
{
    if (sw == dw && sh == dh) {   // dst = src (crn_image_utils.cpp:693-697)
        CRN_CUDA(ctx, cudaMemcpy2DAsync(d_dst, dpitch, d_src, spitch, (size_t)sw * 4, sh, cudaMemcpyDeviceToDevice, ctx->stream));
        return CRN_GPU_OK;
    }
    MipPlan P; P.dw = dw; P.dh = dh;
    mip_plan_build(&P, prm, sw, sh);
    MipTablesDev T;
    HC_RC(mip_tables_upload(ctx, prm, T));
    HC_RC(mip_plan_launch(ctx, prm, P, T, d_src, sw, sh, spitch, d_dst, dpitch));
    CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));         // the plan's host vectors and pooled buffers die here
    return CRN_GPU_OK;
}
bool mip_params_ok(const crn_gpu_resample_params* p)
{
    return p && p->struct_size == sizeof(crn_gpu_resample_params) && p->filter < 5 && p->filter_scale > 0.0f && (p->num_comps == 3 || p->num_comps == 4) &&
           (!p->srgb || p->source_gamma > 0.0f);
}
}  // namespace

int crn_gpu_resample(crn_gpu_ctx* ctx, const crn_gpu_resample_params* params, const void* d_src, uint32_t src_width, uint32_t src_height, uint32_t src_pitch_bytes,
                     void* d_dst, uint32_t dst_width, uint32_t dst_height, uint32_t dst_pitch_bytes)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (!mip_params_ok(params) || !d_src || !d_dst || !src_width || !src_height || !dst_width || !dst_height || src_width > 16384 || src_height > 16384 ||
        dst_width > 16384 || dst_height > 16384 || src_pitch_bytes < src_width * 4u || dst_pitch_bytes < dst_width * 4u || (src_pitch_bytes & 3u) || (dst_pitch_bytes & 3u))
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_resample: bad argument");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    return mip_resample(ctx, params, d_src, src_width, src_height, src_pitch_bytes, d_dst, dst_width, dst_height, dst_pitch_bytes);
}); }

uint32_t crn_gpu_mip_level_count(uint32_t width, uint32_t height, uint32_t min_mip_size, uint32_t max_levels)
{   // mipmapped_texture::generate_mipmaps, crn_mipmapped_texture.cpp:2145-2157
    uint32_t n = 1;
    while (width > min_mip_size || height > min_mip_size) { width >>= 1; height >>= 1; n++; }
    if (max_levels > 0 && n > max_levels) n = max_levels;
    return n;
}

int crn_gpu_generate_mipmaps(crn_gpu_ctx* ctx, const crn_gpu_resample_params* params, const void* d_level0, uint32_t width, uint32_t height, uint32_t pitch_bytes,
                             uint32_t min_mip_size, uint32_t max_levels, void* d_mips, uint64_t capacity, uint32_t* num_levels)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (!mip_params_ok(params) || !d_level0 || !width || !height || width > 16384 || height > 16384 || pitch_bytes < width * 4u || (pitch_bytes & 3u) || !min_mip_size)
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_generate_mipmaps: bad argument");
    const uint32_t n = crn_gpu_mip_level_count(width, height, min_mip_size, max_levels);
    uint64_t need = 0;
    for (uint32_t l = 1; l < n; l++) need += (uint64_t)std::max(1u, width >> l) * std::max(1u, height >> l) * 4;
    if (num_levels) *num_levels = n;
    if (need && (!d_mips || capacity < need)) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_generate_mipmaps: output buffer too small");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    // every level is resampled from level 0 (:2171-2196).  The contributor lists of all levels are built concurrently on host
    // threads (a few hundred thousand libm calls each), then the levels are enqueued back to back and the stream drained once.
    std::vector<MipPlan> plans(n > 1 ? n - 1 : 0);
    for (uint32_t l = 1; l < n; l++) { plans[l - 1].dw = std::max(1u, width >> l); plans[l - 1].dh = std::max(1u, height >> l); }
#ifdef __CUDACC__
    {
        std::vector<std::thread> th;
        for (uint32_t l = 1; l < n; l++) th.emplace_back(mip_plan_build, &plans[l - 1], params, width, height);
        for (auto& t : th) t.join();
    }
#else
    for (uint32_t l = 1; l < n; l++) mip_plan_build(&plans[l - 1], params, width, height);
#endif
    MipTablesDev T;
    HC_RC(mip_tables_upload(ctx, params, T));
    uint8_t* out = static_cast<uint8_t*>(d_mips);
    for (uint32_t l = 1; l < n; l++) {
        MipPlan& P = plans[l - 1];
        HC_RC(mip_plan_launch(ctx, params, P, T, d_level0, width, height, pitch_bytes, out, P.dw * 4));
        if (params->renormalize) HC_RC(crn_gpu_convert_pixels(ctx, out, P.dw, P.dh, P.dw * 4, crn::kConvRenormNormalMap));   // :2207-2208
        out += (size_t)P.dw * P.dh * 4;
    }
    CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CRN_GPU_OK;
}); }

int crn_gpu_generate_mipmaps_host(crn_gpu_ctx* ctx, const crn_gpu_resample_params* params, const void* h_level0, uint32_t width, uint32_t height, uint32_t pitch_bytes,
                                  uint32_t min_mip_size, uint32_t max_levels, void* h_mips, uint64_t capacity, uint32_t* num_levels)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (!h_level0 || !width || !height || pitch_bytes < width * 4u || !min_mip_size) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_generate_mipmaps_host: bad argument");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t n = crn_gpu_mip_level_count(width, height, min_mip_size, max_levels);
    uint64_t need = 0;
    for (uint32_t l = 1; l < n; l++) need += (uint64_t)std::max(1u, width >> l) * std::max(1u, height >> l) * 4;
    if (need && (!h_mips || capacity < need)) { if (num_levels) *num_levels = n; return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_generate_mipmaps_host: output buffer too small"); }
    HcBuf d_in, d_out;
    HC_ALLOC(d_in, (size_t)pitch_bytes * height); HC_ALLOC(d_out, need ? need : 256);
    CRN_CUDA(ctx, cudaMemcpyAsync(d_in.p, h_level0, (size_t)pitch_bytes * height, cudaMemcpyHostToDevice, ctx->stream));
    int rc = crn_gpu_generate_mipmaps(ctx, params, d_in.p, width, height, pitch_bytes, min_mip_size, max_levels, d_out.p, need, num_levels);
    if (rc) return rc;
    if (need) CRN_CUDA(ctx, cudaMemcpyAsync(h_mips, d_out.p, need, cudaMemcpyDeviceToHost, ctx->stream));
    CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CRN_GPU_OK;
}); }

int crn_gpu_blockify(crn_gpu_ctx* ctx, const void* d_rgba, uint32_t width, uint32_t height, uint32_t pitch_bytes, uint32_t pad_pixels, void* d_blocks,
                     uint32_t* blocks_x, uint32_t* blocks_y)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (!d_rgba || !d_blocks || !width || !height || pitch_bytes < width * 4u || (pitch_bytes & 3u) || (pad_pixels != 4 && pad_pixels != 8))
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_blockify: bad argument");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t bw = ((width + pad_pixels - 1) / pad_pixels * pad_pixels) >> 2, bh = ((height + pad_pixels - 1) / pad_pixels * pad_pixels) >> 2;
    if (blocks_x) *blocks_x = bw;
    if (blocks_y) *blocks_y = bh;
    const size_t total = (size_t)bw * bh * 16;
    CRN_LAUNCH(crn::blockify_padded_kernel, (unsigned)((total + 255) / 256), 256, 0, ctx->stream, static_cast<const uint8_t*>(d_rgba), width, height, pitch_bytes, pad_pixels,
               static_cast<uint32_t*>(d_blocks));
    ctx->launches++;
    CRN_CUDA(ctx, cudaGetLastError());
    return CRN_GPU_OK;
}); }

void crn_gpu_default_hc_params(crn_gpu_hc_params* p)
{   // dxt_hc::params::params() (crnlib/crn_dxt_hc.h:105-131)
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->struct_size = sizeof(*p);
    p->format = CRN_GPU_FMT_DXT1; p->num_faces = 1; p->perceptual = 1;
    p->color_endpoint_codebook_size = p->color_selector_codebook_size = p->alpha_endpoint_codebook_size = p->alpha_selector_codebook_size = 3072;
    p->adaptive_tile_color_psnr_derating = 2.0f; p->adaptive_tile_alpha_psnr_derating = 2.0f; p->adaptive_tile_color_alpha_weighting_ratio = 3.0f;
    p->alpha_component_indices[0] = 3; p->alpha_component_indices[1] = 0;
    p->shard_rank = 0; p->shard_count = 1; p->exchange = nullptr; p->exchange_user = nullptr;
}

}  // extern "C"
// crn_gpu_hc_compress with the quality-independent state (tile pass, training sets) kept in *prep across calls on the same blocks
static int hc_compress_prepared(crn_gpu_ctx* ctx, const crn_gpu_hc_params* params, const void* blocks_rgba, int blocks_on_host, crn_gpu_hc** out, HcPrepared* prep)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (out) *out = nullptr;
    if (!params || params->struct_size != sizeof(crn_gpu_hc_params) || !blocks_rgba || !out || !params->num_blocks || !params->num_levels ||
        params->num_levels > 16 || !params->num_faces || params->alpha_component_indices[0] > 3 || params->alpha_component_indices[1] > 3 ||
        !params->color_endpoint_codebook_size || !params->color_selector_codebook_size || !params->alpha_endpoint_codebook_size || !params->alpha_selector_codebook_size ||
        params->color_endpoint_codebook_size > 65535 || params->color_selector_codebook_size > 65535 || params->alpha_endpoint_codebook_size > 65535 ||
        params->alpha_selector_codebook_size > 65535 || !params->shard_count || params->shard_rank >= params->shard_count ||
        (params->shard_count > 1 && !params->exchange))
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_hc_compress: bad argument");
    crn_gpu_hc* H = new (std::nothrow) crn_gpu_hc();
    if (!H) return set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "crn_gpu_hc_compress: out of host memory");
    memset(&H->info, 0, sizeof(H->info));
    H->info.struct_size = sizeof(H->info);
    int rc;
    try { rc = hc_compress_impl(ctx, params, blocks_rgba, blocks_on_host, H, prep); }
    catch (const std::bad_alloc&) { rc = set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "crn_gpu_hc_compress: out of host memory"); }
    ctx->d_cluster_flags = nullptr; ctx->d_cluster_order = nullptr;
    if (rc) { delete H; return rc; }
    *out = H;
    return CRN_GPU_OK;
}); }

extern "C" {
int crn_gpu_hc_compress(crn_gpu_ctx* ctx, const crn_gpu_hc_params* params, const void* blocks_rgba, int blocks_on_host, crn_gpu_hc** out)
{
    return hc_compress_prepared(ctx, params, blocks_rgba, blocks_on_host, out, nullptr);
}

int crn_gpu_hc_get_info(const crn_gpu_hc* hc, crn_gpu_hc_info* info)
{ return crn_guard(nullptr, [&]() -> int {
    if (!hc || !info || info->struct_size != sizeof(crn_gpu_hc_info)) return CRN_GPU_ERR_BAD_PARAM;
    *info = hc->info;
    return CRN_GPU_OK;
}); }
const uint16_t* crn_gpu_hc_endpoint_indices(const crn_gpu_hc* hc) { return hc ? hc->endpoint_indices.data() : nullptr; }
const uint16_t* crn_gpu_hc_selector_indices(const crn_gpu_hc* hc) { return hc ? hc->selector_indices.data() : nullptr; }
const uint32_t* crn_gpu_hc_color_endpoints(const crn_gpu_hc* hc) { return hc ? hc->color_endpoints.data() : nullptr; }
const uint32_t* crn_gpu_hc_alpha_endpoints(const crn_gpu_hc* hc) { return hc ? hc->alpha_endpoints.data() : nullptr; }
const uint32_t* crn_gpu_hc_color_selectors(const crn_gpu_hc* hc) { return hc ? hc->color_selectors.data() : nullptr; }
const uint64_t* crn_gpu_hc_alpha_selectors(const crn_gpu_hc* hc) { return hc ? hc->alpha_selectors.data() : nullptr; }
const uint8_t* crn_gpu_hc_block_encodings(const crn_gpu_hc* hc) { return hc ? hc->block_encodings.data() : nullptr; }
const uint32_t* crn_gpu_hc_tile_indices(const crn_gpu_hc* hc) { return hc ? hc->tile_indices.data() : nullptr; }
void crn_gpu_hc_free(crn_gpu_hc* hc) { delete hc; }

/* ---- .CRN writer back-end (crn_writer.h) ------------------------------------------------------------------- */

void crn_gpu_default_crn_params(crn_gpu_crn_params* p)
{   // crn_comp_params::clear() (inc/crnlib.h:241-285)
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->struct_size = sizeof(*p);
    p->crn_format = 0; p->levels = 1; p->faces = 1;
    p->quality_level = 255; p->perceptual = 1; p->alpha_component = 3;
    p->adaptive_tile_color_psnr_derating = 2.0f; p->adaptive_tile_alpha_psnr_derating = 2.0f;
}

static bool crn_params_ok(const crn_gpu_crn_params* p)
{
    return p && p->struct_size == sizeof(crn_gpu_crn_params) && p->width >= 1 && p->height >= 1 && p->width <= 4096 && p->height <= 4096 &&   // cCRNMaxLevelResolution
           p->levels >= 1 && p->levels <= 16 && (p->faces == 1 || p->faces == 6) && p->quality_level <= 255 && p->alpha_component <= 3;
}

int crn_gpu_crn_hc_params(const crn_gpu_crn_params* p, crn_gpu_hc_params* hp)
{ return crn_guard(nullptr, [&]() -> int {
    if (!crn_params_ok(p) || !hp) return CRN_GPU_ERR_BAD_PARAM;
    crn_gpu_default_hc_params(hp);
    hp->perceptual = p->perceptual ? 1 : 0;
    hp->adaptive_tile_color_psnr_derating = p->adaptive_tile_color_psnr_derating;
    hp->adaptive_tile_alpha_psnr_derating = p->adaptive_tile_alpha_psnr_derating;
    float color_mul = 1.0f;
    float alpha_mul = 1.0f;
    switch (p->crn_format) {                                     // crn_comp.cpp:594-660
    case 0: hp->format = CRN_GPU_FMT_DXT1; break;
    case 2: hp->format = CRN_GPU_FMT_DXT5; hp->alpha_component_indices[0] = p->alpha_component; color_mul = .75f; break;
    case 3:                                                      // DXT5_CCxY: luma in alpha, chroma in red / green (crn_comp.cpp:553-558, :614-625)
        hp->format = CRN_GPU_FMT_DXT5; hp->alpha_component_indices[0] = 3; hp->perceptual = 0; color_mul = 3.5f; alpha_mul = .35f;
        hp->adaptive_tile_color_alpha_weighting_ratio = 1.5f; break;
    case 4: case 5: case 6:                                      // DXT5_xGxR / _xGBR / _AGBR (:626-636)
        hp->format = CRN_GPU_FMT_DXT5; hp->alpha_component_indices[0] = 3; hp->perceptual = 0; break;
    case 7: hp->format = CRN_GPU_FMT_DXN_XY; hp->alpha_component_indices[0] = 0; hp->alpha_component_indices[1] = 1; hp->perceptual = 0; break;
    case 8: hp->format = CRN_GPU_FMT_DXN_YX; hp->alpha_component_indices[0] = 1; hp->alpha_component_indices[1] = 0; hp->perceptual = 0; break;
    case 9: hp->format = CRN_GPU_FMT_DXT5A; hp->alpha_component_indices[0] = p->alpha_component; hp->perceptual = 0; break;
    default: return CRN_GPU_ERR_UNSUPPORTED;                     // DXT3 is refused by the reference too; ETC is not built
    }
    auto clampu = [](uint32_t v, uint32_t lo, uint32_t hi) { return v < lo ? lo : (v > hi ? hi : v); };
    const uint32_t kMin = 8, kMax = 8192;                        // cCRNMinPaletteSize / cCRNMaxPaletteSize
    if (p->palette_sizes[0] && p->palette_sizes[1] && p->palette_sizes[2] && p->palette_sizes[3]) {
        hp->color_endpoint_codebook_size = clampu(p->palette_sizes[0], kMin, kMax); hp->color_selector_codebook_size = clampu(p->palette_sizes[1], kMin, kMax);
        hp->alpha_endpoint_codebook_size = clampu(p->palette_sizes[2], kMin, kMax); hp->alpha_selector_codebook_size = clampu(p->palette_sizes[3], kMin, kMax);
    } else {                                                     // crn_comp.cpp:539-575
        const uint32_t max_entries = clampu(((p->width + 3) / 4) * ((p->height + 3) / 4), kMin, kMax);
        if (p->crn_format == 3) hp->adaptive_tile_color_psnr_derating = 5.0f;    // only with derived palette sizes (crn_comp.cpp:553-558)
        float quality = (float)p->quality_level / 255;
        quality = quality < 0.0f ? 0.0f : (quality > 1.0f ? 1.0f : quality);
        auto size_for = [&](float floor_entries, float power) {
            const float q = powf(quality, power);
            const float a = floor_entries > (float)kMin ? floor_entries : (float)kMin;
            const float v = .5f + (a + ((float)max_entries - a) * q);
            return clampu((uint32_t)v, kMin, kMax);
        };
        hp->color_endpoint_codebook_size = size_for(64, 1.8f * color_mul);
        hp->color_selector_codebook_size = size_for(96, 1.65f * color_mul);
        hp->alpha_endpoint_codebook_size = size_for(24, 2.1f * alpha_mul);
        hp->alpha_selector_codebook_size = size_for(48, 1.65f * alpha_mul);
    }
    hp->num_levels = p->levels; hp->num_faces = p->faces;
    uint32_t total = 0;
    for (uint32_t l = 0; l < p->levels; l++) {                   // crn_comp.cpp:458-466, :706-712
        const uint32_t w = std::max(1u, p->width >> l), h = std::max(1u, p->height >> l);
        hp->levels[l].block_width = ((w + 7) & ~7u) >> 2;
        hp->levels[l].first_block = total;
        hp->levels[l].num_blocks = p->faces * hp->levels[l].block_width * (((h + 7) & ~7u) >> 2);
        hp->levels[l].weight = std::min(12.0f, powf(1.3f, (float)l));
        total += hp->levels[l].num_blocks;
    }
    hp->num_blocks = total;
    return CRN_GPU_OK;
}); }

}  // extern "C"
// The colour palette orderings of the writer on the device (writer_kernels.cuh): one launch, five CTAs.
static bool order_color_on_device(void* user, const uint32_t* ep_lo, const uint32_t* ep_hi, uint32_t n, const uint32_t* row_start, const uint32_t* col, const uint32_t* cnt,
                                  uint32_t selected, const uint32_t base[3], const uint32_t* selectors, uint32_t n_sel, uint16_t* remap4, uint16_t* sel_remap)
{
    crn_gpu_ctx* ctx = static_cast<crn_gpu_ctx*>(user);
    if (!ctx || !n || !n_sel || n > (uint32_t)crn::kOrderMaxN || n_sel > (uint32_t)crn::kOrderMaxN || getenv("CRN_B200_HOST_ORDER")) return false;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return false;
    const uint32_t nt = row_start[n];
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_lo = 0, o_hi = o_lo + al((size_t)n * 4), o_rs = o_hi + al((size_t)n * 4), o_col = o_rs + al((size_t)(n + 1) * 4), o_cnt = o_col + al((size_t)nt * 4),
                 o_sel = o_cnt + al((size_t)nt * 4), o_remap = o_sel + al((size_t)n_sel * 4), o_sremap = o_remap + al((size_t)4 * n * 2), total = o_sremap + al((size_t)n_sel * 2);
    HcBuf d;
    if (d.alloc(ctx, total) != cudaSuccess) return false;
    uint8_t* b = static_cast<uint8_t*>(d.p);
    cudaStream_t st = ctx->stream;
    bool ok = true;
    ok &= cudaMemcpyAsync(b + o_lo, ep_lo, (size_t)n * 4, cudaMemcpyHostToDevice, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(b + o_hi, ep_hi, (size_t)n * 4, cudaMemcpyHostToDevice, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(b + o_rs, row_start, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st) == cudaSuccess;
    if (nt) {
        ok &= cudaMemcpyAsync(b + o_col, col, (size_t)nt * 4, cudaMemcpyHostToDevice, st) == cudaSuccess;
        ok &= cudaMemcpyAsync(b + o_cnt, cnt, (size_t)nt * 4, cudaMemcpyHostToDevice, st) == cudaSuccess;
    }
    ok &= cudaMemcpyAsync(b + o_sel, selectors, (size_t)n_sel * 4, cudaMemcpyHostToDevice, st) == cudaSuccess;
    if (!ok) return false;
    crn::OrderColorJob J;
    J.ep_lo = reinterpret_cast<const uint32_t*>(b + o_lo); J.ep_hi = reinterpret_cast<const uint32_t*>(b + o_hi);
    J.row_start = reinterpret_cast<const uint32_t*>(b + o_rs); J.col = reinterpret_cast<const uint32_t*>(b + o_col); J.cnt = reinterpret_cast<const uint32_t*>(b + o_cnt);
    J.selectors = reinterpret_cast<const uint32_t*>(b + o_sel);
    J.n = n; J.n_sel = n_sel; J.selected = selected;
    J.base[0] = base[0]; J.base[1] = base[1]; J.base[2] = base[2];
    J.remap = reinterpret_cast<uint16_t*>(b + o_remap); J.sel_remap = reinterpret_cast<uint16_t*>(b + o_sremap);
    int threads = crn::kOrderThreads;
    if (const char* t = getenv("CRN_B200_ORDER_THREADS")) { const int v = atoi(t); if (v == 256 || v == 512 || v == 1024) threads = v; }
#ifdef __CUDACC__
    if (cudaFuncSetAttribute(crn::crn_order_color_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(crn::OrderSmem)) != cudaSuccess ||
        cudaFuncSetAttribute(crn::crn_order_color_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(crn::OrderSmem)) != cudaSuccess ||
        cudaFuncSetAttribute(crn::crn_order_color_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(crn::OrderSmem)) != cudaSuccess) return false;
#endif
    if (threads == 256) CRN_LAUNCH(crn::crn_order_color_kernel<256>, 5, 256, sizeof(crn::OrderSmem), st, J);
    else if (threads == 512) CRN_LAUNCH(crn::crn_order_color_kernel<512>, 5, 512, sizeof(crn::OrderSmem), st, J);
    else CRN_LAUNCH(crn::crn_order_color_kernel<1024>, 5, 1024, sizeof(crn::OrderSmem), st, J);
    ctx->launches++;
    ok &= cudaMemcpyAsync(remap4, b + o_remap, (size_t)4 * n * 2, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(sel_remap, b + o_sremap, (size_t)n_sel * 2, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaStreamSynchronize(st) == cudaSuccess;
    return ok && cudaGetLastError() == cudaSuccess;
}

static int crn_write_impl(crn_gpu_ctx* ctx, const crn_gpu_crn_params* p, const crn_gpu_hc_params* hp, const uint16_t* endpoint_indices, const uint16_t* selector_indices,
                          const uint32_t* color_endpoints, uint32_t n_color_endpoints, const uint32_t* alpha_endpoints, uint32_t n_alpha_endpoints,
                          const uint32_t* color_selectors, uint32_t n_color_selectors, const uint64_t* alpha_selectors, uint32_t n_alpha_selectors,
                          void** out_file, uint32_t* out_size);
extern "C" {
int crn_gpu_crn_write(const crn_gpu_crn_params* p, const crn_gpu_hc_params* hp, const uint16_t* endpoint_indices, const uint16_t* selector_indices,
                      const uint32_t* color_endpoints, uint32_t n_color_endpoints, const uint32_t* alpha_endpoints, uint32_t n_alpha_endpoints,
                      const uint32_t* color_selectors, uint32_t n_color_selectors, const uint64_t* alpha_selectors, uint32_t n_alpha_selectors,
                      void** out_file, uint32_t* out_size)
{
    return crn_write_impl(nullptr, p, hp, endpoint_indices, selector_indices, color_endpoints, n_color_endpoints, alpha_endpoints, n_alpha_endpoints, color_selectors, n_color_selectors,
                          alpha_selectors, n_alpha_selectors, out_file, out_size);
}
}  // extern "C"
// ctx != nullptr: the palette orderings run on that context's device; nullptr: host loops only (the public entry point has no context)
static int crn_write_impl(crn_gpu_ctx* ctx, const crn_gpu_crn_params* p, const crn_gpu_hc_params* hp, const uint16_t* endpoint_indices, const uint16_t* selector_indices,
                          const uint32_t* color_endpoints, uint32_t n_color_endpoints, const uint32_t* alpha_endpoints, uint32_t n_alpha_endpoints,
                          const uint32_t* color_selectors, uint32_t n_color_selectors, const uint64_t* alpha_selectors, uint32_t n_alpha_selectors,
                          void** out_file, uint32_t* out_size)
{ return crn_guard(nullptr, [&]() -> int {
    if (out_file) *out_file = nullptr;
    if (out_size) *out_size = 0;
    if (!crn_params_ok(p) || !hp || hp->struct_size != sizeof(crn_gpu_hc_params) || !endpoint_indices || !selector_indices || !out_file || !out_size ||
        hp->num_levels != p->levels || hp->num_faces != p->faces || !hp->num_blocks)
        return CRN_GPU_ERR_BAD_PARAM;
    crnw::Input in;
    memset(&in, 0, sizeof(in));
    in.crn_format = p->crn_format; in.width = p->width; in.height = p->height; in.num_levels = p->levels; in.num_faces = p->faces;
    in.userdata0 = p->userdata0; in.userdata1 = p->userdata1;
    switch (hp->format) {
    case CRN_GPU_FMT_DXT1: in.has_color = true; break;
    case CRN_GPU_FMT_DXT5: in.has_color = true; in.has_alpha0 = true; break;
    case CRN_GPU_FMT_DXT5A: in.has_alpha0 = true; break;
    case CRN_GPU_FMT_DXN_XY: case CRN_GPU_FMT_DXN_YX: in.has_alpha0 = in.has_alpha1 = true; break;
    default: return CRN_GPU_ERR_UNSUPPORTED;
    }
    if ((in.has_color && (!color_endpoints || !color_selectors || !n_color_endpoints || !n_color_selectors || n_color_endpoints > 8192 || n_color_selectors > 8192)) ||
        (in.has_alpha0 && (!alpha_endpoints || !alpha_selectors || !n_alpha_endpoints || !n_alpha_selectors || n_alpha_endpoints > 8192 || n_alpha_selectors > 8192)))
        return CRN_GPU_ERR_BAD_PARAM;
    crnw::Level lv[16];
    for (uint32_t l = 0; l < hp->num_levels; l++) lv[l] = { hp->levels[l].first_block, hp->levels[l].num_blocks, hp->levels[l].block_width };
    in.levels = lv; in.num_blocks = hp->num_blocks;
    in.endpoint_indices = endpoint_indices; in.selector_indices = selector_indices;
    in.color_endpoints = color_endpoints; in.n_color_endpoints = in.has_color ? n_color_endpoints : 0;
    in.color_selectors = color_selectors; in.n_color_selectors = in.has_color ? n_color_selectors : 0;
    in.alpha_endpoints = alpha_endpoints; in.n_alpha_endpoints = in.has_alpha0 ? n_alpha_endpoints : 0;
    in.alpha_selectors = alpha_selectors; in.n_alpha_selectors = in.has_alpha0 ? n_alpha_selectors : 0;
    // every index must address its palette: a bad one would otherwise be read out of bounds
    for (uint32_t b = 0; b < in.num_blocks; b++) {
        const uint16_t* e = endpoint_indices + (size_t)b * 4; const uint16_t* s = selector_indices + (size_t)b * 4;
        if (e[3] > 2 || (in.has_color && (e[0] >= n_color_endpoints || s[0] >= n_color_selectors)) ||
            (in.has_alpha0 && (e[1] >= n_alpha_endpoints || s[1] >= n_alpha_selectors)) || (in.has_alpha1 && (e[2] >= n_alpha_endpoints || s[2] >= n_alpha_selectors)))
            return CRN_GPU_ERR_BAD_DATA;
    }
    crnw::Input::ColorOrderHook hook = { ctx, order_color_on_device };
    in.color_order_hook = ctx ? &hook : nullptr;
    try {
        crnw::Writer w(in);
        std::vector<uint8_t> file;
        if (!w.write(file)) return CRN_GPU_ERR_BAD_DATA;
        void* m = malloc(file.size());
        if (!m) return CRN_GPU_ERR_NO_MEMORY;
        memcpy(m, file.data(), file.size());
        *out_file = m; *out_size = (uint32_t)file.size();
    } catch (const std::bad_alloc&) { return CRN_GPU_ERR_NO_MEMORY; }
    return CRN_GPU_OK;
}); }
extern "C" {

void crn_gpu_free_file(void* file) { free(file); }

// Interpolative search for the quality level closest to a target bitrate (create_compressed_texture,
// crnlib/crn_texture_comp.cpp:120-262): same bracket updates, interpolation, acceptance rule and stop test.
// pass(quality, &file, &size, &rate) produces a malloc'ed file; the best one is returned.
}  // extern "C" (templates need C++ linkage)
struct BitrateBest { float bitrate = 1e+10f; int quality = -1; void* file = nullptr; uint32_t size = 0; };   // survives the reference's second search
template <typename Pass>
static int bitrate_search(crn_gpu_ctx* ctx, float target, Pass pass, BitrateBest& best, float* out_highest)
{
    float cached[256], highest = 0.0f;
    int low = 0, high = 255;
    for (int i = 0; i < 256; i++) cached[i] = -1.0f;
    uint32_t iter = 0;
    bool binary = false;
    while (low <= high) {
        int trial = (low + high) / 2;
        if (iter && !binary) {
            int blo = trial;
            while (cached[blo] < 0 && blo > 0) blo--;
            if (cached[blo] < 0) trial = (int)((float)low + ((float)high - (float)low) * .33f);
            else {
                int bhi = trial + 1;
                if (bhi <= 255) {
                    while (cached[bhi] < 0 && bhi < 255) bhi++;
                    if (cached[bhi] >= 0) {
                        const float rlo = cached[blo], rhi = cached[bhi];
                        if (rlo < rhi && rlo < target && rhi >= target) {
                            const int q = low + (int)(((target - rlo) * (high - low)) / (rhi - rlo));
                            if (q >= low && q <= high) trial = q;
                        }
                    }
                }
            }
        }
        void* file = nullptr; uint32_t size = 0; float rate = 0.0f;
        const int rc = pass((uint32_t)trial, &file, &size, &rate);
        if (rc) { free(best.file); best.file = nullptr; return rc; }
        cached[trial] = rate;
        if (rate > highest) highest = rate;
        if (best.quality < 0 || (rate <= target && best.bitrate > target) ||
            ((rate <= target || best.bitrate > target) && fabsf(rate - target) < fabsf(best.bitrate - target))) {
            best.bitrate = rate; best.quality = trial;
            free(best.file); best.file = file; best.size = size; file = nullptr;
            if (best.bitrate <= target && fabsf(best.bitrate - target) < .005f) break;
        }
        free(file);
        if (rate > target) high = trial - 1; else low = trial + 1;
        if (++iter > 8) binary = true;
    }
    if (best.quality < 0) { free(best.file); best.file = nullptr; return set_err(ctx, CRN_GPU_ERR_BAD_DATA, "bitrate search found nothing"); }
    if (out_highest) *out_highest = highest;
    return CRN_GPU_OK;
}
extern "C" {

int crn_gpu_compress_crn(crn_gpu_ctx* ctx, const crn_gpu_crn_params* p, const void* const* h_images, void** out_file, uint32_t* out_size, float* out_bitrate,
                         uint32_t* out_quality)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (out_file) *out_file = nullptr;
    if (out_size) *out_size = 0;
    if (out_bitrate) *out_bitrate = 0.0f;
    if (out_quality) *out_quality = 0;
    if (!crn_params_ok(p) || !h_images || !out_file || !out_size || (p->shard_count > 1 && (!p->exchange || p->shard_rank >= p->shard_count)))
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_compress_crn: bad argument");
    for (uint32_t i = 0; i < p->faces * p->levels; i++)
        if (!h_images[i]) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_compress_crn: missing image");      // alias_images, crn_comp.cpp:432-435
    crn_gpu_hc_params hp;
    int rc = crn_gpu_crn_hc_params(p, &hp);
    if (rc) return set_err(ctx, rc, "crn_gpu_compress_crn: unsupported format");
    timespec ts_begin; clock_gettime(CLOCK_MONOTONIC, &ts_begin);
    const double t_begin = ts_begin.tv_sec * 1e3 + ts_begin.tv_nsec * 1e-6;
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    // the [block][16] array of all levels and faces (crn_comp.cpp:717-741), gathered on the device from one staging image;
    // it stays in HBM for every trial of the bitrate search (the reference restarts from the pixels on each pass)
    HcBuf d_blocks, d_img;
    if (d_blocks.alloc(ctx, (size_t)hp.num_blocks * 64) != cudaSuccess || d_img.alloc(ctx, (size_t)p->width * p->height * 4) != cudaSuccess)
        return set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "crn_gpu_compress_crn: out of device memory");
    uint64_t texels = 0;
    for (uint32_t l = 0; l < p->levels; l++) {
        const uint32_t w = std::max(1u, p->width >> l), h = std::max(1u, p->height >> l);
        const uint32_t per_face = hp.levels[l].num_blocks / p->faces;
        for (uint32_t f = 0; f < p->faces; f++) {
            CRN_CUDA(ctx, cudaMemcpyAsync(d_img.p, h_images[f * p->levels + l], (size_t)w * h * 4, cudaMemcpyHostToDevice, ctx->stream));
            if (p->crn_format >= 3 && p->crn_format <= 6) {      // cooked into the swizzled layout first (crn_comp.cpp:440-452)
                rc = crn_gpu_convert_pixels(ctx, d_img.p, w, h, w * 4, 1 + 2 * (p->crn_format - 3));
                if (rc) return rc;
            }
            rc = crn_gpu_blockify(ctx, d_img.p, w, h, w * 4, 8, d_blocks.as<uint8_t>() + ((size_t)hp.levels[l].first_block + (size_t)f * per_face) * 64, nullptr, nullptr);
            if (rc) return rc;
            texels += (uint64_t)w * h;
        }
    }
    const bool trace = getenv("CRN_B200_TRACE") != nullptr;
    auto wall_ms = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    if (trace) { cudaStreamSynchronize(ctx->stream); fprintf(stderr, "[crn_b200] compress_crn: upload + block gather %.1f ms\n", wall_ms() - t_begin); }
    // one crn_comp::compress_pass at a quality level: quantise, write, report bits per texel.  The tile pass and the sorted training sets do
    // not depend on the quality level: computed by the first pass, kept in `prepared` for the other trials of a search.
    HcPrepared prepared;
    auto pass = [&](uint32_t quality, void** file, uint32_t* size, float* bitrate) -> int {
        const double tp0 = wall_ms();
        crn_gpu_crn_params q = *p;
        q.quality_level = quality;
        crn_gpu_hc_params qhp;
        int r = crn_gpu_crn_hc_params(&q, &qhp);
        if (r) return r;
        if (p->shard_count > 1) { qhp.shard_rank = p->shard_rank; qhp.shard_count = p->shard_count; qhp.exchange = p->exchange; qhp.exchange_user = p->exchange_user; }
        crn_gpu_hc* H = nullptr;
        r = hc_compress_prepared(ctx, &qhp, d_blocks.p, 0, &H, &prepared);
        if (r) return r;
        const double tp1 = wall_ms();
        r = crn_write_impl(ctx, &q, &qhp, H->endpoint_indices.data(), H->selector_indices.data(), H->color_endpoints.data(), (uint32_t)H->color_endpoints.size(),
                           H->alpha_endpoints.data(), (uint32_t)H->alpha_endpoints.size(), H->color_selectors.data(), (uint32_t)H->color_selectors.size(),
                           H->alpha_selectors.data(), (uint32_t)H->alpha_selectors.size(), file, size);
        crn_gpu_hc_free(H);
        if (r) return set_err(ctx, r, "crn_gpu_compress_crn: the writer rejected the quantiser's output");
        *bitrate = (*size * 8.0f) / (float)texels;                                                                // crn_comp.cpp:1640-1653
        if ((r = progress_tick(ctx, 24, 25, 1, 1)) != CRN_GPU_OK) { free(*file); *file = nullptr; *size = 0; return r; }   // crn_comp.cpp:1600
        if (trace) fprintf(stderr, "[crn_b200] compress_crn pass q%u: quantiser %.1f ms, writer %.1f ms, %u bytes\n", quality, tp1 - tp0, wall_ms() - tp1, *size);
        return CRN_GPU_OK;
    };
    const bool manual = p->palette_sizes[0] && p->palette_sizes[1] && p->palette_sizes[2] && p->palette_sizes[3];
    if (!(p->target_bitrate > 0.0f) || manual) {
        float rate = 0.0f;
        rc = pass(p->quality_level, out_file, out_size, &rate);
        if (rc) return rc;
        if (out_bitrate) *out_bitrate = rate;
        if (out_quality) *out_quality = p->quality_level;
        return CRN_GPU_OK;
    }
    // When even the highest bitrate stays under the target the reference clears cCRNCompFlagHierarchical and searches again
    // (crn_texture_comp.cpp:232-250).  dxt_hc never reads m_hierarchical in this revision (crnlib/crn_dxt_hc.cpp has no use of it), so for a
    // .crn that second search repeats the first one's passes exactly and cannot replace the best (a new best needs a strictly smaller
    // distance to the target): it is not re-run here.
    BitrateBest best;
    rc = bitrate_search(ctx, p->target_bitrate, pass, best, nullptr);
    if (rc) return rc;
    *out_file = best.file; *out_size = best.size;
    if (out_bitrate) *out_bitrate = best.bitrate;
    if (out_quality) *out_quality = (uint32_t)best.quality;
    return CRN_GPU_OK;
}); }

int crn_gpu_crnd_get_texture_info(const void* h_crn, uint32_t crn_size, crn_gpu_texture_info* info)
{ return crn_guard(nullptr, [&]() -> int {
    if (!info || info->struct_size != sizeof(crn_gpu_texture_info)) return CRN_GPU_ERR_BAD_PARAM;
    crn::CrnHeaderInfo h;
    if (!crn::crn_parse_header(static_cast<const uint8_t*>(h_crn), crn_size, h)) return CRN_GPU_ERR_BAD_DATA;
    info->width = h.width; info->height = h.height; info->levels = h.levels; info->faces = h.faces;
    info->bytes_per_block = (h.format == 0 || h.format == 9 || h.format == 10 || h.format == 11 || h.format == 13) ? 8 : 16;
    info->userdata0 = h.userdata0; info->userdata1 = h.userdata1; info->format = h.format;
    return CRN_GPU_OK;
}); }

static int transcode_launch(crn_gpu_ctx* ctx, const crn::TranscodeFile* d_files, uint32_t nfiles, uint32_t row_entries)
{
    // shared-memory row buffers sized by the widest file of the launch: a batch of 1024^2 textures then runs 4 CTAs per SM, not 2
    if (row_entries > crn::kRowbufSmemEntries) row_entries = crn::kRowbufSmemEntries;
    const size_t smem = crn::transcode_smem_bytes(row_entries);
#ifdef __CUDACC__
    if (!ctx->transcode_smem_set) {
        CRN_CUDA(ctx, cudaFuncSetAttribute(crn::transcode_levels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)crn::transcode_smem_bytes(crn::kRowbufSmemEntries)));
        ctx->transcode_smem_set = 1;
    }
#endif
    CRN_LAUNCH(crn::transcode_levels_kernel, nfiles, crn::kTranscodeWarps * 32, smem, ctx->stream, d_files, row_entries);
    ctx->launches++;
    CRN_CUDA(ctx, cudaGetLastError());
    return CRN_GPU_OK;
}

// Levels of at least this many blocks take the table / walk / resolve kernels of transcode_wide.cuh
// (CRN_B200_WIDE_MIN_BLOCKS overrides the default; the CPU tests lower it to cover the path on small files).
static uint32_t wide_min_blocks()
{
    const char* e = getenv("CRN_B200_WIDE_MIN_BLOCKS");
    if (e && *e) return (uint32_t)strtoul(e, nullptr, 10);
    return 4096u;
}

// Launches the transcoder for every active level of `count` textures whose level tables (host_file) are filled in:
// large levels through the wide path (transcode_wide.cuh), the rest through the warp-per-level kernel.  d_files: device
// array the level tables are uploaded to (the texture's own d_file for a single texture).
// The lane-per-stream kernel (transcode_streams.cuh) is OFF unless CRN_B200_STREAMS_MIN names a stream count from which to
// use it: measured on the B200 (profiles/r2a_transcode_lane_per_stream.jsonl) it reaches 6-16 Gtexel/s on 1 K - 16 K file
// batches against 33 Gtexel/s for the CTA-per-file kernels -- with every file of the batch in flight at once the per-file
// tables and palettes (~100-200 KB each) fall out of L2 and every symbol value costs a DRAM round trip.
static uint32_t streams_min()
{
    const char* e = getenv("CRN_B200_STREAMS_MIN");
    if (e && *e) return (uint32_t)strtoul(e, nullptr, 10);
    return 0xFFFFFFFFu;
}

static int transcode_streams(crn_gpu_ctx* ctx, crn_gpu_texture* const* texs, uint32_t count, crn::TranscodeFile* d_files, uint32_t nstreams)
{
    struct Key { uint64_t blocks; uint32_t fmt_class, file, slot; };
    std::vector<Key> keys;
    keys.reserve(nstreams);
    std::vector<crn::TranscodeFile> files(count);
    for (uint32_t t = 0; t < count; t++) {
        const crn::TranscodeFile& hf = texs[t]->host_file;
        files[t] = hf;
        const uint32_t fc = (hf.format == 0) ? 0u : (hf.format == 9 ? 1u : ((hf.format == 7 || hf.format == 8) ? 2u : 3u));
        for (uint32_t slot = 0; slot < 16; slot++) {
            const crn::LevelStream& ls = hf.levels[slot];
            if (!ls.active) continue;
            const uint64_t W = (ls.blocks_x + 1) & ~1u, H = (ls.blocks_y + 1) & ~1u;
            keys.push_back(Key{ W * H * hf.faces, fc, t, slot });
        }
    }
    // lanes of a warp run in lock step: neighbours should be the same format and the same length
    std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) {
        if (a.blocks != b.blocks) return a.blocks > b.blocks;
        if (a.fmt_class != b.fmt_class) return a.fmt_class < b.fmt_class;
        return a.file != b.file ? a.file < b.file : a.slot < b.slot;
    });
    std::vector<crn::StreamDesc> sd(keys.size());
    for (size_t i = 0; i < keys.size(); i++) { sd[i].file = d_files + keys[i].file; sd[i].slot = keys[i].slot; sd[i].pad = 0; }
    int rc = ensure(ctx, &ctx->d_wide, &ctx->d_wide_cap, sizeof(crn::StreamDesc) * sd.size());
    if (rc) return rc;
    CRN_CUDA(ctx, cudaMemcpyAsync(d_files, files.data(), sizeof(crn::TranscodeFile) * count, cudaMemcpyHostToDevice, ctx->stream));
    CRN_CUDA(ctx, cudaMemcpyAsync(ctx->d_wide, sd.data(), sizeof(crn::StreamDesc) * sd.size(), cudaMemcpyHostToDevice, ctx->stream));
    const size_t smem = sizeof(crn::StreamSmem);
#ifdef __CUDACC__
    if (!ctx->streams_smem_set) {
        CRN_CUDA(ctx, cudaFuncSetAttribute(crn::transcode_streams_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->streams_smem_set = 1;
    }
#endif
    const uint32_t n = (uint32_t)sd.size();
    CRN_LAUNCH(crn::transcode_streams_kernel, (n + crn::kStreamThreads - 1) / crn::kStreamThreads, crn::kStreamThreads, smem, ctx->stream,
               static_cast<const crn::StreamDesc*>(ctx->d_wide), n);
    ctx->launches++;
    CRN_CUDA(ctx, cudaGetLastError());
    // the host vectors above are pageable sources of async copies: staged before cudaMemcpyAsync returns
    return CRN_GPU_OK;
}

static int transcode_textures(crn_gpu_ctx* ctx, crn_gpu_texture* const* texs, uint32_t count, crn::TranscodeFile* d_files)
{
    const uint32_t min_blocks = wide_min_blocks();
    {
        uint32_t nstreams = 0;
        for (uint32_t t = 0; t < count; t++)
            for (uint32_t slot = 0; slot < 16; slot++) nstreams += texs[t]->host_file.levels[slot].active ? 1u : 0u;
        if (nstreams >= streams_min()) return transcode_streams(ctx, texs, count, d_files, nstreams);
    }
    std::vector<crn::WideLevel> wide;
    std::vector<crn::TranscodeFile> files(count);
    size_t bytes = 0;
    uint32_t next_cta = 0, small_levels = 0;
    for (uint32_t t = 0; t < count; t++) {
        crn::TranscodeFile& hf = texs[t]->host_file;
        for (uint32_t slot = 0; slot < 16; slot++) {
            crn::LevelStream& ls = hf.levels[slot];
            if (!ls.active) continue;
            const uint32_t W = (ls.blocks_x + 1) & ~1u, H = (ls.blocks_y + 1) & ~1u;
            const uint64_t blocks = (uint64_t)W * H * hf.faces;
            if (blocks < min_blocks || W > crn::kWideMaxW || ls.src_size >= (1u << 27) || blocks >= (1ull << 31)) { small_levels++; continue; }
            crn::WideLevel wl;
            memset(&wl, 0, sizeof(wl));
            wl.file = d_files + t; wl.slot = slot; wl.nbits = ls.src_size * 8; wl.W = W; wl.H = H;
            wl.nrows = H * hf.faces; wl.npairs = wl.nrows * (W / 2);
            wl.ntiles = (wl.nbits + crn::kWideT - 1) / crn::kWideT;
            if (!wl.ntiles) wl.ntiles = 1;
            wl.stride = (uint32_t)((((size_t)wl.ntiles * crn::kWideT + 4 * crn::kWideWB + crn::kWideMW) + 255) & ~(size_t)255);
            wl.first_cta = next_cta; wl.num_ctas = (wl.ntiles + crn::kWideTilesPerCta - 1) / crn::kWideTilesPerCta;
            next_cta += wl.num_ctas;
            crn::wide_format_slots(hf.format, wl.ne, wl.ns, wl.e_model, wl.s_model);
            wl.tab = reinterpret_cast<uint8_t*>(bytes);                  // offsets for now, rebased below
            bytes += (size_t)crn::kWideTabBytes * wl.stride;
            wl.pair_ofs = reinterpret_cast<uint32_t*>(bytes);
            bytes += ((size_t)wl.npairs * 4 + 255) & ~(size_t)255;
            ls.active = 0;                                               // the warp-per-level kernel skips it
            wide.push_back(wl);
        }
        files[t] = hf;
    }
    // pageable sources: cudaMemcpyAsync returns once they are staged, so the vectors may die with this frame
    CRN_CUDA(ctx, cudaMemcpyAsync(d_files, files.data(), sizeof(crn::TranscodeFile) * count, cudaMemcpyHostToDevice, ctx->stream));
    if (small_levels) {
        uint32_t row_entries = 0;
        for (uint32_t t = 0; t < count; t++) row_entries = std::max(row_entries, texs[t]->host_file.rowbuf_total);
        const int rc = transcode_launch(ctx, d_files, count, row_entries);
        if (rc) return rc;
    }
    if (wide.empty()) return CRN_GPU_OK;
    const uint32_t nl = (uint32_t)wide.size();
    const size_t o_progress = (sizeof(crn::WideLevel) * nl + 255) & ~(size_t)255, o_data = (o_progress + 4 * (size_t)nl + 255) & ~(size_t)255;
    int rc = ensure(ctx, &ctx->d_wide, &ctx->d_wide_cap, o_data + bytes);
    if (rc) return rc;
    uint8_t* base = static_cast<uint8_t*>(ctx->d_wide);
    CRN_CUDA(ctx, cudaMemsetAsync(base + o_progress, 0, 4 * (size_t)nl, ctx->stream));        // row counters, one per level
    for (uint32_t i = 0; i < nl; i++) {
        crn::WideLevel& wl = wide[i];
        wl.progress = reinterpret_cast<uint32_t*>(base + o_progress) + i;
        wl.tab = base + o_data + reinterpret_cast<size_t>(wl.tab);
        wl.pair_ofs = reinterpret_cast<uint32_t*>(base + o_data + reinterpret_cast<size_t>(wl.pair_ofs));
    }
    CRN_CUDA(ctx, cudaMemcpyAsync(base, wide.data(), sizeof(crn::WideLevel) * nl, cudaMemcpyHostToDevice, ctx->stream));
    const crn::WideLevel* d_levels = reinterpret_cast<const crn::WideLevel*>(base);
    const size_t smem = sizeof(crn::WideSmemBC);
#ifdef __CUDACC__
    if (!ctx->wide_smem_set) {
        CRN_CUDA(ctx, cudaFuncSetAttribute(crn::transcode_walk_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CRN_CUDA(ctx, cudaFuncSetAttribute(crn::transcode_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CRN_CUDA(ctx, cudaFuncSetAttribute(crn::transcode_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->wide_smem_set = 1;
    }
#endif
    const int pipe = getenv("CRN_B200_WIDE_NOPIPE") ? 0 : 1;
    CRN_LAUNCH(crn::transcode_tables_kernel, next_cta, crn::kWideThreadsA, 0, ctx->stream, d_levels, nl);
    // CRN_B200_WIDE_SPLIT: 1 forces the two-launch form, 0 the fused one (tests; the emulator runs CTAs one after another)
    const char* split_env = getenv("CRN_B200_WIDE_SPLIT");
    const bool fused = split_env ? split_env[0] == '0' : 2 * nl <= (uint32_t)ctx->sm_count;
    if (fused) {
        // one CTA per SM at this shared-memory size: every walker / resolver pair is resident, the resolver trails the walker
        CRN_LAUNCH(crn::transcode_walk_resolve_kernel, 2 * nl, crn::kWideThreadsC, smem, ctx->stream, d_levels, pipe);
        ctx->launches += 2;
    } else {
        CRN_LAUNCH(crn::transcode_walk_kernel, nl, crn::kWideThreadsC, smem, ctx->stream, d_levels, pipe);
        CRN_LAUNCH(crn::transcode_resolve_kernel, nl, crn::kWideThreadsC, smem, ctx->stream, d_levels);
        ctx->launches += 3;
    }
    CRN_CUDA(ctx, cudaGetLastError());
    return CRN_GPU_OK;
}

static int transcode_texture(crn_gpu_texture* tex) { return transcode_textures(tex->ctx, &tex, 1, tex->d_file); }

int crn_gpu_crnd_unpack_begin(crn_gpu_ctx* ctx, const void* h_crn, uint32_t crn_size, crn_gpu_texture** out_tex)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx || !out_tex) return CRN_GPU_ERR_BAD_PARAM;
    *out_tex = nullptr;
    const uint8_t* d = static_cast<const uint8_t*>(h_crn);
    crn::CrnHeaderInfo h;
    if (!crn::crn_parse_header(d, crn_size, h)) return set_err(ctx, CRN_GPU_ERR_BAD_DATA, "crnd_unpack_begin: not a CRN file");
    if (h.format > 9 || h.format == 1) return set_err(ctx, CRN_GPU_ERR_UNSUPPORTED, "crnd_unpack_begin: only DXT1/DXT5*/DXN/DXT5A .crn files are supported");
    const bool has_color = h.format <= 6, has_alpha = h.format != 0;
    if ((has_color && !h.pal_num[0]) || (has_alpha && !h.pal_num[2])) return set_err(ctx, CRN_GPU_ERR_BAD_DATA, "crnd_unpack_begin: missing palette");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));

    // Huffman models: init_tables (:3662-3692) then the model headers of each palette segment
    crn::HostModel hm[crn::kNumModels];
    {
        crn::HostBits b = { d + h.tables_ofs, h.tables_size, 0 };
        bool ok = hm[crn::kDmRef].receive(b);
        if (ok && h.pal_num[0]) ok = hm[crn::kDmColorEp].receive(b) && hm[crn::kDmColorSel].receive(b);
        if (ok && h.pal_num[2]) ok = hm[crn::kDmAlphaEp].receive(b) && hm[crn::kDmAlphaSel].receive(b);
        if (!ok) return set_err(ctx, CRN_GPU_ERR_BAD_DATA, "crnd_unpack_begin: corrupt Huffman tables");
        // Decoded symbols index the palettes directly (selectors) or step through them (endpoint deltas): a model the format needs must
        // exist, and may not name more entries than its palette holds (the reference trusts the file here; we do not).
        auto fits = [&](int m, uint32_t pal) { const size_t n = hm[m].len.size(); return n >= 1 && n <= pal && !hm[m].sorted.empty(); };
        ok = !hm[crn::kDmRef].sorted.empty() && hm[crn::kDmRef].len.size() <= 256;
        if (ok && has_color) ok = fits(crn::kDmColorEp, h.pal_num[0]) && fits(crn::kDmColorSel, h.pal_num[1]);
        if (ok && has_alpha) ok = fits(crn::kDmAlphaEp, h.pal_num[2]) && fits(crn::kDmAlphaSel, h.pal_num[3]);
        if (!ok) return set_err(ctx, CRN_GPU_ERR_BAD_DATA, "crnd_unpack_begin: Huffman models do not match the palettes");
    }
    uint32_t pal_data_ofs[4] = { 0, 0, 0, 0 }, pal_data_bit[4] = { 0, 0, 0, 0 };
    for (int i = 0; i < 4; i++) {
        if (!h.pal_num[i < 2 ? 0 : 2]) continue;
        crn::HostBits b = { d + h.pal_ofs[i], h.pal_size[i], 0 };
        bool ok = true;
        if (i == 0) ok = hm[crn::kDmPalCe0].receive(b) && hm[crn::kDmPalCe1].receive(b);
        else ok = hm[crn::kDmPalCs + (i - 1)].receive(b);
        if (!ok) return set_err(ctx, CRN_GPU_ERR_BAD_DATA, "crnd_unpack_begin: corrupt palette model");
        pal_data_ofs[i] = h.pal_ofs[i] + (uint32_t)(b.pos >> 3);
        pal_data_bit[i] = (uint32_t)(b.pos & 7);
    }
    std::vector<crn::HuffModelDev> dev_models(crn::kNumModels);
    std::vector<uint16_t> pool;
    for (int i = 0; i < crn::kNumModels; i++) hm[i].to_device(dev_models[i], pool);
    if (pool.empty()) pool.push_back(0);

    crn_gpu_texture* t = new (std::nothrow) crn_gpu_texture();
    if (!t) return CRN_GPU_ERR_NO_MEMORY;
    memset(static_cast<void*>(t), 0, sizeof(*t));
    t->ctx = ctx; t->hdr = h;
    t->bytes_per_block = (h.format == 0 || h.format == 9) ? 8 : 16;
    uint32_t rowbuf_total = 0;
    uint64_t ofs = 0;
    for (uint32_t l = 0; l < h.levels; l++) {
        const uint32_t w = h.width >> l ? h.width >> l : 1, hh = h.height >> l ? h.height >> l : 1;
        const uint32_t bx = (w + 3) >> 2, by = (hh + 3) >> 2;
        t->rowbuf_ofs[l] = rowbuf_total;
        rowbuf_total += (bx + 1) & ~1u;
        t->level_ofs[l] = ofs;
        t->level_face_size[l] = (uint64_t)bx * by * t->bytes_per_block;
        ofs += t->level_face_size[l] * h.faces;
    }
    t->total_size = ofs;

    // carve one slab
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t o_bytes = 0, o_models = align(o_bytes + crn_size + 16), o_pool = align(o_models + sizeof(crn::HuffModelDev) * crn::kNumModels);
    size_t o_ce = align(o_pool + pool.size() * 2), o_cs = align(o_ce + 4 * (size_t)(h.pal_num[0] + 1)), o_ae = align(o_cs + 4 * (size_t)(h.pal_num[1] + 1));
    size_t o_as = align(o_ae + 2 * (size_t)(h.pal_num[2] + 1)), o_row = align(o_as + 6 * (size_t)(h.pal_num[3] + 1));
    size_t o_file = align(o_row + 9 * (size_t)rowbuf_total + 16), total = align(o_file + sizeof(crn::TranscodeFile));
    cudaError_t ce = cudaMalloc(&t->slab, total);
    if (ce != cudaSuccess) { delete t; return set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "cudaMalloc", ce); }
    uint8_t* base = static_cast<uint8_t*>(t->slab);
    crn::TranscodeFile& f = t->host_file;
    f.bytes = base + o_bytes;
    f.models = reinterpret_cast<const crn::HuffModelDev*>(base + o_models);
    f.sorted_pool = reinterpret_cast<const uint16_t*>(base + o_pool);
    f.color_endpoints = reinterpret_cast<uint32_t*>(base + o_ce);
    f.color_selectors = reinterpret_cast<uint32_t*>(base + o_cs);
    f.alpha_endpoints = reinterpret_cast<uint16_t*>(base + o_ae);
    f.alpha_selectors = reinterpret_cast<uint16_t*>(base + o_as);
    f.rowbuf_pool = reinterpret_cast<uint2*>(base + o_row);
    f.rowbuf_total = rowbuf_total;
    f.num_color_endpoints = h.pal_num[0]; f.num_color_selectors = h.pal_num[1];
    f.num_alpha_endpoints = h.pal_num[2]; f.num_alpha_selectors = h.pal_num[3];
    for (int i = 0; i < 4; i++) { f.pal_data_ofs[i] = pal_data_ofs[i]; f.pal_data_bit[i] = pal_data_bit[i]; f.pal_size_end[i] = h.pal_ofs[i] + h.pal_size[i]; }
    f.format = h.format; f.faces = h.faces;
    t->d_file = reinterpret_cast<crn::TranscodeFile*>(base + o_file);

    bool fail = false;
    fail |= cudaMemsetAsync(base + o_bytes + crn_size, 0, 16, ctx->stream) != cudaSuccess;
    fail |= cudaMemcpyAsync(base + o_bytes, d, crn_size, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess;
    fail |= cudaMemcpyAsync(base + o_models, dev_models.data(), sizeof(crn::HuffModelDev) * crn::kNumModels, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess;
    fail |= cudaMemcpyAsync(base + o_pool, pool.data(), pool.size() * 2, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess;
    fail |= cudaMemcpyAsync(t->d_file, &f, sizeof(f), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess;
    if (!fail) {
        CRN_LAUNCH(crn::transcode_palettes_kernel, 1, 128, 0, ctx->stream, t->d_file);
        ctx->launches++;
        fail |= cudaGetLastError() != cudaSuccess;
        // the host vectors above die at return: wait for the copies
        fail |= cudaStreamSynchronize(ctx->stream) != cudaSuccess;
    }
    if (fail) { cudaFree(t->slab); delete t; return set_err(ctx, CRN_GPU_ERR_CUDA, "crnd_unpack_begin: upload / palette decode failed"); }
    *out_tex = t;
    return CRN_GPU_OK;
}); }

static void fill_level(crn_gpu_texture* t, uint32_t slot, uint32_t level, uint32_t row_pitch)
{
    const crn::CrnHeaderInfo& h = t->hdr;
    crn::LevelStream& ls = t->host_file.levels[slot];
    const uint32_t w = h.width >> level ? h.width >> level : 1, hh = h.height >> level ? h.height >> level : 1;
    const uint32_t next = level + 1 < h.levels ? h.level_ofs[level + 1] : h.data_size;
    ls.src_ofs = h.level_ofs[level];
    ls.src_size = next > h.level_ofs[level] ? next - h.level_ofs[level] : 0;
    ls.blocks_x = (w + 3) >> 2; ls.blocks_y = (hh + 3) >> 2;
    ls.row_pitch = row_pitch ? row_pitch : ls.blocks_x * t->bytes_per_block;
    ls.rowbuf_ofs = t->rowbuf_ofs[level];
    ls.active = 1;
}

int crn_gpu_crnd_unpack_level(crn_gpu_texture* tex, void* const* d_dst_faces, uint32_t dst_size_in_bytes,
                              uint32_t row_pitch_in_bytes, uint32_t level_index)
{ return crn_guard(nullptr, [&]() -> int {
    if (!tex || !d_dst_faces) return CRN_GPU_ERR_BAD_PARAM;
    crn_gpu_ctx* ctx = tex->ctx;
    if (level_index >= tex->hdr.levels) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crnd_unpack_level: level out of range");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < 16; i++) tex->host_file.levels[i].active = 0;
    fill_level(tex, 0, level_index, row_pitch_in_bytes);
    crn::LevelStream& ls = tex->host_file.levels[0];
    const uint32_t minpitch = ls.blocks_x * tex->bytes_per_block;     // crn_decomp.h:3569-3575
    if (ls.row_pitch < minpitch || (ls.row_pitch & 3) || (uint64_t)dst_size_in_bytes < (uint64_t)ls.row_pitch * ls.blocks_y)
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crnd_unpack_level: bad pitch or destination size");
    for (uint32_t f = 0; f < tex->hdr.faces; f++) {
        if (!d_dst_faces[f]) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crnd_unpack_level: null face pointer");
        ls.dst[f] = (unsigned long long)(uintptr_t)d_dst_faces[f];
    }
    int rc = transcode_texture(tex);
    if (rc) return rc;
    // host_file is reused by the next call: make sure the async copy has read it
    CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CRN_GPU_OK;
}); }

int crn_gpu_crnd_unpack_level_host(crn_gpu_texture* tex, void* const* h_dst_faces, uint32_t dst_size_in_bytes,
                                   uint32_t row_pitch_in_bytes, uint32_t level_index)
{ return crn_guard(tex ? tex->ctx : nullptr, [&]() -> int {
    if (!tex || !h_dst_faces) return CRN_GPU_ERR_BAD_PARAM;
    crn_gpu_ctx* ctx = tex->ctx;
    if (level_index >= tex->hdr.levels) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crnd_unpack_level: level out of range");
    const uint32_t w = tex->hdr.width >> level_index ? tex->hdr.width >> level_index : 1, hh = tex->hdr.height >> level_index ? tex->hdr.height >> level_index : 1;
    const uint32_t bx = (w + 3) >> 2, by = (hh + 3) >> 2, tight = bx * tex->bytes_per_block;
    const uint32_t pitch = row_pitch_in_bytes ? row_pitch_in_bytes : tight;                                      // crn_decomp.h:3569-3575
    if (pitch < tight || (pitch & 3) || (uint64_t)dst_size_in_bytes < (uint64_t)pitch * by)
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crnd_unpack_level: bad pitch or destination size");
    const uint32_t faces = tex->hdr.faces;
    for (uint32_t f = 0; f < faces; f++) if (!h_dst_faces[f]) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crnd_unpack_level: null face pointer");
    const size_t face_bytes = (size_t)tight * by;
    int rc = ensure(ctx, &ctx->d_out, &ctx->d_out_cap, face_bytes * faces);
    if (rc) return rc;
    void* d_faces[6];
    for (uint32_t f = 0; f < faces; f++) d_faces[f] = static_cast<uint8_t*>(ctx->d_out) + face_bytes * f;
    rc = crn_gpu_crnd_unpack_level(tex, d_faces, (uint32_t)face_bytes, tight, level_index);
    if (rc) return rc;
    for (uint32_t f = 0; f < faces; f++)
        CRN_CUDA(ctx, cudaMemcpy2DAsync(h_dst_faces[f], pitch, d_faces[f], tight, tight, by, cudaMemcpyDeviceToHost, ctx->stream));
    CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CRN_GPU_OK;
}); }

uint64_t crn_gpu_crnd_total_size(const crn_gpu_texture* tex) { return tex ? tex->total_size : 0; }
uint64_t crn_gpu_crnd_level_offset(const crn_gpu_texture* tex, uint32_t level_index, uint32_t face_index)
{
    if (!tex || level_index >= tex->hdr.levels || face_index >= tex->hdr.faces) return ~0ull;
    return tex->level_ofs[level_index] + tex->level_face_size[level_index] * face_index;
}

static int prepare_all_levels(crn_gpu_texture* tex, void* d_dst, uint64_t dst_capacity)
{
    if (!d_dst || dst_capacity < tex->total_size) return set_err(tex->ctx, CRN_GPU_ERR_BAD_PARAM, "crnd_unpack_all_levels: destination too small");
    for (int i = 0; i < 16; i++) tex->host_file.levels[i].active = 0;
    for (uint32_t l = 0; l < tex->hdr.levels; l++) {
        fill_level(tex, l, l, 0);
        for (uint32_t f = 0; f < tex->hdr.faces; f++)
            tex->host_file.levels[l].dst[f] = (unsigned long long)(uintptr_t)(static_cast<uint8_t*>(d_dst) + tex->level_ofs[l] + tex->level_face_size[l] * f);
    }
    return CRN_GPU_OK;
}

int crn_gpu_crnd_unpack_all_levels(crn_gpu_texture* tex, void* d_dst, uint64_t dst_capacity)
{ return crn_guard(nullptr, [&]() -> int {
    if (!tex) return CRN_GPU_ERR_BAD_PARAM;
    crn_gpu_ctx* ctx = tex->ctx;
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = prepare_all_levels(tex, d_dst, dst_capacity);
    if (rc) return rc;
    return transcode_texture(tex);
}); }

int crn_gpu_crnd_unpack_all_levels_host(crn_gpu_texture* tex, void* h_dst, uint64_t dst_capacity)
{ return crn_guard(nullptr, [&]() -> int {
    if (!tex || !h_dst) return CRN_GPU_ERR_BAD_PARAM;
    crn_gpu_ctx* ctx = tex->ctx;
    if (dst_capacity < tex->total_size) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crnd_unpack_all_levels_host: destination too small");
    int rc = ensure(ctx, &ctx->d_out, &ctx->d_out_cap, tex->total_size);
    if (rc) return rc;
    rc = crn_gpu_crnd_unpack_all_levels(tex, ctx->d_out, ctx->d_out_cap);
    if (rc) return rc;
    CRN_CUDA(ctx, cudaMemcpyAsync(h_dst, ctx->d_out, tex->total_size, cudaMemcpyDeviceToHost, ctx->stream));
    CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CRN_GPU_OK;
}); }

int crn_gpu_crnd_unpack_batch(crn_gpu_ctx* ctx, crn_gpu_texture* const* textures, uint32_t count, void* const* d_dst,
                              const uint64_t* dst_capacity)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx || !textures || !d_dst || !dst_capacity || !count) return CRN_GPU_ERR_BAD_PARAM;
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ensure(ctx, &ctx->d_files, &ctx->d_files_cap, sizeof(crn::TranscodeFile) * (size_t)count);
    if (rc) return rc;
    for (uint32_t i = 0; i < count; i++) {
        if (!textures[i] || textures[i]->ctx != ctx) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crnd_unpack_batch: texture belongs to another context");
        rc = prepare_all_levels(textures[i], d_dst[i], dst_capacity[i]);
        if (rc) return rc;
    }
    rc = transcode_textures(ctx, textures, count, static_cast<crn::TranscodeFile*>(ctx->d_files));
    if (rc) return rc;
    CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CRN_GPU_OK;
}); }

int crn_gpu_crnd_unpack_end(crn_gpu_texture* tex)
{ return crn_guard(nullptr, [&]() -> int {
    if (!tex) return CRN_GPU_ERR_BAD_PARAM;
    cudaSetDevice(tex->ctx->device);
    cudaStreamSynchronize(tex->ctx->stream);
    cudaFree(tex->slab);
    delete tex;
    return CRN_GPU_OK;
}); }

/* ---- DDS container edge ------------------------------------------------------------------- */

int crn_gpu_dds_header(uint32_t crn_format, uint32_t width, uint32_t height, uint32_t levels, uint32_t faces, void* out_128_bytes)
{ return crn_guard(nullptr, [&]() -> int {   // "DDS " + DDSURFACEDESC2 as mipmapped_texture::write_dds fills it for the block formats (crnlib/crn_mipmapped_texture.cpp:921-1084)
    if (!out_128_bytes || !width || !height || !levels || levels > 16 || (faces != 1 && faces != 6)) return CRN_GPU_ERR_BAD_PARAM;
    auto fourcc = [](char a, char b, char c, char d) { return (uint32_t)(uint8_t)a | ((uint32_t)(uint8_t)b << 8) | ((uint32_t)(uint8_t)c << 16) | ((uint32_t)(uint8_t)d << 24); };
    uint32_t cc, bitcount = 0, bits_per_texel = 8;
    switch (crn_format) {                                        // crn_format -> pixel_format (crn_mipmapped_texture.cpp:2700-2745), then write_dds' switch
    case 0: cc = fourcc('D', 'X', 'T', '1'); bits_per_texel = 4; break;
    case 1: cc = fourcc('D', 'X', 'T', '3'); break;
    case 2: cc = fourcc('D', 'X', 'T', '5'); break;
    case 3: cc = fourcc('D', 'X', 'T', '5'); bitcount = fourcc('C', 'C', 'x', 'Y'); break;
    case 4: cc = fourcc('D', 'X', 'T', '5'); bitcount = fourcc('x', 'G', 'x', 'R'); break;
    case 5: cc = fourcc('D', 'X', 'T', '5'); bitcount = fourcc('x', 'G', 'B', 'R'); break;
    case 6: cc = fourcc('D', 'X', 'T', '5'); bitcount = fourcc('A', 'G', 'B', 'R'); break;
    case 7: cc = fourcc('A', 'T', 'I', '2'); bitcount = fourcc('A', '2', 'X', 'Y'); break;
    case 8: cc = fourcc('A', 'T', 'I', '2'); break;
    case 9: cc = fourcc('A', 'T', 'I', '1'); bits_per_texel = 4; break;
    default: return CRN_GPU_ERR_UNSUPPORTED;
    }
    uint32_t h[32];
    memset(h, 0, sizeof(h));
    h[0] = fourcc('D', 'D', 'S', ' ');
    h[1] = 124;
    h[2] = 0x1u | 0x2u | 0x4u | 0x1000u | 0x80000u;              // CAPS | HEIGHT | WIDTH | PIXELFORMAT | LINEARSIZE
    h[3] = height; h[4] = width;
    h[5] = (((width + 3) & ~3u) * ((height + 3) & ~3u) * bits_per_texel) >> 3;
    h[27] = 0x1000u;                                             // DDSCAPS_TEXTURE
    if (levels > 1) { h[7] = levels; h[2] |= 0x20000u; h[27] |= 0x400000u | 0x8u; }     // MIPMAPCOUNT; MIPMAP | COMPLEX
    if (faces > 1) { h[27] |= 0x8u; h[28] = 0x200u | 0xFC00u; }                          // CUBEMAP + the six face bits
    h[19] = 32; h[20] = 0x4u; h[21] = cc; h[22] = bitcount;      // DDPF_FOURCC
    memcpy(out_128_bytes, h, 128);
    return CRN_GPU_OK;
}); }

void crn_gpu_default_dds_params(crn_gpu_dds_params* p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->struct_size = sizeof(*p);
    p->levels = 1; p->faces = 1; p->quality_level = 255;
    p->hierarchical = 1;
    crn_gpu_default_pack_params(&p->pack);
}

// LZMA-compressed size of a buffer, the DDS path's bitrate measure (dds_comp::compress_pass, crnlib/crn_dds_comp.cpp:291-303 ->
// lzma_codec::pack, crnlib/crn_lzma_codec.cpp:45-112: LZMA SDK level 5 = 16 MiB dictionary, lc 3, lp 0, pb 2, fb 32, bt4, plus a 20-byte
// header).  The reference vendors the LZMA SDK; LZMA is out of this path's scope, so the system's liblzma is loaded at
// run time with the same coder parameters (the two encoders' sizes agree to a fraction of a percent).  0 = liblzma not available.
static uint64_t lzma_packed_size(const void* data, size_t n)
{
    struct Filter { uint64_t id; void* options; };
    typedef int (*preset_fn)(void*, uint32_t);
    typedef int (*encode_fn)(const Filter*, const void*, const uint8_t*, size_t, uint8_t*, size_t*, size_t);
    static preset_fn preset = nullptr; static encode_fn encode = nullptr; static int tried = 0;
    if (!tried) {
        tried = 1;
        void* h = dlopen("liblzma.so.5", RTLD_NOW | RTLD_LOCAL);
        if (h) { preset = (preset_fn)dlsym(h, "lzma_lzma_preset"); encode = (encode_fn)dlsym(h, "lzma_raw_buffer_encode"); }
    }
    if (!preset || !encode || !n) return 0;
    alignas(16) uint8_t opt[512];                                  // lzma_options_lzma (first member: uint32_t dict_size), generously sized
    memset(opt, 0, sizeof(opt));
    if (preset(opt, 5)) return 0;                                  // preset 5: normal mode, bt4, nice_len 32, lc 3 / lp 0 / pb 2
    const uint32_t dict = 1u << 24;
    memcpy(opt, &dict, 4);
    const Filter filters[2] = { { 0x4000000000000001ull, opt }, { ~0ull, nullptr } };   // LZMA_FILTER_LZMA1, LZMA_VLI_UNKNOWN
    const size_t cap = n + (n >> 2) + 4096;
    uint8_t* out = static_cast<uint8_t*>(malloc(cap));
    if (!out) return 0;
    size_t pos = 0;
    const int rc = encode(filters, nullptr, static_cast<const uint8_t*>(data), n, out, &pos, cap);
    free(out);
    return rc == 0 ? (uint64_t)pos + 20 : 0;                       // + sizeof(lzma_codec::header) (crnlib/crn_lzma_codec.h:72-88)
}

uint64_t crn_gpu_lzma_size(const void* data, uint64_t size) { return lzma_packed_size(data, (size_t)size); }

int crn_gpu_compress_dds(crn_gpu_ctx* ctx, const crn_gpu_dds_params* p, const void* const* h_images, void** out_file, uint32_t* out_size)
{
    return crn_gpu_compress_dds_ex(ctx, p, h_images, out_file, out_size, nullptr, nullptr);
}

int crn_gpu_compress_dds_ex(crn_gpu_ctx* ctx, const crn_gpu_dds_params* p, const void* const* h_images, void** out_file, uint32_t* out_size, float* out_bitrate,
                            uint32_t* out_quality)
{ return crn_guard(ctx, [&]() -> int {   // dds_comp::compress_init + convert_to_dxt + compress_pass (crnlib/crn_dds_comp.cpp:148-289)
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (out_file) *out_file = nullptr;
    if (out_size) *out_size = 0;
    if (out_bitrate) *out_bitrate = 0.0f;
    if (out_quality) *out_quality = 0;
    if (!p || p->struct_size != sizeof(crn_gpu_dds_params) || !h_images || !out_file || !out_size || p->width < 1 || p->height < 1 || p->width > 4096 || p->height > 4096 ||
        p->levels < 1 || p->levels > 16 || (p->faces != 1 && p->faces != 6) || p->quality_level > 255 || p->pack.struct_size != sizeof(crn_gpu_pack_params))
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_compress_dds: bad argument");
    const uint32_t count = p->faces * p->levels;
    for (uint32_t i = 0; i < count; i++)
        if (!h_images[i]) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_compress_dds: missing image");     // create_dds_tex, crn_dds_comp.cpp:71-74
    uint32_t fmt;
    switch (p->crn_format) {                                     // pixel_format_helpers::convert_crn_format_to_pixel_format
    case 0: fmt = CRN_GPU_FMT_DXT1; break;
    case 1: fmt = CRN_GPU_FMT_DXT3; break;
    case 2: fmt = CRN_GPU_FMT_DXT5; break;
    case 7: fmt = CRN_GPU_FMT_DXN_XY; break;
    case 8: fmt = CRN_GPU_FMT_DXN_YX; break;
    case 9: fmt = CRN_GPU_FMT_DXT5A; break;
    case 3: case 4: case 5: case 6: fmt = CRN_GPU_FMT_DXT5; break;   // DXT5_CCxY / xGxR / xGBR / AGBR: DXT5 blocks of "cooked" pixels
    default: return set_err(ctx, CRN_GPU_ERR_UNSUPPORTED, "crn_gpu_compress_dds: ETC formats are not built");
    }
    const bool swizzled = p->crn_format >= 3 && p->crn_format <= 6;
    // image_utils::get_image_conversion_type_from_crn_format (create_dds_tex, crn_dds_comp.cpp:90-104); declared beside the other DDS helpers
    const uint32_t cook = swizzled ? 1u + 2u * (p->crn_format - 3u) : 0u;
    crn_gpu_pack_params pack = p->pack;
    if (swizzled) pack.perceptual = 0;                              // mip_level::pack_to_dxt (crn_mipmapped_texture.cpp:146-150), qdxt_pack_init (:2337-2347)
    std::vector<crn_gpu_level_desc> lv(count);                   // face-major: the order write_dds emits and qdxt_pack_init walks
    bool has_alpha = false;
    for (uint32_t f = 0; f < p->faces; f++)
        for (uint32_t l = 0; l < p->levels; l++) {
            const uint32_t w = std::max(1u, p->width >> l), h = std::max(1u, p->height >> l);
            lv[f * p->levels + l] = { h_images[f * p->levels + l], w, h, w * 4 };
            if (!has_alpha && ((fmt == CRN_GPU_FMT_DXT1 && p->dxt1a_for_transparency) || fmt == CRN_GPU_FMT_DXT5A)) {  // image_utils::has_alpha
                const uint8_t* px = static_cast<const uint8_t*>(h_images[f * p->levels + l]);
                for (size_t i = 0, n = (size_t)w * h; i < n; i++) if (px[i * 4 + 3] < 255) { has_alpha = true; break; }
            }
        }
    if (fmt == CRN_GPU_FMT_DXT1 && has_alpha && p->pack.use_both_block_types && p->dxt1a_for_transparency) fmt = CRN_GPU_FMT_DXT1A;   // crn_dds_comp.cpp:243-246
    const uint32_t bpb = crn_gpu_bytes_per_block(fmt);
    uint64_t payload = 0;
    for (const crn_gpu_level_desc& d : lv) payload += (uint64_t)((d.width + 3) >> 2) * ((d.height + 3) >> 2) * bpb;
    if (128 + payload > 0xFFFFFFFFull) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_compress_dds: file would exceed 4 GiB (crn_uint32 size)");
    uint64_t total_texels = 0;                                                                                   // get_total_pixels_in_all_faces_and_mips
    for (const crn_gpu_level_desc& d : lv) total_texels += (uint64_t)d.width * d.height;
    uint8_t header[128];
    int rc = crn_gpu_dds_header(p->crn_format, p->width, p->height, p->levels, p->faces, header);
    if (rc) return set_err(ctx, rc, "crn_gpu_compress_dds: format has no DDS form");
    if (p->hierarchical == 0) pack.non_hierarchical = 1;                                                         // m_q1_params / m_q5_params .m_hierarchical (crn_dds_comp.cpp:160-166)
    crn_gpu_qdxt* q = nullptr;                                                                                   // kept across the passes of a search (m_pQDXT_state)
    // Levels whose pixels change before compression are staged in HBM and converted there (dds_kernels.cuh): the swizzled DXT5 layouts always
    // (`cook`), and an opaque source going to DXT5A block by block, whose alpha becomes its luma (mip_level::pack_to_dxt,
    // crn_mipmapped_texture.cpp:161-162; the clustered path packs the 255s as they are, qdxt_pack_init is called with cook = false).
    HcBuf d_staged; std::vector<crn_gpu_level_desc> staged; uint32_t staged_conv = 0;
    auto stage = [&](uint32_t conv) -> int {
        if (staged_conv == conv) return CRN_GPU_OK;
        CRN_CUDA(ctx, cudaSetDevice(ctx->device));
        if (!d_staged.p) {
            size_t bytes = 0;
            for (const crn_gpu_level_desc& d : lv) bytes += ((size_t)d.pitch_bytes * d.height + 255) & ~(size_t)255;
            if (d_staged.alloc(ctx, bytes) != cudaSuccess) return set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "crn_gpu_compress_dds: out of device memory");
        }
        staged = lv;
        size_t o = 0;
        for (crn_gpu_level_desc& d : staged) {
            uint8_t* dp = static_cast<uint8_t*>(d_staged.p) + o;
            CRN_CUDA(ctx, cudaMemcpyAsync(dp, d.rgba, (size_t)d.pitch_bytes * d.height, cudaMemcpyHostToDevice, ctx->stream));
            int r = crn_gpu_convert_pixels(ctx, dp, d.width, d.height, d.pitch_bytes, conv);
            if (r) return r;
            d.rgba = dp;
            o += ((size_t)d.pitch_bytes * d.height + 255) & ~(size_t)255;
        }
        staged_conv = conv;
        return CRN_GPU_OK;
    };
    // one compress_pass: convert_to_dxt + write_dds (+ the LZMA measurement when a rate is wanted)
    auto pass = [&](uint32_t quality, void** file_out, uint32_t* size_out, float* rate) -> int {
        uint8_t* file = static_cast<uint8_t*>(malloc(128 + payload));
        if (!file) return set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "crn_gpu_compress_dds: out of host memory");
        memcpy(file, header, 128);
        int r = CRN_GPU_OK;
        if (quality == 255 || fmt == CRN_GPU_FMT_DXT3) {                                                         // crn_dds_comp.cpp:150-157: block by block
            uint8_t* dst = file + 128;
            uint64_t done = 0;
            const uint32_t conv = cook ? cook : ((fmt == CRN_GPU_FMT_DXT5A && !has_alpha) ? (uint32_t)crn::kConvYtoA : 0u);
            if (conv) r = stage(conv);
            HcBuf d_blk;
            if (conv && r == CRN_GPU_OK && d_blk.alloc(ctx, (size_t)((lv[0].width + 3) >> 2) * ((lv[0].height + 3) >> 2) * bpb) != cudaSuccess)
                r = set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "crn_gpu_compress_dds: out of device memory");
            for (size_t li = 0; li < lv.size() && r == CRN_GPU_OK; li++) {
                const crn_gpu_level_desc& d = conv ? staged[li] : lv[li];
                r = progress_tick(ctx, 0, 1, (uint32_t)(payload ? done * 100 / payload : 0), 100);               // crn_dds_comp.cpp:130-134, :233-236
                if (r == CRN_GPU_OK && !conv) r = crn_gpu_pack_image_host(ctx, fmt, &pack, d.rgba, d.width, d.height, d.pitch_bytes, dst);
                if (r == CRN_GPU_OK && conv) {
                    r = crn_gpu_pack_image(ctx, fmt, &pack, d.rgba, d.width, d.height, d.pitch_bytes, d_blk.p);
                    if (r == CRN_GPU_OK) {
                        CRN_CUDA(ctx, cudaMemcpyAsync(dst, d_blk.p, (size_t)((d.width + 3) >> 2) * ((d.height + 3) >> 2) * bpb, cudaMemcpyDeviceToHost, ctx->stream));
                        CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                    }
                }
                if (r) break;
                const size_t sz = (size_t)((d.width + 3) >> 2) * ((d.height + 3) >> 2) * bpb;
                dst += sz; done += sz;
            }
            if (r == CRN_GPU_OK) r = progress_tick(ctx, 0, 1, 100, 100);
        } else {                                                                                                 // clustered: qdxt_pack_init once, qdxt_pack per pass
            const bool first = q == nullptr;
            if (first) {
                r = progress_tick(ctx, 0, 2, 0, 100);                                                            // crn_dds_comp.cpp:136-146, :172-188
                if (r == CRN_GPU_OK && cook) r = stage(cook);
                if (r == CRN_GPU_OK) r = crn_gpu_qdxt_init(ctx, fmt, &pack, cook ? staged.data() : lv.data(), count, cook ? 0 : 1, &q);
                // quality -> codebook size curve: deeper for the chroma of CCxY; .75 is for plain DXT5 only (crn_mipmapped_texture.cpp:2533-2542)
                if (r == CRN_GPU_OK && swizzled) q->pow_mul = p->crn_format == 3 ? 1.5f : 1.0f;
                if (r == CRN_GPU_OK && crn_gpu_qdxt_output_size(q) != payload) r = set_err(ctx, CRN_GPU_ERR_BAD_DATA, "crn_gpu_compress_dds: payload size mismatch");
            }
            if (r == CRN_GPU_OK) r = first ? progress_tick(ctx, 1, 2, 0, 100) : progress_tick(ctx, 0, 1, 0, 100);
            if (r == CRN_GPU_OK) r = crn_gpu_qdxt_pack(q, quality, file + 128, 1);
            if (r == CRN_GPU_OK) r = first ? progress_tick(ctx, 1, 2, 100, 100) : progress_tick(ctx, 0, 1, 100, 100);
        }
        if (r) { free(file); return r; }
        if (rate) {
            const uint64_t packed = lzma_packed_size(file, (size_t)(128 + payload));
            *rate = (packed && total_texels) ? (packed * 8.0f) / (float)total_texels : 0.0f;
        }
        *file_out = file; *size_out = (uint32_t)(128 + payload);
        return CRN_GPU_OK;
    };
    // create_compressed_texture (crnlib/crn_texture_comp.cpp:86-118): one pass unless a bitrate target applies
    if (!(p->target_bitrate > 0.0f) || fmt == CRN_GPU_FMT_DXT3) {
        float rate = 0.0f;
        rc = pass(p->quality_level, out_file, out_size, out_bitrate ? &rate : nullptr);
        if (q) crn_gpu_qdxt_free(q);
        if (rc) return rc;
        if (out_bitrate) *out_bitrate = rate;
        if (out_quality) *out_quality = p->quality_level;
        return CRN_GPU_OK;
    }
    if (!lzma_packed_size("crn", 3)) { if (q) crn_gpu_qdxt_free(q); return set_err(ctx, CRN_GPU_ERR_UNSUPPORTED, "crn_gpu_compress_dds: a target bitrate needs liblzma.so.5 for the size measurement"); }
    BitrateBest best;
    float highest = 0.0f;
    rc = bitrate_search(ctx, p->target_bitrate, pass, best, &highest);
    if (rc == CRN_GPU_OK && !pack.non_hierarchical && highest < p->target_bitrate && fabsf(best.bitrate - p->target_bitrate) >= .005f) {
        // "Unable to achieve desired bitrate - disabling adaptive block sizes and retrying search" (crn_texture_comp.cpp:232-250):
        // compress_init again without cCRNCompFlagHierarchical, a second search; the best so far stands unless a pass comes closer
        if (q) { crn_gpu_qdxt_free(q); q = nullptr; }
        pack.non_hierarchical = 1;
        rc = bitrate_search(ctx, p->target_bitrate, pass, best, &highest);
    }
    if (q) crn_gpu_qdxt_free(q);
    if (rc) return rc;
    *out_file = best.file; *out_size = best.size;
    if (out_bitrate) *out_bitrate = best.bitrate;
    if (out_quality) *out_quality = (uint32_t)best.quality;
    return CRN_GPU_OK;
}); }

/* ---- .dds in: read_dds + unpack_from_dxt ------------------------------------------------------------ */
namespace {
struct DdsParsed {
    crn_gpu_dds_desc d;
    bool fourcc;
    uint32_t pitch;                      // of level 0 (bytes per surface for block formats, per line otherwise)
    uint32_t uncook;                     // crn::PixelConversion applied after the unpack, 0 = none
    crn::DdsRawFormat raw;
};
constexpr uint32_t dds_cc(char a, char b, char c, char d) { return (uint32_t)(uint8_t)a | ((uint32_t)(uint8_t)b << 8) | ((uint32_t)(uint8_t)c << 16) | ((uint32_t)(uint8_t)d << 24); }
uint32_t popcount32(uint32_t m) { uint32_t n = 0; while (m) { m &= m - 1; n++; } return n; }          // math::bitmask_size
uint32_t ctz32(uint32_t m) { uint32_t n = 0; if (!m) return 0; while (!(m & 1)) { m >>= 1; n++; } return n; }   // math::bitmask_ofs

// mipmapped_texture::read_dds_internal up to the payload (crnlib/crn_mipmapped_texture.cpp:489-725)
int dds_parse(const void* h_dds, uint32_t size, DdsParsed& P)
{
    memset(&P, 0, sizeof(P));
    P.d.struct_size = sizeof(crn_gpu_dds_desc);
    if (!h_dds || size < 128) return CRN_GPU_ERR_BAD_DATA;
    uint32_t h[32];
    memcpy(h, h_dds, 128);
    if (h[0] != dds_cc('D', 'D', 'S', ' ') || h[1] != 124) return CRN_GPU_ERR_BAD_DATA;
    const uint32_t flags = h[2], height = h[3], width = h[4], caps = h[27], caps2 = h[28];
    const uint32_t pf_flags = h[20], cc = h[21], bitcount = h[22];
    if (!height || !width || height > 8192 || width > 8192) return CRN_GPU_ERR_BAD_DATA;
    uint32_t levels = 1;
    if ((flags & 0x20000u) && (caps & 0x400000u) && h[7]) {
        levels = h[7];
        uint32_t maxm = 1;
        for (uint32_t w = width, hh = height; w > 1 || hh > 1; w >>= 1, hh >>= 1) maxm++;               // utils::compute_max_mips
        if (levels > maxm) return CRN_GPU_ERR_BAD_DATA;
    }
    uint32_t faces = 1;
    if (caps & 0x8u) {
        if (caps2 & 0x200u) { if ((caps2 & 0xFC00u) != 0xFC00u) return CRN_GPU_ERR_UNSUPPORTED; faces = 6; }
        else if (caps2 & 0x200000u) return CRN_GPU_ERR_UNSUPPORTED;                                    // volume textures
    }
    if (pf_flags & 0x20u) return CRN_GPU_ERR_UNSUPPORTED;                                              // palettized
    const uint32_t RGBx = dds_cc('R', 'G', 'B', 'x'), RGBA = dds_cc('R', 'G', 'B', 'A'), Lx = dds_cc('L', 'x', 'x', 'x'), LA = dds_cc('L', 'x', 'x', 'A'), xA = dds_cc('x', 'x', 'x', 'A');
    P.d.width = width; P.d.height = height; P.d.levels = levels; P.d.faces = faces;
    P.d.block_format = 0xFFFFFFFFu;
    P.fourcc = (pf_flags & 0x4u) != 0;
    uint32_t bits_per_pixel = bitcount;
    if (P.fourcc) {
        uint32_t ff = cc, bf, out = RGBA, bpp = 8;
        if (cc == dds_cc('D', 'X', 'T', '1')) { bf = CRN_GPU_FMT_DXT1; out = RGBx; bpp = 4; }
        else if (cc == dds_cc('D', 'X', 'T', '2') || cc == dds_cc('D', 'X', 'T', '3')) { bf = CRN_GPU_FMT_DXT3; ff = dds_cc('D', 'X', 'T', '3'); }
        else if (cc == dds_cc('D', 'X', 'T', '4') || cc == dds_cc('D', 'X', 'T', '5')) {
            bf = CRN_GPU_FMT_DXT5; ff = dds_cc('D', 'X', 'T', '5');
            if (bitcount == dds_cc('C', 'C', 'x', 'Y')) { ff = bitcount; P.uncook = crn::kConvFromCCxY; out = RGBx; }
            else if (bitcount == dds_cc('x', 'G', 'x', 'R')) { ff = bitcount; P.uncook = crn::kConvFromxGxR; out = RGBx; }
            else if (bitcount == dds_cc('x', 'G', 'B', 'R')) { ff = bitcount; P.uncook = crn::kConvFromxGBR; out = RGBx; }
            else if (bitcount == dds_cc('A', 'G', 'B', 'R')) { ff = bitcount; P.uncook = crn::kConvFromAGBR; out = RGBA; }
        } else if (cc == dds_cc('A', 'T', 'I', '2')) {
            if (bitcount == dds_cc('A', '2', 'X', 'Y')) { bf = CRN_GPU_FMT_DXN_XY; ff = dds_cc('A', '2', 'X', 'Y'); } else bf = CRN_GPU_FMT_DXN_YX;
            P.uncook = crn::kConvXYtoXYZ; out = RGBx;
        } else if (cc == dds_cc('A', 'T', 'I', '1')) { bf = CRN_GPU_FMT_DXT5A; bpp = 4; }
        else return CRN_GPU_ERR_UNSUPPORTED;                                                           // ETC family, unknown FOURCCs
        P.d.block_format = bf; P.d.file_format = ff; P.d.pixel_format = out;
        bits_per_pixel = bpp;
    } else if (bitcount < 8 || bitcount > 32 || (bitcount & 7)) return CRN_GPU_ERR_UNSUPPORTED;
    else if (pf_flags & 0x40u) P.d.file_format = (pf_flags & 0x20000u) ? ((pf_flags & 1u) ? LA : Lx) : ((pf_flags & 1u) ? RGBA : RGBx);
    else if (pf_flags & 1u) P.d.file_format = (pf_flags & 0x20000u) ? LA : xA;
    else if (pf_flags & 0x20000u) P.d.file_format = Lx;
    else if (pf_flags & 2u) P.d.file_format = xA;
    else return CRN_GPU_ERR_UNSUPPORTED;
    if (!P.fourcc) P.d.pixel_format = P.d.file_format;
    const uint32_t default_pitch = P.fourcc ? ((((width + 3) & ~3u) * ((height + 3) & ~3u) * bits_per_pixel) >> 3) : ((width * bits_per_pixel) >> 3);
    uint32_t pitch = 0;
    if ((flags & 0x8u) && !(flags & 0x80000u)) pitch = h[5];
    if (!pitch) pitch = default_pitch;
    else if (pitch > default_pitch * 8) return CRN_GPU_ERR_BAD_DATA;
    P.pitch = pitch;
    if (!P.fourcc) {
        P.raw.bytes_per_pixel = bitcount >> 3;
        for (int i = 0; i < 4; i++) { P.raw.mask_size[i] = popcount32(h[23 + i]); P.raw.mask_ofs[i] = ctz32(h[23 + i]); }
        P.raw.luminance = (pf_flags & 0x20000u) ? 1u : 0u;
        if (P.raw.luminance && !P.raw.mask_size[0]) { P.raw.mask_size[0] = bitcount >> 3; if (pf_flags & 1u) P.raw.mask_size[0] /= 2; }   // :720-724 (sic: bytes, not bits)
    }
    return CRN_GPU_OK;
}
}  // namespace

int crn_gpu_dds_get_desc(const void* h_dds, uint32_t dds_size, crn_gpu_dds_desc* desc)
{ return crn_guard(nullptr, [&]() -> int {
    if (!desc || desc->struct_size != sizeof(crn_gpu_dds_desc)) return CRN_GPU_ERR_BAD_PARAM;
    DdsParsed P;
    const int rc = dds_parse(h_dds, dds_size, P);
    if (rc) return rc;
    *desc = P.d;
    return CRN_GPU_OK;
}); }

int crn_gpu_convert_pixels(crn_gpu_ctx* ctx, void* d_rgba, uint32_t width, uint32_t height, uint32_t pitch_bytes, uint32_t conversion)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx || !d_rgba || !width || !height || pitch_bytes < width * 4 || (pitch_bytes & 3) || conversion < 1 || conversion > 11) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_convert_pixels: bad argument");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = (uint64_t)width * height;
    CRN_LAUNCH(crn::pixel_convert_kernel, (uint32_t)((n + 255) / 256), 256, 0, ctx->stream, static_cast<uint8_t*>(d_rgba), width, height, pitch_bytes, conversion);
    ctx->launches++;
    CRN_CUDA(ctx, cudaGetLastError());
    return CRN_GPU_OK;
}); }

int crn_gpu_dds_to_images(crn_gpu_ctx* ctx, const void* h_dds, uint32_t dds_size, void* const* h_images, uint32_t num_images, crn_gpu_dds_desc* desc)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx || !h_images || (desc && desc->struct_size != sizeof(crn_gpu_dds_desc))) return CRN_GPU_ERR_BAD_PARAM;
    DdsParsed P;
    int rc = dds_parse(h_dds, dds_size, P);
    if (rc) return set_err(ctx, rc, "crn_gpu_dds_to_images: not a .dds file this path reads");
    const uint32_t faces = P.d.faces, levels = P.d.levels;
    if (num_images < faces * levels) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_dds_to_images: too few image pointers");
    for (uint32_t i = 0; i < faces * levels; i++) if (!h_images[i]) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_dds_to_images: null image pointer");
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    // payload layout: faces outermost, level 0 of each face padded to `pitch` (:741-776, :787-800)
    const uint32_t bpb = P.fourcc ? crn_gpu_bytes_per_block(P.d.block_format) : 0;
    struct Surf { uint64_t src_ofs, src_bytes, dst_ofs; uint32_t w, h, line_pitch; };
    std::vector<Surf> surfs(faces * levels);
    uint64_t ofs = 128, dst_total = 0;
    for (uint32_t f = 0; f < faces; f++)
        for (uint32_t l = 0; l < levels; l++) {
            Surf& s = surfs[l + levels * f];
            s.w = std::max(1u, P.d.width >> l); s.h = std::max(1u, P.d.height >> l);
            uint64_t actual, stored;
            if (P.fourcc) { actual = (uint64_t)((s.w + 3) >> 2) * ((s.h + 3) >> 2) * bpb; stored = l ? actual : std::max<uint64_t>(actual, P.pitch); s.line_pitch = 0; }
            else { const uint32_t line = s.w * P.raw.bytes_per_pixel; s.line_pitch = l ? line : P.pitch; actual = stored = (uint64_t)s.line_pitch * s.h; if (s.line_pitch < line) return set_err(ctx, CRN_GPU_ERR_BAD_DATA, "crn_gpu_dds_to_images: pitch below the line size"); }
            s.src_ofs = ofs; s.src_bytes = actual; ofs += stored;
            if (s.src_ofs + actual > dds_size) return set_err(ctx, CRN_GPU_ERR_BAD_DATA, "crn_gpu_dds_to_images: truncated file");
            s.dst_ofs = dst_total; dst_total += ((uint64_t)s.w * s.h * 4 + 255) & ~255ull;
        }
    const uint64_t payload = ofs > dds_size ? (uint64_t)dds_size - 128 : ofs - 128;
    rc = ensure(ctx, &ctx->d_in, &ctx->d_in_cap, (size_t)payload + 272);
    if (rc) return rc;
    rc = ensure(ctx, &ctx->d_out, &ctx->d_out_cap, (size_t)dst_total + 256);
    if (rc) return rc;
    uint8_t* d_src = static_cast<uint8_t*>(ctx->d_in);
    uint8_t* d_dst = static_cast<uint8_t*>(ctx->d_out);
    uint32_t* d_flag = reinterpret_cast<uint32_t*>(d_src + (((size_t)payload + 15) & ~(size_t)15));
    CRN_CUDA(ctx, cudaMemcpyAsync(d_src, static_cast<const uint8_t*>(h_dds) + 128, (size_t)payload, cudaMemcpyHostToDevice, ctx->stream));
    CRN_CUDA(ctx, cudaMemsetAsync(d_flag, 0, 4, ctx->stream));
    for (const Surf& s : surfs) {
        const uint8_t* src = d_src + (s.src_ofs - 128);
        uint8_t* dst = d_dst + s.dst_ofs;
        const uint64_t n = (uint64_t)s.w * s.h;
        if (P.fourcc) {
            if ((s.src_ofs - 128) & 7) return set_err(ctx, CRN_GPU_ERR_UNSUPPORTED, "crn_gpu_dds_to_images: surface not 8-byte aligned in the file");
            const uint32_t bx = (s.w + 3) >> 2, by = (s.h + 3) >> 2, nb = bx * by;
            // 16-byte formats are read as ulonglong2: d_in is 256-byte aligned and every surface size is a multiple of the block size
            if (bpb == 16 && ((s.src_ofs - 128) & 15)) return set_err(ctx, CRN_GPU_ERR_UNSUPPORTED, "crn_gpu_dds_to_images: surface not 16-byte aligned in the file");
            if (P.d.block_format == CRN_GPU_FMT_DXT1) {
                CRN_LAUNCH(crn::dxt1_has_alpha_kernel, (nb + 255) / 256, 256, 0, ctx->stream, reinterpret_cast<const unsigned long long*>(src), nb, d_flag);
                ctx->launches++;
            }
            CRN_LAUNCH(crn::unpack_blocks_kernel, (nb + 255) / 256, 256, 0, ctx->stream, reinterpret_cast<const unsigned long long*>(src), P.d.block_format, s.w, s.h, bx, nb, dst, s.w * 4);
            ctx->launches++;
            if (P.uncook) {
                CRN_LAUNCH(crn::pixel_convert_kernel, (uint32_t)((n + 255) / 256), 256, 0, ctx->stream, dst, s.w, s.h, s.w * 4, P.uncook);
                ctx->launches++;
            }
        } else {
            CRN_LAUNCH(crn::dds_raw_pixels_kernel, (uint32_t)((n + 255) / 256), 256, 0, ctx->stream, src, s.line_pitch, s.w, s.h, P.raw, reinterpret_cast<uint32_t*>(dst));
            ctx->launches++;
        }
    }
    CRN_CUDA(ctx, cudaGetLastError());
    uint32_t flag = 0;
    for (size_t i = 0; i < surfs.size(); i++)
        CRN_CUDA(ctx, cudaMemcpyAsync(h_images[i], d_dst + surfs[i].dst_ofs, (size_t)surfs[i].w * surfs[i].h * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CRN_CUDA(ctx, cudaMemcpyAsync(&flag, d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CRN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (P.d.block_format == CRN_GPU_FMT_DXT1 && flag) {                                                // change_dxt1_to_dxt1a (:866-882)
        P.d.block_format = CRN_GPU_FMT_DXT1A; P.d.file_format = dds_cc('D', 'X', '1', 'A'); P.d.pixel_format = dds_cc('R', 'G', 'B', 'A');
    }
    if (desc) *desc = P.d;
    return CRN_GPU_OK;
}); }

int crn_gpu_prepare_mip_source(crn_gpu_ctx* ctx, const crn_gpu_mip_source_params* sp, const crn_gpu_resample_params* mip, uint32_t faces, uint32_t width, uint32_t height,
                               const void* const* h_level0_faces, void** out_faces, uint32_t* out_width, uint32_t* out_height, uint32_t* out_changed)
{ return crn_guard(ctx, [&]() -> int {
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (!sp || sp->struct_size != sizeof(crn_gpu_mip_source_params) || !mip || mip->struct_size != sizeof(crn_gpu_resample_params) || (faces != 1 && faces != 6) || !width || !height ||
        width > 4096 || height > 4096 || !h_level0_faces || !out_faces || !out_width || !out_height || sp->scale_mode > 5)
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_prepare_mip_source: bad argument");
    for (uint32_t f = 0; f < faces; f++) { if (!h_level0_faces[f]) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_prepare_mip_source: missing image"); out_faces[f] = nullptr; }
    if (out_changed) *out_changed = 0;
    // the working image of face 0 while cropping (cubemaps are never cropped): origin + size inside the source
    uint32_t ox = 0, oy = 0, cw = width, ch = height;
    bool cropped = false;
    if (sp->window_right > sp->window_left && sp->window_bottom > sp->window_top && faces == 1) {          // rect::is_empty, :392-408
        const uint32_t x = sp->window_left, y = sp->window_top;
        if (x < cw && y < ch) { ox = x; oy = y; cw = sp->window_right - sp->window_left; ch = sp->window_bottom - sp->window_top; cropped = true; }   // image::extract_block refuses an origin outside
    }
    int new_w = (int)cw, new_h = (int)ch;
    const bool clamp = sp->clamp_width && sp->clamp_height;
    if (clamp && (new_w > (int)sp->clamp_width || new_h > (int)sp->clamp_height) && !sp->clamp_scale && faces == 1) {      // :413-433: clamp by cropping at the origin
        new_w = (int)std::min<uint32_t>(sp->clamp_width, (uint32_t)new_w); new_h = (int)std::min<uint32_t>(sp->clamp_height, (uint32_t)new_h);
        // mipmapped_texture::crop(0, 0, ...) of the (possibly already cropped) image
        cw = (uint32_t)new_w; ch = (uint32_t)new_h; cropped = true;
    }
    auto is_pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    auto lower = [](int v) { int r = 1; while (r * 2 <= v) r *= 2; return r; };
    auto upper = [&](int v) { if (is_pow2(v)) return v; int r = 1; while (r < v) r *= 2; return r; };
    if (sp->scale_mode) {                                                                                  // :435-499
        const bool p2 = is_pow2(new_w) && is_pow2(new_h);
        switch (sp->scale_mode) {
        case 1: new_w = (int)(uint32_t)sp->scale_x; new_h = (int)(uint32_t)sp->scale_y; break;
        case 2: new_w = (int)(uint32_t)(sp->scale_x * (float)new_w + .5f); new_h = (int)(uint32_t)(sp->scale_y * (float)new_h + .5f); break;
        case 3: if (!p2) { new_w = lower(new_w); new_h = lower(new_h); } break;
        case 4: if (!p2) {
                    const int lw = lower(new_w), lh = lower(new_h), uw = upper(new_w), uh = upper(new_h);
                    new_w = labs(new_w - lw) < labs(new_w - uw) ? lw : uw;
                    new_h = labs(new_h - lh) < labs(new_h - uh) ? lh : uh;
                }
                break;
        case 5: if (!p2) { new_w = upper(new_w); new_h = upper(new_h); } break;
        }
    }
    if (clamp && (new_w > (int)sp->clamp_width || new_h > (int)sp->clamp_height) && sp->clamp_scale) {       // :501-511
        new_w = (int)std::min<uint32_t>(sp->clamp_width, (uint32_t)new_w); new_h = (int)std::min<uint32_t>(sp->clamp_height, (uint32_t)new_h);
    }
    new_w = std::min(std::max(new_w, 1), 4096); new_h = std::min(std::max(new_h, 1), 4096);                  // cCRNMaxLevelResolution
    const bool resize = new_w != (int)cw || new_h != (int)ch || (mip->renormalize && sp->rtopmip);
    *out_width = (uint32_t)new_w; *out_height = (uint32_t)new_h;
    if (!cropped && !resize) return CRN_GPU_OK;
    if (out_changed) *out_changed = 1;
    CRN_CUDA(ctx, cudaSetDevice(ctx->device));
    crn_gpu_resample_params rp = *mip;
    rp.filter_scale = 1.0f; rp.wrapping = 0;                                                                 // :522-529 (m_wrapping is cleared whenever the texture has faces, i.e. always)
    auto release = [&]() { for (uint32_t f = 0; f < faces; f++) { free(out_faces[f]); out_faces[f] = nullptr; } };
    for (uint32_t f = 0; f < faces; f++) {
        // crop on the host side of the copy: rows of the window, clamped reads past the source edge (extract_block's get_clamped)
        std::vector<uint8_t> win;
        const uint8_t* src = static_cast<const uint8_t*>(h_level0_faces[f]);
        const uint8_t* img = src; uint32_t iw = width, ih = height;
        if (cropped) {
            win.resize((size_t)cw * ch * 4);
            for (uint32_t y = 0; y < ch; y++) {
                const uint32_t sy = std::min(oy + y, height - 1);
                for (uint32_t x = 0; x < cw; x++) memcpy(&win[((size_t)y * cw + x) * 4], src + ((size_t)sy * width + std::min(ox + x, width - 1)) * 4, 4);
            }
            img = win.data(); iw = cw; ih = ch;
        }
        uint8_t* dst = static_cast<uint8_t*>(malloc((size_t)new_w * new_h * 4));
        if (!dst) { release(); return set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "crn_gpu_prepare_mip_source: out of host memory"); }
        out_faces[f] = dst;
        if (!resize) { memcpy(dst, img, (size_t)iw * ih * 4); continue; }
        if (!rp.num_comps) {                                                                                 // is_component_valid(3): any alpha below 255 in the source
            bool has_alpha = false;
            for (uint32_t g = 0; g < faces && !has_alpha; g++) {
                const uint8_t* px = static_cast<const uint8_t*>(h_level0_faces[g]);
                for (size_t i = 0, n = (size_t)width * height; i < n; i++) if (px[i * 4 + 3] < 255) { has_alpha = true; break; }
            }
            rp.num_comps = has_alpha ? 4 : 3;
        }
        HcBuf d_in, d_out;
        if (d_in.alloc(ctx, (size_t)iw * ih * 4) != cudaSuccess || d_out.alloc(ctx, (size_t)new_w * new_h * 4) != cudaSuccess) { release(); return set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "crn_gpu_prepare_mip_source: out of device memory"); }
        int rc = CRN_GPU_OK;
        if (cudaMemcpyAsync(d_in.p, img, (size_t)iw * ih * 4, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = CRN_GPU_ERR_CUDA;
        if (rc == CRN_GPU_OK) rc = crn_gpu_resample(ctx, &rp, d_in.p, iw, ih, iw * 4, d_out.p, (uint32_t)new_w, (uint32_t)new_h, (uint32_t)new_w * 4);
        if (rc == CRN_GPU_OK && rp.renormalize) rc = crn_gpu_convert_pixels(ctx, d_out.p, (uint32_t)new_w, (uint32_t)new_h, (uint32_t)new_w * 4, crn::kConvRenormNormalMap);
        if (rc == CRN_GPU_OK && (cudaMemcpyAsync(dst, d_out.p, (size_t)new_w * new_h * 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess)) rc = CRN_GPU_ERR_CUDA;
        if (rc) { release(); return rc == CRN_GPU_ERR_CUDA ? set_err(ctx, rc, "crn_gpu_prepare_mip_source: copy") : rc; }
    }
    return CRN_GPU_OK;
}); }

int crn_gpu_compress_mip_chain(crn_gpu_ctx* ctx, uint32_t file_type, const crn_gpu_crn_params* cp, const crn_gpu_dds_params* dp, const crn_gpu_resample_params* mip,
                               uint32_t min_mip_size, uint32_t max_levels, const void* const* h_level0_faces, void** out_file, uint32_t* out_size)
{ return crn_guard(ctx, [&]() -> int {   // crn_compress(const crn_comp_params&, const crn_mipmap_params&, ...) (inc/crnlib.h:614): create_texture_mipmaps in generate mode
    // (crnlib/crn_texture_comp.cpp:352-575) -> generate_mipmaps -> the compressor of the file type
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (out_file) *out_file = nullptr;
    if (out_size) *out_size = 0;
    const bool crn = file_type == 0;
    if (file_type > 1 || (crn ? !crn_params_ok(cp) : (!dp || dp->struct_size != sizeof(crn_gpu_dds_params))) || !h_level0_faces || !out_file || !out_size ||
        (mip && mip->struct_size != sizeof(crn_gpu_resample_params)))
        return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_compress_mip_chain: bad argument");
    const uint32_t width = crn ? cp->width : dp->width, height = crn ? cp->height : dp->height, faces = crn ? cp->faces : dp->faces;
    if (!width || !height || width > 4096 || height > 4096 || (faces != 1 && faces != 6)) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_compress_mip_chain: bad size");
    for (uint32_t f = 0; f < faces; f++) if (!h_level0_faces[f]) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_compress_mip_chain: missing image");
    crn_gpu_resample_params rp;
    if (mip) rp = *mip;
    else { crn_gpu_default_resample_params(&rp); rp.filter = 4; rp.filter_scale = .9f; rp.srgb = 1; rp.source_gamma = 2.2f; rp.wrapping = 0; rp.num_comps = 0; }   // crn_mipmap_params::clear()
    if (!rp.num_comps) {                                         // alpha is filtered only when the source has any (is_component_valid(3))
        bool has_alpha = false;
        for (uint32_t f = 0; f < faces && !has_alpha; f++) {
            const uint8_t* px = static_cast<const uint8_t*>(h_level0_faces[f]);
            for (size_t i = 0, n = (size_t)width * height; i < n; i++) if (px[i * 4 + 3] < 255) { has_alpha = true; break; }
        }
        rp.num_comps = has_alpha ? 4 : 3;
    }
    const uint32_t levels = crn_gpu_mip_level_count(width, height, min_mip_size ? min_mip_size : 1, max_levels ? std::min(max_levels, 16u) : 16u);
    size_t mip_bytes = 0;
    for (uint32_t l = 1; l < levels; l++) mip_bytes += (size_t)std::max(1u, width >> l) * std::max(1u, height >> l) * 4;
    std::vector<std::vector<uint8_t>> chains(faces);
    std::vector<const void*> images((size_t)faces * levels);
    for (uint32_t f = 0; f < faces; f++) {
        chains[f].resize(mip_bytes ? mip_bytes : 1);
        uint32_t n = 0;
        int rc = crn_gpu_generate_mipmaps_host(ctx, &rp, h_level0_faces[f], width, height, width * 4, min_mip_size ? min_mip_size : 1, max_levels ? std::min(max_levels, 16u) : 16u,
                                               chains[f].data(), chains[f].size(), &n);
        if (rc) return rc;
        if (n != levels) return set_err(ctx, CRN_GPU_ERR_BAD_DATA, "crn_gpu_compress_mip_chain: level count mismatch");
        images[(size_t)f * levels] = h_level0_faces[f];
        size_t at = 0;
        for (uint32_t l = 1; l < levels; l++) { images[(size_t)f * levels + l] = chains[f].data() + at; at += (size_t)std::max(1u, width >> l) * std::max(1u, height >> l) * 4; }
    }
    if (crn) { crn_gpu_crn_params q = *cp; q.levels = levels; return crn_gpu_compress_crn(ctx, &q, images.data(), out_file, out_size, nullptr, nullptr); }
    crn_gpu_dds_params q = *dp; q.levels = levels;
    return crn_gpu_compress_dds(ctx, &q, images.data(), out_file, out_size);
}); }

int crn_gpu_crn_to_dds(crn_gpu_ctx* ctx, const void* h_crn, uint32_t crn_size, void** out_file, uint32_t* out_size)
{ return crn_guard(ctx, [&]() -> int {   // crn_decompress_crn_to_dds (crnlib/crnlib.cpp:269-291): transcode every level on the device, lay the faces out DDS-style
    if (!ctx) return CRN_GPU_ERR_BAD_PARAM;
    if (out_file) *out_file = nullptr;
    if (out_size) *out_size = 0;
    if (!h_crn || !out_file || !out_size) return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_crn_to_dds: bad argument");
    crn_gpu_texture_info ti; ti.struct_size = sizeof(ti);
    int rc = crn_gpu_crnd_get_texture_info(h_crn, crn_size, &ti);
    if (rc) return set_err(ctx, rc, "crn_gpu_crn_to_dds: not a CRN file");
    uint8_t header[128];
    rc = crn_gpu_dds_header(ti.format, ti.width, ti.height, ti.levels, ti.faces, header);
    if (rc) return set_err(ctx, rc, "crn_gpu_crn_to_dds: format has no DDS form here");
    crn_gpu_texture* tex = nullptr;
    rc = crn_gpu_crnd_unpack_begin(ctx, h_crn, crn_size, &tex);
    if (rc) return rc;
    const uint64_t total = crn_gpu_crnd_total_size(tex);
    if (128 + total > 0xFFFFFFFFull) { crn_gpu_crnd_unpack_end(tex); return set_err(ctx, CRN_GPU_ERR_BAD_PARAM, "crn_gpu_crn_to_dds: file would exceed 4 GiB (crn_uint32 size)"); }
    uint8_t* file = static_cast<uint8_t*>(malloc(128 + total));
    uint8_t* tmp = static_cast<uint8_t*>(malloc(total ? total : 1));
    if (!file || !tmp) { free(file); free(tmp); crn_gpu_crnd_unpack_end(tex); return set_err(ctx, CRN_GPU_ERR_NO_MEMORY, "crn_gpu_crn_to_dds: out of host memory"); }
    rc = crn_gpu_crnd_unpack_all_levels_host(tex, tmp, total);
    if (rc == CRN_GPU_OK) {
        memcpy(file, header, 128);
        uint8_t* dst = file + 128;
        for (uint32_t f = 0; f < ti.faces; f++)                  // write_dds: faces outermost (crn_mipmapped_texture.cpp:1093-1094)
            for (uint32_t l = 0; l < ti.levels; l++) {
                const uint32_t w = std::max(1u, ti.width >> l), h = std::max(1u, ti.height >> l);
                const size_t bytes = (size_t)((w + 3) >> 2) * ((h + 3) >> 2) * ti.bytes_per_block;
                memcpy(dst, tmp + crn_gpu_crnd_level_offset(tex, l, f), bytes);
                dst += bytes;
            }
        *out_file = file; *out_size = (uint32_t)(128 + total);
        file = nullptr;
    }
    free(file); free(tmp);
    crn_gpu_crnd_unpack_end(tex);
    return rc;
}); }

}  // extern "C"
