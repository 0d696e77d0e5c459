// tests/cusim/cusim.h -- TEST INFRASTRUCTURE ONLY: a tiny SIMT emulator.
//
// The development container has nvcc but no GPU.  To debug the sm_100a kernels
// bit-for-bit against the oracle before spending GPU minutes, the SAME .cu/.cuh
// sources are also compiled by g++ against this header.  Every CUDA thread of a
// CTA becomes a fiber (hand-rolled x86-64 context switch); warp collectives
// (__shfl*_sync, __ballot_sync, __reduce_*_sync, __syncwarp ...) and
// __syncthreads() are rendez-vous points between fibers, honouring the member
// mask exactly like independent thread scheduling does.  CTAs run one after
// another, so `__shared__` maps to function-local static storage.
//
// The emulated build (libcrn_b200_sim.so) exists only for tests/ (-m "not gpu");
// the product library is built by nvcc, contains none of this, and the Python
// package refuses to load anything but the nvcc build.
#pragma once
#ifdef __CUDACC__
#error "cusim.h is for the g++ emulation build only"
#endif
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <sys/mman.h>
#include <algorithm>
#include <functional>
#include <type_traits>
#include <vector>

#define CUSIM 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __constant__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#ifndef __restrict__
#define __restrict__ __restrict
#endif

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uchar4 { unsigned char x, y, z, w; };
struct ushort2 { unsigned short x, y; };
struct __align__(8) uint2 { unsigned x, y; };
struct __align__(16) uint4 { unsigned x, y, z, w; };
struct __align__(8) int2 { int x, y; };
struct __align__(16) int4 { int x, y, z, w; };
struct __align__(8) float2 { float x, y; };
struct __align__(16) float4 { float x, y, z, w; };
struct __align__(16) ulonglong2 { unsigned long long x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r = { x, y }; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r = { x, y, z, w }; return r; }
static inline int2 make_int2(int x, int y) { int2 r = { x, y }; return r; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 r = { x, y, z, w }; return r; }
static inline float2 make_float2(float x, float y) { float2 r = { x, y }; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = { x, y, z, w }; return r; }
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { uchar4 r = { x, y, z, w }; return r; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { ulonglong2 r = { x, y }; return r; }

extern "C" void cusim_swap(void** save_sp, void* load_sp);

namespace cusim {

enum State { RUNNABLE, WAIT_WARP, WAIT_CTA, DONE };
enum Op { OP_SYNCWARP, OP_SHFL_IDX, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_BALLOT, OP_ANY, OP_ALL,
          OP_RED_MIN_U, OP_RED_MAX_U, OP_RED_MIN_S, OP_RED_MAX_S, OP_RED_ADD, OP_RED_AND, OP_RED_OR, OP_RED_XOR, OP_MATCH_ANY };

struct Fiber {
    void* sp;
    char* stack;
    State state;
    int op;
    unsigned mask;
    uint64_t val;
    int arg, width;
    uint64_t result;
    uint3 tid;
    unsigned lin;  // linear thread id in CTA
};

struct Cta {
    std::vector<Fiber> fibers;
    void* sched_sp;
    Fiber* cur;
    uint3 bid;
    dim3 bdim, gdim;
    unsigned nthreads;
    void* dyn_smem;
    const std::function<void()>* body;
    size_t stack_bytes;
    char* stack_pool;
    size_t stack_pool_bytes;
    unsigned long long collectives;
};

inline Cta g_cta;

inline Fiber* cur() { return g_cta.cur; }
inline void* dyn_smem() { return g_cta.dyn_smem; }

[[noreturn]] inline void die(const char* msg)
{
    Fiber* f = g_cta.cur;
    fprintf(stderr, "cusim: %s (block %u,%u,%u thread %u)\n", msg, g_cta.bid.x, g_cta.bid.y, g_cta.bid.z, f ? f->lin : 0u);
    abort();
}

inline void to_sched()
{
    Fiber* f = g_cta.cur;
    cusim_swap(&f->sp, g_cta.sched_sp);
}

inline void fiber_main()
{
    (*g_cta.body)();
    g_cta.cur->state = DONE;
    to_sched();
    die("resumed a finished fiber");
}

// Try to complete the warp collective lane `f` is waiting on.  Returns true if completed.
inline bool try_complete_warp(Fiber* f)
{
    Cta& c = g_cta;
    const unsigned warp_base = f->lin & ~31u;
    const unsigned mask = f->mask;
    for (unsigned l = 0; l < 32; l++) {
        if (!(mask >> l & 1)) continue;
        unsigned t = warp_base + l;
        if (t >= c.nthreads) die("collective mask names a lane beyond the CTA");
        Fiber& g = c.fibers[t];
        if (g.state == DONE) die("collective mask names an exited lane");
        if (g.state != WAIT_WARP) return false;
        // a member may still be parked at an EARLIER collective with a different mask / op (it will get
        // here once that one completes); a genuine mismatch ends up in the deadlock detector.
        if (g.mask != mask || g.op != f->op) return false;
    }
    c.collectives++;
    // all members present: compute results
    uint64_t red = 0;
    bool first = true;
    const int op = f->op;
    if (op >= OP_BALLOT && op != OP_MATCH_ANY) {
        for (unsigned l = 0; l < 32; l++) {
            if (!(mask >> l & 1)) continue;
            uint64_t v = c.fibers[warp_base + l].val;
            switch (op) {
            case OP_BALLOT: red |= (uint64_t)(v != 0) << l; break;
            case OP_ANY: red |= (v != 0); break;
            case OP_ALL: red = first ? (v != 0) : (red & (uint64_t)(v != 0)); break;
            case OP_RED_MIN_U: red = first ? v : std::min<uint64_t>(red, v); break;
            case OP_RED_MAX_U: red = first ? v : std::max<uint64_t>(red, v); break;
            case OP_RED_MIN_S: red = first ? v : (uint64_t)std::min<int64_t>((int64_t)red, (int64_t)v); break;
            case OP_RED_MAX_S: red = first ? v : (uint64_t)std::max<int64_t>((int64_t)red, (int64_t)v); break;
            case OP_RED_ADD: red += v; break;
            case OP_RED_AND: red = first ? v : (red & v); break;
            case OP_RED_OR: red |= v; break;
            case OP_RED_XOR: red ^= v; break;
            }
            first = false;
        }
    }
    for (unsigned l = 0; l < 32; l++) {
        if (!(mask >> l & 1)) continue;
        Fiber& g = c.fibers[warp_base + l];
        switch (op) {
        case OP_SYNCWARP: g.result = 0; break;
        case OP_SHFL_IDX: case OP_SHFL_UP: case OP_SHFL_DOWN: case OP_SHFL_XOR: {
            const int w = g.width;
            const int seg = (int)l & ~(w - 1);
            int src;
            bool own = false;
            if (op == OP_SHFL_IDX) src = seg | (g.arg & (w - 1));
            else if (op == OP_SHFL_UP) { src = (int)l - g.arg; if (src < seg) own = true; }
            else if (op == OP_SHFL_DOWN) { src = (int)l + g.arg; if (src >= seg + w) own = true; }
            else { src = (int)l ^ g.arg; if (src >= seg + w) own = true; }
            if (own) g.result = g.val;
            else {
                if (!(mask >> src & 1)) die("shuffle reads a lane outside the member mask");
                g.result = c.fibers[warp_base + src].val;
            }
            break;
        }
        case OP_MATCH_ANY: {
            uint64_t m = 0;
            for (unsigned k = 0; k < 32; k++)
                if ((mask >> k & 1) && c.fibers[warp_base + k].val == g.val) m |= 1ull << k;
            g.result = m;
            break;
        }
        default: g.result = red; break;
        }
    }
    for (unsigned l = 0; l < 32; l++)
        if (mask >> l & 1) c.fibers[warp_base + l].state = RUNNABLE;
    return true;
}

inline uint64_t warp_collective(int op, unsigned mask, uint64_t val, int arg = 0, int width = 32)
{
    Fiber* f = g_cta.cur;
    const unsigned lane = f->lin & 31;
    if (!(mask >> lane & 1)) die("calling lane is not in its own member mask");
    f->op = op; f->mask = mask; f->val = val; f->arg = arg; f->width = width;
    f->state = WAIT_WARP;
    if (!try_complete_warp(f)) {
        to_sched();
        if (f->state != RUNNABLE) die("fiber resumed while still waiting");
    }
    return f->result;
}

inline void cta_barrier()
{
    Cta& c = g_cta;
    Fiber* f = c.cur;
    f->state = WAIT_CTA;
    bool all = true;
    for (unsigned t = 0; t < c.nthreads; t++)
        if (c.fibers[t].state != WAIT_CTA && c.fibers[t].state != DONE) { all = false; break; }
    if (all) {
        for (unsigned t = 0; t < c.nthreads; t++)
            if (c.fibers[t].state == WAIT_CTA) c.fibers[t].state = RUNNABLE;
        return;
    }
    to_sched();
}

inline void run_cta()
{
    Cta& c = g_cta;
    for (unsigned t = 0; t < c.nthreads; t++) {
        Fiber& f = c.fibers[t];
        f.stack = c.stack_pool + (size_t)t * c.stack_bytes;
        uint64_t* top = (uint64_t*)(f.stack + c.stack_bytes);
        top -= 8;  // r15 r14 r13 r12 rbx rbp ret dummy
        for (int i = 0; i < 6; i++) top[i] = 0;
        top[6] = (uint64_t)(void (*)())fiber_main;
        top[7] = 0;
        f.sp = top;
        f.state = RUNNABLE;
        f.lin = t;
        f.tid.x = t % c.bdim.x;
        f.tid.y = (t / c.bdim.x) % c.bdim.y;
        f.tid.z = t / (c.bdim.x * c.bdim.y);
    }
    for (;;) {
        bool progressed = false, alive = false;
        for (unsigned t = 0; t < c.nthreads; t++) {
            Fiber& f = c.fibers[t];
            if (f.state == DONE) continue;
            alive = true;
            if (f.state == WAIT_WARP) { c.cur = &f; try_complete_warp(&f); }
            if (f.state != RUNNABLE) continue;
            c.cur = &f;
            cusim_swap(&c.sched_sp, f.sp);
            progressed = true;
        }
        if (!alive) break;
        if (!progressed) {
            // maybe a CTA barrier became releasable because the last non-waiting thread exited
            bool all = true, any = false;
            for (unsigned t = 0; t < c.nthreads; t++) {
                if (c.fibers[t].state == WAIT_CTA) any = true;
                else if (c.fibers[t].state != DONE) all = false;
            }
            if (all && any) {
                for (unsigned t = 0; t < c.nthreads; t++)
                    if (c.fibers[t].state == WAIT_CTA) c.fibers[t].state = RUNNABLE;
                continue;
            }
            c.cur = nullptr;
            for (unsigned t = 0; t < c.nthreads; t++)
                fprintf(stderr, "  thread %u state %d op %d mask %08x\n", t, (int)c.fibers[t].state, c.fibers[t].op, c.fibers[t].mask);
            die("deadlock: no runnable fiber");
        }
    }
    c.cur = nullptr;
}

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body)
{
    Cta& c = g_cta;
    c.nthreads = block.x * block.y * block.z;
    c.bdim = block; c.gdim = grid;
    c.body = &body;
    c.stack_bytes = 256 * 1024;
    size_t need = c.stack_bytes * c.nthreads;
    if (need > c.stack_pool_bytes) {
        if (c.stack_pool) munmap(c.stack_pool, c.stack_pool_bytes);
        c.stack_pool = (char*)mmap(nullptr, need, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (c.stack_pool == MAP_FAILED) { perror("cusim mmap"); abort(); }
        c.stack_pool_bytes = need;
    }
    c.fibers.resize(c.nthreads);
    std::vector<unsigned char> smem(smem_bytes + 128);
    c.dyn_smem = (void*)(((uintptr_t)smem.data() + 127) & ~(uintptr_t)127);
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                c.bid.x = bx; c.bid.y = by; c.bid.z = bz;
                run_cta();
            }
}

template <typename T> inline uint64_t to_bits(T v)
{
    static_assert(sizeof(T) <= 8, "shuffle operand too wide");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T> inline T from_bits(uint64_t b)
{
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

}  // namespace cusim

#define threadIdx (cusim::g_cta.cur->tid)
#define blockIdx (cusim::g_cta.bid)
#define blockDim (cusim::g_cta.bdim)
#define gridDim (cusim::g_cta.gdim)
#define warpSize 32

static inline void __syncthreads() { cusim::cta_barrier(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { cusim::warp_collective(cusim::OP_SYNCWARP, mask, 0); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline unsigned __activemask() { cusim::die("__activemask is not emulated (its result is scheduling dependent)"); }

template <typename T> static inline T __shfl_sync(unsigned m, T v, int src, int w = 32)
{ return cusim::from_bits<T>(cusim::warp_collective(cusim::OP_SHFL_IDX, m, cusim::to_bits(v), src, w)); }
template <typename T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d, int w = 32)
{ return cusim::from_bits<T>(cusim::warp_collective(cusim::OP_SHFL_UP, m, cusim::to_bits(v), (int)d, w)); }
template <typename T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d, int w = 32)
{ return cusim::from_bits<T>(cusim::warp_collective(cusim::OP_SHFL_DOWN, m, cusim::to_bits(v), (int)d, w)); }
template <typename T> static inline T __shfl_xor_sync(unsigned m, T v, int lm, int w = 32)
{ return cusim::from_bits<T>(cusim::warp_collective(cusim::OP_SHFL_XOR, m, cusim::to_bits(v), lm, w)); }
static inline unsigned __ballot_sync(unsigned m, int p) { return (unsigned)cusim::warp_collective(cusim::OP_BALLOT, m, p != 0); }
static inline int __any_sync(unsigned m, int p) { return (int)cusim::warp_collective(cusim::OP_ANY, m, p != 0); }
static inline int __all_sync(unsigned m, int p) { return (int)cusim::warp_collective(cusim::OP_ALL, m, p != 0); }
static inline unsigned __reduce_min_sync(unsigned m, unsigned v) { return (unsigned)cusim::warp_collective(cusim::OP_RED_MIN_U, m, v); }
static inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return (unsigned)cusim::warp_collective(cusim::OP_RED_MAX_U, m, v); }
static inline int __reduce_min_sync(unsigned m, int v) { return (int)(int64_t)cusim::warp_collective(cusim::OP_RED_MIN_S, m, (uint64_t)(int64_t)v); }
static inline int __reduce_max_sync(unsigned m, int v) { return (int)(int64_t)cusim::warp_collective(cusim::OP_RED_MAX_S, m, (uint64_t)(int64_t)v); }
static inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return (unsigned)cusim::warp_collective(cusim::OP_RED_ADD, m, v); }
static inline int __reduce_add_sync(unsigned m, int v) { return (int)cusim::warp_collective(cusim::OP_RED_ADD, m, (uint64_t)(int64_t)v); }
static inline unsigned __reduce_and_sync(unsigned m, unsigned v) { return (unsigned)cusim::warp_collective(cusim::OP_RED_AND, m, v); }
static inline unsigned __reduce_or_sync(unsigned m, unsigned v) { return (unsigned)cusim::warp_collective(cusim::OP_RED_OR, m, v); }
static inline unsigned __reduce_xor_sync(unsigned m, unsigned v) { return (unsigned)cusim::warp_collective(cusim::OP_RED_XOR, m, v); }
template <typename T> static inline unsigned __match_any_sync(unsigned m, T v)
{ return (unsigned)cusim::warp_collective(cusim::OP_MATCH_ANY, m, cusim::to_bits(v)); }

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned __brev(unsigned v)
{
    v = (v >> 16) | (v << 16);
    v = ((v & 0xff00ff00u) >> 8) | ((v & 0x00ff00ffu) << 8);
    v = ((v & 0xf0f0f0f0u) >> 4) | ((v & 0x0f0f0f0fu) << 4);
    v = ((v & 0xccccccccu) >> 2) | ((v & 0x33333333u) << 2);
    v = ((v & 0xaaaaaaaau) >> 1) | ((v & 0x55555555u) << 1);
    return v;
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s)
{
    uint64_t ab = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)(ab >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
static inline float __int_as_float(int v) { return cusim::from_bits<float>((uint32_t)v); }
static inline int __float_as_int(float v) { return (int)cusim::to_bits(v); }
static inline unsigned __float_as_uint(float v) { return (unsigned)cusim::to_bits(v); }
static inline float __uint_as_float(unsigned v) { return cusim::from_bits<float>(v); }
static inline double __longlong_as_double(long long v) { return cusim::from_bits<double>((uint64_t)v); }
static inline long long __double_as_longlong(double v) { return (long long)cusim::to_bits(v); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline int __float2int_rz(float a) { return (int)a; }
static inline int __double2int_rz(double a) { return (int)a; }
using std::min;
using std::max;
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }

// sequential emulation -> atomics are plain read-modify-writes
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicSub(T* p, T v) { T o = *p; *p = o - v; return o; }
template <typename T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <typename T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

// kernel launch + dynamic shared memory, spelled the same way in both builds (see csrc/launch.h)
#define CRN_LAUNCH(kernel, grid, block, smem, stream, ...) \
    cusim::launch(dim3(grid), dim3(block), (smem), [&]() { kernel(__VA_ARGS__); })
#define CRN_DYN_SMEM(type, name) type* name = (type*)cusim::dyn_smem()
