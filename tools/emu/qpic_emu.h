// DEVELOPMENT / TEST TOOL ONLY -- NOT a CPU fallback and never loaded by the product.
//
// The build container has nvcc but no GPU.  To debug kernel LOGIC (indexing, barriers, shuffles) before
// spending GPU-box minutes, this header lets the very same .cu sources compile with g++ (-DQPIC_EMU) into a
// separate debug library, tools/emu/_build/libqpic_emu_debug.so, in which
//   * a kernel launch runs its blocks one after another;
//   * the threads of a block are cooperative fibers (ucontext); __syncthreads / warp collectives
//     yield until every live thread of the block / warp has arrived;
//   * cudaMalloc & friends are malloc/memcpy, streams / events / graphs are no-ops.
// Threads never run concurrently, so data races are NOT detected here (compute-sanitizer on the GPU box
// does that).  quetzalpic_b200's loader only ever opens libqpic_b200.so, which requires a CUDA device.
#pragma once

#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

// ---------------------------------------------------------------- vector types / qualifiers
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3 {
    unsigned x, y, z;
};
struct double2 {
    double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct int2 {
    int x, y;
};
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __constant__ static

namespace qemu {

struct Fiber {
    ucontext_t ctx;
    char * stack = nullptr;
    bool done = false;
    bool started = false;
    uint3 tid{0, 0, 0};
    int linear = 0;
    // collective state
    unsigned long long bar_gen_seen = 0;
    alignas(16) unsigned char xchg[16];
};

struct Block {
    dim3 grid, block;
    uint3 bid{0, 0, 0};
    std::vector<Fiber> fibers;
    int nthreads = 0;
    int live = 0;
    int bar_arrived = 0;
    unsigned long long bar_gen = 0;
    std::vector<int> warp_arrived;
    std::vector<unsigned long long> warp_gen;
    std::vector<int> warp_live;
    char * dyn_smem = nullptr;
    bool flat = false;  // barrier-free kernel executed as plain calls
    std::function<void()> body;
    ucontext_t sched;
};

extern thread_local Block * g_block;
extern thread_local Fiber * g_fiber;

void launch(dim3 grid, dim3 block, size_t smem, bool flat, const std::function<void()> & body);
void yield_fiber();
void sync_threads();
void sync_warp();

template <class T>
inline T warp_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 16, "exchange slot too small");
    Block * b = g_block;
    Fiber * f = g_fiber;
    if (b->flat) {
        std::fprintf(stderr, "qemu: warp collective inside a kernel launched as FLAT\n");
        std::abort();
    }
    std::memcpy(f->xchg, &v, sizeof(T));
    sync_warp();
    const int warp_base = (f->linear / 32) * 32;
    int src = warp_base + (src_lane & 31);
    T r = v;
    if (src < b->nthreads && !b->fibers[src].done) std::memcpy(&r, b->fibers[src].xchg, sizeof(T));
    sync_warp();
    return r;
}

}  // namespace qemu

#define threadIdx (qemu::g_fiber->tid)
#define blockIdx (qemu::g_block->bid)
#define blockDim (qemu::g_block->block)
#define gridDim (qemu::g_block->grid)
#define warpSize 32

static inline void __syncthreads() { qemu::sync_threads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { qemu::sync_warp(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}

static inline unsigned __activemask() { return 0xffffffffu; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __ballot_sync(unsigned, int pred);
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
template <class T>
static inline T __shfl_sync(unsigned, T v, int lane, int = 32) {
    return qemu::warp_exchange(v, lane);
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int mask, int = 32) {
    return qemu::warp_exchange(v, (qemu::g_fiber->linear & 31) ^ mask);
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned delta, int = 32) {
    const int lane = qemu::g_fiber->linear & 31;
    const int src = lane + (int)delta;
    T r = qemu::warp_exchange(v, src < 32 ? src : lane);
    return src < 32 ? r : v;
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned delta, int = 32) {
    const int lane = qemu::g_fiber->linear & 31;
    const int src = lane - (int)delta;
    T r = qemu::warp_exchange(v, src >= 0 ? src : lane);
    return src >= 0 ? r : v;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned bits = 0;
    for (int l = 0; l < 32; l++) {
        const int p = qemu::warp_exchange(pred ? 1 : 0, l);
        // lanes beyond the block / already finished return their own value from warp_exchange; mask them
        const int idx = (qemu::g_fiber->linear / 32) * 32 + l;
        if (idx < qemu::g_block->nthreads && !qemu::g_block->fibers[idx].done && p) bits |= (1u << l);
    }
    return bits;
}

// ---------------------------------------------------------------- atomics / loads / math
template <class T>
static inline T atomicAdd(T * addr, T v) {
    T old = *addr;
    *addr = old + v;
    return old;
}
static inline int atomicOr(int * addr, int v) {
    int old = *addr;
    *addr = old | v;
    return old;
}
static inline unsigned long long atomicOr(unsigned long long * addr, unsigned long long v) {
    unsigned long long old = *addr;
    *addr = old | v;
    return old;
}
static inline int atomicMax(int * addr, int v) {
    int old = *addr;
    if (v > old) *addr = v;
    return old;
}
static inline unsigned long long atomicMax(unsigned long long * addr, unsigned long long v) {
    unsigned long long old = *addr;
    if (v > old) *addr = v;
    return old;
}
static inline unsigned int __umulhi(unsigned int a, unsigned int b) { return (unsigned int)(((unsigned long long)a * b) >> 32); }
static inline long long __double_as_longlong(double v) {
    long long r;
    std::memcpy(&r, &v, sizeof(r));
    return r;
}
static inline int atomicCAS(int * addr, int compare, int v) {
    int old = *addr;
    if (old == compare) *addr = v;
    return old;
}
static inline int atomicExch(int * addr, int v) {
    int old = *addr;
    *addr = v;
    return old;
}
template <class T>
static inline T __ldg(const T * p) {
    return *p;
}
template <class T>
static inline T __ldcs(const T * p) {
    return *p;
}
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline int __double2int_rd(double a) { return (int)std::floor(a); }
static inline int __double2int_rz(double a) { return (int)a; }

// ---------------------------------------------------------------- runtime API subset
typedef int cudaError_t;
typedef void * cudaStream_t;
typedef void * cudaEvent_t;
typedef void * cudaGraph_t;
typedef void * cudaGraphExec_t;
#define cudaSuccess 0
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1, cudaStreamCaptureModeRelaxed = 2 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };

static inline const char * cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int * d) {
    *d = 0;
    return 0;
}
static inline cudaError_t cudaGetDeviceCount(int * n) {
    *n = 1;
    return 0;
}
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaMalloc(void ** p, size_t n) {
    *p = std::malloc(n ? n : 1);
    return *p ? 0 : 2;
}
template <class T>
static inline cudaError_t cudaMalloc(T ** p, size_t n) {
    return cudaMalloc(reinterpret_cast<void **>(p), n);
}
static inline cudaError_t cudaFree(void * p) {
    std::free(p);
    return 0;
}
static inline cudaError_t cudaMallocHost(void ** p, size_t n) { return cudaMalloc(p, n); }
template <class T>
static inline cudaError_t cudaMallocHost(T ** p, size_t n) {
    return cudaMalloc(reinterpret_cast<void **>(p), n);
}
static inline cudaError_t cudaFreeHost(void * p) { return cudaFree(p); }
static inline cudaError_t cudaMemcpy(void * d, const void * s, size_t n, cudaMemcpyKind) {
    std::memmove(d, s, n);
    return 0;
}
static inline cudaError_t cudaMemcpyAsync(void * d, const void * s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
    std::memmove(d, s, n);
    return 0;
}
static inline cudaError_t cudaMemset(void * d, int v, size_t n) {
    std::memset(d, v, n);
    return 0;
}
static inline cudaError_t cudaMemsetAsync(void * d, int v, size_t n, cudaStream_t = nullptr) {
    std::memset(d, v, n);
    return 0;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t * s, unsigned) {
    *s = nullptr;
    return 0;
}
static inline cudaError_t cudaStreamCreate(cudaStream_t * s) {
    *s = nullptr;
    return 0;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t * e) {
    *e = nullptr;
    return 0;
}
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t * e, unsigned) {
    *e = nullptr;
    return 0;
}
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float * ms, cudaEvent_t, cudaEvent_t) {
    *ms = 0.f;
    return 0;
}
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) {
    return 0;
}

#define QPIC_LAUNCH(kernel, grid, block, smem, stream, ...) \
    qemu::launch((grid), (block), (smem), false, [&]() { kernel(__VA_ARGS__); })
// kernels without barriers / warp collectives may be launched FLAT: threads run as plain calls (fast)
#define QPIC_LAUNCH_FLAT(kernel, grid, block, smem, stream, ...) \
    qemu::launch((grid), (block), (smem), true, [&]() { kernel(__VA_ARGS__); })
#define QPIC_DYN_SMEM(type, name) type * name = reinterpret_cast<type *>(qemu::g_block->dyn_smem)
