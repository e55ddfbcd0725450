// simt.h — a minimal SIMT emulator: TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Lets the CUDA sources of lattice_based_tagger_b200/csrc compile with g++ (-DLT_SIMT_EMU) and run on
// the host, so that the kernel LOGIC (enumeration order, tie-breaking, fp64 association, buffer
// overflow paths) can be checked against the oracle in the CPU test suite and debugged without a
// GPU.  The product never loads the library built from this header: `_native.load()` only opens
// `liblt_b200.so` (nvcc, sm_100a) and fails loudly when that is missing.  Nothing here is timed.
//
// Model: a launch runs its blocks one after another; the threads of a block are cooperative
// fibers (hand-written x86-64 context switch).  A fiber runs until it reaches a warp collective
// (`*_sync`, `__syncwarp`) or `__syncthreads`, where it waits for the other lanes named in the
// mask; the last lane to arrive computes every lane's result.  Collectives with different masks
// may be pending in one warp at the same time (divergent code).  Memory is sequentially
// consistent by construction, atomics are plain read-modify-writes.  A deadlock (every live fiber
// waiting) aborts with a message.
// LT_SIMT_MEMCHECK=1: every device buffer and the dynamic shared memory of every launch end at an
// inaccessible page (simt.cpp), so an access past the end of either faults at the access and is
// reported with its block and thread — the pool's GPU boxes refuse compute-sanitizer, this is the
// memory check the kernels get.
#pragma once
#ifndef LT_SIMT_EMU
#error "simt.h is only for -DLT_SIMT_EMU builds (tests)"
#endif

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>
// (every standard header the CUDA sources include is pulled in above, BEFORE the qualifier macros: libstdc++
// spells __attribute__((__noinline__)) itself)

// ---- qualifiers -----------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) alignas(n)
#define __shared__ static              // blocks run one at a time: a function-local static IS block-shared

// ---- vector types ---------------------------------------------------------------------------------
struct uint2 { unsigned int x, y; };
struct alignas(16) uint4 { unsigned int x, y, z, w; };
struct uint3 { unsigned int x, y, z; };
struct dim3 { unsigned int x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a) : x(a) {} };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

namespace simt {

constexpr int kWarp = 32;

enum Op { OP_SYNC, OP_SHFL, OP_SHFL_UP, OP_SHFL_XOR, OP_BALLOT, OP_MATCH, OP_RED_ADD, OP_RED_MAX, OP_RED_OR, OP_BLOCK };

struct Pending {           // one collective in flight inside a warp, keyed by its mask
    uint32_t mask = 0;
    uint32_t arrived = 0;
    uint32_t gen = 0;
    int op = -1;
    const char* file = nullptr;      // call site of the first lane to arrive: every lane must come from the same one
    int line = 0;
    uint64_t val[kWarp];
    int arg[kWarp];
    uint64_t res[kWarp];
};

struct Warp {
    std::vector<Pending> pending;
    uint32_t exited = 0;
};

struct Fiber {
    void* sp = nullptr;
    unsigned char* stack = nullptr;
    bool done = false;
    uint3 tid{0, 0, 0};
    int lane = 0, warp = 0;
};

struct State {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    void* sched_sp = nullptr;
    Fiber* cur = nullptr;
    uint3 bid{0, 0, 0};
    dim3 bdim, gdim;
    unsigned char* smem = nullptr;
    const std::function<void()>* body = nullptr;
    // __syncthreads
    uint32_t block_arrived = 0, block_gen = 0;
    uint64_t progress = 0;              // bumped whenever any fiber makes progress (deadlock detection)
};

State& state();
void yield();
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body);
uint64_t collective(int op, uint32_t mask, uint64_t value, int arg, const char* file, int line);
void block_barrier();
inline unsigned char* dyn_smem() { return state().smem; }
void* dev_alloc(size_t n);
void dev_free(void* p);

struct TidProxy { operator uint3() const { return state().cur->tid; } };

}  // namespace simt

#define threadIdx (simt::state().cur->tid)
#define blockIdx (simt::state().bid)
#define blockDim (simt::state().bdim)
#define gridDim (simt::state().gdim)

// dynamic shared memory: `LT_DYN_SMEM(name);` in the kernels
#define LT_DYN_SMEM(name) unsigned char* name = simt::dyn_smem()
#define LT_LAUNCH(kernel, grid, block, smem, stream, ...) \
    simt::launch(dim3((unsigned)(grid)), dim3((unsigned)(block)), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })

// ---- warp collectives -----------------------------------------------------------------------------
#define SIMT_SITE const char* file_ = __builtin_FILE(), int line_ = __builtin_LINE()
static inline void __syncwarp(unsigned mask = 0xFFFFFFFFu, SIMT_SITE) { simt::collective(simt::OP_SYNC, mask, 0, 0, file_, line_); }
static inline void __syncthreads() { simt::block_barrier(); }
static inline unsigned __ballot_sync(unsigned mask, int pred, SIMT_SITE) { return (unsigned)simt::collective(simt::OP_BALLOT, mask, pred ? 1 : 0, 0, file_, line_); }
static inline int __any_sync(unsigned mask, int pred, SIMT_SITE) { return __ballot_sync(mask, pred, file_, line_) != 0; }
static inline int __all_sync(unsigned mask, int pred, SIMT_SITE) { return __ballot_sync(mask, pred, file_, line_) == mask; }
static inline unsigned __activemask() { return 0xFFFFFFFFu; }

template <typename T>
static inline T simt_shfl(int op, unsigned mask, T v, int arg, const char* file_, int line_) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    raw = simt::collective(op, mask, raw, arg, file_, line_);
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32, SIMT_SITE) { (void)width; return simt_shfl(simt::OP_SHFL, mask, v, src & 31, file_, line_); }
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32, SIMT_SITE) { (void)width; return simt_shfl(simt::OP_SHFL_UP, mask, v, (int)d, file_, line_); }
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int m, int width = 32, SIMT_SITE) { (void)width; return simt_shfl(simt::OP_SHFL_XOR, mask, v, m, file_, line_); }
template <typename T> static inline unsigned __match_any_sync(unsigned mask, T v, SIMT_SITE) {
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    return (unsigned)simt::collective(simt::OP_MATCH, mask, raw, 0, file_, line_);
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v, SIMT_SITE) { return (unsigned)simt::collective(simt::OP_RED_ADD, mask, v, 0, file_, line_); }
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v, SIMT_SITE) { return (unsigned)simt::collective(simt::OP_RED_MAX, mask, v, 0, file_, line_); }
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v, SIMT_SITE) { return (unsigned)simt::collective(simt::OP_RED_OR, mask, v, 0, file_, line_); }

// device-side min / max are global functions in CUDA
template <typename T> static inline T min(T a, T b) { return b < a ? b : a; }
template <typename T> static inline T max(T a, T b) { return a < b ? b : a; }

// ---- scalar intrinsics ------------------------------------------------------------------------------
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
static inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t bits = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double r; memcpy(&r, &bits, 8); return r;
}
static inline int __float2int_rz(float f) { return (int)f; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }

// ---- atomics (fibers are cooperative: plain read-modify-write) ---------------------------------------
template <typename T, typename U> static inline T atomicAdd(T* p, U v) { T old = *p; *p = (T)(old + (T)v); return old; }
template <typename T, typename U> static inline T atomicMin(T* p, U v) { T old = *p; if ((T)v < old) *p = (T)v; return old; }
template <typename T, typename U> static inline T atomicMax(T* p, U v) { T old = *p; if ((T)v > old) *p = (T)v; return old; }
template <typename T, typename U> static inline T atomicOr(T* p, U v) { T old = *p; *p = (T)(old | (T)v); return old; }
template <typename T, typename U, typename V> static inline T atomicCAS(T* p, U cmp, V v) { T old = *p; if (old == (T)cmp) *p = (T)v; return old; }
template <typename T, typename U> static inline T atomicExch(T* p, U v) { T old = *p; *p = (T)v; return old; }

// ---- host runtime stubs ---------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef struct simt_stream* cudaStream_t;
typedef struct simt_event { double t; }* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int multiProcessorCount; char name[64]; };

static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
// (LT_SIMT_SMS: the SM count the host code sizes its grids by — 2 unless a test asks for a wider device)
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { const char* sms_ = getenv("LT_SIMT_SMS"); p->multiProcessorCount = sms_ ? atoi(sms_) : 2; strcpy(p->name, "simt-emu"); return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "simt-emu error"; }
// (simt::dev_alloc: plain aligned memory filled with 0xCD, or — LT_SIMT_MEMCHECK=1 — a mapping that ends at an
// inaccessible guard page, so that a kernel reading or writing past a device buffer faults at the access)
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = simt::dev_alloc(n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { simt::dev_free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline double simt_now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new simt_event{0.0}; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = simt_now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 1; return cudaSuccess; }
