// CPU stand-in for <cuda_runtime.h>, for the TEST SUITE ONLY (tools/emu/README.md).
//
// It lets the translation units of icqsol_b200/csrc be compiled by g++ (after
// tools/emu/translate.py has rewritten the <<<...>>> launches) and run on the host, so that
// the kernels' indexing, partition and reduction logic can be checked where there is no
// GPU.  One CUDA thread = one cooperative fiber; the blocks of a launch run one after the
// other; "device memory" is host memory; streams are synchronous.  It says nothing about
// races, memory ordering, performance -- or about PTX: code behind `#ifdef ICQB_EMU` replaces
// the inline-PTX helpers.  The product never loads this build: it has a different file name
// and only tests/test_emulated_kernels.py asks for it.
#pragma once

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <functional>

#ifndef ICQB_EMU
#error "tools/emu/include is for the emulation build only (-DICQB_EMU)"
#endif

// ---- qualifiers --------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
// one copy per OS thread: the multi-rank tests run one rank per thread
#define __shared__ static thread_local
#define __constant__ const
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types / launch geometry ------------------------------------------------------
struct alignas(16) double2 {
    double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct alignas(16) float4 {
    float x, y, z, w;
};

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3 {
    unsigned x, y, z;
};

namespace emu {
struct ThreadCtx {
    uint3 tid, bid;
    dim3 bdim, gdim;
};
ThreadCtx& cur();                 // the running fiber's indices
void yield();                     // let the other threads of the block run
void syncthreads();
void syncwarp();
void* dyn_smem();                 // dynamic shared memory of the running block
uint64_t* warp_slots();           // 32 exchange slots of the running warp
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
}  // namespace emu

#define threadIdx (emu::cur().tid)
#define blockIdx (emu::cur().bid)
#define blockDim (emu::cur().bdim)
#define gridDim (emu::cur().gdim)
#define EMU_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kernel(__VA_ARGS__); })

// ---- device intrinsics -------------------------------------------------------------------
static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::syncwarp(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}

template <class T>
static inline T __ldg(const T* p) {
    return *p;
}

template <class T>
static inline T __ldcs(const T* p) {
    return *p;
}

template <class T>
static inline T emu_shfl(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    uint64_t* slot = emu::warp_slots();
    const int lane = (int)((emu::cur().tid.x + emu::cur().bdim.x * (emu::cur().tid.y +
                            emu::cur().bdim.y * emu::cur().tid.z)) & 31);
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    slot[lane] = bits;
    emu::syncwarp();
    const uint64_t got = slot[src & 31];
    emu::syncwarp();
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int off) {
    const int lane = (int)((emu::cur().tid.x + emu::cur().bdim.x * (emu::cur().tid.y +
                            emu::cur().bdim.y * emu::cur().tid.z)) & 31);
    return emu_shfl(v, lane ^ off);
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src) {
    return emu_shfl(v, src);
}

// warp vote: every lane of the (full) warp deposits its predicate, all read all 32
static inline int __all_sync(unsigned, int pred) {
    uint64_t* slot = emu::warp_slots();
    const int lane = (int)((emu::cur().tid.x + emu::cur().bdim.x * (emu::cur().tid.y +
                            emu::cur().bdim.y * emu::cur().tid.z)) & 31);
    slot[lane] = pred ? 1u : 0u;
    emu::syncwarp();
    int all = 1;
    for (int l = 0; l < 32; ++l) all &= (int)slot[l];
    emu::syncwarp();
    return all;
}

static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    const unsigned long long old = *p;
    *p = old + v;
    return old;
}
static inline unsigned int atomicAdd(unsigned int* p, unsigned int v) {
    const unsigned int old = *p;
    *p = old + v;
    return old;
}
static inline int atomicAdd(int* p, int v) {
    const int old = *p;
    *p = old + v;
    return old;
}
static inline double atomicAdd(double* p, double v) {
    const double old = *p;
    *p = old + v;
    return old;
}

// explicitly rounded arithmetic (the build uses -ffp-contract=off)
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }

// CUDA's overloaded min/max in device code
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline long long min(long long a, int b) { return a < b ? a : b; }
static inline long long max(long long a, int b) { return a > b ? a : b; }
static inline long long min(int a, long long b) { return a < b ? a : b; }
static inline long long max(int a, long long b) { return a > b ? a : b; }
static inline long min(long a, long b) { return a < b ? a : b; }
static inline long max(long a, long b) { return a > b ? a : b; }
static inline long long min(long long a, long b) { return a < b ? a : (long long)b; }
static inline long long max(long long a, long b) { return a > b ? a : (long long)b; }
static inline long long min(long a, long long b) { return a < b ? (long long)a : b; }
static inline long long max(long a, long long b) { return a > b ? (long long)a : b; }
static inline double min(double a, double b) { return a < b ? a : b; }
static inline double max(double a, double b) { return a > b ? a : b; }

// ---- runtime API (synchronous, one "device") -------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
typedef struct emuStream* cudaStream_t;
typedef struct emuEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice,
                      cudaMemcpyHostToHost, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int* d);
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int dev);
cudaError_t cudaGetLastError();
const char* cudaGetErrorString(cudaError_t e);
cudaError_t emuMalloc(void** p, size_t bytes);
template <class T>
static inline cudaError_t cudaMalloc(T** p, size_t bytes) {
    return emuMalloc(reinterpret_cast<void**>(p), bytes);
}
template <class T>
static inline cudaError_t cudaMallocHost(T** p, size_t bytes) {
    return emuMalloc(reinterpret_cast<void**>(p), bytes);
}
cudaError_t cudaFree(void* p);
// stream-ordered allocation (ICQB_POOL_ALLOC): the same allocator, synchronously
typedef struct emuMemPool* cudaMemPool_t;
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4 };
static inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* pool, int) {
    *pool = nullptr;
    return cudaSuccess;
}
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) {
    return cudaSuccess;
}
template <class T>
static inline cudaError_t cudaMallocAsync(T** p, size_t bytes, cudaStream_t) {
    return emuMalloc(reinterpret_cast<void**>(p), bytes);
}
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { return cudaFree(p); }
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t = nullptr);
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width,
                              size_t height, cudaMemcpyKind, cudaStream_t = nullptr);
cudaError_t cudaMemsetAsync(void* p, int value, size_t bytes, cudaStream_t = nullptr);
cudaError_t cudaStreamCreate(cudaStream_t* s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
// "IPC": the ranks of the emulation are threads of one process, a handle is the pointer itself
struct cudaIpcMemHandle_t {
    char reserved[64];
};
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
    memset(h, 0, sizeof(*h));
    memcpy(h, &p, sizeof(p));
    return cudaSuccess;
}
static inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
    memcpy(p, &h, sizeof(*p));
    return cudaSuccess;
}
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
cudaError_t cudaMemset(void* p, int value, size_t bytes);
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) {
    return cudaSuccess;
}
