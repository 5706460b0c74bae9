// CTA emulator (TEST INFRASTRUCTURE ONLY -- never compiled into libaiolos_b200.so).
//
// The dev container has no GPU, so kernel numerics cannot be debugged there with CUDA.  This header lets g++ compile
// the very same kernel sources (aiolos_b200/csrc/*.cuh) and the host engine (engine.cu) into a separate test library,
// libaiolos_emu.so, in which one thread block is run as a set of cooperative fibers (ucontext) on one OS thread:
//   __syncthreads()            -> fiber yields until every live fiber of the block has arrived
//   __shfl_*_sync / *_or       -> exchange through a slot array bracketed by warp-/block-level rendezvous
//   __shared__ / dynamic smem  -> static storage / a per-launch buffer poisoned with NaNs (catches uninitialised reads)
//   cudaMalloc/cudaMemcpy/...  -> malloc/memcpy; streams and events are dummies
// Blocks run one after the other, so the emulator is deterministic.  It exists so that `pytest -m "not gpu"` can compare
// the kernels' logic with the oracle on the CPU; the product library is built by nvcc only and has no CPU path.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>

#define AIO_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __managed__
#define __grid_constant__
#define __launch_bounds__(...)
#define __shared__ static
#define __restrict__

struct alignas(16) double2 { double x, y; };
inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint3_emu { unsigned x, y, z; };

namespace emu {
extern uint3_emu tIdx, bIdx, bDim, gDim;
extern void* dyn_smem;
extern size_t dyn_smem_bytes;           // size of the current launch's dynamic shared memory (bounds checks)
void sync_block();                       // __syncthreads
void sync_warp();                        // rendezvous of the live fibers of the calling warp
extern double xd[1024];                  // exchange slots
extern int xi[1024];
void run_grid(void (*thunk)(void*), void* args, dim3 grid, dim3 block, size_t smem);
template <typename... A> struct Pack;
template <typename P> struct Call { void (*k)(P); P p; };
template <typename P> void thunk1(void* v) { auto* c = (Call<P>*)v; c->k(c->p); }
template <typename P> void launch(void (*k)(P), dim3 grid, dim3 block, size_t smem, const P& p) {
    Call<P> c{k, p};
    run_grid(&thunk1<P>, &c, grid, block, smem);
}
}  // namespace emu

#define threadIdx emu::tIdx
#define blockIdx emu::bIdx
#define blockDim emu::bDim
#define gridDim emu::gDim

inline void __syncthreads() { emu::sync_block(); }
inline int __syncthreads_or(int v) {
    emu::xi[threadIdx.x] = v; emu::sync_block();
    int r = 0; for (unsigned k = 0; k < blockDim.x; k++) r |= emu::xi[k];
    emu::sync_block(); return r;
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu::sync_warp(); }
template <typename T> inline T __ldg(const T* p) { return *p; }
inline double __shfl_xor_sync(unsigned, double v, int o) {
    const unsigned t = threadIdx.x; emu::xd[t] = v; emu::sync_warp();
    const unsigned src = (t & ~31u) | ((t ^ (unsigned)o) & 31u);
    double r = (src < blockDim.x) ? emu::xd[src] : v; emu::sync_warp(); return r;
}
inline double __shfl_up_sync(unsigned, double v, int o) {
    const unsigned t = threadIdx.x; emu::xd[t] = v; emu::sync_warp();
    double r = ((int)(t & 31u) - o >= 0) ? emu::xd[t - o] : v; emu::sync_warp(); return r;
}
inline double __shfl_down_sync(unsigned, double v, int o) {
    const unsigned t = threadIdx.x; emu::xd[t] = v; emu::sync_warp();
    double r = ((t & 31u) + o < 32u && t + o < blockDim.x) ? emu::xd[t + o] : v; emu::sync_warp(); return r;
}
inline double __shfl_sync(unsigned, double v, int lane) {
    const unsigned t = threadIdx.x; emu::xd[t] = v; emu::sync_warp();
    const unsigned src = (t & ~31u) | ((unsigned)lane & 31u);
    double r = (src < blockDim.x) ? emu::xd[src] : v; emu::sync_warp(); return r;
}
inline int __shfl_xor_sync(unsigned, int v, int o) {
    const unsigned t = threadIdx.x; emu::xi[t] = v; emu::sync_warp();
    const unsigned src = (t & ~31u) | ((t ^ (unsigned)o) & 31u);
    int r = (src < blockDim.x) ? emu::xi[src] : v; emu::sync_warp(); return r;
}
inline int __shfl_sync(unsigned, int v, int lane) {
    const unsigned t = threadIdx.x; emu::xi[t] = v; emu::sync_warp();
    const unsigned src = (t & ~31u) | ((unsigned)lane & 31u);
    int r = (src < blockDim.x) ? emu::xi[src] : v; emu::sync_warp(); return r;
}
inline unsigned __ballot_sync(unsigned, int pred) {
    const unsigned t = threadIdx.x; emu::xi[t] = pred ? 1 : 0; emu::sync_warp();
    unsigned m = 0; const unsigned base = t & ~31u;
    for (unsigned k = 0; k < 32 && base + k < blockDim.x; k++) if (emu::xi[base + k]) m |= (1u << k);
    emu::sync_warp(); return m;
}
inline int __ffs(unsigned v) { return v ? __builtin_ffs((int)v) : 0; }
inline double __drcp_rn(double x) { return 1.0 / x; }
inline double __dsqrt_rn(double x) { return std::sqrt(x); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
using std::isnan;
using std::min;
using std::max;

// ---- host API subset used by engine.cu ------------------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct { double t; }* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97, cudaDevAttrMultiProcessorCount = 16,
       cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
double emu_now_ms();
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)malloc(sizeof(double)); return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = emu_now_ms(); return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return 0; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
template <typename T> inline cudaError_t cudaMallocHost(T** p, size_t n) { *p = (T*)malloc(n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; r++) memcpy((char*)d + r * dp, (const char*)s + r * sp, w);
    return 0;
}
inline cudaError_t cudaDeviceGetAttribute(int* v, int attr, int) {
    *v = (attr == cudaDevAttrMultiProcessorCount) ? 148 : 232448; return 0;
}
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
