/* TEST INFRASTRUCTURE ONLY.
 *
 * cusim.h -- a small SIMT emulator so that the product's .cu kernels can be compiled with g++ and run
 * on a CPU-only box, lane for lane: every CUDA thread of a block is a fiber with its own stack;
 * warp collectives (__shfl*_sync, __ballot_sync, __reduce_*_sync, __syncwarp) and __syncthreads() are
 * rendezvous points between fibers.  Blocks of a launch run sequentially (optionally on several OS
 * threads, one block at a time each).  The emulator checks LOGIC (indexing, collectives, control
 * flow, tie-breaks); it knows nothing about memory-model races or performance.
 *
 * Used only by tests/ (make cusim).  The product library is compiled by nvcc for sm_100a and never
 * sees this file.
 */
#ifndef CUSIM_H
#define CUSIM_H
#ifndef BWAG_CUSIM
#define BWAG_CUSIM 1
#endif

#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <functional>
#include <algorithm>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; } __attribute__((aligned(16)));
struct int2 { int x, y; };
struct int4 { int x, y, z, w; } __attribute__((aligned(16)));
struct ulonglong2 { unsigned long long x, y; } __attribute__((aligned(16)));
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { uint4 r = {a, b, c, d}; return r; }
static inline int4 make_int4(int a, int b, int c, int d) { int4 r = {a, b, c, d}; return r; }
static inline uint2 make_uint2(unsigned a, unsigned b) { uint2 r = {a, b}; return r; }
static inline int2 make_int2(int a, int b) { int2 r = {a, b}; return r; }

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
extern thread_local unsigned char *cusim_dyn_smem;
static const int warpSize = 32;

/* ---- fiber runtime (cusim.cpp) ---- */
void cusim_yield();
uint64_t cusim_warp_gather(unsigned mask, uint64_t v, uint64_t out[32]); /* all participating lanes deposit v; everyone gets the 32 slots */
void cusim_block_barrier();
void cusim_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
int cusim_lane();

/* ---- warp collectives ---- */
template <typename T> static inline uint64_t cusim_bits(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, "shuffle payload too large"); memcpy(&u, &v, sizeof(T)); return u; }
template <typename T> static inline T cusim_unbits(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
	uint64_t all[32];
	int lane = cusim_lane(), base = lane & ~(width - 1);
	cusim_warp_gather(mask, cusim_bits(v), all);
	return cusim_unbits<T>(all[base + (src & (width - 1))]);
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32)
{
	uint64_t all[32];
	int lane = cusim_lane(), base = lane & ~(width - 1), s = lane - (int)d;
	cusim_warp_gather(mask, cusim_bits(v), all);
	return s < base ? v : cusim_unbits<T>(all[s]);
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32)
{
	uint64_t all[32];
	int lane = cusim_lane(), base = lane & ~(width - 1), s = lane + (int)d;
	cusim_warp_gather(mask, cusim_bits(v), all);
	return s >= base + width ? v : cusim_unbits<T>(all[s]);
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32)
{
	uint64_t all[32];
	int lane = cusim_lane();
	(void)width;
	cusim_warp_gather(mask, cusim_bits(v), all);
	return cusim_unbits<T>(all[lane ^ x]);
}
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
	uint64_t all[32];
	unsigned r = 0;
	cusim_warp_gather(mask, pred ? 1 : 0, all);
	for (int i = 0; i < 32; ++i) if ((mask >> i & 1) && all[i]) r |= 1u << i;
	return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { uint64_t all[32]; cusim_warp_gather(mask, 0, all); }
static inline unsigned __activemask() { return 0xffffffffu; }
static inline int __reduce_max_sync(unsigned mask, int v) { uint64_t all[32]; cusim_warp_gather(mask, cusim_bits(v), all); int r = v; for (int i = 0; i < 32; ++i) if (mask >> i & 1) r = std::max(r, cusim_unbits<int>(all[i])); return r; }
static inline int __reduce_min_sync(unsigned mask, int v) { uint64_t all[32]; cusim_warp_gather(mask, cusim_bits(v), all); int r = v; for (int i = 0; i < 32; ++i) if (mask >> i & 1) r = std::min(r, cusim_unbits<int>(all[i])); return r; }
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) { uint64_t all[32]; cusim_warp_gather(mask, v, all); unsigned r = v; for (int i = 0; i < 32; ++i) if (mask >> i & 1) r = std::max(r, (unsigned)all[i]); return r; }
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) { uint64_t all[32]; cusim_warp_gather(mask, v, all); unsigned r = v; for (int i = 0; i < 32; ++i) if (mask >> i & 1) r = std::min(r, (unsigned)all[i]); return r; }
static inline int __reduce_add_sync(unsigned mask, int v) { uint64_t all[32]; cusim_warp_gather(mask, cusim_bits(v), all); int r = 0; for (int i = 0; i < 32; ++i) if (mask >> i & 1) r += cusim_unbits<int>(all[i]); return r; }
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { uint64_t all[32]; cusim_warp_gather(mask, v, all); unsigned r = 0; for (int i = 0; i < 32; ++i) if (mask >> i & 1) r += (unsigned)all[i]; return r; }
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) { uint64_t all[32]; cusim_warp_gather(mask, v, all); unsigned r = 0; for (int i = 0; i < 32; ++i) if (mask >> i & 1) r |= (unsigned)all[i]; return r; }
static inline void __syncthreads() { cusim_block_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

/* ---- integer intrinsics ---- */
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= (x >> i & 1) << (31 - i); return r; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { uint64_t v = (uint64_t)hi << 32 | lo; return (unsigned)(v >> (s & 31)); }
static inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned s) { uint64_t v = (uint64_t)hi << 32 | lo; if (s > 32) s = 32; return (unsigned)((v << s) >> 32); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { uint64_t v = (uint64_t)hi << 32 | lo; return (unsigned)((v << (s & 31)) >> 32); }
static inline int __vimax3_s32(int a, int b, int c) { return std::max(a, std::max(b, c)); }
static inline int __vimin3_s32(int a, int b, int c) { return std::min(a, std::min(b, c)); }
static inline int __viaddmax_s32(int a, int b, int c) { return std::max(a + b, c); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
template <typename T> static inline void __stcs(T *p, T v) { *p = v; }
template <typename T> static inline void __stcg(T *p, T v) { *p = v; }
using std::max;
using std::min;
static inline long long max(long long a, int b) { return a > b ? a : b; }
static inline long long min(long long a, int b) { return a < b ? a : b; }

/* ---- atomics: fibers never pre-empt each other, OS threads run different blocks -> use real atomics ---- */
static inline int atomicAdd(int *p, int v) { return __sync_fetch_and_add(p, v); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __sync_fetch_and_add(p, v); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __sync_fetch_and_add(p, v); }
static inline int atomicMax(int *p, int v) { int o = *p; while (o < v && !__sync_bool_compare_and_swap(p, o, v)) o = *p; return o; }
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; while (o < v && !__sync_bool_compare_and_swap(p, o, v)) o = *p; return o; }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __sync_fetch_and_or(p, v); }
static inline int atomicExch(int *p, int v) { return __sync_lock_test_and_set(p, v); }
static inline int atomicCAS(int *p, int c, int v) { return __sync_val_compare_and_swap(p, c, v); }

/* ---- the slice of the CUDA runtime the product's host code uses ---- */
typedef int cudaError_t;
typedef struct cusim_stream *cudaStream_t;
typedef struct cusim_event *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNoDevice = 100, cudaErrorNotReady = 600 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaHostAllocDefault = 0, cudaEventDefault = 0, cudaStreamNonBlocking = 1 };
struct cudaDeviceProp { int multiProcessorCount; size_t totalGlobalMem; char name[256]; int major, minor; size_t sharedMemPerBlockOptin; };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned f) { (void)f; return cudaMallocHost(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind k) { (void)k; memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t st = 0) { (void)k; (void)st; memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st = 0) { (void)st; memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned f) { (void)f; *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { (void)s; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t s) { (void)s; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = 0; return cudaSuccess; }
enum { cudaEventBlockingSync = 1, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned f) { (void)f; *e = 0; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { (void)e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = 0) { (void)e; (void)s; return cudaSuccess; }
static inline cudaError_t cudaStreamQuery(cudaStream_t s) { (void)s; return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t e) { (void)e; return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { (void)a; (void)b; *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { (void)d; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t e) { (void)e; return "cusim"; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d) { (void)d; memset(p, 0, sizeof(*p)); p->multiProcessorCount = 2; p->totalGlobalMem = (size_t)8 << 30; strcpy(p->name, "cusim"); p->major = 10; p->sharedMemPerBlockOptin = 227 << 10; return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F f, cudaFuncAttribute a, int v) { (void)f; (void)a; (void)v; return cudaSuccess; }
enum cudaLimit { cudaLimitMaxL2FetchGranularity = 5 };
static inline cudaError_t cudaDeviceSetLimit(cudaLimit l, size_t v) { (void)l; (void)v; return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t *fr, size_t *tot) { *fr = (size_t)32 << 30; *tot = (size_t)32 << 30; return cudaSuccess; }

#endif
