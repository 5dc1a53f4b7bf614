/*
 * tests/cusim/cusim.h -- HOST EMULATION OF THE CUDA SUBSET USED BY dumphfdl_b200/csrc (TEST INFRASTRUCTURE).
 *
 * There is no GPU in the development container, so the kernels' *logic* (indexing, shared-memory
 * choreography, state machines) is exercised on the CPU by compiling the very same .cu sources
 * as C++ with -DHFDL_CUSIM: every CUDA thread of a block becomes a host thread, __syncthreads()
 * is a real barrier, warp shuffles go through a per-warp exchange buffer, and the handful of
 * CUDA runtime calls the host code makes map onto malloc/memcpy.  The result is
 * tests/cusim/libhfdl_cusim.so, loaded ONLY by the `-m "not gpu"` logic tests.
 *
 * It is NOT a CPU fallback: libhfdl_b200.so (the product) is built by nvcc for sm_100a, contains no
 * host implementation of any kernel, and refuses to run without a CUDA device.
 */
#pragma once
#ifndef HFDL_CUSIM
#error "cusim.h is only for the -DHFDL_CUSIM test build"
#endif
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <thread>
#include <vector>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <atomic>

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint3_ { unsigned x, y, z; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct short2 { short x, y; };
struct uchar2 { unsigned char x, y; };
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
static inline int2 make_int2(int a, int b) { int2 r; r.x = a; r.y = b; return r; }
static inline uint2 make_uint2(unsigned a, unsigned b) { uint2 r; r.x = a; r.y = b; return r; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __constant__ static

#ifndef CUSIM_YIELDS
#define CUSIM_YIELDS 200
#endif
namespace cusim {
class Barrier {
	std::atomic<unsigned> count{0}, gen{0}; unsigned n;
public:
	explicit Barrier(unsigned n_) : n(n_) {}
	void wait() {
		unsigned g = gen.load(std::memory_order_acquire);
		if(count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
			count.store(0, std::memory_order_relaxed);
			gen.store(g + 1, std::memory_order_release);
			gen.notify_all();
		} else {
			unsigned spins = 0;
			while(gen.load(std::memory_order_acquire) == g) {
				if(++spins <= 64) continue;
				if(spins <= 64 + CUSIM_YIELDS) std::this_thread::yield();
				else gen.wait(g, std::memory_order_acquire);            // a long wait (idle lanes, whole-CTA barriers): sleep on the futex
			}
		}
	}
};
struct WarpCtx { Barrier bar; uint64_t xch[2][32] = {}; explicit WarpCtx(unsigned n) : bar(n) {} };      // two exchange buffers, used alternately
struct BlockCtx { Barrier *bar; std::vector<WarpCtx *> warps; unsigned char *dyn_smem; };
extern thread_local uint3_ t_threadIdx, t_blockIdx;
extern thread_local dim3 t_blockDim, t_gridDim;
extern thread_local BlockCtx *t_ctx;
extern thread_local WarpCtx *t_warp;
extern thread_local unsigned t_lane, t_xpar;

template <typename F> void launch(dim3 grid, dim3 block, size_t smem, F body) {
	unsigned nth = block.x * block.y * block.z;
	Barrier bar(nth);
	BlockCtx ctx;
	ctx.bar = &bar;
	unsigned nwarps = (nth + 31) / 32;
	for(unsigned w = 0; w < nwarps; w++) {
		unsigned members = (w == nwarps - 1) ? nth - 32 * w : 32;
		ctx.warps.push_back(new WarpCtx(members));
	}
	ctx.dyn_smem = (unsigned char *)aligned_alloc(128, ((smem + 127) / 128 + 1) * 128);
	auto worker = [&](unsigned tid) {
		t_ctx = &ctx; t_warp = ctx.warps[tid / 32]; t_lane = tid % 32; t_xpar = 0;
		t_blockDim = block; t_gridDim = grid;
		t_threadIdx.x = tid % block.x; t_threadIdx.y = (tid / block.x) % block.y; t_threadIdx.z = tid / (block.x * block.y);
		for(unsigned bz = 0; bz < grid.z; bz++)
			for(unsigned by = 0; by < grid.y; by++)
				for(unsigned bx = 0; bx < grid.x; bx++) {
					t_blockIdx.x = bx; t_blockIdx.y = by; t_blockIdx.z = bz;
					body();
					bar.wait();
				}
	};
	if(nth == 1) { worker(0); }
	else {
		std::vector<std::thread> th;
		for(unsigned t = 0; t < nth; t++) th.emplace_back(worker, t);
		for(auto &t : th) t.join();
	}
	for(auto w : ctx.warps) delete w;
	free(ctx.dyn_smem);
}
// One barrier per exchange: every lane writes buffer p, all meet, every lane reads buffer p; the next exchange uses buffer
// 1 - p, and buffer p is only written again two exchanges later -- after a barrier every lane reaches only once it has
// finished reading here.  (All lanes of a warp take part in every exchange, so their parities stay in step.)
template <typename T> static inline T shfl_any(T v, int src) {
	static_assert(sizeof(T) <= 8, "shfl size");
	uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
	const unsigned p = t_xpar; t_xpar ^= 1u;
	t_warp->xch[p][t_lane] = raw;
	t_warp->bar.wait();
	uint64_t got = t_warp->xch[p][src & 31];
	T r; memcpy(&r, &got, sizeof(T)); return r;
}
}  // namespace cusim

#define threadIdx (cusim::t_threadIdx)
#define blockIdx (cusim::t_blockIdx)
#define blockDim (cusim::t_blockDim)
#define gridDim (cusim::t_gridDim)
#define __shared__ static
static inline void __syncthreads() { cusim::t_ctx->bar->wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cusim::t_warp->bar.wait(); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int w = 32) { int base = (int)cusim::t_lane & ~(w - 1); return cusim::shfl_any(v, base + (src & (w - 1))); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return cusim::shfl_any(v, (int)cusim::t_lane ^ m); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d, int w = 32) { int l = (int)cusim::t_lane, s = l + d; return cusim::shfl_any(v, (s & ~(w - 1)) != (l & ~(w - 1)) ? l : s); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d, int = 32) { int s = (int)cusim::t_lane - d; return cusim::shfl_any(v, s < 0 ? (int)cusim::t_lane : s); }
static inline unsigned __ballot_sync(unsigned, int pred) {
	const unsigned p = cusim::t_xpar; cusim::t_xpar ^= 1u;
	cusim::t_warp->xch[p][cusim::t_lane] = pred ? 1u : 0u;
	cusim::t_warp->bar.wait();
	unsigned r = 0;
	for(int i = 0; i < 32; i++) r |= (unsigned)(cusim::t_warp->xch[p][i] & 1u) << i;
	return r;
}
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for(int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i); return r; }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline void sincospif(float x, float *s, float *c) { double a = M_PI * (double)x; *s = (float)sin(a); *c = (float)cos(a); }
static inline void sincospi(double x, double *s, double *c) { double a = M_PI * x; *s = sin(a); *c = cos(a); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline float4 make_float4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
#define CUSIM_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(cusim::t_ctx->dyn_smem)

/* ---------------- minimal CUDA runtime stand-in for the host code ---------------- */
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef struct cusim_event { double t; } *cudaEvent_t;
enum { cudaSuccess = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaEventDefault = 0, cudaEventDisableTiming = 2 };
static inline const char *cudaGetErrorString(cudaError_t) { return "cusim"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = aligned_alloc(256, ((n + 255) / 256 + 1) * 256); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, int) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, int, cudaStream_t) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDeviceCount(int *n) { const char *e = getenv("HFDL_CUSIM_DEVICES"); *n = e ? atoi(e) : 1; return 0; }      // several emulated "devices" share the host memory
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new cusim_event{0}; return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
template <typename T> static inline cudaError_t cudaFuncSetAttribute(T, int, int) { return 0; }
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; size_t totalGlobalMem; };
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { memset(p, 0, sizeof(*p)); strcpy(p->name, "cusim"); p->major = 10; p->multiProcessorCount = 148; return 0; }
enum { cudaErrorNotReady = 600 };
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return 0; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return 0; }
static inline cudaError_t cudaMemcpyPeerAsync(void *d, int, const void *s, int, size_t n, cudaStream_t) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return 0; }
static inline cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 1; return 0; }
