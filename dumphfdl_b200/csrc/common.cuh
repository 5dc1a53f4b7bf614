// dumphfdl_b200/csrc/common.cuh -- shared definitions for the sm_100a kernels and their host driver.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <string.h>
#ifdef HFDL_CUSIM
#include "cusim.h"     // tests/cusim: host emulation for logic tests only (never part of the product build)
#define HFDL_LAUNCH(kernel, grid, block, smem, stream, ...) cusim::launch(grid, block, smem, [&] { kernel(__VA_ARGS__); })
#define HFDL_DYN_SMEM(type, name) CUSIM_DYN_SMEM(type, name)
#else
#include <cuda_runtime.h>
#define HFDL_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define HFDL_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw_[]; type *name = reinterpret_cast<type *>(name##_raw_)
#endif

#define HFDL_CHECK(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) { \
	fprintf(stderr, "hfdl_b200: CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(e_), __FILE__, __LINE__, #call); return -1; } } while(0)

typedef float2 cf;

// fast transcendental wrappers: MUFU-based intrinsics on the device, libm under host emulation
#ifdef HFDL_CUSIM
static inline float hfdl_log2_fast(float x) { return log2f(x); }
static inline float hfdl_exp2_fast(float x) { return exp2f(x); }
static inline int hfdl_round_pos(float x) { return (int)floorf(x + 0.5f); }
static inline void hfdl_sincos_fast(float x, float *s, float *c) { sincosf(x, s, c); }
#else
__device__ __forceinline__ float hfdl_log2_fast(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float hfdl_exp2_fast(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ int hfdl_round_pos(float x) { return __float2int_rd(x + 0.5f); }      // == roundf for 0 <= x < 2^22
__device__ __forceinline__ void hfdl_sincos_fast(float x, float *s, float *c) { __sincosf(x, s, c); }
#endif

__host__ __device__ __forceinline__ cf cmul(cf a, cf b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__host__ __device__ __forceinline__ cf cmulc(cf a, cf b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }   // a*conj(b)
__host__ __device__ __forceinline__ cf cadd(cf a, cf b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cf csub(cf a, cf b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cf cscale(cf a, float s) { return make_float2(a.x * s, a.y * s); }

// cp.async (LDGSTS) global->shared prefetch; immediate copies under host emulation
#ifdef HFDL_CUSIM
static inline void hfdl_cp_async8(void *sdst, const void *gsrc) { memcpy(sdst, gsrc, 8); }
static inline void hfdl_cp_async4(void *sdst, const void *gsrc) { memcpy(sdst, gsrc, 4); }
static inline void hfdl_cp_async16(void *sdst, const void *gsrc) { memcpy(sdst, gsrc, 16); }
static inline void hfdl_cp_async_commit() {}
template <int N> static inline void hfdl_cp_async_wait() {}
#else
__device__ __forceinline__ void hfdl_cp_async8(void *sdst, const void *gsrc) {
	unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void hfdl_cp_async4(void *sdst, const void *gsrc) {
	unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void hfdl_cp_async16(void *sdst, const void *gsrc) {
	unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void hfdl_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void hfdl_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

// TMA bulk copies (cp.async.bulk, global -> shared, completion on an mbarrier): one elected lane moves a whole
// contiguous tile with a single instruction; the consumer side waits on the mbarrier's phase.  Under host emulation
// the copy is an immediate memcpy and the barrier a generation counter.
#ifdef HFDL_CUSIM
typedef struct { volatile unsigned long long phase; } hfdl_mbar_t;
static inline void hfdl_mbar_init(hfdl_mbar_t *b, unsigned) { b->phase = 0; }
static inline void hfdl_mbar_expect_tx(hfdl_mbar_t *, unsigned) {}
static inline void hfdl_bulk_g2s(void *sdst, const void *gsrc, unsigned bytes, hfdl_mbar_t *) { memcpy(sdst, gsrc, bytes); }
static inline void hfdl_mbar_arrive_emul(hfdl_mbar_t *b) { __atomic_thread_fence(__ATOMIC_SEQ_CST); b->phase = b->phase + 1; }
static inline bool hfdl_mbar_try_wait(hfdl_mbar_t *b, unsigned parity) { return ((b->phase & 1ull) != (unsigned long long)parity); }
static inline void hfdl_fence_proxy_async() {}
static inline void hfdl_fence_mbar_init() {}
#else
typedef unsigned long long hfdl_mbar_t;
__device__ __forceinline__ void hfdl_mbar_init(hfdl_mbar_t *b, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void hfdl_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void hfdl_mbar_expect_tx(hfdl_mbar_t *b, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hfdl_bulk_g2s(void *sdst, const void *gsrc, unsigned bytes, hfdl_mbar_t *b) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void hfdl_mbar_arrive_emul(hfdl_mbar_t *) {}
__device__ __forceinline__ bool hfdl_mbar_try_wait(hfdl_mbar_t *b, unsigned parity) {
	unsigned ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		: "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(b)), "r"(parity) : "memory");
	return ok != 0;
}
// generic-proxy accesses to shared memory (the consumers' loads) are ordered before later async-proxy writes (the next bulk copy)
__device__ __forceinline__ void hfdl_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

#ifdef HFDL_CUSIM
static inline long long hfdl_clock() { return 0; }
#else
__device__ __forceinline__ long long hfdl_clock() { return clock64(); }
#endif
// polite spinning while one warp waits for another
#ifdef HFDL_CUSIM
static inline void hfdl_cusim_spin(int line) {      // test build only: yield + a watchdog that names a stuck spin loop
	static thread_local unsigned long long n = 0;
	std::this_thread::yield();
	if((++n % 50000000ull) == 0 && getenv("HFDL_CUSIM_WATCHDOG")) fprintf(stderr, "cusim: thread %u block %u still spinning at line %d\n", threadIdx.x, blockIdx.x, line);
}
#define HFDL_SPIN_PAUSE() hfdl_cusim_spin(__LINE__)
#define HFDL_SPIN_PAUSE_LONG() hfdl_cusim_spin(__LINE__)
#else
#define HFDL_SPIN_PAUSE() asm volatile("nanosleep.u32 20;" ::: "memory")
// a waiter that is many symbols ahead of what it waits for (timing warp with a full output ring, loader with chunks in
// flight): it shares a scheduler with another channel's demodulator warp, so it should not poll every few dozen cycles
#ifndef HFDL_SPIN_LONG_NS
#define HFDL_SPIN_LONG_NS 400
#endif
#define HFDL_SPIN_STR2(x) #x
#define HFDL_SPIN_STR(x) HFDL_SPIN_STR2(x)
#define HFDL_SPIN_PAUSE_LONG() asm volatile("nanosleep.u32 " HFDL_SPIN_STR(HFDL_SPIN_LONG_NS) ";" ::: "memory")
#endif

// sample formats (src/input-common.h sample_format)
enum { HFDL_SFMT_CU8 = 1, HFDL_SFMT_CS16 = 2, HFDL_SFMT_CF32 = 3 };

static inline int hfdl_ilog2(int64_t x) { int l = 0; while(((int64_t)1 << l) < x) l++; return l; }
