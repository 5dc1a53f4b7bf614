// dumphfdl_b200/csrc/loop_kernel.cuh -- K8b-K11: the feedback part of the per-channel HFDL demodulator
// (hfdl.c:707-891): symbol-timing recursion, Costas loop, T/2 LMS equaliser, M-PSK slicer, sampler and framer.
//
// Three specialised warps per channel that talk through shared-memory rings; two or four channels per CTA (see "CTA
// layouts" below).  Every feedback loop of the reference is a strictly sequential float recurrence, so the design goal
// is the SHORTEST DEPENDENT CHAIN and the FEWEST INSTRUCTIONS on each sequential warp; everything that is not on a chain
// is moved to another warp or to other lanes:
//   warp 2 "loader"  streams the precomputed filter-bank rows (bank_kernel) and AGC levels of the channel from HBM
//                    into a 128-sample shared-memory ring (one TMA bulk copy per 32-sample chunk, completion on mbarriers, 3 chunks in flight).
//   warp 1 "timing"  symsync_crcf_step: iterates per OUTPUT (not per input sample): arm lookup in the ring, timing
//                    error detector + loop filter on every second output, tau/arm update, skip to the input sample
//                    of the next output.  Publishes {symbol, AGC level, tag} entries into a 64-entry output ring.
//                    It runs ahead of the demodulator; symsync_crcf_reset requests (framer reset, Costas blow-up,
//                    13-frame timeout: hfdl.c:711-715,746-752,968-991) roll it back to the reset point.
//   warp 0 "demod"   consumes the entries in order: Costas rotation, equaliser, slicer, sampler, framer.  The 15
//                    equaliser taps live one per lane (both half-warps hold a copy).  Between framer events it runs
//                    demod_run<MODE, ARITY>: straight-line code per symbol in which only the two newest taps follow
//                    the rotation; the partial sum over the older taps is reduced with shuffles OFF the phase ->
//                    decision -> phase chain (frozen-weight runs: one symbol ahead; LMS runs: pipelined across the
//                    symbol boundary).
// Ring entries are 16 bytes of data plus a separate validity tag (reset generation + lap) that is written after the
// data and read before it, with the data load address-dependent on the tag: see "output ring" below.
#pragma once
#include "common.cuh"

#define HFDL_LK_WARPS 3                  // warps per channel: demodulator, timing, loader
#define HFDL_LK_NCH_MAX 4                // most channels a CTA serves (layouts below)
#define HFDL_LK_CH_THREADS (32 * HFDL_LK_WARPS)
// CTA layouts (template parameters of loop_kernel; a warp sits on scheduler warp % 4):
//   <4, false>  four channels, warps [d0 t0 l0 d1 t1 l1 ...]: every scheduler serves one demodulator, one timing and one
//               loader warp of different channels.  Quarter of the SMs; the demodulator warp shares its scheduler and
//               shared-memory pipeline with a timing warp (1.41 ms per cfg-3 batch alone on the GPU).
//   <2, true>   two channels, role-major with two idle warps: [d0 d1 t0 t1 - - l0 l1] -- each demodulator warp has a
//               scheduler of its own (1.14 ms alone, like one channel per CTA: 1.15 ms).  Half of the SMs; used with the
//               padded shared-memory request when the other stages have SMs to spare (multi-GPU).
//   <2, false>  two channels, warps [d0 t0 l0 d1 t1 l1] (host emulation: fewer host threads per CTA).
__host__ __device__ constexpr int lk_threads(int nch, bool role_major) { return role_major ? 256 : HFDL_LK_CH_THREADS * nch; }
#define HFDL_LK_RING 64                  // output entries the timing warp may run ahead of the demodulator warp
#ifndef HFDL_LK_BR
#define HFDL_LK_BR 128                   // bank ring, input samples (power of two, >= 4 loader chunks)
#endif
#define HFDL_LK_CH 32                    // input samples per loader chunk
#define HFDL_LK_INFLIGHT 3               // loader chunks in flight
#define HFDL_LK_SMEM_CH (HFDL_LK_BR * 32 * 8)             // dynamic shared memory per channel: its bank ring
// Channel packing: the sequential warps of a channel are latency-bound and leave their SM's issue slots almost empty, but
// they are slowed by ~40 % when the schedulers they sit on also serve the FFT / filter-bank CTAs of the other pipeline
// stages (loop_kernel alone: 1.25 ms per cfg-3 batch; beside the other stages, one channel per SM: 1.75 ms).  Four
// channels per CTA put each kind of warp on its own scheduler (warp w: channel w / 3, role w % 3 -> scheduler w % 4), use
// a quarter of the SMs, and fill those SMs' register file so that no FFT CTA can be co-resident there.
#define HFDL_LK_MAXN (1 << 20)           // input samples per launch (tag layout)

struct LoopArgs {
	const cf *bank; long long bank_stride;
	const cf *mfo; long long mfo_stride;
	const float *lvl; long long lvl_stride;
	int n_samples, n_channels;
	DemodState *state;
	const DemodTables *tab;
	cf *datasym; int nslots;   // [C][nslots][HFDL_DATA_SYMS_MAX]
	FrameRec *frames; int *nframes; int max_frames;
	int cap_channel; cf *cap_eq; int *cap_cnt; int cap_max;      // f_eq_out checkpoint of one channel
	long long *dbg_cycles;     // diagnostics only: [C][8] = timing-warp total / blocked, demod-warp total / waiting, polls (ring full / loader), outputs
	int debug_mode;
};

__device__ __forceinline__ void bits_push(unsigned *b, unsigned bit) {
	b[3] = ((b[3] << 1) | (b[2] >> 31)) & 0x7FFFFFFFu;
	b[2] = (b[2] << 1) | (b[1] >> 31);
	b[1] = (b[1] << 1) | (b[0] >> 31);
	b[0] = (b[0] << 1) | (bit & 1u);
}
__device__ __forceinline__ int bits_corr(const unsigned *a, const unsigned *b) {   // equal positions of 127
	return 127 - (__popc(a[0] ^ b[0]) + __popc(a[1] ^ b[1]) + __popc(a[2] ^ b[2]) + __popc((a[3] ^ b[3]) & 0x7FFFFFFFu));
}

__device__ __forceinline__ void ss_reset(DemodState &S) {     // symsync_crcf_reset: mf window, timing state, loop filter
	S.ss_since_reset = 0;                                     // the mf-arm window is cleared, the dmf one is not
	S.ss_rate = 1.5f; S.ss_del = 1.5f;                        // k / k_out = 3/2
	S.ss_b = 0; S.ss_tau = 0.f; S.ss_q = 0.f; S.ss_q_hat = 0.f; S.ss_decim_counter = 0;
	S.ss_v[0] = S.ss_v[1] = S.ss_v[2] = 0.f;
}

// Equaliser state of the demodulator warp.  Lane l (and lane l+16) owns tap l & 15: its weight, its window element
// (x[0] oldest .. x[14] newest) and |x|^2.  w13 / w14 are warp-uniform copies of the weights of the two newest taps.
struct EqL { cf w, x; float x2; cf w13, w14; };

__device__ __forceinline__ cf conj_mul(cf w, cf v) {          // conj(w) * v
	return make_float2(fmaf(w.x, v.x, w.y * v.y), fmaf(w.x, v.y, -(w.y * v.x)));
}
__device__ __forceinline__ cf half_warp_sum(cf p) {           // sum over the 16 lanes of a half-warp, result in every lane
#pragma unroll
	for(int m = 8; m >= 1; m >>= 1) {
		p.x += __shfl_xor_sync(0xffffffffu, p.x, m);
		p.y += __shfl_xor_sync(0xffffffffu, p.y, m);
	}
	return p;
}
__device__ __forceinline__ void eq_reset(DemodState &S, const DemodTables &T, EqL &E, int l16) {    // eqlms_cccf_reset
	E.w = l16 < HFDL_EQ_LEN ? T.eq_h0[l16] : make_float2(0.f, 0.f);
	E.w13 = T.eq_h0[13]; E.w14 = T.eq_h0[14];
	E.x = make_float2(0.f, 0.f); E.x2 = 0.f;
	S.eq_count = 0; S.eq_buf_full = 0; S.eq_x2_sum = 0.f;
}
__device__ __forceinline__ void eq_push(DemodState &S, EqL &E, cf r, int l16) {                     // eqlms_cccf_push
	const float x2n = r.x * r.x + r.y * r.y;
	const float x20 = __shfl_sync(0xffffffffu, E.x2, 0, 16);
	const float xsx = __shfl_down_sync(0xffffffffu, E.x.x, 1, 16), xsy = __shfl_down_sync(0xffffffffu, E.x.y, 1, 16);
	const float x2s = __shfl_down_sync(0xffffffffu, E.x2, 1, 16);
	E.x = l16 == 14 ? r : make_float2(xsx, xsy);
	E.x2 = l16 == 14 ? x2n : x2s;
	S.eq_x2_sum = S.eq_x2_sum + x2n - x20;
	S.eq_count++;
}
__device__ __forceinline__ cf eq_execute(const EqL &E, int l16) {                                   // eqlms_cccf_execute
	cf p = conj_mul(E.w, E.x);
	if(l16 >= HFDL_EQ_LEN) p = make_float2(0.f, 0.f);
	return half_warp_sum(p);
}
// eqlms_cccf_step(d, d_hat = s): w += mu * conj(d - d_hat) * x / sum|x|^2, mu = 0.1 (hfdl.c:496,730-733)
__device__ __forceinline__ void eq_step(DemodState &S, EqL &E, float d, cf s, cf x13, cf x14) {
	bool run = true;
	if(!S.eq_buf_full) { if(S.eq_count < HFDL_EQ_LEN) run = false; else S.eq_buf_full = 1; }
	if(run) {
		const float inv = __fdividef(1.0f, S.eq_x2_sum);
		const cf t = make_float2(0.1f * (d - s.x) * inv, 0.1f * s.y * inv);
		const cf u = cmul(t, E.x), u13 = cmul(t, x13), u14 = cmul(t, x14);
		E.w.x += u.x; E.w.y += u.y;
		E.w13.x += u13.x; E.w13.y += u13.y;
		E.w14.x += u14.x; E.w14.y += u14.y;
	}
}
__device__ __forceinline__ void framer_reset(DemodState &S, const DemodTables &T, EqL &E, int l16) {   // hfdl.c:968-991
	S.fr_state = HF_A1; S.symbols_wanted = 1; S.search_retries = 0; S.cur_arity = 1;
	S.train_bits_total = S.train_bits_bad = 0; S.T_idx = 0; S.cur_buf = 0;
	eq_reset(S, T, E, l16);
	S.data_n = 0; S.training_n = 0;
	ss_reset(S);
	S.s_state = HS_EMIT_BITS; S.bitmask = 0;
}

// hard decision of liquid's modem_demodulate for BPSK / PSK4 / PSK8 (gray-coded symbol) + re-modulated point
__device__ __forceinline__ unsigned modem_demod(int m, cf x, const cf (*psk)[8], cf *x_hat) {
	unsigned sym;
	if(m == 1) {
		sym = (x.x > 0.f) ? 0u : 1u;
		*x_hat = make_float2(sym ? -1.0f : 1.0f, 0.f);
	} else {
		// nearest constellation angle k*2pi/M: the same decision regions as liquid's atan2 + linear search
		unsigned k;
		const float ax = fabsf(x.x), ay = fabsf(x.y);
		if(m == 2) {
			k = (ax >= ay) ? (x.x > 0.f ? 0u : 2u) : (x.y > 0.f ? 1u : 3u);
		} else {
			const float t8 = 0.41421356237f;                 // tan(pi/8)
			if(ay < t8 * ax) k = x.x > 0.f ? 0u : 4u;
			else if(ax < t8 * ay) k = x.y > 0.f ? 2u : 6u;
			else k = x.x > 0.f ? (x.y > 0.f ? 1u : 7u) : (x.y > 0.f ? 3u : 5u);
		}
		sym = k ^ (k >> 1);
		*x_hat = psk[m][sym];
	}
	return sym;
}

#define HFDL_UNLIKELY(x) __builtin_expect(!!(x), 0)
#define HFDL_LIKELY(x) __builtin_expect(!!(x), 1)
// a value polled from shared memory is made warp-uniform (lane 0's view) so that every lane takes the same branch
#define HFDL_UNI(v) __shfl_sync(0xffffffffu, (int)(v), 0)

// ---- output ring ---------------------------------------------------------------------------------------------
// Entry i of the ring is 16 bytes of data {sym.re, sym.im, AGC level, info} in lk_ring[i] plus a validity tag in
// lk_tags[i].  info = more[20] | k[19:0]: "another output of the same input sample follows", input-sample index of
// the output.  tag = gen[10:4] | lap[3:0]: reset generation of the timing warp, (sequence number / ring size) mod 16.
// The producer writes the data, then (after a fence) the tag; a consumer reads the tag first and the data with a
// load whose ADDRESS depends on the tag value, so "tag valid" implies "data complete" without assuming that a
// 16-byte shared-memory access is single-copy atomic between warps (it is not: a combined {data, tag} vector was
// observed torn when the consumer runs right behind the producer).
#define HFDL_LK_TAG_INVALID 0x7FFFFFFFu          // bit 31 clear like every real tag (see lk_pair_load)
__device__ __forceinline__ unsigned lk_info(int more, int k) { return ((unsigned)(more & 1) << 20) | (unsigned)k; }
__device__ __forceinline__ unsigned lk_tag_hi(int gen, int seq) { return (((unsigned)gen & 0x7Fu) << 4) | (((unsigned)seq >> 6) & 0xFu); }
__device__ __forceinline__ bool lk_tag_ok(unsigned tag, int gen, int seq) { return tag == lk_tag_hi(gen, seq); }

__device__ __forceinline__ float costas_wrap_fwd(float phi) {      // (double)phi > M_PI  <=>  phi > 3.1415925f; 2*pi split hi+lo
	const float dn = (phi - 6.2831855f) + 1.7484555e-7f, up = (phi + 6.2831855f) - 1.7484555e-7f;
	float w_ = phi;
	w_ = phi > 3.1415925f ? dn : w_;
	w_ = phi < -3.1415925f ? up : w_;
	return w_;
}
// Costas step + rotation of one symsync output (hfdl.c:250-267,709-710)
__device__ __forceinline__ cf costas_rotate(DemodState &S, float re, float im) {
	S.c_phi = costas_wrap_fwd(S.c_phi + S.c_dphi);
	float sn, cs;
	hfdl_sincos_fast(S.c_phi, &sn, &cs);
	return make_float2(re * cs + im * sn, im * cs - re * sn);
}

// Scheduling fence for the single sequential warp: both values are complete before anything that is derived from
// them afterwards is issued.  Used to keep the shuffle stages of a reduction apart from their consumers so that
// the shuffle latency overlaps with an independent dependent chain (ptxas otherwise packs producer and consumer
// together and the in-order warp eats the full latency of every stage).
#if defined(HFDL_CUSIM) || defined(HFDL_NO_ORDER)
#define HFDL_ORDER2(a, b) do { } while(0)
#else
#define HFDL_ORDER2(a, b) asm volatile("" : "+f"(a), "+f"(b))
#endif

// ---- shared memory: one LkShared per channel of the CTA.  The names below resolve through the reference `sh` that every
// function touching the rings takes (or declares) -------------------------------------------------------------------
struct LkShared {
	float4 ring[HFDL_LK_RING];                   // output ring data: {sym.re, sym.im, AGC level, info}
	__align__(8) unsigned tags[HFDL_LK_RING];    // output ring validity tags (read as aligned pairs)
	__align__(16) float lvl[HFDL_LK_BR];         // AGC level ring (1/g after the sample's update)
	cf psk[4][8];
	cf train[16];
	volatile int loaded, end_seq, done;
	__align__(8) hfdl_mbar_t mbar[HFDL_LK_INFLIGHT];      // completion barriers of the loader's bulk copies
	__align__(8) volatile int tailv[2];          // {sequence number, input-sample index} the demodulator warp has passed
	volatile int reset_gen, reset_k, reset_seq, ack_gen;
};
__shared__ __align__(16) LkShared lk_sh[HFDL_LK_NCH_MAX];
#define lk_ring (sh.ring)
#define lk_tags (sh.tags)
#define lk_lvl (sh.lvl)
#define lk_psk (sh.psk)
#define lk_train (sh.train)
#define lk_loaded (sh.loaded)
#define lk_end_seq (sh.end_seq)
#define lk_done (sh.done)
#define lk_mbar (sh.mbar)
#define lk_tailv (sh.tailv)
#define lk_tail (sh.tailv[0])
#define lk_tail_k (sh.tailv[1])
#define lk_reset_gen (sh.reset_gen)
#define lk_reset_k (sh.reset_k)
#define lk_reset_seq (sh.reset_seq)
#define lk_ack_gen (sh.ack_gen)
__device__ __forceinline__ void lk_tail_publish(LkShared &sh, int seq, int k) {      // one 8-byte store on the hot path
#ifdef HFDL_CUSIM
	lk_tailv[0] = seq; lk_tailv[1] = k;
#else
	const unsigned sa = (unsigned)__cvta_generic_to_shared(const_cast<int *>(lk_tailv));
	asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(sa), "r"(seq), "r"(k) : "memory");
#endif
}

// Output-ring accessors; all accesses are volatile (the other warp changes the ring behind the compiler's back).
#ifdef HFDL_CUSIM
static inline void lk_ring_store(LkShared &sh, int i, float x, float y, float z, unsigned info, unsigned tag) {
	volatile float *p = reinterpret_cast<volatile float *>(&lk_ring[i]);
	p[0] = x; p[1] = y; p[2] = z; reinterpret_cast<volatile unsigned *>(p)[3] = info;
	__atomic_thread_fence(__ATOMIC_RELEASE);
	reinterpret_cast<volatile unsigned *>(lk_tags)[i] = tag;
}
// Host threads are not in lockstep, so the emulation reads through lane 0 and broadcasts (called by all 32 lanes).
static inline unsigned lk_ring_tag(LkShared &sh, int i) {
	unsigned t = reinterpret_cast<volatile unsigned *>(lk_tags)[i];
	__atomic_thread_fence(__ATOMIC_ACQUIRE);
	return __shfl_sync(0xffffffffu, t, 0);
}
static inline float4 lk_ring_load(LkShared &sh, int i, unsigned = 0u) {
	volatile float *p = reinterpret_cast<volatile float *>(&lk_ring[i]);
	float4 r;
	r.x = p[0]; r.y = p[1]; r.z = p[2]; r.w = p[3];
	r.x = __shfl_sync(0xffffffffu, r.x, 0); r.y = __shfl_sync(0xffffffffu, r.y, 0); r.z = __shfl_sync(0xffffffffu, r.z, 0);
	r.w = __shfl_sync(0xffffffffu, r.w, 0);
	return r;
}
static inline void lk_pair_load(LkShared &sh, int seq, float4 &e0, float4 &e1, unsigned &t0, unsigned &t1) {
	const int i = seq & (HFDL_LK_RING - 1);
	t0 = lk_ring_tag(sh, i); t1 = lk_ring_tag(sh, i + 1);
	e0 = lk_ring_load(sh, i); e1 = lk_ring_load(sh, i + 1);
}
#else
__device__ __forceinline__ void lk_ring_store(LkShared &sh, int i, float x, float y, float z, unsigned info, unsigned tag) {
	const unsigned sa = (unsigned)__cvta_generic_to_shared(&lk_ring[i]);
	const unsigned ta = (unsigned)__cvta_generic_to_shared(&lk_tags[i]);
	asm volatile("st.volatile.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sa), "f"(x), "f"(y), "f"(z), "f"(__uint_as_float(info)) : "memory");
	__threadfence_block();          // (measured: dropping the fence or using st.release.cta instead changes the kernel time by < 2 %)
	asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(ta), "r"(tag) : "memory");
}
__device__ __forceinline__ unsigned lk_ring_tag(LkShared &sh, int i) {
	const unsigned ta = (unsigned)__cvta_generic_to_shared(&lk_tags[i]);
	unsigned t;
	asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(t) : "r"(ta) : "memory");
	return t;
}
// data of entry i; `dep` is a value that is 0 (bit 31 of a tag) but only known at run time: it makes the load
// address -- and with it the load -- depend on the tag that was read before
__device__ __forceinline__ float4 lk_ring_load(LkShared &sh, int i, unsigned dep = 0u) {
	const unsigned sa = (unsigned)__cvta_generic_to_shared(&lk_ring[i]) + (dep << 4);
	float4 r;
	asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(sa) : "memory");
	return r;
}
// tags and data of the pair (seq, seq + 1), seq even: one 8-byte tag load, then the two data loads behind it
__device__ __forceinline__ void lk_pair_load(LkShared &sh, int seq, float4 &e0, float4 &e1, unsigned &t0, unsigned &t1) {
	const int i = seq & (HFDL_LK_RING - 1);
	const unsigned ta = (unsigned)__cvta_generic_to_shared(&lk_tags[i]);
	asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(t0), "=r"(t1) : "r"(ta) : "memory");
	e0 = lk_ring_load(sh, i, t0 >> 31);
	e1 = lk_ring_load(sh, i + 1, t1 >> 31);
}
#endif

// Cold path of the demodulator warp: both outputs of the next symbol are not in the ring yet.  Polls until they are
// (returns true) or until the timing warp has declared the end of the batch before them (false).  Kept out of line
// so that the hot loop is straight fall-through code.
#ifdef HFDL_INLINE_WAIT
#define HFDL_WAIT_ATTR __forceinline__
#else
#define HFDL_WAIT_ATTR __noinline__
#endif
__device__ HFDL_WAIT_ATTR bool lk_wait_pair(LkShared &sh, int gen, int seq, long long *p_twait) {
	bool ok = false;
	const long long t0 = hfdl_clock();
	for(;;) {
		const int ack = HFDL_UNI(lk_ack_gen), end_seq = HFDL_UNI(lk_end_seq);
		const unsigned t0w = (unsigned)HFDL_UNI(lk_ring_tag(sh, seq & (HFDL_LK_RING - 1)));
		const unsigned t1w = (unsigned)HFDL_UNI(lk_ring_tag(sh, (seq + 1) & (HFDL_LK_RING - 1)));
		if(lk_tag_ok(t0w, gen, seq) && lk_tag_ok(t1w, gen, seq + 1)) { ok = true; break; }
		if(ack == gen && seq + 1 >= end_seq) break;
		HFDL_SPIN_PAUSE();
	}
	*p_twait += hfdl_clock() - t0;
	return ok;
}

// ---- fast runs -------------------------------------------------------------------------------------------------
// Between two framer events the framer only counts symbols down (hfdl.c:774-777) and, inside a frame, neither the
// noise-floor clock nor any loop reset can fire (both need FRAMER_A1_SEARCH).  demod_run() therefore processes
// up to `nsym` whole symbols (an even + an odd symsync output each) as straight-line code specialised on the sampler
// mode and the modulation, and hands control back to the generic per-output path for the symbol that triggers
// the next framer event.  RUN_A1 is the preamble search (hfdl.c:779-793): every symbol is correlated against the A
// sequence, the noise-floor clock ticks per input sample, and the run ends when A1 is found or a loop reset is due.
enum { RUN_BITS = 0, RUN_TRAIN = 1, RUN_DATA = 2, RUN_SKIP = 3, RUN_A1 = 4 };

template <int MODE, int ARITY>
__device__ __forceinline__ int demod_run(LkShared &sh, DemodState &S, EqL &E, int &seq, int &k_prev, const int gen, const int nsym,
		cf *dsym, const int lane, unsigned &symcnt, long long *p_twait, const bool cap, cf *cap_eq, int &cap_n, const int cap_max,
		const float *lvl, const unsigned *A_bits, float &last_lvl, unsigned &last_info) {
	const int l16 = lane & 15;
	int done = 0;
	unsigned bacc = 0; int nacc = 0;             // RUN_BITS: bits collected since the last merge into S.bits
	// E.x stays the canonical 15-element window (lane l16 = element l16).  The equaliser output of a symbol is
	//   s = sum_{i<13} conj(w_i) x_{i+2}  +  conj(w_13) r0 + conj(w_14) r1          (r0, r1: the symbol's two rotated inputs)
	// and only the last two terms depend on the current Costas phase.  P = the first sum is prepared ahead:
	//  * LMS runs (weights change every symbol): after the update, from the window shifted by two (xs);
	//  * all other runs (weights frozen): P = Q + conj(w_11) r0' + conj(w_12) r1' with r0', r1' the previous symbol's
	//    inputs and Q = sum_{i<11} conj(w_i) x_{i+2} -- the latter is known a whole symbol earlier, so its shuffle
	//    reduction is off the phase -> decision -> phase chain.  wq = the weights moved up by four lanes, so that Q of
	//    the NEXT symbol is a plain lane-wise product with the current window.
	constexpr bool LMS = (MODE == RUN_TRAIN);
	float drop0 = 0.f, drop1 = 0.f, x2s = 0.f;
	if(LMS) { drop0 = __shfl_sync(0xffffffffu, E.x2, 0, 16); drop1 = __shfl_sync(0xffffffffu, E.x2, 1, 16); x2s = __shfl_down_sync(0xffffffffu, E.x2, 2, 16); }
	cf xs = make_float2(__shfl_down_sync(0xffffffffu, E.x.x, 2, 16), __shfl_down_sync(0xffffffffu, E.x.y, 2, 16));
	cf P, pl;                                     // pl (LMS runs): P after the shuffle stages 8 and 4; stages 2 and 1 follow at
	{	                                          // the top of the next symbol, overlapped with its Costas step
		cf p = conj_mul(E.w, xs); if(l16 >= 13) p = make_float2(0.f, 0.f);
		p.x += __shfl_xor_sync(0xffffffffu, p.x, 8); p.y += __shfl_xor_sync(0xffffffffu, p.y, 8);
		p.x += __shfl_xor_sync(0xffffffffu, p.x, 4); p.y += __shfl_xor_sync(0xffffffffu, p.y, 4);
		pl = p;
		p.x += __shfl_xor_sync(0xffffffffu, p.x, 2); p.y += __shfl_xor_sync(0xffffffffu, p.y, 2);
		p.x += __shfl_xor_sync(0xffffffffu, p.x, 1); p.y += __shfl_xor_sync(0xffffffffu, p.y, 1);
		P = p;
	}
	cf wq = make_float2(__shfl_up_sync(0xffffffffu, E.w.x, 4, 16), __shfl_up_sync(0xffffffffu, E.w.y, 4, 16));
	if(l16 < 4 || l16 >= HFDL_EQ_LEN) wq = make_float2(0.f, 0.f);
	const cf w11 = make_float2(__shfl_sync(0xffffffffu, E.w.x, 11, 16), __shfl_sync(0xffffffffu, E.w.y, 11, 16));
	const cf w12 = make_float2(__shfl_sync(0xffffffffu, E.w.x, 12, 16), __shfl_sync(0xffffffffu, E.w.y, 12, 16));
	float4 e0, e1; unsigned t0, t1;
	lk_pair_load(sh, seq, e0, e1, t0, t1);
	// Validity of the two prefetched entries is a warp vote: the lanes are not guaranteed to be converged at the
	// prefetch, so the decision must not depend on one lane's view (a lane that saw a not-yet-valid entry sends the
	// whole warp through the reload).  seq is even here: seq and seq + 1 lie in the same lap of the ring.
#define HFDL_PAIR_VALID() __all_sync(0xffffffffu, (t0 == lk_tag_hi(gen, seq)) & (t1 == lk_tag_hi(gen, seq)))
	bool stop = false;
	for(;;) {
		if(!HFDL_PAIR_VALID()) {
			// both outputs of the symbol are not there yet: wait, or leave when the batch ends before them
			if(!lk_wait_pair(sh, gen, seq, p_twait)) break;
			lk_pair_load(sh, seq, e0, e1, t0, t1);
		}
	  // hot loop: one symbol per iteration, left only at the end of the run or when the ring runs dry (the back edge is
	  // its only taken branch)
	  for(;;) {
		// Costas step of both outputs, rotation, and -- in frozen-weight runs -- Q of the NEXT symbol from the current
		// window.  Q is independent of everything else in this iteration; its four shuffle stages are spread over the
		// phase -> sincos -> rotation -> decision chain (HFDL_ORDER2) so that their latency is hidden by that chain.
		cf q = make_float2(0.f, 0.f);
		float qa = 0.f, qb = 0.f;
		if(!LMS) {
			q = conj_mul(wq, E.x);
			qa = __shfl_xor_sync(0xffffffffu, q.x, 8); qb = __shfl_xor_sync(0xffffffffu, q.y, 8);
			xs = make_float2(__shfl_down_sync(0xffffffffu, E.x.x, 2, 16), __shfl_down_sync(0xffffffffu, E.x.y, 2, 16));
		}
		if(LMS) { qa = __shfl_xor_sync(0xffffffffu, pl.x, 2); qb = __shfl_xor_sync(0xffffffffu, pl.y, 2); }
		// branch-free wrap: a rarely-taken branch here splits the block and costs more than the eight selects
		const float phi0 = costas_wrap_fwd(S.c_phi + S.c_dphi);
		float phi1 = costas_wrap_fwd(phi0 + S.c_dphi);
		if(!LMS) {
			HFDL_ORDER2(qa, phi1);
			q.x += qa; q.y += qb;
			qa = __shfl_xor_sync(0xffffffffu, q.x, 4); qb = __shfl_xor_sync(0xffffffffu, q.y, 4);
		} else {
			HFDL_ORDER2(qa, phi1);
			pl.x += qa; pl.y += qb;
			qa = __shfl_xor_sync(0xffffffffu, pl.x, 1); qb = __shfl_xor_sync(0xffffffffu, pl.y, 1);
		}
		float sn0, cs0, sn1, cs1;
		hfdl_sincos_fast(phi0, &sn0, &cs0);
		hfdl_sincos_fast(phi1, &sn1, &cs1);
		if(!LMS) {
			HFDL_ORDER2(qa, sn1);
			q.x += qa; q.y += qb;
			qa = __shfl_xor_sync(0xffffffffu, q.x, 2); qb = __shfl_xor_sync(0xffffffffu, q.y, 2);
		}
		S.c_phi = phi1;
		const cf r0 = make_float2(e0.x * cs0 + e0.y * sn0, e0.y * cs0 - e0.x * sn0);
		cf r1 = make_float2(e1.x * cs1 + e1.y * sn1, e1.y * cs1 - e1.x * sn1);
		if(!LMS) {
			HFDL_ORDER2(qa, r1.x);
			q.x += qa; q.y += qb;
			qa = __shfl_xor_sync(0xffffffffu, q.x, 1); qb = __shfl_xor_sync(0xffffffffu, q.y, 1);
		} else {
			HFDL_ORDER2(qa, r1.x);
			P = make_float2(pl.x + qa, pl.y + qb);
		}
		const float lvl1 = e1.z;
		const int k1 = (int)(__float_as_uint(e1.w) & 0xFFFFFu);
		last_lvl = lvl1; last_info = __float_as_uint(e1.w);
		// ---- eqlms_cccf_push x2 + execute: the 13 older taps are already summed in P
		cf s;
		{
			const cf a = conj_mul(E.w13, r0), b = conj_mul(E.w14, r1);
			s = make_float2((P.x + a.x) + b.x, (P.y + a.y) + b.y);
		}
		if(LMS) S.eq_count += 2;                   // (frozen-weight runs: counters are advanced by `done` after the loop)
		const bool is13 = (l16 == 13), is14 = (l16 == 14);
		seq += 2;
		done++;
		const bool last = (done == nsym);
		if(LMS) {
			const float x2a = r0.x * r0.x + r0.y * r0.y, x2b = r1.x * r1.x + r1.y * r1.y;
			S.eq_x2_sum = S.eq_x2_sum + x2a - drop0;
			S.eq_x2_sum = S.eq_x2_sum + x2b - drop1;
			// window update without branches: lanes 13 / 14 take the two new elements
			E.x.x = is13 ? r0.x : xs.x; E.x.y = is13 ? r0.y : xs.y; E.x2 = is13 ? x2a : x2s;
			E.x.x = is14 ? r1.x : E.x.x; E.x.y = is14 ? r1.y : E.x.y; E.x2 = is14 ? x2b : E.x2;
			// eqlms_cccf_step(T_seq[bitmask&1][T_idx], s)  hfdl.c:730-733
			const float d = (((0x9AFu >> (HFDL_T_LEN - 1 - S.T_idx)) ^ S.bitmask) & 1u) ? -1.0f : 1.0f;
			{	// eqlms_cccf_step with a full window (the dispatcher guarantees S.eq_buf_full)
				const float inv = __fdividef(1.0f, S.eq_x2_sum);
				const cf t = make_float2(0.1f * (d - s.x) * inv, 0.1f * s.y * inv);
				const cf u = cmul(t, E.x), u13 = cmul(t, r0), u14 = cmul(t, r1);
				E.w.x += u.x; E.w.y += u.y;
				E.w13.x += u13.x; E.w13.y += u13.y;
				E.w14.x += u14.x; E.w14.y += u14.y;
			}
			S.T_idx++;
			// prefetch the next symbol's entries, pre-shift the window for it and reduce its 13 known taps
			lk_pair_load(sh, seq, e0, e1, t0, t1);
			drop0 = __shfl_sync(0xffffffffu, E.x2, 0, 16); drop1 = __shfl_sync(0xffffffffu, E.x2, 1, 16);
			xs = make_float2(__shfl_down_sync(0xffffffffu, E.x.x, 2, 16), __shfl_down_sync(0xffffffffu, E.x.y, 2, 16));
			x2s = __shfl_down_sync(0xffffffffu, E.x2, 2, 16);
			pl = conj_mul(E.w, xs); if(l16 >= 13) pl = make_float2(0.f, 0.f);
			qa = __shfl_xor_sync(0xffffffffu, pl.x, 8); qb = __shfl_xor_sync(0xffffffffu, pl.y, 8);
		} else {
			HFDL_ORDER2(qa, s.x);
			const cf Qn = make_float2(q.x + qa, q.y + qb);
			const cf a = conj_mul(w11, r0), b = conj_mul(w12, r1);
			P = make_float2((Qn.x + a.x) + b.x, (Qn.y + a.y) + b.y);
			E.x.x = is13 ? r0.x : xs.x; E.x.y = is13 ? r0.y : xs.y;
			E.x.x = is14 ? r1.x : E.x.x; E.x.y = is14 ? r1.y : E.x.y;
			lk_pair_load(sh, seq, e0, e1, t0, t1);
		}
		// ---- slicer, Costas adjust
		if(cap & (lane == 0) & (cap_n < cap_max)) cap_eq[cap_n] = s;
		cap_n += cap ? 1 : 0;
		cf x_hat;
		unsigned bits = modem_demod(ARITY, s, sh.psk, &x_hat);
		float err = s.y * x_hat.x - s.x * x_hat.y;
		err = 0.5f * (fabsf(err + 1.0f) - fabsf(err - 1.0f));
		S.c_phi += 0.1f * err;
		S.c_dphi += (0.047f * 0.1f * 0.1f) * err;
		if(LMS) {                                  // stages 8 -> 4 of the next P behind the decision chain
			HFDL_ORDER2(qa, S.c_phi);
			pl.x += qa; pl.y += qb;
			qa = __shfl_xor_sync(0xffffffffu, pl.x, 4); qb = __shfl_xor_sync(0xffffffffu, pl.y, 4);
		}
		stop = last;
		if(MODE == RUN_BITS) {
			bacc = (bacc << 1) | ((bits ^ S.bitmask) & 1u);
			if(++nacc == 32) {         // merge 32 bits at once: a whole-word shift of the 127-bit register
				S.bits[3] = S.bits[2] & 0x7FFFFFFFu; S.bits[2] = S.bits[1]; S.bits[1] = S.bits[0]; S.bits[0] = bacc;
				nacc = 0; bacc = 0;
			}
		} else if(MODE == RUN_TRAIN) {
			const bool room = S.training_n < HFDL_T_LEN;
			if(room & (lane == 0)) lk_train[S.training_n] = s;
			S.training_n += room ? 1 : 0;
		} else if(MODE == RUN_DATA) {
			const bool room = S.data_n < HFDL_DATA_SYMS_MAX;
			if(room & (lane == 0)) dsym[S.data_n] = s;
			S.data_n += room ? 1 : 0;
		} else if(MODE == RUN_A1) {
			// noise-floor clock: one tick per input sample (k_prev, k1], update on every 256th (hfdl.c:700-706); the
			// number of updates due is floor((clk_after + 1) / 256) - floor((clk_before + 1) / 256), almost always 0
			const int d = k1 - k_prev;
			const unsigned clk_after = S.nf_clk + (unsigned)d;
			const bool nf_due = ((clk_after + 1u) >> 8) != ((S.nf_clk + 1u) >> 8);
			bits_push(S.bits, bits ^ S.bitmask);
			// |2*eq/127 - 1| > 0.36 in float arithmetic  <=>  eq <= 40 or eq >= 87 (all 128 values enumerated)
			const int eq = bits_corr(A_bits, S.bits);
			const bool hit = (unsigned)(eq - 41) > 45u;
			const bool blow = fabsf(S.c_dphi) > 0.25f;
			// one rarely taken branch for the three rare events (a lone warp pays ~20 cycles per branch it has to resolve)
			if(HFDL_UNLIKELY(nf_due | hit | blow)) {
				if(nf_due) {
					unsigned j = (0xFFu - (S.nf_clk & 0xFFu)) & 0xFFu;
					if(j == 0u) j = 256u;
					for(; (int)j <= d; j += 256u)
						S.noise_floor = 0.65f * S.noise_floor + 0.35f * fminf(S.noise_floor, lvl[k_prev + (int)j]) + 1e-6f;
				}
				if(hit) {                                     // A1 found (hfdl.c:779-793)
					S.st_a1++;
					S.bitmask = eq >= 87 ? 0u : ~0u;
					S.signal_level = lvl1;
					S.frame_symbol_cnt = 1.0f;
					S.symbols_wanted = HFDL_A_LEN;
					S.search_retries = 0;
					S.fr_state = HF_A2;
				}
				stop = stop | hit | blow;                     // Costas blow-up: the generic path resets the loops
			}
			S.nf_clk = clk_after;
		}
		if(MODE != RUN_A1) S.signal_level += lvl1;    // in-frame: the SUM of the AGC levels (the mean is taken at the frame end)
		if(LMS) { HFDL_ORDER2(qa, S.signal_level); pl.x += qa; pl.y += qb; }
		k_prev = k1;
		if(lane == 0) lk_tail_publish(sh, seq, k_prev);
		if(stop | !HFDL_PAIR_VALID()) break;
	  }
		if(stop) break;
	}
#undef HFDL_PAIR_VALID
	// per-symbol counters of the run
	symcnt += (unsigned)done;
	S.symsync_out_idx += 2u * (unsigned)done;
	if(!LMS) S.eq_count += 2 * done;
	if(MODE != RUN_A1) S.frame_symbol_cnt += (float)done;     // exact: whole numbers far below 2^24
	if(!LMS && done > 0) {             // the |x|^2 bookkeeping of eqlms_cccf_push was skipped: rebuild it from the window
		E.x2 = E.x.x * E.x.x + E.x.y * E.x.y;
		float t = l16 < HFDL_EQ_LEN ? E.x2 : 0.f;
#pragma unroll
		for(int m = 8; m >= 1; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
		S.eq_x2_sum = t;
	}
	if(MODE == RUN_BITS) {             // merge the remaining nacc < 32 bits
		for(int i = nacc - 1; i >= 0; i--) bits_push(S.bits, bacc >> i);
	}
	if(MODE != RUN_A1) S.symbols_wanted -= done;
	return done;
}

// diagnostics wrapper (HFDL_B200_DEBUG): cycles and symbols per run mode, slots 12.. of the per-channel counters
template <int MODE, int ARITY>
__device__ __forceinline__ int demod_run_timed(LkShared &sh, long long *dbg, int c, DemodState &S, EqL &E, int &seq, int &k_prev, const int gen, const int nsym,
		cf *dsym, const int lane, unsigned &symcnt, long long *p_twait, const bool cap, cf *cap_eq, int &cap_n, const int cap_max,
		const float *lvl, const unsigned *A_bits, float &last_lvl, unsigned &last_info) {
	if(!dbg) return demod_run<MODE, ARITY>(sh, S, E, seq, k_prev, gen, nsym, dsym, lane, symcnt, p_twait, cap, cap_eq, cap_n, cap_max, lvl, A_bits, last_lvl, last_info);
	const long long t0 = hfdl_clock(), w0 = *p_twait;
	const int did = demod_run<MODE, ARITY>(sh, S, E, seq, k_prev, gen, nsym, dsym, lane, symcnt, p_twait, cap, cap_eq, cap_n, cap_max, lvl, A_bits, last_lvl, last_info);
	if(lane == 0) {
		const int slot = 12 + 2 * (MODE == RUN_DATA ? 4 + ARITY : MODE);      // BITS 0, TRAIN 1, SKIP 3, A1 4, DATA arity 1..3 -> 5..7
		dbg[c * 32 + slot] += (hfdl_clock() - t0) - (*p_twait - w0);
		dbg[c * 32 + slot + 1] += did;
	}
	return did;
}

// loop_kernel<NCH, ROLE_MAJOR>: grid = ceil(C / NCH), block = lk_threads(NCH, ROLE_MAJOR) (per channel: a demodulator, a
// timing and a loader warp; see "CTA layouts" above), dynamic smem = one bank ring per channel (+ padding).
template <int NCH, bool ROLE_MAJOR>
__global__ void __launch_bounds__(lk_threads(NCH, ROLE_MAJOR)) loop_kernel(LoopArgs a) {
	HFDL_DYN_SMEM(cf, s_bank_all);                  // [NCH][HFDL_LK_BR][32]: arms 0..15 matched, 16..31 derivative
	int slot, warp, lane;
	bool idle_warp = false;
	if(ROLE_MAJOR) {
		const int wcta = threadIdx.x >> 5;
		idle_warp = (wcta == 4 || wcta == 5);
		warp = wcta < 2 ? 0 : (wcta < 4 ? 1 : 2); slot = wcta & 1; lane = threadIdx.x & 31;
	} else {
		slot = threadIdx.x / HFDL_LK_CH_THREADS;
		const int t = threadIdx.x - slot * HFDL_LK_CH_THREADS;
		warp = t >> 5; lane = t & 31;
	}
	const int tid = warp * 32 + lane;               // thread index within the channel's three warps
	const int c = blockIdx.x * NCH + slot;
	const bool live = c < a.n_channels && !idle_warp;
	LkShared &sh = lk_sh[slot];
	cf *s_bank = s_bank_all + slot * (HFDL_LK_BR * 32);
	const DemodTables &T = *a.tab;
	const float *lvl = a.lvl + (long long)(live ? c : 0) * a.lvl_stride;
	const int N = a.n_samples;
	if(tid < HFDL_LK_RING) { lk_ring[tid] = make_float4(0.f, 0.f, 0.f, 0.f); lk_tags[tid] = HFDL_LK_TAG_INVALID; }
	if(tid < 32) lk_psk[tid >> 3][tid & 7] = T.psk[tid >> 3][tid & 7];
	if(tid == 0) {
		for(int i = 0; i < HFDL_LK_INFLIGHT; i++) hfdl_mbar_init(&lk_mbar[i], 1);
		hfdl_fence_mbar_init();
		lk_loaded = 0; lk_tail = live ? (int)(a.state[c].symsync_out_idx & 1u) : 0; lk_tail_k = -1; lk_end_seq = 0x7fffffff; lk_done = 0;
		lk_reset_gen = 0; lk_reset_k = 0; lk_reset_seq = 0; lk_ack_gen = 0;
	}
	__syncthreads();
	if(!live) return;                               // idle warps; the last CTA of a channel count that is not a multiple of NCH

	if(warp == 2) {
		// =========================== loader warp ===========================
		// One TMA bulk copy (cp.async.bulk) per chunk of HFDL_LK_CH samples: 8 KiB of filter-bank rows + the chunk's AGC
		// levels, issued by lane 0, HFDL_LK_INFLIGHT chunks in flight, each completing on its own mbarrier.
		const cf *bank = a.bank + (long long)c * a.bank_stride * 32;
		const int nchunks = (N + HFDL_LK_CH - 1) / HFDL_LK_CH;
		// Issue and completion are polled independently: a chunk waiting for ring space must not hold back the
		// publication of the chunks that are already in flight (the demodulator warp only advances -- and frees ring
		// space -- when the timing warp sees those samples).
		if(lane == 0) {
			int issued = 0, completed = 0;
			while(completed < nchunks) {
				bool progress = false;
				if(issued < nchunks && issued - completed < HFDL_LK_INFLIGHT) {
					const int n0 = issued * HFDL_LK_CH;
					// ring space: samples the demodulator warp has not passed yet must stay (the timing warp may be rolled back to them)
					if(n0 + HFDL_LK_CH <= lk_tail_k + 1 + HFDL_LK_BR - HFDL_LK_CH) {
						__threadfence_block();
						hfdl_fence_proxy_async();                  // the consumers' reads of this ring slot come before the copy's writes
						const int cnt = (N - n0 < HFDL_LK_CH) ? (N - n0) : HFDL_LK_CH;
						const unsigned bank_bytes = (unsigned)cnt * 256u, lvl_bytes = (unsigned)((cnt + 3) & ~3) * 4u;
						hfdl_mbar_t *mb = &lk_mbar[issued % HFDL_LK_INFLIGHT];      // chunk issued-INFLIGHT has completed: the barrier is free
						hfdl_mbar_expect_tx(mb, bank_bytes + lvl_bytes);
						hfdl_bulk_g2s(s_bank + (n0 & (HFDL_LK_BR - 1)) * 32, bank + (long long)n0 * 32, bank_bytes, mb);
						hfdl_bulk_g2s(&lk_lvl[n0 & (HFDL_LK_BR - 1)], &lvl[n0], lvl_bytes, mb);
						hfdl_mbar_arrive_emul(mb);
						issued++; progress = true;
					}
				}
				if(completed < issued) {
					hfdl_mbar_t *mb = &lk_mbar[completed % HFDL_LK_INFLIGHT];
					const unsigned parity = (unsigned)(completed / HFDL_LK_INFLIGHT) & 1u;
					if(hfdl_mbar_try_wait(mb, parity)) {
						__threadfence_block();
						const int ready = (completed + 1) * HFDL_LK_CH;
						lk_loaded = ready < N ? ready : N;
						completed++; progress = true;
					}
				}
				if(!progress) {
					if(lk_done) break;
					HFDL_SPIN_PAUSE_LONG();
				}
			}
		}
	} else if(warp == 1) {
		// =========================== timing warp (producer) ===========================
		DemodState S = a.state[c];
		const cf *mfo = a.mfo + (long long)c * a.mfo_stride + HFDL_MFO_HIST;
		const float ss_a1 = T.ss_a1, ss_a2 = T.ss_a2, ss_b0 = T.ss_b0, ss_radj = T.ss_rate_adj;
		int kn = 0;                         // next input sample (generic stepping), or the sample of the next output (fast loop)
		bool mid = false;                   // generic stepping: sample kn has been consumed, its outputs are being produced
		// sequence number of the next output.  It starts with the parity of the stream's output counter, so that the two
		// outputs of a symbol are always the ring pair (even, odd): the demodulator warp reads them as one aligned pair
		int seq = (int)(S.symsync_out_idx & 1u);
		int my_gen = 0;
		bool finished = false;
		long long t_begin = hfdl_clock(), t_wait = 0, t_blocked = 0, n_full = 0, n_starved = 0, n_fast = 0, n_gen_out = 0, n_exit_k = 0, n_exit_seq = 0;
		// one symsync output at input sample k_, arm index b_ (symsync_crcf_step body); TED_ = timing-error detector runs
#define HFDL_SS_TED(mf_, row_, bb_) do { \
			const cf dmf_ = (row_)[16 + (bb_)]; \
			const float q_ = fminf(fmaxf(mf_.x * dmf_.x + mf_.y * dmf_.y, -1.0f), 1.0f);     /* Re(conj(mf)*dmf), clipped */ \
			S.ss_q = q_; \
			S.ss_v[2] = S.ss_v[1]; S.ss_v[1] = S.ss_v[0]; \
			S.ss_v[0] = q_ - ss_a1 * S.ss_v[1] - ss_a2 * S.ss_v[2]; \
			S.ss_q_hat = ss_b0 * S.ss_v[0]; \
			S.ss_rate += ss_radj * S.ss_q_hat; \
			S.ss_del = S.ss_rate + S.ss_q_hat; \
		} while(0)
		for(;;) {
			// ---------------- control: resets, ring room, loader progress, end of batch ----------------
			const int p_gen = HFDL_UNI(lk_reset_gen);
			if(HFDL_UNLIKELY(p_gen != my_gen)) {            // symsync_crcf_reset posted by the demodulator warp: roll back
				my_gen = p_gen;
				__threadfence_block();
				kn = HFDL_UNI(lk_reset_k) + 1; seq = HFDL_UNI(lk_reset_seq); mid = false;
				ss_reset(S);
				finished = false;
				__syncwarp();
				if(lane == 0) { lk_end_seq = 0x7fffffff; __threadfence_block(); lk_ack_gen = my_gen; }
				continue;
			}
			if(HFDL_UNLIKELY(finished)) {
				if(HFDL_UNI(lk_done)) break;
				const long long t0 = hfdl_clock(); HFDL_SPIN_PAUSE(); t_wait += hfdl_clock() - t0;
				continue;
			}
			if(!mid && kn >= N) {                           // every input sample of the batch has been processed
				finished = true;
				__syncwarp();
				if(lane == 0) { __threadfence_block(); lk_end_seq = seq; }
				continue;
			}
			const int seq_lim = HFDL_UNI(lk_tail) + HFDL_LK_RING - 2;       // entries [tail, seq) are unread
			int k_lim = HFDL_UNI(lk_loaded);
			if(seq >= seq_lim || (!mid && kn >= k_lim)) {
				if(!t_blocked) t_blocked = hfdl_clock();
				if(seq >= seq_lim) { n_full++; HFDL_SPIN_PAUSE_LONG(); }      // a full ring: dozens of symbols ahead of the demodulator
				else { n_starved++; HFDL_SPIN_PAUSE(); }
				continue;
			}
			if(t_blocked) { t_wait += hfdl_clock() - t_blocked; t_blocked = 0; }
			__threadfence_block();

			if(S.ss_since_reset >= HFDL_SS_SUB && S.ss_decim_counter == 1u && S.ss_b >= 0) {
				// ---------------- fast loop: two outputs (one symbol) per iteration, straight-line ----------------
				// A lone warp pays ~20 cycles for every branch it has to resolve, so the pair body has only rarely
				// taken ones: limit refreshes (once per loader chunk / when the ring is full) and the del < 1 case.
				// Every lane stores the (identical) ring entry: no lane-0 branch.
				float tau = S.ss_tau;
				int k = kn + (S.ss_b >> 4), b = S.ss_b & 15;
				tau -= (float)(S.ss_b >> 4);
				int odd = 0;                     // 0: the next output is the non-TED one of the pair
				int rare = 0;
				int k_lim2 = k_lim, seq_lim2 = seq_lim;
				const int seq_in = seq;
				for(;;) {
					if(HFDL_UNLIKELY(k + 4 > k_lim2)) {       // both outputs of the pair lie within 3 samples (del ~ 1.5)
						k_lim2 = HFDL_UNI(lk_loaded); __threadfence_block();
						if(k + 4 > k_lim2) break;
					}
					if(HFDL_UNLIKELY(seq + 2 > seq_lim2)) {
						seq_lim2 = HFDL_UNI(lk_tail) + HFDL_LK_RING - 2;
						if(seq + 2 > seq_lim2) break;
					}
					{	// first output of the symbol: no timing-error detector
						const cf *row = s_bank + (k & (HFDL_LK_BR - 1)) * 32;
						const cf mf = row[b];
						const float level = lk_lvl[k & (HFDL_LK_BR - 1)];
						tau += S.ss_del;
						const int bi = hfdl_round_pos(tau * (float)HFDL_SS_NPFB);
						lk_ring_store(sh, seq & (HFDL_LK_RING - 1), mf.x * 0.33333334f, mf.y * 0.33333334f, level, lk_info(bi < HFDL_SS_NPFB, k), lk_tag_hi(my_gen, seq));
						if(HFDL_UNLIKELY(bi < HFDL_SS_NPFB)) { seq++; odd = 1; b = bi; rare = 1; break; }     // del < 1: another output of the same sample
						const int m = bi >> 4;
						k += m; tau -= (float)m; b = bi & 15;
						if(HFDL_UNLIKELY(k >= k_lim2)) { seq++; odd = 1; break; }      // (cannot happen while del < 3)
					}
					{	// second output: timing-error detector + loop filter
						const cf *row = s_bank + (k & (HFDL_LK_BR - 1)) * 32;
						const cf mf = row[b];
						const float level = lk_lvl[k & (HFDL_LK_BR - 1)];
						HFDL_SS_TED(mf, row, b);
						tau += S.ss_del;
						const int bi = hfdl_round_pos(tau * (float)HFDL_SS_NPFB);
						lk_ring_store(sh, (seq + 1) & (HFDL_LK_RING - 1), mf.x * 0.33333334f, mf.y * 0.33333334f, level, lk_info(bi < HFDL_SS_NPFB, k), lk_tag_hi(my_gen, seq + 1));
						seq += 2;
						if(HFDL_UNLIKELY(bi < HFDL_SS_NPFB)) { b = bi; rare = 1; break; }
						const int m = bi >> 4;
						k += m; tau -= (float)m; b = bi & 15;
					}
				}
				const int k_lim_seen = k_lim2;
				n_fast++; if(k + 4 > k_lim_seen) n_exit_k++; else n_exit_seq++;
				// back to the generic representation
				S.ss_tau = tau; S.ss_b = b; S.ss_decim_counter = odd ? 2u : 1u;
				kn = k; mid = rare != 0;
				if(seq != seq_in || rare) continue;
				// no pair fitted: the ring is full or the loader is less than 4 samples ahead -> poll again; only the last
				// samples of the batch (everything is loaded) go through the generic stepping below
				if(seq + 2 > seq_lim2 || k_lim2 < N || kn >= N) {
					if(kn < N) { if(seq + 2 > seq_lim2) HFDL_SPIN_PAUSE_LONG(); else HFDL_SPIN_PAUSE(); }
					continue;
				}
			}
			// ---------------- generic stepping (after a reset, odd alignment, rare cases) ----------------
			if(!mid) {
				if(S.ss_since_reset < HFDL_SS_SUB) S.ss_since_reset++;      // the push itself happened in bank_kernel
				mid = true;
			}
			if(S.ss_b < HFDL_SS_NPFB) {                                     // while(b < npfb) { output ... }
				const int bb = S.ss_b < 0 ? 0 : S.ss_b;
				const cf *row = s_bank + (kn & (HFDL_LK_BR - 1)) * 32;
				cf mf = row[bb];
				if(S.ss_since_reset < HFDL_SS_SUB) {        // window still filling after a reset: only samples pushed since then count
					mf = make_float2(0.f, 0.f);
					const float *h = T.ss_mf[bb];
					for(int j = (int)S.ss_since_reset - 1; j >= 0; j--) { const cf v = mfo[kn - j]; mf.x += h[j] * v.x; mf.y += h[j] * v.y; }
				}
				if(S.ss_decim_counter == 2u) { S.ss_decim_counter = 0; HFDL_SS_TED(mf, row, bb); }
				S.ss_decim_counter++;
				S.ss_tau += S.ss_del;
				S.ss_b = hfdl_round_pos(S.ss_tau * (float)HFDL_SS_NPFB);
				if(lane == 0) lk_ring_store(sh, seq & (HFDL_LK_RING - 1), mf.x * 0.33333334f, mf.y * 0.33333334f, lk_lvl[kn & (HFDL_LK_BR - 1)], lk_info(S.ss_b < HFDL_SS_NPFB, kn), lk_tag_hi(my_gen, seq));
				seq++; n_gen_out++;
			} else {
				S.ss_tau -= 1.0f; S.ss_b -= HFDL_SS_NPFB;                   // ... then tau -= 1, b -= npfb
				kn++; mid = false;
			}
		}
#undef HFDL_SS_TED
		if(a.dbg_cycles && lane == 0) {
			a.dbg_cycles[c * 32 + 0] += hfdl_clock() - t_begin; a.dbg_cycles[c * 32 + 1] += t_wait;
			a.dbg_cycles[c * 32 + 4] += n_full; a.dbg_cycles[c * 32 + 5] += n_starved; a.dbg_cycles[c * 32 + 6] += seq;
			a.dbg_cycles[c * 32 + 7] += n_fast; a.dbg_cycles[c * 32 + 8] += n_gen_out; a.dbg_cycles[c * 32 + 9] += n_exit_k; a.dbg_cycles[c * 32 + 10] += n_exit_seq;
		}
		// timing-loop state after the last input sample of the batch (kn may lie beyond it: undo the skipped samples)
		__syncwarp();
		if(lane == 0) {
			const int over = kn - N;
			DemodState *G = &a.state[c];
			G->ss_since_reset = S.ss_since_reset; G->ss_decim_counter = S.ss_decim_counter; G->ss_rate = S.ss_rate; G->ss_del = S.ss_del;
			G->ss_tau = S.ss_tau + (float)over; G->ss_bf = 0.f; G->ss_q = S.ss_q; G->ss_q_hat = S.ss_q_hat; G->ss_b = S.ss_b + HFDL_SS_NPFB * over;
			G->ss_v[0] = S.ss_v[0]; G->ss_v[1] = S.ss_v[1]; G->ss_v[2] = S.ss_v[2];
		}
	} else {
		// =========================== demodulator warp (consumer) ===========================
		DemodState S = a.state[c];
		const int l16 = lane & 15;
		const bool cap = (c == a.cap_channel);
		int cap_n_eq = cap ? a.cap_cnt[1] : 0;
		cf *dsym = a.datasym + ((long long)c * a.nslots + S.slot) * HFDL_DATA_SYMS_MAX;
		const unsigned long long cnt_base = S.sample_cnt;
		unsigned symcnt = (unsigned)S.symbol_cnt;
		unsigned A_bits[4];
#pragma unroll
		for(int i = 0; i < 4; i++) A_bits[i] = T.A_bits[i];
		EqL E;
		E.w = l16 < HFDL_EQ_LEN ? a.state[c].eq_w[l16] : make_float2(0.f, 0.f);
		E.x = l16 < HFDL_EQ_LEN ? a.state[c].eq_win[l16] : make_float2(0.f, 0.f);
		E.x2 = l16 < HFDL_EQ_LEN ? a.state[c].eq_x2[l16] : 0.f;
		E.w13 = a.state[c].eq_w[13]; E.w14 = a.state[c].eq_w[14];
		if(lane < 16) lk_train[lane] = lane < HFDL_T_LEN ? a.state[c].training[lane] : make_float2(0.f, 0.f);
		__syncwarp();
#define HFDL_NF_TICK(sidx) do { if(S.fr_state == HF_A1) { if((++S.nf_clk & 0xFFu) == 0xFFu) \
			S.noise_floor = 0.65f * S.noise_floor + 0.35f * fminf(S.noise_floor, lvl[sidx]) + 1e-6f; } } while(0)     /* hfdl.c:700-706 */
		int seq = (int)(S.symsync_out_idx & 1u), k_prev = -1, gen = 0;     // same start as the timing warp (pair alignment)
		float last_lvl = 0.f; unsigned last_info = 0u;                     // AGC level / info word of the last output a run consumed
		bool reset_pending = false;
		long long t_begin = hfdl_clock(), t_wait = 0;
		// symsync_crcf_reset happened while processing the outputs of input sample k: outputs of that sample that were
		// already produced kept their (old-state) value; the timing warp restarts with sample k+1
		auto post_reset = [&](int k) {
			gen = (gen + 1) % 127;
			__syncwarp();
			if(lane == 0) { lk_reset_k = k; lk_reset_seq = seq; __threadfence_block(); lk_reset_gen = gen; }
			reset_pending = false;
		};
		// framer FSM, entered at the symbol on which the symbols_wanted countdown expires (hfdl.c:779-891); k / level:
		// input-sample index and AGC level of that symbol's last output
		auto framer_event = [&](const int k, const float level) {
			switch(S.fr_state) {
			case HF_A1: {
				float corr = 2.0f * (float)bits_corr(A_bits, S.bits) / 127.0f - 1.0f;
				if(fabsf(corr) > 0.36f) {
					S.st_a1++;
					S.bitmask = corr > 0.f ? 0u : ~0u;
					S.signal_level = level;
					S.frame_symbol_cnt = 1.0f;
					S.symbols_wanted = HFDL_A_LEN;
					S.search_retries = 0;
					S.fr_state = HF_A2;
				}
				break; }
			case HF_A2: {
				float corr = 2.0f * (float)bits_corr(A_bits, S.bits) / 127.0f - 1.0f;
				if(fabsf(corr) > 0.3f) {
					S.a2_sample_cnt = cnt_base + (unsigned long long)k;
					S.freq_err_hz = (float)((double)(S.c_dphi * 1800.0f) / (2.0 * M_PI));   // hfdl.c:812
					S.st_a2++;
					S.symbols_wanted = 127;
					S.search_retries = 0;
					S.fr_state = HF_M1;
				} else if(++S.search_retries >= 3) {
					framer_reset(S, T, E, l16); reset_pending = true;
				}
				break; }
			case HF_M1: {
				float max_corr = 0.f; int max_idx = -1;
				for(int idx = 0; idx < 8; idx++) {
					float corr = fabsf(2.0f * (float)bits_corr(T.M1_bits[idx], S.bits) / 127.0f - 1.0f);
					if(corr > max_corr) { max_corr = corr; max_idx = idx; }
				}
				if(max_corr > 0.3f) {
					S.st_m1++;
					S.data_segment_cnt = T.mode_segments[max_idx];
					S.data_arity = T.mode_arity[max_idx];
					S.M1 = max_idx;
					S.symbols_wanted = 15;
					S.search_retries = 0;
					S.fr_state = HF_M2_SKIP;
					S.s_state = HS_SKIP;
				} else {
					S.st_m1_fail++;                  // statsd demod.preamble.errors.M1_not_found (hfdl.c:840)
					framer_reset(S, T, E, l16); reset_pending = true;
				}
				break; }
			case HF_M2_SKIP:
				S.training_n = 0;
				S.symbols_wanted = HFDL_T_LEN;
				S.eq_train_seq_cnt = 9;
				S.fr_state = HF_EQ_TRAIN;
				S.s_state = HS_EMIT_SYMBOLS;
				break;
			case HF_EQ_TRAIN: {
				__syncwarp();
				unsigned tseq = 0;                       // compute_train_bit_error_cnt hfdl.c:952-966
#pragma unroll
				for(int j = 0; j < HFDL_T_LEN; j++) {
					unsigned bit = (lk_train[j].x > 0.f) ? 0u : 1u;
					bit ^= (S.bitmask & 1u);
					tseq = (tseq << 1) | bit;
				}
				__syncwarp();
				S.train_bits_total += HFDL_T_LEN;
				S.train_bits_bad += __popc(0x9AFu ^ tseq);
				S.training_n = 0;
				if(S.eq_train_seq_cnt > 1) {
					S.eq_train_seq_cnt--;
					S.symbols_wanted = HFDL_T_LEN;
					S.T_idx = 0;
				} else if(S.data_segment_cnt > 0) {
					S.symbols_wanted = 15;
					S.fr_state = HF_DATA_1;
					S.cur_arity = S.data_arity;
					S.cur_buf = 1;
				} else {                                 // end of frame: hand the symbols to fec_kernel
					int q = 0;
					if(lane == 0) q = atomicAdd(a.nframes, 1);
					q = __shfl_sync(0xffffffffu, q, 0);
					if(q < a.max_frames && lane == 0) {
						FrameRec fr;
						fr.channel = c; fr.slot = S.slot; fr.M1 = S.M1; fr.bitmask = S.bitmask;
						fr.freq_err_hz = S.freq_err_hz; fr.signal_level = S.signal_level / S.frame_symbol_cnt; fr.noise_floor = S.noise_floor;
						fr.sample_cnt_a2 = S.a2_sample_cnt; fr.sample_cnt_end = cnt_base + (unsigned long long)k;
						fr.train_bits_bad = S.train_bits_bad; fr.train_bits_total = S.train_bits_total;
						a.frames[q] = fr;
					}
					S.st_frames++;
					S.slot = (S.slot + 1) % a.nslots;
					dsym = a.datasym + ((long long)c * a.nslots + S.slot) * HFDL_DATA_SYMS_MAX;
					framer_reset(S, T, E, l16); reset_pending = true;
					symcnt = 0;
				}
				break; }
			case HF_DATA_1:
				S.symbols_wanted = 15;
				S.fr_state = HF_DATA_2;
				break;
			case HF_DATA_2:
				S.data_segment_cnt--;
				S.cur_arity = 1;
				S.cur_buf = 0;
				S.fr_state = HF_EQ_TRAIN;
				S.eq_train_seq_cnt = 1;
				S.symbols_wanted = HFDL_T_LEN;
				S.T_idx = 0;
				break;
			}
		};
		for(;;) {
			// fast path: a run of whole symbols up to AND including the symbol of the next framer event
#define HFDL_RUN(MODE, AR) demod_run_timed<MODE, AR>(sh, a.dbg_cycles, c, S, E, seq, k_prev, gen, nsym, dsym, lane, symcnt, &t_wait, cap, a.cap_eq, cap_n_eq, a.cap_max, lvl, A_bits, last_lvl, last_info)
			if(S.fr_state == HF_A1 && !(S.symsync_out_idx & 1u) && !reset_pending && a.debug_mode < 2
					&& fabsf(S.c_dphi) <= 0.25f && symcnt + 1u < 13u * HFDL_SINGLE_SLOT_FRAME_LEN) {
				const int nsym = (int)(13u * HFDL_SINGLE_SLOT_FRAME_LEN - 1u - symcnt);     // the symbol of the 13-frame timeout goes the generic way
				if(HFDL_RUN(RUN_A1, 1) > 0) continue;
			} else if(S.fr_state > HF_A1 && S.symbols_wanted >= 1 && !(S.symsync_out_idx & 1u) && !reset_pending && a.debug_mode < 2) {
				// DATA_1's framer event only re-arms the countdown for DATA_2 (hfdl.c:878-881): both halves of a data
				// segment go through one run
				const bool fuse = (S.fr_state == HF_DATA_1);
				const int nsym = S.symbols_wanted + (fuse ? 15 : 0);
				int did;
				if(S.s_state == HS_EMIT_BITS) did = HFDL_RUN(RUN_BITS, 1);
				else if(S.s_state == HS_SKIP) did = HFDL_RUN(RUN_SKIP, 1);
				else if(S.cur_buf == 0) { if(!S.eq_buf_full) goto generic_path; did = HFDL_RUN(RUN_TRAIN, 1); }
				else if(S.cur_arity == 1) did = HFDL_RUN(RUN_DATA, 1);
				else if(S.cur_arity == 2) did = HFDL_RUN(RUN_DATA, 2);
				else did = HFDL_RUN(RUN_DATA, 3);
				if(did > 0) {
					if(fuse && S.symbols_wanted <= 0) { S.fr_state = HF_DATA_2; S.symbols_wanted += 15; }      // the DATA_1 event, in passing
					if(S.symbols_wanted == 0) {            // the countdown expired on the run's last symbol: framer event
						S.symbols_wanted = 1;
						framer_event(k_prev, last_lvl);
						if(HFDL_UNLIKELY(reset_pending) && !((last_info >> 20) & 1u)) post_reset(k_prev);
					}
					continue;
				}
			}
#undef HFDL_RUN
		generic_path:
			// ---- generic path: one symsync output
			unsigned tagw;
			bool have = false;
			for(;;) {
				tagw = (unsigned)HFDL_UNI(lk_ring_tag(sh, seq & (HFDL_LK_RING - 1)));
				if(lk_tag_ok(tagw, gen, seq)) { have = true; break; }
				if(HFDL_UNI(lk_ack_gen) == gen && seq >= HFDL_UNI(lk_end_seq)) break;
				const long long t0 = hfdl_clock(); HFDL_SPIN_PAUSE(); t_wait += hfdl_clock() - t0;
			}
			if(!have) break;                                          // end of batch: every output consumed
			const float4 ent = lk_ring_load(sh, seq & (HFDL_LK_RING - 1), tagw >> 31);
			tagw = __float_as_uint(ent.w);
			const int k = (int)(tagw & 0xFFFFFu);
			const bool more = (tagw >> 20) & 1u;
			const float level = ent.z;
			if(HFDL_UNLIKELY(a.debug_mode == 3)) {      // diagnostics: drain the ring without demodulating (speed of the timing warp alone)
				k_prev = k; seq++;
				if(lane == 0) { lk_tail = seq; lk_tail_k = k; }
				continue;
			}
			// noise-floor clock ticks once per input sample, before that sample's outputs (hfdl.c:700)
			if(HFDL_UNLIKELY(S.fr_state == HF_A1)) { for(int sidx = k_prev + 1; sidx <= k; sidx++) HFDL_NF_TICK(sidx); }
			k_prev = k;
			do {
				// ---- Costas step + rotate (hfdl.c:250-294,709-715)
				const cf r = costas_rotate(S, ent.x, ent.y);
				if(HFDL_UNLIKELY(S.fr_state == HF_A1 && fabsf(S.c_dphi) > 0.25f)) {
					S.c_phi = S.c_dphi = 0.f;
					reset_pending = true;                              // symsync_crcf_reset
				}
				eq_push(S, E, r, l16);
				if(!(S.symsync_out_idx & 1u)) break;
				const cf s = eq_execute(E, l16);
				if(S.fr_state == HF_EQ_TRAIN) {        // eqlms_cccf_step(T_seq[bitmask&1][T_idx], s)  hfdl.c:730-733
					float d = ((0x9AFu >> (HFDL_T_LEN - 1 - S.T_idx)) & 1u) ? -1.0f : 1.0f;
					if(S.bitmask & 1u) d = -d;
					const cf x13 = make_float2(__shfl_sync(0xffffffffu, E.x.x, 13, 16), __shfl_sync(0xffffffffu, E.x.y, 13, 16));
					eq_step(S, E, d, s, x13, r);
					S.T_idx++;
				}
				if(HFDL_UNLIKELY(cap)) { if(lane == 0 && cap_n_eq < a.cap_max) a.cap_eq[cap_n_eq] = s; cap_n_eq++; }
				cf x_hat;
				unsigned bits = modem_demod(S.cur_arity, s, sh.psk, &x_hat);
				// ---- costas adjust with the modem's phase error Im(r*conj(x_hat)) (hfdl.c:738,276-281)
				float err = s.y * x_hat.x - s.x * x_hat.y;
				err = 0.5f * (fabsf(err + 1.0f) - fabsf(err - 1.0f));     // branchless_limit, hfdl.c:269-274
				S.c_phi += 0.1f * err;
				S.c_dphi += (0.047f * 0.1f * 0.1f) * err;

				symcnt++;
				if(HFDL_UNLIKELY(S.fr_state == HF_A1 && symcnt >= 13u * HFDL_SINGLE_SLOT_FRAME_LEN)) {
					symcnt = 0;
					S.c_phi = S.c_dphi = 0.f;
					reset_pending = true;
				}
				if(S.s_state == HS_EMIT_BITS) {
					bits ^= S.bitmask;
					for(int bb = 0; bb < S.cur_arity; bb++, bits >>= 1) bits_push(S.bits, bits);
				} else if(S.s_state == HS_EMIT_SYMBOLS) {
					if(S.cur_buf == 0) {
						if(S.training_n < HFDL_T_LEN) { if(lane == 0) lk_train[S.training_n] = s; S.training_n++; }
					} else {
						if(S.data_n < HFDL_DATA_SYMS_MAX) { if(lane == 0) dsym[S.data_n] = s; S.data_n++; }
					}
				}
				if(S.fr_state > HF_A1) {
					S.signal_level += level;               // running mean of hfdl.c:768 kept as sum / count
					S.frame_symbol_cnt += 1.0f;
				}
				if(S.symbols_wanted > 1) { S.symbols_wanted--; break; }

				framer_event(k, level);
			} while(0);
			S.symsync_out_idx++;
			seq++;
			__syncwarp();
			if(lane == 0) { lk_tail = seq; lk_tail_k = k; }
			if(HFDL_UNLIKELY(reset_pending) && !more) post_reset(k);
		}
		if(a.dbg_cycles && lane == 0) { a.dbg_cycles[c * 32 + 2] += hfdl_clock() - t_begin; a.dbg_cycles[c * 32 + 3] += t_wait; }
		for(int sidx = k_prev + 1; sidx < N; sidx++) HFDL_NF_TICK(sidx);      // input samples after the last output
#undef HFDL_NF_TICK
		S.sample_cnt = cnt_base + (unsigned long long)N;
		S.symbol_cnt = symcnt;
		__syncwarp();
		DemodState *G = &a.state[c];
		if(lane < HFDL_EQ_LEN) { G->eq_w[lane] = E.w; G->eq_win[lane] = E.x; G->eq_x2[lane] = E.x2; G->training[lane] = lk_train[lane]; }
		__syncwarp();
		if(lane == 0) {
			// everything except the timing-loop fields (owned by the timing warp) and the per-lane equaliser arrays (written above)
			G->eq_x2_sum = S.eq_x2_sum; G->eq_count = S.eq_count; G->eq_buf_full = S.eq_buf_full;
			G->c_phi = S.c_phi; G->c_dphi = S.c_dphi;
			for(int i = 0; i < 4; i++) G->bits[i] = S.bits[i];
			G->training_n = S.training_n; G->data_n = S.data_n; G->cur_buf = S.cur_buf; G->slot = S.slot;
			G->symbol_cnt = S.symbol_cnt; G->sample_cnt = S.sample_cnt; G->a2_sample_cnt = S.a2_sample_cnt;
			G->s_state = S.s_state; G->fr_state = S.fr_state; G->data_arity = S.data_arity; G->cur_arity = S.cur_arity;
			G->symbols_wanted = S.symbols_wanted; G->search_retries = S.search_retries; G->eq_train_seq_cnt = S.eq_train_seq_cnt;
			G->data_segment_cnt = S.data_segment_cnt; G->train_bits_total = S.train_bits_total; G->train_bits_bad = S.train_bits_bad;
			G->T_idx = S.T_idx; G->M1 = S.M1; G->bitmask = S.bitmask; G->symsync_out_idx = S.symsync_out_idx;
			G->freq_err_hz = S.freq_err_hz; G->signal_level = S.signal_level; G->noise_floor = S.noise_floor;
			G->nf_clk = S.nf_clk; G->frame_symbol_cnt = S.frame_symbol_cnt;
			G->st_a1 = S.st_a1; G->st_a2 = S.st_a2; G->st_m1 = S.st_m1; G->st_frames = S.st_frames; G->st_m1_fail = S.st_m1_fail;
			if(cap) a.cap_cnt[1] = cap_n_eq;
			__threadfence_block();
			lk_done = 1;
		}
	}
}
