// dumphfdl_b200/csrc/demod_kernels.cuh -- per-channel HFDL demodulator + framer (K6-K11) and
// FEC (K12-K14) kernels for sm_100a.
//
// The reference runs the whole per-sample chain of hfdl_decoder_thread (hfdl.c:685-893) in one loop.
// Its data dependencies split it into a feed-forward pipeline of three kernels:
//   agc_kernel   K6      AGC (agc_crcf_execute, hfdl.c:686): a strictly sequential 2-state nonlinear
//                        recurrence per channel, nothing downstream feeds back into it -> one thread per
//                        channel runs the bare recurrence and writes gain-controlled samples + 1/g.
//   bank_kernel  K7+K8a  matched filter (firfilt_crcf, hfdl.c:694-695) and the symbol synchroniser's 16-arm
//                        polyphase matched / derivative filter banks (symsync_crcf, hfdl.c:503,707) evaluated
//                        for EVERY sample and EVERY arm: pure FIR work, parallel over samples x arms.
//   loop_kernel  K8b-K11 the feedback part: timing loop picks an arm per output, Costas rotation, T/2 LMS
//                        equaliser, M-PSK slicer, sampler and the framer FSM (hfdl.c:708-891).  One warp
//                        per channel; lanes hold a prefetched ring of bank rows (lane = arm), the arm the
//                        loop selects is fetched with one warp shuffle, so the sequential critical path has
//                        no dependent global load.
// fec_kernel is decode_user_data (hfdl.c:993-1056): descramble + soft demod, 40-row deinterleaver
// as a closed-form scatter/gather, chip averaging for r=1/4, K=7 Viterbi bit-exact with
// libfec/viterbi27_port.c (one warp: 32 butterflies in 32 lanes), byte reversal and the frame check
// (pdu.c:68-79, mpdu.c:56-85, spdu.c:55-64).
#pragma once
#include "common.cuh"

#define HFDL_MF_TAPS 19
#define HFDL_SS_NPFB 16
#define HFDL_SS_SUB 18
#define HFDL_EQ_LEN 15
#define HFDL_T_LEN 15
#define HFDL_A_LEN 127
#define HFDL_DATA_SYMS_MAX 5040
#define HFDL_MAX_PDU 945
#define HFDL_SINGLE_SLOT_FRAME_LEN 4219     // hfdl.c:41
#define HFDL_FRAME_SLOTS 4                  // data-symbol buffers per channel (frames in flight per batch)
#define HFDL_AGC_HIST (HFDL_MF_TAPS - 1 + HFDL_SS_SUB - 1)   // AGC-output samples kept in front of a batch (35)
#define HFDL_MFO_HIST (HFDL_SS_SUB - 1)                     // matched-filter outputs kept in front of a batch (17)
#define HFDL_BANK_TILE 64
#define HFDL_LOOP_CH 32                     // samples per shared-memory prefetch chunk of loop_kernel

enum { HS_EMIT_BITS = 1, HS_EMIT_SYMBOLS = 2, HS_SKIP = 3 };
enum { HF_A1 = 1, HF_A2, HF_M1, HF_M2_SKIP, HF_EQ_TRAIN, HF_DATA_1, HF_DATA_2 };

// constants shared by all channels (designed on the host at create time, see design.hpp)
struct DemodTables {
	float mf[HFDL_MF_TAPS];                          // hfdl.c:148-154
	float ss_mf[HFDL_SS_NPFB][HFDL_SS_SUB];          // symsync matched-filter bank
	float ss_dmf[HFDL_SS_NPFB][HFDL_SS_SUB];         // derivative bank
	float ss_b0, ss_a1, ss_a2, ss_rate_adj;          // timing loop filter (lf_bw 0.001, hfdl.c:504)
	cf eq_h0[HFDL_EQ_LEN];                           // eqlms lowpass initial weights (hfdl.c:495)
	unsigned A_bits[4];                              // 127-bit templates, bit 0 of word 0 = newest
	unsigned M1_bits[8][4];
	cf psk[4][8];                                    // [arity][symbol] constellation (liquid modem PSK, gray coded)
	unsigned char scr[120];                          // scrambler bits (hfdl.c:333-345)
	int mode_arity[8], mode_segments[8], mode_code_rate[8], mode_col_shift[8];   // hfdl.c:81-138
};

struct AgcState { float g, y2; };

struct DemodState {                                  // loop_kernel state carried between batches
	unsigned ss_since_reset;                         // pushes since the last symsync reset (saturates at 18)
	unsigned ss_decim_counter;
	float ss_rate, ss_del, ss_tau, ss_bf, ss_q, ss_q_hat;
	int ss_b;
	float ss_v[3];
	float c_phi, c_dphi;
	cf eq_w[HFDL_EQ_LEN], eq_win[HFDL_EQ_LEN];      // eq_win[0] oldest
	float eq_x2[HFDL_EQ_LEN], eq_x2_sum;
	unsigned eq_count; int eq_buf_full;
	unsigned bits[4];                                // 127-bit shift register
	cf training[HFDL_T_LEN]; int training_n;
	int data_n, cur_buf, slot;
	unsigned long long symbol_cnt, sample_cnt, a2_sample_cnt;
	int s_state, fr_state, data_arity, cur_arity;
	int symbols_wanted, search_retries, eq_train_seq_cnt, data_segment_cnt;
	int train_bits_total, train_bits_bad, T_idx, M1;
	unsigned bitmask, symsync_out_idx;
	float freq_err_hz, signal_level, noise_floor;
	unsigned nf_clk; float frame_symbol_cnt;
	int st_a1, st_a2, st_m1, st_frames;
};

struct FrameRec {          // one completed frame handed from loop_kernel to fec_kernel
	int channel, slot, M1;
	unsigned bitmask;
	float freq_err_hz, signal_level, noise_floor;
	unsigned long long sample_cnt_a2, sample_cnt_end;
	int train_bits_bad, train_bits_total;
};

struct PduRec {            // what the host turns into hfdl_pdu_metadata + octet_string (hfdl.c:1058-1080)
	int channel, M1, len, crc_good;
	float freq_err_hz, signal_level, noise_floor;
	unsigned long long sample_cnt_a2, sample_cnt_end;
	int train_bits_bad, train_bits_total;
	unsigned char octets[HFDL_MAX_PDU + 3];
};

// ======================================================================================
// K6: AGC.  grid = C blocks of 32 threads, lane 0 runs the recurrence of channel blockIdx.x.
//   y = x*g;  y2' = (1-a)*y2' + a*|y|^2;  if(y2' > 1e-6) g *= exp(-0.5*a*ln y2');  g = min(g, 1e6)   (a = 0.01)
// ======================================================================================
struct AgcArgs {
	const cf *rs; long long rs_stride; int n_samples;
	AgcState *state;
	cf *agc_out; long long agc_stride;       // [C][HFDL_AGC_HIST + n]
	float *lvl; long long lvl_stride;        // [C][n]  1/g after the update (agc_crcf_get_signal_level)
};

#define HFDL_AGC_CH 64
__global__ void __launch_bounds__(32) agc_kernel(AgcArgs a) {
	__shared__ cf s_in[2][HFDL_AGC_CH];
	__shared__ cf s_out[HFDL_AGC_CH];
	__shared__ float s_lvl[HFDL_AGC_CH];
	const int c = blockIdx.x, lane = threadIdx.x;
	const cf *x = a.rs + (long long)c * a.rs_stride;
	cf *out = a.agc_out + (long long)c * a.agc_stride + HFDL_AGC_HIST;
	float *lvl = a.lvl + (long long)c * a.lvl_stride;
	float g = a.state[c].g, y2 = a.state[c].y2;
	const float alpha = 0.01f;
	const int N = a.n_samples;
	const int nchunks = (N + HFDL_AGC_CH - 1) / HFDL_AGC_CH;
	for(int pre = 0; pre < 2; pre++) {
		for(int i = lane; i < HFDL_AGC_CH; i += 32) { int n = pre * HFDL_AGC_CH + i; if(n < N) hfdl_cp_async8(&s_in[pre][i], &x[n]); }
		hfdl_cp_async_commit();
	}
	for(int j = 0; j < nchunks; j++) {
		hfdl_cp_async_wait<1>();
		__syncwarp();
		const int n0 = j * HFDL_AGC_CH;
		const int cnt = (N - n0 < HFDL_AGC_CH) ? (N - n0) : HFDL_AGC_CH;
		if(lane == 0) {
			const cf *in = s_in[j & 1];
#pragma unroll 4
			for(int i = 0; i < cnt; i++) {
				cf xv = in[i];
				cf r = make_float2(xv.x * g, xv.y * g);
				float p = r.x * r.x + r.y * r.y;
				// (1.0 - alpha)*y2 + alpha*p with the product term kept exact: y2 - alpha*y2
				y2 = fmaf(alpha, p, fmaf(-alpha, y2, y2));
				// g *= exp(-0.5*alpha*ln(y2)) == 2^(-0.5*alpha*log2(y2)): MUFU.LG2 + MUFU.EX2 on the critical chain
				const float e = hfdl_exp2_fast(-0.5f * alpha * hfdl_log2_fast(y2));
				g = fminf((y2 > 1e-6f) ? g * e : g, 1e6f);
				s_out[i] = r;
				s_lvl[i] = __fdividef(1.0f, g);
			}
		}
		__syncwarp();
		for(int i = lane; i < cnt; i += 32) { out[n0 + i] = s_out[i]; lvl[n0 + i] = s_lvl[i]; }
		__syncwarp();
		{
			const int nn0 = (j + 2) * HFDL_AGC_CH;
			for(int i = lane; i < HFDL_AGC_CH; i += 32) { int n = nn0 + i; if(n < N) hfdl_cp_async8(&s_in[j & 1][i], &x[n]); }
			hfdl_cp_async_commit();
		}
	}
	hfdl_cp_async_wait<0>();
	if(lane == 0) { a.state[c].g = g; a.state[c].y2 = y2; }
}

// ======================================================================================
// K7+K8a: matched filter + symsync filter banks.  grid = (ceil(n/TILE), C), 256 threads.
// bank[c][n][arm]: arm 0..15 = matched-filter arm output (before the 1/k scaling), 16..31 = derivative arm.
// ======================================================================================
struct BankArgs {
	const cf *agc_out; long long agc_stride; int n_samples;
	cf *mfo; long long mfo_stride;           // [C][HFDL_MFO_HIST + n]   matched filter output (f_mf_out)
	cf *bank; long long bank_stride;         // [C][n][32]
	const DemodTables *tab;
};

__global__ void __launch_bounds__(256) bank_kernel(BankArgs a) {
	__shared__ cf s_agc[HFDL_BANK_TILE + HFDL_AGC_HIST];
	__shared__ cf s_mfo[HFDL_BANK_TILE + HFDL_MFO_HIST];
	__shared__ float s_mf[HFDL_MF_TAPS];
	const int c = blockIdx.y;
	const int n0 = blockIdx.x * HFDL_BANK_TILE;
	const int tid = threadIdx.x;
	const DemodTables &T = *a.tab;
	const cf *ag = a.agc_out + (long long)c * a.agc_stride + HFDL_AGC_HIST;     // ag[n], n >= -35 valid
	for(int i = tid; i < HFDL_BANK_TILE + HFDL_AGC_HIST; i += blockDim.x) {
		int n = n0 - HFDL_AGC_HIST + i;
		s_agc[i] = (n < a.n_samples) ? ag[n] : make_float2(0.f, 0.f);
	}
	if(tid < HFDL_MF_TAPS) s_mf[tid] = T.mf[tid];
	__syncthreads();
	// matched filter for samples n0-17 .. n0+TILE-1 (the 17 extra ones re-create the bank's input history)
	for(int i = tid; i < HFDL_BANK_TILE + HFDL_MFO_HIST; i += blockDim.x) {
		const cf *w = s_agc + i + (HFDL_AGC_HIST - HFDL_MFO_HIST);     // sample n = n0-17+i  <->  s_agc[i+18]
		float re = 0.f, im = 0.f;
#pragma unroll
		for(int k = HFDL_MF_TAPS - 1; k >= 0; k--) { re += s_mf[k] * w[-k].x; im += s_mf[k] * w[-k].y; }   // oldest first
		s_mfo[i] = make_float2(re, im);
		int n = n0 - HFDL_MFO_HIST + i;
		if(i >= HFDL_MFO_HIST && n < a.n_samples) a.mfo[(long long)c * a.mfo_stride + HFDL_MFO_HIST + n] = make_float2(re, im);
	}
	__syncthreads();
	const int lane = tid & 31, warp = tid >> 5;
	float h[HFDL_SS_SUB];
#pragma unroll
	for(int j = 0; j < HFDL_SS_SUB; j++) h[j] = (lane < 16) ? T.ss_mf[lane][j] : T.ss_dmf[lane - 16][j];
	for(int s = warp; s < HFDL_BANK_TILE; s += 8) {
		int n = n0 + s;
		if(n >= a.n_samples) break;
		const cf *w = s_mfo + s + HFDL_MFO_HIST;
		float re = 0.f, im = 0.f;
#pragma unroll
		for(int j = HFDL_SS_SUB - 1; j >= 0; j--) { re += h[j] * w[-j].x; im += h[j] * w[-j].y; }
		a.bank[((long long)c * a.bank_stride + n) * 32 + lane] = make_float2(re, im);
	}
}

// carries the tails of the per-channel work streams in front of the next batch
__global__ void demod_carry(cf *agc_out, long long agc_stride, cf *mfo, long long mfo_stride, long long n_new) {
	int c = blockIdx.x, t = threadIdx.x;
	cf *ra = agc_out + (long long)c * agc_stride, *rm = mfo + (long long)c * mfo_stride;
	cf va = make_float2(0.f, 0.f), vm = va;
	if(t < HFDL_AGC_HIST) va = ra[n_new + t];
	if(t < HFDL_MFO_HIST) vm = rm[n_new + t];
	__syncthreads();
	if(t < HFDL_AGC_HIST) ra[t] = va;
	if(t < HFDL_MFO_HIST) rm[t] = vm;
}

// ======================================================================================
// K8b-K11: timing loop, Costas, equaliser, slicer, framer.  One warp per channel; every lane runs the
// same (warp-uniform) scalar state machine, lanes differ only in the bank arm they prefetch.
// ======================================================================================
struct LoopArgs {
	const cf *bank; long long bank_stride;
	const cf *mfo; long long mfo_stride;
	const float *lvl; long long lvl_stride;
	int n_samples;
	DemodState *state;
	const DemodTables *tab;
	cf *datasym;               // [C][HFDL_FRAME_SLOTS][HFDL_DATA_SYMS_MAX]
	FrameRec *frames; int *nframes; int max_frames;
	int cap_channel; cf *cap_eq; int *cap_cnt; int cap_max;      // f_eq_out checkpoint of one channel
	long long *dbg_cycles;     // diagnostics only: [C][4] = timing-warp total / waiting-for-room, demod-warp total / waiting-for-outputs
	int debug_mode;            // diagnostics only (HFDL_B200_DEBUG): 1 = demodulator warp drains the ring without processing
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
// equaliser window / |x|^2 delay line live in a 16-slot shared-memory ring while the kernel runs (ep = oldest slot).
// Every lane owns a private column (slot*32 + lane): lanes run the same state machine but need not stay in lockstep.
struct EqRing { cf *win; float *x2; int ep; };
#define EQS(i) ((i) * 32)
__device__ __forceinline__ void eq_reset(DemodState &S, const DemodTables &T, EqRing &E) {
#pragma unroll
	for(int i = 0; i < HFDL_EQ_LEN; i++) S.eq_w[i] = T.eq_h0[i];
#pragma unroll
	for(int i = 0; i < 16; i++) { E.win[EQS(i)] = make_float2(0.f, 0.f); E.win[EQS(i + 16)] = make_float2(0.f, 0.f); E.x2[EQS(i)] = 0.f; }
	E.ep = 0;
	S.eq_count = 0; S.eq_buf_full = 0; S.eq_x2_sum = 0.f;
}
__device__ __forceinline__ void framer_reset(DemodState &S, const DemodTables &T, EqRing &E) {   // hfdl.c:968-991
	S.fr_state = HF_A1; S.symbols_wanted = 1; S.search_retries = 0; S.cur_arity = 1;
	S.train_bits_total = S.train_bits_bad = 0; S.T_idx = 0; S.cur_buf = 0;
	eq_reset(S, T, E);
	S.data_n = 0; S.training_n = 0;
	ss_reset(S);
	S.s_state = HS_EMIT_BITS; S.bitmask = 0;
}

// hard decision of liquid's modem_demodulate for BPSK / PSK4 / PSK8 (gray-coded symbol) + re-modulated point
__device__ __forceinline__ unsigned modem_demod(int m, cf x, const DemodTables &T, cf *x_hat) {
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
		*x_hat = T.psk[m][sym];
	}
	return sym;
}

// branch layout hints: on a lone warp every TAKEN branch costs a fetch bubble, so rare paths go out of line
#define HFDL_UNLIKELY(x) __builtin_expect(!!(x), 0)
#define HFDL_LIKELY(x) __builtin_expect(!!(x), 1)

// ---- fast in-frame runs -------------------------------------------------------------------------------------
// Between two framer events the framer only counts symbols down (hfdl.c:774-777) and, inside a frame, neither the
// noise-floor clock nor any loop reset can fire (both need FRAMER_A1_SEARCH).  demod_run() therefore processes
// `nsym` whole symbols (an even + an odd symsync output each) as straight-line code specialised on the sampler
// mode and the modulation, and hands control back to the generic per-output path for the symbol that triggers
// the next framer event.  Arithmetic is expression-for-expression the same as in the generic path.
enum { RUN_BITS = 0, RUN_TRAIN = 1, RUN_DATA = 2, RUN_SKIP = 3 };

__device__ __forceinline__ cf costas_rotate_push(DemodState &S, EqRing &E, cf so) {
	S.c_phi += S.c_dphi;
	{
		const float dn = (S.c_phi - 6.2831855f) + 1.7484555e-7f, up = (S.c_phi + 6.2831855f) - 1.7484555e-7f;
		S.c_phi = S.c_phi > 3.1415925f ? dn : (S.c_phi < -3.1415925f ? up : S.c_phi);
	}
	float sn, cs;
	hfdl_sincos_fast(S.c_phi, &sn, &cs);
	cf r = make_float2(so.x * cs + so.y * sn, so.y * cs - so.x * sn);
	const int wp = (E.ep + 15) & 15;
	float x2n = r.x * r.x + r.y * r.y, x20 = E.x2[EQS(E.ep)];
	E.win[EQS(wp)] = r;
	E.win[EQS(wp + 16)] = r;
	E.x2[EQS(wp)] = x2n;
	E.ep = (E.ep + 1) & 15;
	S.eq_x2_sum = S.eq_x2_sum + x2n - x20;
	S.eq_count++;
	return r;
}

template <int MODE, int ARITY>
__device__ __forceinline__ int demod_run(DemodState &S, EqRing &E, const DemodTables &T, const float4 *ring,
		volatile int *p_head, volatile int *p_end, volatile int *p_tail, int &seq, int &k_prev, int nsym,
		cf *s_train, cf *dsym, int lane, unsigned &symcnt, long long *p_twait) {
	int done = 0;
	int c_head = __shfl_sync(0xffffffffu, (int)*p_head, 0);
	while(done < nsym) {
		if(HFDL_UNLIKELY(c_head - seq < 2)) {
			c_head = __shfl_sync(0xffffffffu, (int)*p_head, 0);
			if(c_head - seq < 2) {
				if(seq + 1 >= __shfl_sync(0xffffffffu, (int)*p_end, 0)) break;     // the batch ends inside this run
				{ long long t0 = hfdl_clock(); HFDL_SPIN_PAUSE(); *p_twait += hfdl_clock() - t0; }
				continue;
			}
			__threadfence_block();
		}
		const float4 e0 = ring[seq & 63], e1 = ring[(seq + 1) & 63];
		(void)costas_rotate_push(S, E, make_float2(e0.x, e0.y));
		const cf r = costas_rotate_push(S, E, make_float2(e1.x, e1.y));
		// ---- eqlms_cccf_execute
		cf s = make_float2(0.f, 0.f);
		cf wv[HFDL_EQ_LEN];
		{
			const cf *wb = E.win + EQS(E.ep);
#pragma unroll
			for(int j = 0; j < HFDL_EQ_LEN - 1; j++) wv[j] = wb[EQS(j)];
			wv[HFDL_EQ_LEN - 1] = r;
			cf s2 = make_float2(0.f, 0.f);
#pragma unroll
			for(int j = 0; j < HFDL_EQ_LEN; j++) {
				cf w = S.eq_w[j], v = wv[j];
				if(j & 1) { s2.x = fmaf(w.x, v.x, fmaf(w.y, v.y, s2.x)); s2.y = fmaf(w.x, v.y, fmaf(-w.y, v.x, s2.y)); }
				else { s.x = fmaf(w.x, v.x, fmaf(w.y, v.y, s.x)); s.y = fmaf(w.x, v.y, fmaf(-w.y, v.x, s.y)); }
			}
			s.x += s2.x; s.y += s2.y;
		}
		if(MODE == RUN_TRAIN) {        // eqlms_cccf_step(T_seq[bitmask&1][T_idx], s)  hfdl.c:730-733
			float d = ((0x9AFu >> (HFDL_T_LEN - 1 - S.T_idx)) & 1u) ? -1.0f : 1.0f;
			if(S.bitmask & 1u) d = -d;
			bool run = true;
			if(HFDL_UNLIKELY(!S.eq_buf_full)) { if(S.eq_count < HFDL_EQ_LEN) run = false; else S.eq_buf_full = 1; }
			if(run) {
				const float inv = 1.0f / S.eq_x2_sum;
				cf t = make_float2(0.1f * (d - s.x) * inv, 0.1f * s.y * inv);
#pragma unroll
				for(int j = 0; j < HFDL_EQ_LEN; j++) {
					cf uu = cmul(t, wv[j]);
					S.eq_w[j].x += uu.x;
					S.eq_w[j].y += uu.y;
				}
			}
			S.T_idx++;
		}
		cf x_hat;
		unsigned bits = modem_demod(ARITY, s, T, &x_hat);
		float err = s.y * x_hat.x - s.x * x_hat.y;
		err = 0.5f * (fabsf(err + 1.0f) - fabsf(err - 1.0f));
		S.c_phi += 0.1f * err;
		S.c_dphi += (0.047f * 0.1f * 0.1f) * err;
		symcnt++;
		if(MODE == RUN_BITS) {
			bits ^= S.bitmask;
			bits_push(S.bits, bits);
		} else if(MODE == RUN_TRAIN) {
			if(S.training_n < HFDL_T_LEN) { s_train[EQS(S.training_n)] = s; S.training_n++; }
		} else if(MODE == RUN_DATA) {
			if(S.data_n < HFDL_DATA_SYMS_MAX) { if(lane == 0) dsym[S.data_n] = s; S.data_n++; }
		}
		S.signal_level = __fdividef(S.signal_level * S.frame_symbol_cnt + e1.z, S.frame_symbol_cnt + 1.0f);
		S.frame_symbol_cnt += 1.0f;
		k_prev = __float_as_int(e1.w);
		seq += 2;
		S.symsync_out_idx += 2;
		done++;
		if((done & 3) == 0) { __syncwarp(); if(lane == 0) *p_tail = seq; }
	}
	__syncwarp();
	if(lane == 0) *p_tail = seq;
	S.symbols_wanted -= done;
	return done;
}

#define HFDL_PBATCH 12          // input samples the timing warp handles between two polls of the ring indices
#define HFDL_RING 64           // symsync outputs the timing warp may run ahead of the demodulator warp
// a value polled from shared memory is made warp-uniform (lane 0's view) so that every lane takes the same branch
#define HFDL_UNI(v) __shfl_sync(0xffffffffu, (int)(v), 0)
// loop_kernel: 2 warps per channel.
//   warp 1 ("timing")  runs the symbol-timing recursion (symsync_crcf_step: arm selection, timing-error detector,
//                      loop filter) AHEAD of the demodulator and publishes its outputs into a shared-memory ring.
//                      That recursion does not depend on Costas/equaliser/framer -- except when they reset it
//                      (symsync_crcf_reset from framer_reset, Costas blow-up, 13-frame timeout: hfdl.c:711-715,
//                      746-752, 968-991).  Resets are rare (about once per frame): the demodulator warp posts
//                      (sample, sequence) of the reset and the timing warp rolls back to that point.
//   warp 0 ("demod")   consumes the outputs in order: Costas, equaliser, slicer, sampler, framer.
// Inside a warp every lane runs the same scalar program (warp-uniform); lane 0 publishes ring indices.
__global__ void __launch_bounds__(64) loop_kernel(LoopArgs a) {
	const int c = blockIdx.x;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const DemodTables &T = *a.tab;
	DemodState S = a.state[c];
	const float *lvl = a.lvl + (long long)c * a.lvl_stride;
	const int N = a.n_samples;
	__shared__ float4 s_ring[HFDL_RING];            // {sym.re, sym.im, AGC level, input sample index (int bits)}
	__shared__ volatile int s_head, s_tail, s_end_seq, s_done;
	__shared__ volatile int s_reset_gen, s_reset_k, s_reset_seq, s_ack_gen;
	__shared__ cf s_bank[2][HFDL_LOOP_CH][32];
	__shared__ float s_lvl[2][HFDL_LOOP_CH];
	__shared__ cf s_eqwin[32 * 32];
	__shared__ float s_eqx2[16 * 32];
	__shared__ cf s_train_all[16 * 32];
	if(threadIdx.x == 0) { s_head = 0; s_tail = 0; s_end_seq = 0x7fffffff; s_done = 0; s_reset_gen = 0; s_reset_k = 0; s_reset_seq = 0; s_ack_gen = 0; }
	__syncthreads();

	if(warp == 1) {
		// =========================== timing warp (producer) ===========================
		const cf *bank = a.bank + (long long)c * a.bank_stride * 32;
		const cf *mfo = a.mfo + (long long)c * a.mfo_stride + HFDL_MFO_HIST;
		const float ss_a1 = T.ss_a1, ss_a2 = T.ss_a2, ss_b0 = T.ss_b0, ss_radj = T.ss_rate_adj;
		int chunk = 0, chunk_end = 0;       // chunk readable in s_bank[chunk & 1]; (re)primed by HFDL_STAGE
		int k = -1;                         // last input sample consumed
		int seq = 0;                        // sequence number of the next output
		int my_gen = 0;
		bool finished = false, need_stage = true;
		long long t_begin = hfdl_clock(), t_wait = 0;
#define HFDL_STAGE_CHUNK(ch_, buf_) do { \
			const int nn0_ = (ch_) * HFDL_LOOP_CH; \
			for(int i_ = 0; i_ < HFDL_LOOP_CH; i_++) { int n_ = nn0_ + i_; if(n_ < N) hfdl_cp_async8(&s_bank[buf_][i_][lane], &bank[(long long)n_ * 32 + lane]); } \
			{ int n_ = nn0_ + lane; if(lane < HFDL_LOOP_CH && n_ < N) hfdl_cp_async4(&s_lvl[buf_][lane], &lvl[n_]); } \
			hfdl_cp_async_commit(); } while(0)
#define HFDL_SS_CONSUME() do { if(S.ss_since_reset < HFDL_SS_SUB) S.ss_since_reset++; } while(0)
#define HFDL_SS_OUTPUT(dst) do { \
			const int bb_ = S.ss_b < 0 ? 0 : S.ss_b; \
			cf mf_ = row[bb_]; \
			if(HFDL_UNLIKELY(S.ss_since_reset < HFDL_SS_SUB)) { /* window still filling after a reset: only samples pushed since then count */ \
				mf_ = make_float2(0.f, 0.f); \
				const float *h_ = T.ss_mf[bb_]; \
				for(int j_ = (int)S.ss_since_reset - 1; j_ >= 0; j_--) { cf v_ = mfo[k - j_]; mf_.x += h_[j_] * v_.x; mf_.y += h_[j_] * v_.y; } \
			} \
			dst = make_float2(mf_.x * 0.33333334f, mf_.y * 0.33333334f);      /* output scaled by 1/k, k = 3 samples/symbol */ \
			if(S.ss_decim_counter == 2u) { \
				S.ss_decim_counter = 0; \
				const cf dmf_ = row[16 + bb_]; \
				float q_ = fminf(fmaxf(mf_.x * dmf_.x + mf_.y * dmf_.y, -1.0f), 1.0f);     /* Re(conj(mf)*dmf), clipped */ \
				S.ss_q = q_; \
				S.ss_v[2] = S.ss_v[1]; S.ss_v[1] = S.ss_v[0]; \
				S.ss_v[0] = q_ - ss_a1 * S.ss_v[1] - ss_a2 * S.ss_v[2]; \
				S.ss_q_hat = ss_b0 * S.ss_v[0]; \
				S.ss_rate += ss_radj * S.ss_q_hat; \
				S.ss_del = S.ss_rate + S.ss_q_hat; \
			} \
			S.ss_decim_counter++; \
			S.ss_tau += S.ss_del; \
			S.ss_b = hfdl_round_pos(S.ss_tau * (float)HFDL_SS_NPFB);     /* bf = tau*npfb > 0 here: == (int)roundf(bf) */ \
		} while(0)
		for(;;) {
			const int p_gen = HFDL_UNI(s_reset_gen), p_tail = HFDL_UNI(s_tail), p_done = HFDL_UNI(s_done);
			if(HFDL_UNLIKELY(p_gen != my_gen)) {            // symsync_crcf_reset posted by the demodulator warp: roll back
				my_gen = p_gen;
				__threadfence_block();
				k = HFDL_UNI(s_reset_k); seq = HFDL_UNI(s_reset_seq);
				ss_reset(S);
				finished = false; need_stage = true;
				__syncwarp();
				if(lane == 0) { s_end_seq = 0x7fffffff; s_head = seq; __threadfence_block(); s_ack_gen = my_gen; }
				continue;
			}
			if(HFDL_UNLIKELY(finished)) { if(p_done) break; long long t0 = hfdl_clock(); HFDL_SPIN_PAUSE(); t_wait += hfdl_clock() - t0; continue; }
			if(seq - p_tail > HFDL_RING - HFDL_PBATCH - 4) { long long t0 = hfdl_clock(); HFDL_SPIN_PAUSE(); t_wait += hfdl_clock() - t0; continue; }      // not enough room: wait (keeps polling for resets)
			// several input samples are processed between polls / publications (the demodulator warp is behind anyway)
			for(int pb = 0; pb < HFDL_PBATCH; pb++) {      // one input sample per iteration (2 of 3 yield an output)
				const int kn = k + 1;
				if(HFDL_UNLIKELY(kn >= N)) { finished = true; break; }
				k = kn;
				if(HFDL_UNLIKELY(need_stage)) {                           // (re)prime the two staging buffers at the chunk of sample k
					hfdl_cp_async_wait<0>();
					__syncwarp();
					chunk = k / HFDL_LOOP_CH; chunk_end = (chunk + 1) * HFDL_LOOP_CH;
					HFDL_STAGE_CHUNK(chunk, chunk & 1);
					HFDL_STAGE_CHUNK(chunk + 1, (chunk + 1) & 1);
					hfdl_cp_async_wait<1>();
					__syncwarp();
					need_stage = false;
				}
				while(HFDL_UNLIKELY(k >= chunk_end)) {                    // move to the next staged chunk, refill the one just left
					__syncwarp();
					HFDL_STAGE_CHUNK(chunk + 2, chunk & 1);
					hfdl_cp_async_wait<1>();
					__syncwarp();
					chunk++; chunk_end += HFDL_LOOP_CH;
				}
				HFDL_SS_CONSUME();                       // the push itself happened in bank_kernel
				if(S.ss_b < HFDL_SS_NPFB) {              // symsync_crcf_step: while(b < npfb) { output ... }
					const int ii = k - (chunk_end - HFDL_LOOP_CH);
					const cf *row = s_bank[chunk & 1][ii];
					const float level = s_lvl[chunk & 1][ii];           // 1/g after this sample's AGC update
					cf sym0, sym1 = make_float2(0.f, 0.f);
					int produced = 1;
					HFDL_SS_OUTPUT(sym0);
					if(HFDL_UNLIKELY(S.ss_b < HFDL_SS_NPFB)) {                // further outputs of the same input sample: del < 1, rare
						HFDL_SS_OUTPUT(sym1);
						produced = 2;
						while(S.ss_b < HFDL_SS_NPFB) { cf drop; HFDL_SS_OUTPUT(drop); (void)drop; }
					}
					// all outputs of one input sample are published together
					s_ring[seq & (HFDL_RING - 1)] = make_float4(sym0.x, sym0.y, level, __int_as_float(k));
					if(HFDL_UNLIKELY(produced == 2)) s_ring[(seq + 1) & (HFDL_RING - 1)] = make_float4(sym1.x, sym1.y, level, __int_as_float(k));
					seq += produced;
				}
				S.ss_tau -= 1.0f; S.ss_b -= HFDL_SS_NPFB;      // ... then tau -= 1, b -= npfb
			}   // pb
			__threadfence_block();
			__syncwarp();
			if(lane == 0) { s_head = seq; if(finished) { __threadfence_block(); s_end_seq = seq; } }
		}
#undef HFDL_STAGE_CHUNK
#undef HFDL_SS_CONSUME
#undef HFDL_SS_OUTPUT
		hfdl_cp_async_wait<0>();
		if(a.dbg_cycles && lane == 0) { a.dbg_cycles[c * 4 + 0] += hfdl_clock() - t_begin; a.dbg_cycles[c * 4 + 1] += t_wait; }
	} else {
		// =========================== demodulator warp (consumer) ===========================
		const bool cap = (c == a.cap_channel);
		int cap_n_eq = cap ? a.cap_cnt[1] : 0;
		cf *dsym = a.datasym + ((long long)c * HFDL_FRAME_SLOTS + S.slot) * HFDL_DATA_SYMS_MAX;
		const unsigned long long cnt_base = S.sample_cnt;
		unsigned symcnt = (unsigned)S.symbol_cnt;
		unsigned A_bits[4];
#pragma unroll
		for(int i = 0; i < 4; i++) A_bits[i] = T.A_bits[i];
		EqRing E;
		E.win = s_eqwin + lane; E.x2 = s_eqx2 + lane; E.ep = 0;
		cf *s_train = s_train_all + lane;
		for(int j = 0; j < 16; j++) {
			cf v = j < HFDL_EQ_LEN ? a.state[c].eq_win[j] : make_float2(0.f, 0.f);
			E.win[EQS(j)] = v; E.win[EQS(j + 16)] = v;
			E.x2[EQS(j)] = j < HFDL_EQ_LEN ? a.state[c].eq_x2[j] : 0.f;
			s_train[EQS(j)] = j < HFDL_T_LEN ? a.state[c].training[j] : make_float2(0.f, 0.f);
		}
#define HFDL_NF_TICK(sidx) do { if(S.fr_state == HF_A1) { if((++S.nf_clk & 0xFFu) == 0xFFu) \
			S.noise_floor = 0.65f * S.noise_floor + 0.35f * fminf(S.noise_floor, lvl[sidx]) + 1e-6f; } } while(0)     /* hfdl.c:700-706 */
		int seq = 0, k_prev = -1, gen = 0, wait_seq = 0x7fffffff;
		long long t_begin = hfdl_clock(), t_wait = 0;
		for(;;) {
			if(HFDL_UNLIKELY(seq >= wait_seq)) {                      // outputs from here on must come from the re-started timing loop
				while(HFDL_UNI(s_ack_gen) != gen) { HFDL_SPIN_PAUSE(); }
				__threadfence_block();
				wait_seq = 0x7fffffff;
			}
			// fast path: a run of whole symbols up to (not including) the symbol of the next framer event
			if(!cap && a.debug_mode != 1 && a.debug_mode != 2 && S.fr_state > HF_A1 && S.symbols_wanted > 1 && !(S.symsync_out_idx & 1u) && wait_seq == 0x7fffffff) {
				const int nsym = S.symbols_wanted - 1;
				int did;
				const float4 *rg = s_ring;
				if(S.s_state == HS_EMIT_BITS) did = demod_run<RUN_BITS, 1>(S, E, T, rg, &s_head, &s_end_seq, &s_tail, seq, k_prev, nsym, s_train, dsym, lane, symcnt, &t_wait);
				else if(S.s_state == HS_SKIP) did = demod_run<RUN_SKIP, 1>(S, E, T, rg, &s_head, &s_end_seq, &s_tail, seq, k_prev, nsym, s_train, dsym, lane, symcnt, &t_wait);
				else if(S.cur_buf == 0) did = demod_run<RUN_TRAIN, 1>(S, E, T, rg, &s_head, &s_end_seq, &s_tail, seq, k_prev, nsym, s_train, dsym, lane, symcnt, &t_wait);
				else if(S.cur_arity == 1) did = demod_run<RUN_DATA, 1>(S, E, T, rg, &s_head, &s_end_seq, &s_tail, seq, k_prev, nsym, s_train, dsym, lane, symcnt, &t_wait);
				else if(S.cur_arity == 2) did = demod_run<RUN_DATA, 2>(S, E, T, rg, &s_head, &s_end_seq, &s_tail, seq, k_prev, nsym, s_train, dsym, lane, symcnt, &t_wait);
				else did = demod_run<RUN_DATA, 3>(S, E, T, rg, &s_head, &s_end_seq, &s_tail, seq, k_prev, nsym, s_train, dsym, lane, symcnt, &t_wait);
				if(did > 0) continue;
			}
			int c_head = HFDL_UNI(s_head);
			while(HFDL_UNLIKELY(c_head <= seq) && seq < HFDL_UNI(s_end_seq)) { long long t0 = hfdl_clock(); HFDL_SPIN_PAUSE(); c_head = HFDL_UNI(s_head); t_wait += hfdl_clock() - t0; }
			__threadfence_block();
			if(HFDL_UNLIKELY(c_head <= seq)) break;                   // end of batch: every output consumed
			const float4 ent = s_ring[seq & (HFDL_RING - 1)];
			if(a.debug_mode == 1) { seq++; __syncwarp(); if(lane == 0) s_tail = seq; continue; }
			const int k = __float_as_int(ent.w);
			const float level = ent.z;
			// noise-floor clock ticks once per input sample, before that sample's outputs (hfdl.c:700)
			if(HFDL_UNLIKELY(S.fr_state == HF_A1)) { for(int sidx = k_prev + 1; sidx <= k; sidx++) HFDL_NF_TICK(sidx); }
			k_prev = k;
			bool reset_req = false;
			for(int i = 0; i < 1; i++, S.symsync_out_idx++) {
				// ---- Costas step + rotate (hfdl.c:250-294,709-715)
				S.c_phi += S.c_dphi;
				// (double)phi > M_PI  <=>  phi > 3.1415925f (largest float below pi); 2*pi split hi+lo
				{
					const float dn = (S.c_phi - 6.2831855f) + 1.7484555e-7f, up = (S.c_phi + 6.2831855f) - 1.7484555e-7f;
					S.c_phi = S.c_phi > 3.1415925f ? dn : (S.c_phi < -3.1415925f ? up : S.c_phi);
				}
				float sn, cs;
				hfdl_sincos_fast(S.c_phi, &sn, &cs);
				const cf so = make_float2(ent.x, ent.y);
				cf r = make_float2(so.x * cs + so.y * sn, so.y * cs - so.x * sn);
				if(HFDL_UNLIKELY(S.fr_state == HF_A1 && fabsf(S.c_dphi) > 0.25f)) {
					S.c_phi = S.c_dphi = 0.f;
					{ ss_reset(S); reset_req = true; }
				}
				// ---- eqlms_cccf_push (mirrored ring: slot w and w+16 hold the same element, so the 15-element
				//      window is always the contiguous run [ep, ep+14])
				{
					const int wp = (E.ep + 15) & 15;
					float x2n = r.x * r.x + r.y * r.y, x20 = E.x2[EQS(E.ep)];
					E.win[EQS(wp)] = r;
					E.win[EQS(wp + 16)] = r;
					E.x2[EQS(wp)] = x2n;
					E.ep = (E.ep + 1) & 15;
					S.eq_x2_sum = S.eq_x2_sum + x2n - x20;
					S.eq_count++;
				}
				if(!(S.symsync_out_idx & 1u)) continue;
				// ---- eqlms_cccf_execute: y = sum conj(w[i]) * x[i]
				cf s = make_float2(0.f, 0.f);
				cf wv[HFDL_EQ_LEN];
				{
					const cf *wb = E.win + EQS(E.ep);
	#pragma unroll
					for(int j = 0; j < HFDL_EQ_LEN - 1; j++) wv[j] = wb[EQS(j)];
					wv[HFDL_EQ_LEN - 1] = r;
					cf s2 = make_float2(0.f, 0.f);          // two accumulator pairs shorten the dependent FMA chain
	#pragma unroll
					for(int j = 0; j < HFDL_EQ_LEN; j++) {
						cf w = S.eq_w[j], v = wv[j];
						if(j & 1) { s2.x = fmaf(w.x, v.x, fmaf(w.y, v.y, s2.x)); s2.y = fmaf(w.x, v.y, fmaf(-w.y, v.x, s2.y)); }
						else { s.x = fmaf(w.x, v.x, fmaf(w.y, v.y, s.x)); s.y = fmaf(w.x, v.y, fmaf(-w.y, v.x, s.y)); }
					}
					s.x += s2.x; s.y += s2.y;
				}
				if(S.fr_state == HF_EQ_TRAIN) {        // eqlms_cccf_step(T_seq[bitmask&1][T_idx], s)  hfdl.c:730-733
					float d = ((0x9AFu >> (HFDL_T_LEN - 1 - S.T_idx)) & 1u) ? -1.0f : 1.0f;
					if(S.bitmask & 1u) d = -d;
					bool run = true;
					if(HFDL_UNLIKELY(!S.eq_buf_full)) { if(S.eq_count < HFDL_EQ_LEN) run = false; else S.eq_buf_full = 1; }
					if(run) {
						const float inv = 1.0f / S.eq_x2_sum;
						cf t = make_float2(0.1f * (d - s.x) * inv, 0.1f * s.y * inv);      // mu * conj(d - d_hat) / sum|x|^2, mu = 0.1 (hfdl.c:496)
	#pragma unroll
						for(int j = 0; j < HFDL_EQ_LEN; j++) {
							cf uu = cmul(t, wv[j]);
							S.eq_w[j].x += uu.x;
							S.eq_w[j].y += uu.y;
						}
					}
					S.T_idx++;
				}
				if(HFDL_UNLIKELY(cap)) { if(lane == 0 && cap_n_eq < a.cap_max) a.cap_eq[cap_n_eq] = s; cap_n_eq++; }
				cf x_hat;
				unsigned bits = modem_demod(S.cur_arity, s, T, &x_hat);
				// ---- costas adjust with the modem's phase error Im(r*conj(x_hat)) (hfdl.c:738,276-281)
				float err = s.y * x_hat.x - s.x * x_hat.y;
				err = 0.5f * (fabsf(err + 1.0f) - fabsf(err - 1.0f));     // branchless_limit, hfdl.c:269-274
				S.c_phi += 0.1f * err;
				S.c_dphi += (0.047f * 0.1f * 0.1f) * err;

				symcnt++;
				if(HFDL_UNLIKELY(S.fr_state == HF_A1 && symcnt >= 13u * HFDL_SINGLE_SLOT_FRAME_LEN)) {
					symcnt = 0;
					S.c_phi = S.c_dphi = 0.f;
					{ ss_reset(S); reset_req = true; }
				}
				if(HFDL_UNLIKELY(S.s_state == HS_EMIT_BITS)) {
					bits ^= S.bitmask;
					for(int bb = 0; bb < S.cur_arity; bb++, bits >>= 1) bits_push(S.bits, bits);
				} else if(S.s_state == HS_EMIT_SYMBOLS) {
					if(S.cur_buf == 0) {
						if(S.training_n < HFDL_T_LEN) { s_train[EQS(S.training_n)] = s; S.training_n++; }
					} else {
						if(S.data_n < HFDL_DATA_SYMS_MAX) { if(lane == 0) dsym[S.data_n] = s; S.data_n++; }
					}
				}
				if(HFDL_LIKELY(S.fr_state > HF_A1)) {
					S.signal_level = __fdividef(S.signal_level * S.frame_symbol_cnt + level, S.frame_symbol_cnt + 1.0f);
					S.frame_symbol_cnt += 1.0f;
				}
				if(HFDL_LIKELY(S.symbols_wanted > 1)) { S.symbols_wanted--; continue; }

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
						{ framer_reset(S, T, E); reset_req = true; }
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
						{ framer_reset(S, T, E); reset_req = true; }
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
					unsigned tseq = 0;                       // compute_train_bit_error_cnt hfdl.c:952-966
	#pragma unroll
					for(int j = 0; j < HFDL_T_LEN; j++) {
						unsigned bit = (s_train[EQS(j)].x > 0.f) ? 0u : 1u;
						bit ^= (S.bitmask & 1u);
						tseq = (tseq << 1) | bit;
					}
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
							fr.freq_err_hz = S.freq_err_hz; fr.signal_level = S.signal_level; fr.noise_floor = S.noise_floor;
							fr.sample_cnt_a2 = S.a2_sample_cnt; fr.sample_cnt_end = cnt_base + (unsigned long long)k;
							fr.train_bits_bad = S.train_bits_bad; fr.train_bits_total = S.train_bits_total;
							a.frames[q] = fr;
						}
						S.st_frames++;
						S.slot = (S.slot + 1) % HFDL_FRAME_SLOTS;
						dsym = a.datasym + ((long long)c * HFDL_FRAME_SLOTS + S.slot) * HFDL_DATA_SYMS_MAX;
						{ framer_reset(S, T, E); reset_req = true; }
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
			}

			if(HFDL_UNLIKELY(reset_req)) {
				// symsync_crcf_reset happened while processing output seq (input sample k): outputs of the same
				// input sample already produced keep their (old-state) value, the timing warp restarts after them
				int e2 = seq;
				while(e2 + 1 < c_head && __float_as_int(s_ring[(e2 + 1) & (HFDL_RING - 1)].w) == k) e2++;
				__syncwarp();
				gen++;
				if(lane == 0) { s_reset_k = k; s_reset_seq = e2 + 1; __threadfence_block(); s_reset_gen = gen; }
				wait_seq = e2 + 1;
			}
			seq++;
			__syncwarp();
			if(lane == 0) s_tail = seq;
		}
		if(a.dbg_cycles && lane == 0) { a.dbg_cycles[c * 4 + 2] += hfdl_clock() - t_begin; a.dbg_cycles[c * 4 + 3] += t_wait; }
		for(int sidx = k_prev + 1; sidx < N; sidx++) HFDL_NF_TICK(sidx);      // input samples after the last output
#undef HFDL_NF_TICK
		S.sample_cnt = cnt_base + (unsigned long long)N;
		S.symbol_cnt = symcnt;
		__syncwarp();
		if(lane == 0) {
			for(int j = 0; j < HFDL_EQ_LEN; j++) { S.eq_win[j] = E.win[EQS(E.ep + j)]; S.eq_x2[j] = E.x2[EQS((E.ep + j) & 15)]; }
			for(int j = 0; j < HFDL_T_LEN; j++) S.training[j] = s_train[EQS(j)];
			// everything except the timing-loop fields, which the timing warp owns
			DemodState *G = &a.state[c];
			DemodState O = *G;
			S.ss_since_reset = O.ss_since_reset; S.ss_decim_counter = O.ss_decim_counter; S.ss_rate = O.ss_rate; S.ss_del = O.ss_del;
			S.ss_tau = O.ss_tau; S.ss_bf = O.ss_bf; S.ss_q = O.ss_q; S.ss_q_hat = O.ss_q_hat; S.ss_b = O.ss_b;
			S.ss_v[0] = O.ss_v[0]; S.ss_v[1] = O.ss_v[1]; S.ss_v[2] = O.ss_v[2];
			*G = S;
			if(cap) a.cap_cnt[1] = cap_n_eq;
			__threadfence_block();
			s_done = 1;
		}
	}
	__syncthreads();
	if(warp == 1 && lane == 0) {       // timing-loop state after the last input sample of the batch
		DemodState *G = &a.state[c];
		G->ss_since_reset = S.ss_since_reset; G->ss_decim_counter = S.ss_decim_counter; G->ss_rate = S.ss_rate; G->ss_del = S.ss_del;
		G->ss_tau = S.ss_tau; G->ss_bf = 0.f; G->ss_q = S.ss_q; G->ss_q_hat = S.ss_q_hat; G->ss_b = S.ss_b;
		G->ss_v[0] = S.ss_v[0]; G->ss_v[1] = S.ss_v[1]; G->ss_v[2] = S.ss_v[2];
	}
}

// ======================================================================================
// FEC: one warp per completed frame.
// ======================================================================================
struct FecArgs {
	const FrameRec *frames; const int *nframes; int max_frames;
	const cf *datasym;                 // [C][SLOTS][5040]; or a bare symbol array when frames[].channel/slot are 0
	const DemodTables *tab;
	PduRec *pdus;                      // [max_frames]
	unsigned char *soft_out;           // optional [max_frames][15120] soft bits in push order (debug/parity)
	const unsigned char *vin_direct;   // stage entry hfdl_b200_viterbi27: [max_frames][2*vin_nbits] soft bytes, skips demod/deinterleave
	int vin_nbits;
};

#define HFDL_FEC_VIN_MAX 15120
#define HFDL_FEC_SMEM (HFDL_FEC_VIN_MAX + 7566 * 8)

__device__ __forceinline__ unsigned crc16_x25_update(unsigned crc, unsigned byte) {    // crc.c:4-47, bitwise form
	crc ^= byte;
	for(int b = 0; b < 8; b++) crc = (crc & 1u) ? ((crc >> 1) ^ 0x8408u) : (crc >> 1);
	return crc;
}
__device__ __forceinline__ int fcs_check(const unsigned char *buf, unsigned hdr_len) {  // pdu.c:68-79
	unsigned crc = 0xFFFFu;
	for(unsigned i = 0; i < hdr_len; i++) crc = crc16_x25_update(crc, buf[i]);
	crc ^= 0xFFFFu;
	unsigned rx = (unsigned)buf[hdr_len] | ((unsigned)buf[hdr_len + 1] << 8);
	return rx == crc;
}
__device__ __forceinline__ int pdu_crc_good(const unsigned char *buf, unsigned len) {   // pdu.c:104, mpdu.c:56-85, spdu.c:55-64
	if(len < 1) return 0;
	if(buf[0] & 1u) {
		unsigned hdr_len;
		if(buf[0] & 0x2u) hdr_len = 6 + ((buf[0] >> 2) & 0xFu);
		else {
			unsigned ac = ((buf[0] & 0x70u) >> 4) + 1;
			hdr_len = 2;
			for(unsigned i = 0; i < ac; i++) {
				if(len < hdr_len + 2) return 0;
				hdr_len += 2 + (buf[hdr_len + 1] >> 4);
			}
		}
		if(len < hdr_len + 2) return 0;
		return fcs_check(buf, hdr_len);
	}
	if(len < 66) return 0;
	return fcs_check(buf, 64u);
}

__global__ void __launch_bounds__(32) fec_kernel(FecArgs a) {
	HFDL_DYN_SMEM(unsigned char, sm);
	const int q = blockIdx.x;
	int nfr = *a.nframes;
	if(nfr > a.max_frames) nfr = a.max_frames;
	if(q >= nfr) return;
	const int lane = threadIdx.x;
	const DemodTables &T = *a.tab;
	const FrameRec fr = a.frames[q];
	const int M1 = fr.M1;
	const int arity = T.mode_arity[M1], code_rate = T.mode_code_rate[M1], shift = T.mode_col_shift[M1];
	const int nsym = T.mode_segments[M1] * 30;
	const int nenc = nsym * arity;
	const int ncol = nenc / 40;
	const cf *sym = a.datasym + ((long long)fr.channel * HFDL_FRAME_SLOTS + fr.slot) * HFDL_DATA_SYMS_MAX;
	unsigned char *vin = sm;                                  // [<=15120] Viterbi input
	unsigned char *table = sm + HFDL_FEC_VIN_MAX;             // [40][ncol], later overlaid by the decisions
	uint2 *dec = reinterpret_cast<uint2 *>(sm + HFDL_FEC_VIN_MAX);   // [nbits+6] (.x even states, .y odd states)
	const float pol = (fr.bitmask & 1u) ? -1.0f : 1.0f;
	int vin_len = (code_rate == 4) ? nenc / 2 : nenc;

	if(a.vin_direct) {
		vin_len = 2 * a.vin_nbits;
		for(int i = lane; i < vin_len; i += 32) vin[i] = a.vin_direct[(long long)q * vin_len + i];
	} else {
	// ---- descramble + soft demod + deinterleaver push (hfdl.c:1008-1018,387-399): closed form of the push
	// sequence: soft bit i lands in row i%40, column (i/40 - shift*i) mod ncol
	for(int i = lane; i < nsym; i += 32) {
		float flip = (T.scr[i % 120] ? -1.0f : 1.0f) * pol;
		cf x = make_float2(sym[i].x * flip, sym[i].y * flip);
		unsigned char soft[3];
		if(arity == 1) {
			float LLR = -2.0f * x.x * 4.0f;
			int sb = (int)__fadd_rn(__fmul_rn(LLR, 16.0f), 127.0f);
			sb = sb > 255 ? 255 : (sb < 0 ? 0 : sb);
			soft[0] = (unsigned char)sb;
		} else {
			cf xh;
			unsigned s = modem_demod(arity, x, T, &xh);
			if(arity == 2) {
				soft[0] = (s & 2u) ? 255 : 0; soft[1] = (s & 1u) ? 255 : 0;
			} else {
				// liquid modem_demodulate_soft_table, p = 2 nearest neighbours (the adjacent PSK8 points)
				const float gamma = 1.2f * 8.0f;
				float dmin0[3], dmin1[3];
				for(int k = 0; k < 3; k++) dmin0[k] = dmin1[k] = 4.0f;
				float ex = __fsub_rn(x.x, xh.x), ey = __fsub_rn(x.y, xh.y);
				float d = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
				for(int k = 0; k < 3; k++) { if((s >> (2 - k)) & 1u) dmin1[k] = d; else dmin0[k] = d; }
				unsigned g = s; { unsigned mm = g >> 1; while(mm) { g ^= mm; mm >>= 1; } }   // gray decode
				for(int n = 0; n < 2; n++) {
					unsigned gg = (g + (n == 0 ? 1u : 7u)) & 7u;
					unsigned nb = gg ^ (gg >> 1);
					cf p = T.psk[3][nb];
					ex = __fsub_rn(x.x, p.x); ey = __fsub_rn(x.y, p.y);
					d = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
					for(int k = 0; k < 3; k++) {
						if((nb >> (2 - k)) & 1u) { if(d < dmin1[k]) dmin1[k] = d; }
						else { if(d < dmin0[k]) dmin0[k] = d; }
					}
				}
				for(int k = 0; k < 3; k++) {
					int sb = (int)__fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(dmin0[k], dmin1[k]), gamma), 16.0f), 127.0f);
					sb = sb > 255 ? 255 : (sb < 0 ? 0 : sb);
					soft[k] = (unsigned char)sb;
				}
			}
		}
		for(int j = 0; j < arity; j++) {
			int p = i * arity + j;
			int row = p % 40;
			int col = (int)(((long long)(p / 40) - (long long)shift * p) % ncol);
			if(col < 0) col += ncol;
			table[row * ncol + col] = soft[j];
			if(a.soft_out) a.soft_out[(long long)q * HFDL_FEC_VIN_MAX + p] = soft[j];
		}
	}
	__syncwarp();
	// ---- deinterleaver pop (hfdl.c:401-409): pop j reads row 9j%40, column j/40; r=1/4 averages chip pairs
	for(int i = lane; i < vin_len; i += 32) {
		if(code_rate == 4) {
			int j0 = 2 * i, j1 = 2 * i + 1;
			unsigned A = table[((9 * j0) % 40) * ncol + j0 / 40], B = table[((9 * j1) % 40) * ncol + j1 / 40];
			vin[i] = (unsigned char)((A & B) + ((A ^ B) >> 1));
		} else {
			vin[i] = table[((9 * i) % 40) * ncol + i / 40];
		}
	}
	}   // !vin_direct
	__syncwarp();
	// ---- Viterbi K=7 (viterbi27_port.c:147-221): lane i owns butterfly i -> new states 2i, 2i+1
	const int nbits = vin_len / 2;
	const unsigned bt0 = (__popc((2u * lane) & 0x6du) & 1) ? 255u : 0u;    // set_viterbi27_polynomial, :81-89
	const unsigned bt1 = (__popc((2u * lane) & 0x4fu) & 1) ? 255u : 0u;
	unsigned m_even = 63u, m_odd = 63u;                  // metrics of states 2*lane, 2*lane+1 (init_viterbi27 :65-79)
	if(lane == 0) m_even = 0u;
	for(int t = 0; t < nbits; t++) {
		unsigned s0 = vin[2 * t], s1 = vin[2 * t + 1];
		// old[i] and old[i+32] for butterfly i=lane: state i is held by lane i>>1 (even/odd slot i&1)
		unsigned src_lo = (unsigned)lane >> 1, src_hi = 16u + ((unsigned)lane >> 1);
		unsigned lo_e = __shfl_sync(0xffffffffu, m_even, src_lo), lo_o = __shfl_sync(0xffffffffu, m_odd, src_lo);
		unsigned hi_e = __shfl_sync(0xffffffffu, m_even, src_hi), hi_o = __shfl_sync(0xffffffffu, m_odd, src_hi);
		unsigned old_i = (lane & 1) ? lo_o : lo_e;
		unsigned old_i32 = (lane & 1) ? hi_o : hi_e;
		unsigned metric = (bt0 ^ s0) + (bt1 ^ s1);
		unsigned a0 = old_i + metric, b0 = old_i32 + (510u - metric);
		unsigned d0 = ((int)(a0 - b0) > 0) ? 1u : 0u;
		m_even = d0 ? b0 : a0;
		unsigned a1 = old_i + (510u - metric), b1 = old_i32 + metric;
		unsigned d1 = ((int)(a1 - b1) > 0) ? 1u : 0u;
		m_odd = d1 ? b1 : a1;
		unsigned de = __ballot_sync(0xffffffffu, d0), dod = __ballot_sync(0xffffffffu, d1);
		if(lane == 0) dec[t] = make_uint2(de, dod);
	}
	__syncwarp();
	// ---- chainback from state 0, reading 6 steps ahead; the 6 steps past the end were never written
	// by the reference (calloc'd zero) -> zeros here (viterbi27_port.c:105-134)
	PduRec *out = &a.pdus[q];
	const int out_octets = nbits / 8 + ((nbits % 8) ? 1 : 0);
	if(lane == 0) {
		unsigned endstate = 0;
		for(int i = 0; i < out_octets; i++) out->octets[i] = 0;
		for(int n = nbits - 1; n >= 0; n--) {
			unsigned st = endstate >> 2;
			unsigned k = 0;
			if(n + 6 < nbits) {
				uint2 d = dec[n + 6];
				k = (((st & 1u) ? d.y : d.x) >> (st >> 1)) & 1u;
			}
			endstate = ((endstate >> 1) | (k << 7)) & 0xFFu;
			out->octets[n >> 3] = (unsigned char)endstate;
		}
		if(!a.vin_direct) for(int i = 0; i < out_octets; i++)  // REVERSE_BYTE (util.h:109, hfdl.c:1051-1053)
			out->octets[i] = (unsigned char)(__brev((unsigned)out->octets[i]) >> 24);
		out->channel = fr.channel; out->M1 = M1; out->len = out_octets;
		out->freq_err_hz = fr.freq_err_hz; out->signal_level = fr.signal_level; out->noise_floor = fr.noise_floor;
		out->sample_cnt_a2 = fr.sample_cnt_a2; out->sample_cnt_end = fr.sample_cnt_end;
		out->train_bits_bad = fr.train_bits_bad; out->train_bits_total = fr.train_bits_total;
		out->crc_good = pdu_crc_good(out->octets, (unsigned)out_octets);
	}
}
