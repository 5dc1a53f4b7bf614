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
#define HFDL_FRAME_SLOTS_MIN 4              // data-symbol buffers per channel: run-time value, >= 2 x (frames a batch can end per channel)
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

struct AgcState { float g, y2; };      // g: POWER gain g^2 of agc_crcf (see agc_kernel), y2: its energy estimate

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
	int st_a1, st_a2, st_m1, st_frames, st_m1_fail;      // st_m1_fail: statsd demod.preamble.errors.M1_not_found (hfdl.c:840)
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
	// front parser (pdu.c:104,123, mpdu.c:56-134, spdu.c:55-64, lpdu.c:124-150): what the statsd counters count
	int frame_status;          // 0 good, 1 bad_fcs, 2 too_short
	int direction;             // 1 air2gnd (downlink MPDU), 0 gnd2air (uplink MPDU / SPDU); valid when frame_status == 0
	int lpdus_processed, lpdus_good, lpdus_bad_fcs, lpdus_too_short;
	unsigned long long lpdu_good_mask;     // bit j: the j-th LPDU handed to lpdu_parse had a good FCS
	unsigned char octets[HFDL_MAX_PDU + 3];
};

// ======================================================================================
// K6: AGC (liquid agc_crcf_execute, hfdl.c:686).  grid = C blocks of one warp.
//   y = x*g;  y2' = (1-a)*y2' + a*|y|^2;  if(y2' > 1e-6) g *= exp(-0.5*a*ln y2');  g = min(g, 1e6)   (a = 0.01)
// The recurrence is strictly sequential, so only its dependent chain stays on lane 0, restated in the power gain
// G = g^2 so that the sample itself is off the chain:
//   y2' = (a*|x|^2)*G + (1-a)*y2';   G *= 2^(-a*log2 y2');   G = min(G, 1e12)
// (FFMA -> MUFU.LG2 -> FMUL -> MUFU.EX2 -> FMUL -> FMNMX per sample).  |x|^2 before and sqrt(G), x*g, 1/g after the
// chain are computed by all 32 lanes per chunk.
// ======================================================================================
struct AgcArgs {
	const cf *rs; long long rs_stride; int n_samples;
	AgcState *state;
	cf *agc_out; long long agc_stride;       // [C][HFDL_AGC_HIST + n]
	float *lvl; long long lvl_stride;        // [C][n]  1/g after the update (agc_crcf_get_signal_level)
};

#define HFDL_AGC_CH 128
__global__ void __launch_bounds__(32) agc_kernel(AgcArgs a) {
	__shared__ cf s_in[2][HFDL_AGC_CH];
	__shared__ float s_px[HFDL_AGC_CH];
	__shared__ float s_G[HFDL_AGC_CH + 1];       // s_G[i + 1] = G after sample i; s_G[0] = G before the chunk
	const int c = blockIdx.x, lane = threadIdx.x;
	const cf *x = a.rs + (long long)c * a.rs_stride;
	cf *out = a.agc_out + (long long)c * a.agc_stride + HFDL_AGC_HIST;
	float *lvl = a.lvl + (long long)c * a.lvl_stride;
	float G = a.state[c].g, y2 = a.state[c].y2;  // AgcState.g holds the power gain g^2
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
		const cf *in = s_in[j & 1];
		for(int i = lane; i < cnt; i += 32) { const cf xv = in[i]; s_px[i] = alpha * (xv.x * xv.x + xv.y * xv.y); }
		__syncwarp();
		if(lane == 0) {
			s_G[0] = G;
#pragma unroll 4
			for(int i = 0; i < cnt; i++) {
				// (1.0 - alpha)*y2 with the product term kept exact: y2 - alpha*y2
				y2 = fmaf(s_px[i], G, fmaf(-alpha, y2, y2));
				const float e = hfdl_exp2_fast(-alpha * hfdl_log2_fast(y2));
				G = fminf((y2 > 1e-6f) ? G * e : G, 1e12f);
				s_G[i + 1] = G;
			}
		}
		__syncwarp();
		for(int i = lane; i < cnt; i += 32) {
			const float gb = sqrtf(s_G[i]);              // gain applied to sample i (before its update)
			const cf xv = in[i];
			out[n0 + i] = make_float2(xv.x * gb, xv.y * gb);
			lvl[n0 + i] = 1.0f / sqrtf(s_G[i + 1]);
		}
		__syncwarp();
		{
			const int nn0 = (j + 2) * HFDL_AGC_CH;
			for(int i = lane; i < HFDL_AGC_CH; i += 32) { int n = nn0 + i; if(n < N) hfdl_cp_async8(&s_in[j & 1][i], &x[n]); }
			hfdl_cp_async_commit();
		}
	}
	hfdl_cp_async_wait<0>();
	if(lane == 0) { a.state[c].g = G; a.state[c].y2 = y2; }
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

// carries the tails of the per-channel work streams of the previous batch (set "src", n_prev samples) in front of the
// arrays of the next batch (set "dst").  With n_prev = 0 (first batch) the zeroed history of the other set is copied.
__global__ void demod_carry(const cf *agc_src, cf *agc_dst, long long agc_stride, const cf *mfo_src, cf *mfo_dst, long long mfo_stride, long long n_prev) {
	int c = blockIdx.x, t = threadIdx.x;
	if(t < HFDL_AGC_HIST) agc_dst[(long long)c * agc_stride + t] = agc_src[(long long)c * agc_stride + n_prev + t];
	if(t < HFDL_MFO_HIST) mfo_dst[(long long)c * mfo_stride + t] = mfo_src[(long long)c * mfo_stride + n_prev + t];
}

#include "loop_kernel.cuh"     // K8b-K11: timing loop, Costas, equaliser, slicer, framer

// ======================================================================================
// FEC: one warp per completed frame.
// ======================================================================================
struct FecArgs {
	const FrameRec *frames; const int *nframes; int max_frames;
	const cf *datasym; int nslots;     // [C][nslots][5040]; or a bare symbol array when frames[].channel/slot are 0
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
// Front parser of one PDU, run by a whole warp: lane 0 walks the MPDU header exactly like mpdu_parse /
// parse_lpdu_list (mpdu.c:56-134,136-159) and lists the LPDUs in shared memory; the lanes then check the LPDU frame
// check sequences in parallel (lpdu_parse, lpdu.c:124-150).  SPDUs carry no LPDUs (spdu.c:55-64).
struct LpduDesc { unsigned short off, len; };
__device__ __forceinline__ void pdu_front_parse(const unsigned char *buf, unsigned len, PduRec *out, LpduDesc *list /* smem [64] */, int *s_n, int lane) {
	int status = 0, dir = 0;
	if(lane == 0) {
		int n = 0;
		if(len < 1) status = 2;
		else if(buf[0] & 1u) {                                   // IS_MPDU, pdu.c:104
			unsigned hdr_len, lpdu_cnt = 0, ac = 0;
			if(buf[0] & 0x2u) { dir = 1; lpdu_cnt = (buf[0] >> 2) & 0xFu; hdr_len = 6 + lpdu_cnt; }
			else {
				ac = ((buf[0] & 0x70u) >> 4) + 1;
				hdr_len = 2;
				for(unsigned i = 0; i < ac && status == 0; i++) {
					if(len < hdr_len + 2) { status = 2; break; }
					lpdu_cnt = buf[hdr_len + 1] >> 4;
					hdr_len += 2 + lpdu_cnt;
				}
			}
			if(status == 0 && len < hdr_len + 2) status = 2;
			if(status == 0 && !fcs_check(buf, hdr_len)) status = 1;
			if(status == 0) {
				unsigned dp = hdr_len + 2;                          // first data octet of the first LPDU
				if(dir == 1) {
					unsigned hp = 6;
					for(unsigned j = 0; j < lpdu_cnt; j++) {
						unsigned ll = (unsigned)buf[hp] + 1u;
						if(dp + ll > len) break;                    // truncated: parse_lpdu_list returns -1
						if(n < 64) { list[n].off = (unsigned short)dp; list[n].len = (unsigned short)ll; n++; }
						dp += ll; hp++;
					}
				} else {
					unsigned hp = 2;
					bool stop = false;
					for(unsigned i = 0; i < ac && !stop; i++) {
						hp++;                                       // aircraft id
						unsigned cnt = (buf[hp++] >> 4) & 0xFu;
						for(unsigned j = 0; j < cnt; j++) {
							unsigned ll = (unsigned)buf[hp + j] + 1u;
							if(dp + ll > len) { stop = true; break; }
							if(n < 64) { list[n].off = (unsigned short)dp; list[n].len = (unsigned short)ll; n++; }
							dp += ll;
						}
						hp += cnt;
					}
				}
			}
		} else {
			if(len < 66) status = 2;                              // spdu.c:12,55-59
			else if(!fcs_check(buf, 64u)) status = 1;
		}
		*s_n = n;
	}
	__syncwarp();
	const int n = *s_n;
	unsigned long long good = 0ull; int ngood = 0, nbad = 0, nshort = 0;
	for(int j0 = 0; j0 < n; j0 += 32) {
		const int j = j0 + lane;
		int st = -1;                                                // -1 none, 0 good, 1 bad fcs, 2 too short
		if(j < n) {
			const unsigned ll = list[j].len;
			if(ll < 3) st = 2;                                      // lpdu.c:137-142
			else st = fcs_check(buf + list[j].off, ll - 2) ? 0 : 1;
		}
		const unsigned mg = __ballot_sync(0xffffffffu, st == 0), mb = __ballot_sync(0xffffffffu, st == 1), ms = __ballot_sync(0xffffffffu, st == 2);
		good |= (unsigned long long)mg << j0;
		ngood += __popc(mg); nbad += __popc(mb); nshort += __popc(ms);
	}
	if(lane == 0) {
		out->frame_status = status; out->direction = dir;
		out->lpdus_processed = n; out->lpdus_good = ngood; out->lpdus_bad_fcs = nbad; out->lpdus_too_short = nshort;
		out->lpdu_good_mask = good;
		out->crc_good = status == 0;
	}
}

// stage entry hfdl_b200_pdu_front_parse: one warp per PDU record (octets + len already in place)
__global__ void __launch_bounds__(32) front_kernel(PduRec *pdus, int n) {
	__shared__ LpduDesc list[64];
	__shared__ int s_n;
	if((int)blockIdx.x < n) pdu_front_parse(pdus[blockIdx.x].octets, (unsigned)pdus[blockIdx.x].len, &pdus[blockIdx.x], list, &s_n, threadIdx.x);
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
	const cf *sym = a.datasym + ((long long)fr.channel * a.nslots + fr.slot) * HFDL_DATA_SYMS_MAX;
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
			unsigned s = modem_demod(arity, x, T.psk, &xh);
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
			const unsigned p = (unsigned)(i * arity + j);
			const unsigned row = p % 40u;
			// column (p/40 - shift*p) mod ncol in 32-bit unsigned arithmetic (shift*p < 2^19)
			const unsigned col = ((p / 40u) % (unsigned)ncol + (unsigned)ncol - ((unsigned)shift * p) % (unsigned)ncol) % (unsigned)ncol;
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
		// Sequential by nature (each step's state selects the next decision bit), so the dependent chain is kept to
		// shift / mask / or: the decision words of eight steps are loaded ahead of the chain, and an octet is stored once,
		// when its last bit (n % 8 == 0) has been shifted in -- the value the reference's per-step store leaves behind.
		unsigned endstate = 0;
		for(int i = 0; i < out_octets; i++) out->octets[i] = 0;
		int n = nbits - 1;
		for(; (n & 7) != 7 && n >= 0; n--) {                 // ragged top octet (nbits not a multiple of 8)
			const unsigned st = endstate >> 2;
			unsigned k = 0;
			if(n + 6 < nbits) { const uint2 d = dec[n + 6]; k = (((st & 1u) ? d.y : d.x) >> (st >> 1)) & 1u; }
			endstate = ((endstate >> 1) | (k << 7)) & 0xFFu;
			if((n & 7) == 0) out->octets[n >> 3] = (unsigned char)endstate;
		}
		for(; n >= 7; n -= 8) {
			uint2 d[8];
#pragma unroll
			for(int q = 0; q < 8; q++) d[q] = (n - q + 6 < nbits) ? dec[n - q + 6] : make_uint2(0u, 0u);
#pragma unroll
			for(int q = 0; q < 8; q++) {
				const unsigned st = endstate >> 2;
				const unsigned k = (((st & 1u) ? d[q].y : d[q].x) >> (st >> 1)) & 1u;
				endstate = ((endstate >> 1) | (k << 7)) & 0xFFu;
			}
			out->octets[(n - 7) >> 3] = (unsigned char)endstate;
		}
		if(!a.vin_direct) for(int i = 0; i < out_octets; i++)  // REVERSE_BYTE (util.h:109, hfdl.c:1051-1053)
			out->octets[i] = (unsigned char)(__brev((unsigned)out->octets[i]) >> 24);
		out->channel = fr.channel; out->M1 = M1; out->len = out_octets;
		out->freq_err_hz = fr.freq_err_hz; out->signal_level = fr.signal_level; out->noise_floor = fr.noise_floor;
		out->sample_cnt_a2 = fr.sample_cnt_a2; out->sample_cnt_end = fr.sample_cnt_end;
		out->train_bits_bad = fr.train_bits_bad; out->train_bits_total = fr.train_bits_total;
	}
	__syncwarp();
	__threadfence_block();
	{
		// the decisions are dead: the front parser's LPDU list overlays them
		LpduDesc *list = reinterpret_cast<LpduDesc *>(sm + HFDL_FEC_VIN_MAX);
		int *s_n = reinterpret_cast<int *>(sm + HFDL_FEC_VIN_MAX + 64 * sizeof(LpduDesc));
		pdu_front_parse(out->octets, (unsigned)out_octets, out, list, s_n, lane);
	}
}
