// dumphfdl_b200/csrc/frontend.cu -- host driver + C ABI (include/hfdl_b200.h) of the B200 front-end.
// Owns the CUDA streams, the HBM layout and the batch schedule (a four-stream software pipeline across batches,
// see enqueue_batch); all arithmetic on the sample path runs
// in the kernels of ddc_kernels.cuh / demod_kernels.cuh.  There is no CPU implementation of the path
// in this library: without a CUDA device hfdl_b200_create() fails.
#include <vector>
#include <deque>
#include <string>
#include <thread>
#include <mutex>
#include <algorithm>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"
#include "ddc_kernels.cuh"
#include "demod_kernels.cuh"
#include "design.hpp"
#include "../../include/hfdl_b200.h"

#define CK(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) { \
	fprintf(stderr, "hfdl_b200: CUDA error '%s' at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return -1; } } while(0)

namespace {

enum { LK_PACK4 = 0, LK_ROLE2 = 1, LK_PAIR2 = 2 };
enum { KC_FFT1 = 0, KC_FFT2, KC_FFT3, KC_CHAN, KC_RESAMP, KC_AGC, KC_BANK, KC_LOOP, KC_FEC, KC_PACK, KC_COUNT };
const char *kc_names[KC_COUNT] = { "fft_pass1", "fft_pass2", "fft_pass3", "chan_extract", "resamp", "agc", "bank", "loop", "fec", "slice_pack" };

struct ProfRec { int cls; cudaEvent_t e0, e1; };
#ifndef HFDL_NSETS
#define HFDL_NSETS 4        // batches in flight: sets of every buffer a later pipeline stage reads
#endif

FftPlan make_plan(int N) {
	FftPlan p;
	memset(&p, 0, sizeof(p));
	p.N = N; p.lgN = hfdl_ilog2(N);
	int lg = p.lgN;
	if(lg <= 12) { p.P = 1; p.lgL[0] = lg; }
	else if(lg <= 18) { p.P = 2; p.lgL[0] = lg / 2; p.lgL[1] = lg - p.lgL[0]; }
	else { p.P = 3; p.lgL[0] = lg / 3; p.lgL[1] = (lg - p.lgL[0]) / 2; p.lgL[2] = lg - p.lgL[0] - p.lgL[1]; }
	if(p.P == 3) {
		// experiments: HFDL_B200_FFT_PLAN="a,b" = log2 of the first two pass lengths of a three-pass plan
		const char *e = getenv("HFDL_B200_FFT_PLAN");
		int a = 0, b = 0;
		if(e && sscanf(e, "%d,%d", &a, &b) == 2 && a >= 6 && a <= 9 && b >= 6 && b <= 9 && lg - a - b >= 6 && lg - a - b <= 9) { p.lgL[0] = a; p.lgL[1] = b; p.lgL[2] = lg - a - b; }
	}
	// natural-order, out-of-place last pass (fft_last_pass_nat): needs a register-resident last pass (L = 32*B,
	// B = 2..16) and at least one tile of 256/B consecutive k1 rows
	if(p.P >= 2 && !getenv("HFDL_B200_SMEM_FFT")) {
		const int lgb = p.lgL[p.P - 1] - 5;
		if(lgb >= 1 && lgb <= 4 && p.lgL[0] >= 8 - lgb) p.natural = 1;
	}
	return p;
}

int tile_for(int lgL) {          // columns (or rows) per CTA: keep the tile <= 64 KiB
#ifndef HFDL_FFT_TILE_CAP
#define HFDL_FFT_TILE_CAP 16
#endif
	int t = (512 * HFDL_FFT_TILE_CAP) >> lgL;
	if(t > HFDL_FFT_TILE_CAP) t = HFDL_FFT_TILE_CAP;
	if(t < 1) t = 1;
	return t;
}

struct FftEngine {
	cf *d_tw = nullptr;
	bool attrs_set = false;
	int init() {
		std::vector<cf> tw(HFDL_TWN);
		for(int i = 0; i < HFDL_TWN; i++) {
			double a = -2.0 * M_PI * (double)i / (double)HFDL_TWN;
			tw[i] = make_float2((float)cos(a), (float)sin(a));
		}
		CK(cudaMalloc((void **)&d_tw, sizeof(cf) * HFDL_TWN));
		CK(cudaMemcpy(d_tw, tw.data(), sizeof(cf) * HFDL_TWN, cudaMemcpyHostToDevice));
		CK(cudaFuncSetAttribute(fft_col_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
		CK(cudaFuncSetAttribute(fft_col_pass_reg<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
		CK(cudaFuncSetAttribute(fft_col_pass_reg<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
		CK(cudaFuncSetAttribute(fft_col_pass_reg<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
		CK(cudaFuncSetAttribute(fft_col_pass_reg<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
		CK(cudaFuncSetAttribute(fft_row_pass_reg<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
		CK(cudaFuncSetAttribute(fft_row_pass_reg<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
		CK(cudaFuncSetAttribute(fft_row_pass_reg<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
		CK(cudaFuncSetAttribute(fft_row_pass_reg<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
		CK(cudaFuncSetAttribute(fft_last_pass_nat<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
		CK(cudaFuncSetAttribute(fft_last_pass_nat<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
		CK(cudaFuncSetAttribute(fft_last_pass_nat<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
		CK(cudaFuncSetAttribute(fft_last_pass_nat<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
		CK(cudaFuncSetAttribute(fft_row_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
		CK(cudaFuncSetAttribute(chan_extract, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
		CK(cudaFuncSetAttribute(fec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HFDL_FEC_SMEM));
		CK(cudaFuncSetAttribute(loop_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
		CK(cudaFuncSetAttribute(loop_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
		CK(cudaFuncSetAttribute(loop_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
		return 0;
	}
	void destroy() { if(d_tw) cudaFree(d_tw); d_tw = nullptr; }
};

}  // namespace

struct hfdl_b200_frontend {
	hfdl_b200_config_t cfg;
	std::vector<int32_t> freqs;
	hfdl_design::Geometry g;
	std::vector<hfdl_design::ChannelGeom> chg;
	FftPlan plan;
	int C = 0, Bmax = 0, sfmt = 0, bps = 0, out_per_block = 0;
	float resamp_rate = 0;
	// streams: front (H2D, FFT, channeliser, resampler) | agc + bank | loop | fec + D2H.  One launch per stage per batch;
	// the stages of consecutive batches overlap (front/agc/bank of batch i+1 beside loop of batch i beside fec of i-1).
	cudaStream_t stream = nullptr, stream2 = nullptr, st_loop = nullptr, st_fec = nullptr, st_stats = nullptr, st_h2d = nullptr;
	cudaEvent_t ev_front[HFDL_NSETS] = { nullptr }, ev_bank[HFDL_NSETS] = { nullptr }, ev_loop[HFDL_NSETS] = { nullptr }, ev_fec_done[HFDL_NSETS] = { nullptr };
	cudaEvent_t ev_h2d = nullptr;
	struct Flight { bool busy = false; } flight[HFDL_NSETS];
	std::recursive_mutex mtx;       // every public entry point: the frontend may be driven and queried from different threads
	bool peer_enabled = false, h2d_pending = false;
	// loop_kernel: CTA layout (LK_PACK4 / LK_ROLE2 / LK_PAIR2, loop_kernel.cuh) and extra dynamic shared memory requested
	// on top of the bank rings (padding keeps other stages' CTAs off its SMs)
	int loop_layout = LK_PACK4; size_t loop_smem_pad_to = 0;
	bool loop_auto = true;
	bool failed = false;            // a CUDA call failed mid-pipeline: every later call returns -1
	int Bsub = 1;                   // blocks per FFT sub-batch (intermediate spectra stay in L2)
	long long n_out_prev = 0;       // resampled samples of the previous batch (carry source)
	long long batch_seq = 0;        // batches enqueued so far; set p = batch_seq % HFDL_NSETS
	int nslots = HFDL_FRAME_SLOTS_MIN;
	FftEngine fft;
	// device memory
	cf *d_work = nullptr, *d_spec = nullptr;       // FFT workspace (passes in place) / natural-order spectra (plan.natural)
	unsigned *d_mask = nullptr; double spec_fill = 1.0;   // spectrum granules the channels read (bit set) / their share of the band
	// sharded spectrum (hfdl_b200_set_exchange): the channel list of the whole job; channel j belongs to rank j % xr_ranks
	int xr_ranks = 0, xr_nall = 0; int *d_all_offsetbin = nullptr;
	cudaEvent_t ev_ext = nullptr;      // recorded on the caller's stream: the slices of the next batch have arrived
	void *d_ring = nullptr; long long ring_len = 0;
	cf *d_tapslice = nullptr; int *d_offsetbin = nullptr; float *d_dsa_rate = nullptr;
	cf *d_bb = nullptr; long long bb_stride = 0;
	cf *d_rs[HFDL_NSETS] = { nullptr }; long long rs_stride = 0; float *d_rs_h = nullptr;
	DemodTables *d_tab = nullptr; DemodState *d_state = nullptr; AgcState *d_agc_state = nullptr; cf *d_datasym = nullptr;
	cf *d_agc[HFDL_NSETS] = { nullptr }, *d_mfo[HFDL_NSETS] = { nullptr }, *d_bank[HFDL_NSETS] = { nullptr }; float *d_lvl[HFDL_NSETS] = { nullptr };
	long long agc_stride = 0, mfo_stride = 0;      // work arrays of the demodulator stages, one set per batch parity
	long long cap_n = 0;            // AGC/MF checkpoint samples captured so far
	FrameRec *d_frames[HFDL_NSETS] = { nullptr }; int *d_nframes[HFDL_NSETS] = { nullptr }; PduRec *d_pdus[HFDL_NSETS] = { nullptr }; int max_frames = 0;
	cf *d_cap_agc = nullptr, *d_cap_mf = nullptr, *d_cap_eq = nullptr; int *d_cap_cnt = nullptr;
	cf *d_tmp = nullptr; long long tmp_len = 0;
	long long *d_dbg = nullptr;
	// host state
	PduRec *h_pdus[HFDL_NSETS] = { nullptr }; int *h_nframes[HFDL_NSETS] = { nullptr };
	long long fed = 0;              // samples pushed so far (host-fed path)
	long long blocks_done = 0;      // overlap-save blocks processed
	unsigned long long rs_phi0 = 0; unsigned rs_step = 0;
	int last_nblocks = 0, last_nout = 0, last_sub_blocks = 0, debug_mode = 0;
	std::deque<hfdl_b200_pdu_t> pduq;
	struct ChanTally { long long processed = 0, good = 0, bad_fcs = 0, too_short = 0, air2gnd = 0, gnd2air = 0, lpdus_processed = 0, lpdus_good = 0, lpdus_bad = 0, lpdus_short = 0; };
	std::vector<ChanTally> tally;
	long long launches = 0;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	bool profiling = false;
	std::vector<ProfRec> prof;
	float prof_ms[KC_COUNT] = { 0 }; int prof_n[KC_COUNT] = { 0 };
};

namespace {

// Entry-point guard: serialises the public calls on one frontend (the block thread pushes while a stats thread reads)
// and makes the frontend's device current for the calling host thread (the current device is per thread), restoring
// the caller's device afterwards.
struct ApiGuard {
	std::unique_lock<std::recursive_mutex> lk;
	int prev = -1;
	bool ok = true;
	explicit ApiGuard(hfdl_b200_frontend *fe) : lk(fe->mtx) {
		if(cudaGetDevice(&prev) != cudaSuccess) prev = -1;
		if(prev != fe->cfg.device) { if(cudaSetDevice(fe->cfg.device) != cudaSuccess) ok = false; }
		else prev = -1;
	}
	~ApiGuard() { if(prev >= 0) cudaSetDevice(prev); }
};
#define HFDL_API(fe, errval) if(!(fe)) return errval; ApiGuard guard_(fe); if(!guard_.ok || (fe)->failed) return errval

inline void prof_begin(hfdl_b200_frontend *fe, int cls, ProfRec &r) {
	r.cls = -1;
	if(!fe || !fe->profiling) return;
	r.cls = cls;
	cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
	cudaEventRecord(r.e0, fe->stream);
}
inline void prof_end(hfdl_b200_frontend *fe, ProfRec &r) {
	if(r.cls < 0) return;
	cudaEventRecord(r.e1, fe->stream);
	fe->prof.push_back(r);
}
inline void prof_begin2(hfdl_b200_frontend *fe, int cls, ProfRec &r, cudaStream_t st) {
	r.cls = -1;
	if(!fe || !fe->profiling) return;
	r.cls = cls;
	cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
	cudaEventRecord(r.e0, st);
}
inline void prof_end2(hfdl_b200_frontend *fe, ProfRec &r, cudaStream_t st) {
	if(r.cls < 0) return;
	cudaEventRecord(r.e1, st);
	fe->prof.push_back(r);
}

// forward FFT of nb windows described by src into work (scrambled layout)
// forward FFT of nb windows: passes in place in `work`; when pl.natural the last pass writes the natural-order
// spectrum to `spec` (else the digit-scrambled spectrum stays in `work`)
int run_fft(hfdl_b200_frontend *fe, const FftEngine &eng, const FftPlan &pl, const RawSource &src, cf *work, cf *spec, int nb, cudaStream_t st, const unsigned *mask = nullptr) {
	int inner = pl.N;
	int outer = 1;
	for(int p = 0; p < pl.P; p++) {
		int lgL = pl.lgL[p], L = 1 << lgL;
		inner /= L;
		ProfRec pr;
		prof_begin2(fe, KC_FFT1 + p, pr, st);
		if(p < pl.P - 1) {
			ColPassArgs a;
			a.src = src; a.work = work; a.tw = eng.d_tw; a.N = pl.N; a.lgL = lgL; a.inner = inner; a.lgInner = hfdl_ilog2(inner);
			a.T = tile_for(lgL); if(a.T > inner) a.T = inner;
			a.lgT = hfdl_ilog2(a.T);
			a.first = (p == 0);
			const int lgb = lgL - 5;              // register-resident pass: L = 32 * B, tile of 8192 / L columns
			if(!getenv("HFDL_B200_SMEM_FFT") && lgb >= 1 && lgb <= 4 && (256 >> lgb) <= inner) {
				dim3 grid((unsigned)(outer * (inner / (256 >> lgb))), (unsigned)nb);
				const size_t smem = sizeof(cf) * 8192;
				switch(lgb) {
				case 1: HFDL_LAUNCH(fft_col_pass_reg<1>, grid, dim3(256), smem, st, a); break;
				case 2: HFDL_LAUNCH(fft_col_pass_reg<2>, grid, dim3(256), smem, st, a); break;
				case 3: HFDL_LAUNCH(fft_col_pass_reg<3>, grid, dim3(256), smem, st, a); break;
				default: HFDL_LAUNCH(fft_col_pass_reg<4>, grid, dim3(256), smem, st, a); break;
				}
			} else {
				dim3 grid((unsigned)(outer * (inner / a.T)), (unsigned)nb);
				size_t smem = sizeof(cf) * (size_t)L * a.T;
				HFDL_LAUNCH(fft_col_pass, grid, dim3(HFDL_FFT_THREADS), smem, st, a);
			}
		} else {
			RowPassArgs a;
			a.src = src; a.work = work; a.tw = eng.d_tw; a.N = pl.N; a.lgL = lgL; a.R = tile_for(lgL);
			int rows = pl.N / L;
			if(a.R > rows) a.R = rows;
			a.lgR = hfdl_ilog2(a.R);
			a.first = (p == 0);
			a.out = nullptr; a.L1 = 0; a.mid = 0; a.mask = mask;
			const int lgb = lgL - 5;
			if(pl.natural) {
				a.out = spec; a.L1 = 1 << pl.lgL[0]; a.mid = pl.P == 3 ? (1 << pl.lgL[1]) : 1;
				dim3 grid((unsigned)(rows / (256 >> lgb)), (unsigned)nb);
				const size_t smem = sizeof(cf) * (size_t)(256 >> lgb) * (size_t)(33 * (1 << lgb) + 1);
				switch(lgb) {
				case 1: HFDL_LAUNCH(fft_last_pass_nat<1>, grid, dim3(256), smem, st, a); break;
				case 2: HFDL_LAUNCH(fft_last_pass_nat<2>, grid, dim3(256), smem, st, a); break;
				case 3: HFDL_LAUNCH(fft_last_pass_nat<3>, grid, dim3(256), smem, st, a); break;
				default: HFDL_LAUNCH(fft_last_pass_nat<4>, grid, dim3(256), smem, st, a); break;
				}
			} else if(!getenv("HFDL_B200_SMEM_FFT") && lgb >= 1 && lgb <= 4 && (256 >> lgb) <= rows) {
				dim3 grid((unsigned)(rows / (256 >> lgb)), (unsigned)nb);
				const size_t smem = sizeof(cf) * 256 * 33;
				switch(lgb) {
				case 1: HFDL_LAUNCH(fft_row_pass_reg<1>, grid, dim3(256), smem, st, a); break;
				case 2: HFDL_LAUNCH(fft_row_pass_reg<2>, grid, dim3(256), smem, st, a); break;
				case 3: HFDL_LAUNCH(fft_row_pass_reg<3>, grid, dim3(256), smem, st, a); break;
				default: HFDL_LAUNCH(fft_row_pass_reg<4>, grid, dim3(256), smem, st, a); break;
				}
			} else {
				dim3 grid((unsigned)(rows / a.R), (unsigned)nb);
				size_t smem = sizeof(cf) * (size_t)L * a.R;
				HFDL_LAUNCH(fft_row_pass, grid, dim3(HFDL_FFT_THREADS), smem, st, a);
			}
		}
		prof_end2(fe, pr, st);
		if(fe) fe->launches++;
		outer *= L;
	}
	CK(cudaGetLastError());
	return 0;
}

int bytes_per_sample(int sfmt) { return sfmt == HFDL_SFMT_CF32 ? 8 : (sfmt == HFDL_SFMT_CS16 ? 4 : 2); }

void to_pdu(hfdl_b200_frontend *fe, const PduRec &r, hfdl_b200_pdu_t &p) {    // dispatch_pdu, hfdl.c:1058-1080
	static const int ar[8] = { 1, 1, 2, 3, 1, 1, 2, 3 }, cr[8] = { 4, 2, 2, 2, 4, 2, 2, 2 };
	memset(&p, 0, sizeof(p));
	p.version = 1;
	p.freq = fe->freqs[r.channel];
	p.bit_rate = 1800 * ar[r.M1] / cr[r.M1] * 30 / (30 + 15);
	p.freq_err_hz = r.freq_err_hz;
	p.rssi = 20.0f * log10f(r.signal_level);
	p.noise_floor = 20.0f * log10f(r.noise_floor);
	p.slot = r.M1 < 4 ? 'S' : 'D';
	p.M1 = r.M1; p.crc_good = r.crc_good;
	p.train_bits_bad = r.train_bits_bad; p.train_bits_total = r.train_bits_total;
	p.sample_cnt_a2 = r.sample_cnt_a2; p.sample_cnt_end = r.sample_cnt_end;
	p.rx_time_s = (double)r.sample_cnt_a2 / 5400.0 - (448.0 + 2 * 127.0) / 1800.0;
	p.signal_level = r.signal_level; p.noise_floor_lin = r.noise_floor;
	p.len = r.len;
	memcpy(p.octets, r.octets, (size_t)r.len);
	p.frame_status = r.frame_status; p.direction = r.direction;
	p.lpdus_processed = r.lpdus_processed; p.lpdus_good = r.lpdus_good; p.lpdus_bad_fcs = r.lpdus_bad_fcs; p.lpdus_too_short = r.lpdus_too_short;
	p.lpdu_good_mask = r.lpdu_good_mask;
	hfdl_b200_frontend::ChanTally &t = fe->tally[(size_t)r.channel];      // the statsd counters of pdu.c / mpdu.c / spdu.c / lpdu.c
	t.processed++;
	if(r.frame_status == 0) { t.good++; if(r.direction) t.air2gnd++; else t.gnd2air++; }
	else if(r.frame_status == 1) t.bad_fcs++;
	else t.too_short++;
	t.lpdus_processed += r.lpdus_processed; t.lpdus_good += r.lpdus_good; t.lpdus_bad += r.lpdus_bad_fcs; t.lpdus_short += r.lpdus_too_short;
}

// Results of one finished batch: PDU records of set q -> host queue.  wait = false: only if the batch has finished.
int collect_batch(hfdl_b200_frontend *fe, int q, bool wait = true) {
	if(!fe->flight[q].busy) return 0;
	if(!wait) {
		cudaError_t e = cudaEventQuery(fe->ev_fec_done[q]);
		if(e == cudaErrorNotReady) return 0;
		if(e != cudaSuccess) { fprintf(stderr, "hfdl_b200: CUDA error '%s' while polling\n", cudaGetErrorString(e)); return -1; }
	} else {
		CK(cudaEventSynchronize(fe->ev_fec_done[q]));
	}
	fe->flight[q].busy = false;
	int nfr = *fe->h_nframes[q];
	if(nfr > fe->max_frames) {
		fprintf(stderr, "hfdl_b200: frame queue overflow (%d > %d), frames dropped\n", nfr, fe->max_frames);
		nfr = fe->max_frames;
	}
	if(nfr > 0) {
		const PduRec *hp = fe->h_pdus[q];
		// canonical order within a batch: (end sample, channel) -- the reference emits in thread-race order
		std::vector<int> order((size_t)nfr);
		for(int i = 0; i < nfr; i++) order[(size_t)i] = i;
		std::sort(order.begin(), order.end(), [&](int x, int y) {
			const PduRec &a = hp[x], &b = hp[y];
			if(a.sample_cnt_end != b.sample_cnt_end) return a.sample_cnt_end < b.sample_cnt_end;
			return a.channel < b.channel;
		});
		for(int i : order) { hfdl_b200_pdu_t p; to_pdu(fe, hp[i], p); fe->pduq.push_back(p); }
	}
	return 0;
}

// Everything enqueued so far has finished and its PDUs are in the host queue (oldest batch first).
int drain(hfdl_b200_frontend *fe) {
	for(int k = 0; k < HFDL_NSETS; k++)          // oldest batch first: set (batch_seq + k) % NSETS
		if(collect_batch(fe, (int)((fe->batch_seq + k) % HFDL_NSETS))) return -1;
	CK(cudaStreamSynchronize(fe->stream));
	CK(cudaStreamSynchronize(fe->stream2));
	CK(cudaStreamSynchronize(fe->st_loop));
	CK(cudaStreamSynchronize(fe->st_fec));
	return 0;
}

// One batch of nb overlap-save blocks, fully asynchronous; ONE launch per demodulator stage per batch.  Stages and
// the stream each runs on:
//   front  (stream)   per sub-batch of Bsub blocks: FFT passes + chan_extract (the intermediate spectra of a sub-batch
//                     fit the L2: only the raw samples come from HBM); then resamp -> d_rs[p]
//   agc    (stream2)  demod_carry (history from set p^1), agc_kernel, bank_kernel -> work set p
//   loop   (st_loop)  loop_kernel over the whole batch                  (the latency-bound stage: sets the pace)
//   fec    (st_fec)   fec_kernel, D2H of the PDU records of set p
// Consecutive batches use alternate sets (p = batch parity) of every buffer a later stage reads, so front + agc + bank
// of batch i+1 run beside loop of batch i and fec of batch i-1; events order the reuse of a set two batches later.
// The host collects the PDUs of batch i-1 after it has enqueued batch i.
int run_batch_impl(hfdl_b200_frontend *fe, const RawSource &src0, int nb, const cf *slices) {
	const int p = (int)(fe->batch_seq % HFDL_NSETS), pp = (p + HFDL_NSETS - 1) % HFDL_NSETS, p2 = (p + HFDL_NSETS - 2) % HFDL_NSETS;
	if(collect_batch(fe, p)) return -1;          // batch i-NSETS (normally collected long ago): every set-p buffer is free
	cudaStream_t st = fe->stream, st2 = fe->stream2, stl = fe->st_loop, stf = fe->st_fec;
	const auto &g = fe->g;
	hfdl_b200_frontend::Flight &cur = fe->flight[p];
	ProfRec pr;
	// FFT passes over sub-batches of Bsub blocks (the in-place intermediate stays in L2 between the passes); the last pass
	// writes the natural-order spectrum of block b at d_spec + b*N -- only the granules some channel's slice reads
	// (sharded spectrum: the pass-band slices of the batch arrive from the exchange instead, once ev_ext has fired)
	if(slices) CK(cudaStreamWaitEvent(st, fe->ev_ext, 0));
	for(int b0 = 0; b0 < nb && !slices; b0 += fe->Bsub) {
		const int nsb = std::min(fe->Bsub, nb - b0);
		RawSource src = src0;
		src.pos0 = src0.pos0 + (long long)b0 * src0.block_stride;
		cf *spec = fe->plan.natural ? fe->d_spec + (long long)b0 * g.fft_size : nullptr;
		cf *work = fe->plan.natural ? fe->d_work : fe->d_work + (long long)b0 * g.fft_size;
		if(run_fft(fe, fe->fft, fe->plan, src, work, spec, nsb, st, fe->d_mask)) return -1;
	}
	{
		ChanArgs a;
		a.work = fe->plan.natural ? fe->d_spec : fe->d_work; a.slices = slices; a.tapslice = fe->d_tapslice; a.offsetbin = fe->d_offsetbin; a.dsa_rate = fe->d_dsa_rate;
		a.bb = fe->d_bb; a.tw = fe->fft.d_tw; a.pl = fe->plan;
		a.M = g.fft_inv_size; a.lgM = hfdl_ilog2(g.fft_inv_size); a.scrap = g.scrap; a.post_dec = g.post_decimation;
		a.out_per_block = fe->out_per_block; a.bb_stride = fe->bb_stride;
		a.out_index0 = fe->blocks_done * (long long)fe->out_per_block;
		a.block0 = 0;
		a.inv_norm = 1.0f / (float)(g.pre_decimation * g.fft_inv_size);
		prof_begin(fe, KC_CHAN, pr);
		HFDL_LAUNCH(chan_extract, dim3((unsigned)fe->C, (unsigned)nb), dim3(HFDL_FFT_THREADS), sizeof(cf) * (size_t)a.M, st, a);
		prof_end(fe, pr);
		fe->launches++;
	}
	fe->last_sub_blocks = nb;
	const long long n_in = (long long)nb * fe->out_per_block;
	int n_out = 0;
	{
		unsigned long long span = (unsigned long long)n_in << 24;
		if(fe->rs_phi0 < span) n_out = (int)((span - fe->rs_phi0 + fe->rs_step - 1) / fe->rs_step);
		ResampArgs a;
		a.bb = fe->d_bb; a.bb_stride = fe->bb_stride; a.rs = fe->d_rs[p]; a.rs_stride = fe->rs_stride; a.h = fe->d_rs_h;
		a.phi0 = fe->rs_phi0; a.step = fe->rs_step; a.n_out = n_out;
		if(n_out > 0) {
			CK(cudaStreamWaitEvent(st, fe->ev_bank[p], 0));              // agc of batch i-2 has read d_rs[p]
			prof_begin(fe, KC_RESAMP, pr);
			HFDL_LAUNCH(resamp_kernel, dim3((unsigned)((n_out + 255) / 256), (unsigned)fe->C), dim3(256), 0, st, a);
			prof_end(fe, pr);
			fe->launches++;
		}
		fe->rs_phi0 = fe->rs_phi0 + (unsigned long long)n_out * fe->rs_step - span;
		HFDL_LAUNCH(bb_carry, dim3((unsigned)fe->C), dim3(32), 0, st, fe->d_bb, fe->bb_stride, n_in);
		fe->launches++;
	}
	CK(cudaEventRecord(fe->ev_front[p], st));
	// ---- agc + bank: work set p (free once loop of batch i-2 is done)
	CK(cudaStreamWaitEvent(st2, fe->ev_front[p], 0));
	CK(cudaStreamWaitEvent(st2, fe->ev_loop[p], 0));
	// the last HIST samples of the previous batch's AGC / matched-filter outputs (set p^1) go in front of set p
	HFDL_LAUNCH(demod_carry, dim3((unsigned)fe->C), dim3(64), 0, st2, fe->d_agc[pp], fe->d_agc[p], fe->agc_stride, fe->d_mfo[pp], fe->d_mfo[p], fe->mfo_stride, fe->n_out_prev);
	fe->launches++;
	if(n_out > 0) {
		AgcArgs a;
		a.rs = fe->d_rs[p]; a.rs_stride = fe->rs_stride; a.n_samples = n_out; a.state = fe->d_agc_state;
		a.agc_out = fe->d_agc[p]; a.agc_stride = fe->agc_stride; a.lvl = fe->d_lvl[p]; a.lvl_stride = fe->rs_stride;
		prof_begin2(fe, KC_AGC, pr, st2);
		HFDL_LAUNCH(agc_kernel, dim3((unsigned)fe->C), dim3(32), 0, st2, a);
		prof_end2(fe, pr, st2);
		BankArgs b;
		b.agc_out = fe->d_agc[p]; b.agc_stride = fe->agc_stride; b.n_samples = n_out; b.mfo = fe->d_mfo[p]; b.mfo_stride = fe->mfo_stride;
		b.bank = fe->d_bank[p]; b.bank_stride = fe->rs_stride; b.tab = fe->d_tab;
		prof_begin2(fe, KC_BANK, pr, st2);
		HFDL_LAUNCH(bank_kernel, dim3((unsigned)((n_out + HFDL_BANK_TILE - 1) / HFDL_BANK_TILE), (unsigned)fe->C), dim3(256), 0, st2, b);
		prof_end2(fe, pr, st2);
		fe->launches += 2;
		if(fe->cfg.capture_channel >= 0 && fe->cap_n < fe->cfg.capture_max) {       // f_agc_out / f_mf_out checkpoints
			long long n = std::min<long long>(n_out, fe->cfg.capture_max - fe->cap_n);
			int cc = fe->cfg.capture_channel;
			CK(cudaMemcpyAsync(fe->d_cap_agc + fe->cap_n, fe->d_agc[p] + (long long)cc * fe->agc_stride + HFDL_AGC_HIST, sizeof(cf) * (size_t)n, cudaMemcpyDeviceToDevice, st2));
			CK(cudaMemcpyAsync(fe->d_cap_mf + fe->cap_n, fe->d_mfo[p] + (long long)cc * fe->mfo_stride + HFDL_MFO_HIST, sizeof(cf) * (size_t)n, cudaMemcpyDeviceToDevice, st2));
		}
		if(fe->cfg.capture_channel >= 0) fe->cap_n += n_out;
	}
	CK(cudaEventRecord(fe->ev_bank[p], st2));
	// ---- loop: frame records of set p are free once fec of batch i-2 is done
	CK(cudaStreamWaitEvent(stl, fe->ev_bank[p], 0));
	CK(cudaStreamWaitEvent(stl, fe->ev_fec_done[p], 0));
	// data-symbol slots: a channel has room for the frames of two batches, so fec of batch i-2 must have read its frames
	CK(cudaStreamWaitEvent(stl, fe->ev_fec_done[p2], 0));
	CK(cudaMemsetAsync(fe->d_nframes[p], 0, sizeof(int), stl));
	if(n_out > 0) {
		LoopArgs l;
		l.bank = fe->d_bank[p]; l.bank_stride = fe->rs_stride; l.mfo = fe->d_mfo[p]; l.mfo_stride = fe->mfo_stride;
		l.lvl = fe->d_lvl[p]; l.lvl_stride = fe->rs_stride; l.n_samples = n_out; l.n_channels = fe->C;
		l.state = fe->d_state; l.tab = fe->d_tab; l.datasym = fe->d_datasym; l.nslots = fe->nslots;
		l.frames = fe->d_frames[p]; l.nframes = fe->d_nframes[p]; l.max_frames = fe->max_frames;
		l.cap_channel = fe->cfg.capture_channel; l.cap_eq = fe->d_cap_eq; l.cap_cnt = fe->d_cap_cnt; l.cap_max = fe->cfg.capture_max;
		l.debug_mode = fe->debug_mode;
		l.dbg_cycles = fe->d_dbg;
		prof_begin2(fe, KC_LOOP, pr, stl);
		const int nch = fe->loop_layout == LK_PACK4 ? 4 : 2;
		const size_t smem = std::max((size_t)nch * HFDL_LK_SMEM_CH, fe->loop_smem_pad_to);
		const dim3 grid((unsigned)((fe->C + nch - 1) / nch));
		if(fe->loop_layout == LK_PACK4) HFDL_LAUNCH((loop_kernel<4, false>), grid, dim3(lk_threads(4, false)), smem, stl, l);
		else if(fe->loop_layout == LK_ROLE2) HFDL_LAUNCH((loop_kernel<2, true>), grid, dim3(lk_threads(2, true)), smem, stl, l);
		else HFDL_LAUNCH((loop_kernel<2, false>), grid, dim3(lk_threads(2, false)), smem, stl, l);
		prof_end2(fe, pr, stl);
		fe->launches++;
	}
	CK(cudaEventRecord(fe->ev_loop[p], stl));
	{
		CK(cudaStreamWaitEvent(stf, fe->ev_loop[p], 0));
		FecArgs a;
		a.frames = fe->d_frames[p]; a.nframes = fe->d_nframes[p]; a.max_frames = fe->max_frames; a.datasym = fe->d_datasym; a.nslots = fe->nslots;
		a.tab = fe->d_tab; a.pdus = fe->d_pdus[p]; a.soft_out = nullptr; a.vin_direct = nullptr; a.vin_nbits = 0;
		prof_begin2(fe, KC_FEC, pr, stf);
		HFDL_LAUNCH(fec_kernel, dim3((unsigned)fe->max_frames), dim3(32), HFDL_FEC_SMEM, stf, a);
		prof_end2(fe, pr, stf);
		fe->launches++;
		CK(cudaGetLastError());
		CK(cudaMemcpyAsync(fe->h_nframes[p], fe->d_nframes[p], sizeof(int), cudaMemcpyDeviceToHost, stf));
		CK(cudaMemcpyAsync(fe->h_pdus[p], fe->d_pdus[p], sizeof(PduRec) * (size_t)fe->max_frames, cudaMemcpyDeviceToHost, stf));
		CK(cudaEventRecord(fe->ev_fec_done[p], stf));
	}
	cur.busy = true;
	fe->batch_seq++;
	fe->blocks_done += nb;
	fe->n_out_prev = n_out;
	fe->last_nblocks = nb; fe->last_nout = n_out;
	// while this batch runs, pick up what has finished (oldest first; stops at the first batch still running)
	for(int k = 1; k < HFDL_NSETS; k++) {
		const int q = (p + k) % HFDL_NSETS;
		if(!fe->flight[q].busy) continue;
		if(collect_batch(fe, q, false)) return -1;
		if(fe->flight[q].busy) break;
	}
	return 0;
}

// A failure in the middle of enqueueing leaves events unrecorded and counters half advanced: stop everything and
// refuse further work instead of running on inconsistent pipeline state.
int run_batch(hfdl_b200_frontend *fe, const RawSource &src, int nb, const cf *slices = nullptr) {
	if(fe->failed) return -1;
	if(run_batch_impl(fe, src, nb, slices) == 0) return 0;
	fe->failed = true;
	cudaStream_t sts[4] = { fe->stream, fe->stream2, fe->st_loop, fe->st_fec };
	for(cudaStream_t q : sts) if(q) cudaStreamSynchronize(q);
	for(int q = 0; q < HFDL_NSETS; q++) fe->flight[q].busy = false;
	fprintf(stderr, "hfdl_b200: batch failed, frontend disabled\n");
	return -1;
}

// granules of the spectrum some channel's pass-band slice reads (chan_extract / slice_pack: bins offsetbin - M/2 ..
// offsetbin + M/2 - 1 mod N); the last FFT pass stores only those.  With a capture channel (parity / debug: the
// CP_SPECTRUM checkpoint wants every bin) there is no mask and everything is stored.
int build_spec_mask(hfdl_b200_frontend *fe, const std::vector<int> &offsetbin) {
	const long long N = fe->g.fft_size, M = fe->g.fft_inv_size;
	const long long ngran = N >> HFDL_SPEC_LG_GRAN;
	std::vector<unsigned> m((size_t)((ngran + 31) / 32), 0u);
	for(int ob : offsetbin) {
		const long long off = ob;
		for(long long k = off - M / 2; k < off + M / 2; k += HFDL_SPEC_GRAN) {
			const long long gidx = (((k % N) + N) % N) >> HFDL_SPEC_LG_GRAN;
			m[(size_t)(gidx >> 5)] |= 1u << (gidx & 31);
		}
		const long long last = ((((off + M / 2 - 1) % N) + N) % N) >> HFDL_SPEC_LG_GRAN;
		m[(size_t)(last >> 5)] |= 1u << (last & 31);
	}
	if(!fe->d_mask) CK(cudaMalloc((void **)&fe->d_mask, sizeof(unsigned) * m.size()));
	CK(cudaMemcpy(fe->d_mask, m.data(), sizeof(unsigned) * m.size(), cudaMemcpyHostToDevice));
	long long set = 0;
	for(unsigned w : m) set += __builtin_popcount(w);
	fe->spec_fill = (double)set / (double)ngran;
	return 0;
}

int compute_tapslices(hfdl_b200_frontend *fe) {
	// fft_channelizer_create (fastddc.c:217-252): taps -> N-point forward FFT; keep the M bins of the slice
	const auto &g = fe->g;
	const int N = g.fft_size, M = g.fft_inv_size, C = fe->C;
	int chunk = std::min(fe->Bsub, C);
	cf *d_in = nullptr;
	CK(cudaMalloc((void **)&d_in, sizeof(cf) * (size_t)N * chunk));
	std::vector<std::vector<std::complex<float>>> taps((size_t)chunk);
	for(int c0 = 0; c0 < C; c0 += chunk) {
		int nc = std::min(chunk, C - c0);
		std::vector<std::thread> th;
		for(int i = 0; i < nc; i++) th.emplace_back([&, i] {
			float fs = fe->chg[(size_t)(c0 + i)].freq_shift;
			float half = 0.5f / g.decimation;
			hfdl_design::bandpass_taps(taps[(size_t)i], g.taps_length, (-fs) - half, (-fs) + half);
		});
		for(auto &t : th) t.join();
		CK(cudaMemsetAsync(d_in, 0, sizeof(cf) * (size_t)N * nc, fe->stream));
		for(int i = 0; i < nc; i++)
			CK(cudaMemcpyAsync(d_in + (size_t)i * N, taps[(size_t)i].data(), sizeof(cf) * (size_t)g.taps_length, cudaMemcpyHostToDevice, fe->stream));
		RawSource src;
		src.base = d_in; src.ring_len = (long long)N * nc; src.pos0 = 0; src.ring_origin = 0; src.block_stride = N; src.sfmt = HFDL_SFMT_CF32;
		if(run_fft(nullptr, fe->fft, fe->plan, src, fe->d_work, fe->d_spec, nc, fe->stream, fe->d_mask)) return -1;
		HFDL_LAUNCH(tapslice_gather, dim3((unsigned)((M + 255) / 256), (unsigned)nc), dim3(256), 0, fe->stream,
			fe->plan.natural ? fe->d_spec : fe->d_work, fe->plan, M, fe->d_offsetbin, c0, fe->d_tapslice);
		CK(cudaGetLastError());
		CK(cudaStreamSynchronize(fe->stream));
	}
	cudaFree(d_in);
	return 0;
}

}  // namespace

extern "C" {

int32_t hfdl_b200_device_count(void) {
	int n = 0;
	if(cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

int32_t hfdl_b200_pdu_len(int32_t M1) { return (M1 < 0 || M1 > 7) ? -1 : hfdl_design::pdu_len(M1); }

int32_t hfdl_b200_create(hfdl_b200_frontend_t **out, const hfdl_b200_config_t *cfg) {
	if(!out || !cfg || cfg->nfreq < 1 || !cfg->freqs_hz) { fprintf(stderr, "hfdl_b200_create: bad arguments\n"); return -1; }
	if(hfdl_b200_device_count() < 1) { fprintf(stderr, "hfdl_b200_create: no CUDA device -- this library has no CPU path\n"); return -1; }
	if(cfg->sample_format < HFDL_SFMT_CU8 || cfg->sample_format > HFDL_SFMT_CF32) { fprintf(stderr, "hfdl_b200_create: bad sample format\n"); return -1; }
	hfdl_b200_frontend *fe = new hfdl_b200_frontend();
	fe->cfg = *cfg;
	fe->freqs.assign(cfg->freqs_hz, cfg->freqs_hz + cfg->nfreq);
	fe->cfg.freqs_hz = fe->freqs.data();
	fe->C = cfg->nfreq;
	fe->tally.resize((size_t)fe->C);
	fe->sfmt = cfg->sample_format;
	fe->bps = bytes_per_sample(fe->sfmt);
	if(!hfdl_design::geometry_init(fe->g, cfg->sample_rate)) { fprintf(stderr, "hfdl_b200_create: unsupported sample rate %d\n", cfg->sample_rate); delete fe; return -1; }
	const auto &g = fe->g;
	// check_frequency_span (main.c:214-226)
	for(int i = 0; i < fe->C; i++) {
		if(abs(cfg->centerfreq_hz - fe->freqs[(size_t)i]) >= cfg->sample_rate / 2) {
			fprintf(stderr, "hfdl_b200_create: channel %d Hz too far from the centre frequency %d Hz\n", fe->freqs[(size_t)i], cfg->centerfreq_hz);
			delete fe; return -1;
		}
		fe->chg.push_back(hfdl_design::channel_geom(g, cfg->sample_rate, cfg->centerfreq_hz, fe->freqs[(size_t)i]));
	}
	fe->out_per_block = g.post_input_size / g.post_decimation;
	if(g.post_input_size % g.post_decimation != 0 || hfdl_ilog2(g.fft_inv_size) > 12 || g.fft_size > (1 << 27)) {
		fprintf(stderr, "hfdl_b200_create: geometry outside the supported range\n"); delete fe; return -1;
	}
	fe->resamp_rate = (float)(1800 * 3) / ((float)cfg->sample_rate / (float)g.decimation);     // hfdl.c:471
	if(!(fe->resamp_rate >= 0.5f && fe->resamp_rate <= 1.0f)) { fprintf(stderr, "hfdl_b200_create: resampling rate %f outside [0.5,1]\n", fe->resamp_rate); delete fe; return -1; }
	fe->plan = make_plan(g.fft_size);
	fe->Bmax = cfg->max_blocks_per_batch > 0 ? cfg->max_blocks_per_batch : std::max(1, std::min(64, (int)((1024ll << 20) / ((long long)g.fft_size * 8))));
	// loop_kernel addresses the resampled samples of one launch (= one batch) with 20 bits (HFDL_LK_MAXN)
	if((long long)fe->Bmax * fe->out_per_block >= (long long)HFDL_LK_MAXN) fe->Bmax = (int)((HFDL_LK_MAXN - 1) / fe->out_per_block);
	// FFT sub-batches: the passes and chan_extract of Bsub blocks run back to back so that the intermediate and the
	// final spectra (2 x Bsub x N x 8 bytes) stay in the 126 MB L2 instead of making a round trip through HBM
	{
		const char *e = getenv("HFDL_B200_FFT_SUB_MB");
		const long long budget = (e ? atoll(e) : 80) << 20;                 // bytes of spectrum per sub-batch
		fe->Bsub = (int)std::max<long long>(1, std::min<long long>(fe->Bmax, budget / ((long long)g.fft_size * 8)));
	}
	{ const char *dbg = getenv("HFDL_B200_DEBUG"); fe->debug_mode = dbg ? atoi(dbg) : 0; }
	{
		// loop_kernel's CTA layout and SMs.  Asking for more shared memory than the bank rings need keeps CTAs with a sizeable
		// shared-memory footprint (FFT passes, chan_extract, fec) away from the latency-bound warps.  Defaults: one GPU --
		// four channels per CTA, no padding (the FFT stream needs the SMs); sharded spectrum over several GPUs
		// (hfdl_b200_set_exchange) -- two channels per CTA, demodulator warps on schedulers of their own, padded.
		// HFDL_B200_LOOP_LAYOUT = pack4 | role2 | pair2 and HFDL_B200_LOOP_SMEM_KB (0 = no padding) override.
		const char *lay = getenv("HFDL_B200_LOOP_LAYOUT"), *e = getenv("HFDL_B200_LOOP_SMEM_KB");
		fe->loop_auto = (lay == nullptr && e == nullptr);
#ifdef HFDL_CUSIM
		fe->loop_layout = LK_PAIR2;                    // host emulation: fewer host threads per CTA
#endif
		if(lay) fe->loop_layout = !strcmp(lay, "role2") ? LK_ROLE2 : (!strcmp(lay, "pair2") ? LK_PAIR2 : LK_PACK4);
		long kb = e ? atol(e) : 0;
		if(kb > 216) kb = 216;
		fe->loop_smem_pad_to = (size_t)(kb > 0 ? kb : 0) * 1024;
	}
	if(cfg->capture_channel >= fe->C) fe->cfg.capture_channel = -1;
	if(fe->cfg.capture_max < 0) fe->cfg.capture_max = 0;

#define CKD(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) { fprintf(stderr, "hfdl_b200_create: CUDA error '%s' (%s)\n", cudaGetErrorString(e_), #call); hfdl_b200_destroy(fe); return -1; } } while(0)
	CKD(cudaSetDevice(cfg->device));
	CKD(cudaStreamCreateWithFlags(&fe->stream, cudaStreamNonBlocking));
	CKD(cudaStreamCreateWithFlags(&fe->stream2, cudaStreamNonBlocking));
	CKD(cudaStreamCreateWithFlags(&fe->st_loop, cudaStreamNonBlocking));
	CKD(cudaStreamCreateWithFlags(&fe->st_fec, cudaStreamNonBlocking));
	CKD(cudaStreamCreateWithFlags(&fe->st_stats, cudaStreamNonBlocking));
	CKD(cudaStreamCreateWithFlags(&fe->st_h2d, cudaStreamNonBlocking));
	CKD(cudaEventCreateWithFlags(&fe->ev_h2d, cudaEventDisableTiming));
	CKD(cudaEventCreateWithFlags(&fe->ev_ext, cudaEventDisableTiming));
	for(int q = 0; q < HFDL_NSETS; q++) {
		CKD(cudaEventCreateWithFlags(&fe->ev_front[q], cudaEventDisableTiming));
		CKD(cudaEventCreateWithFlags(&fe->ev_bank[q], cudaEventDisableTiming));
		CKD(cudaEventCreateWithFlags(&fe->ev_loop[q], cudaEventDisableTiming));
		CKD(cudaEventCreateWithFlags(&fe->ev_fec_done[q], cudaEventDisableTiming));
	}
	if(fe->fft.init()) { hfdl_b200_destroy(fe); return -1; }
	const int C = fe->C, N = g.fft_size, M = g.fft_inv_size, B = fe->Bmax;
	// natural-order plans: the in-place workspace only holds one sub-batch, the spectra of the whole batch go to d_spec
	CKD(cudaMalloc((void **)&fe->d_work, sizeof(cf) * (size_t)N * (fe->plan.natural ? std::max(fe->Bsub, 1) : B)));
	if(fe->plan.natural) CKD(cudaMalloc((void **)&fe->d_spec, sizeof(cf) * (size_t)N * B));
	if(fe->plan.natural && fe->cfg.capture_channel < 0 && !getenv("HFDL_B200_NO_SPEC_MASK")) {
		std::vector<int> ob((size_t)C);
		for(int i = 0; i < C; i++) ob[(size_t)i] = fe->chg[(size_t)i].offsetbin;
		if(build_spec_mask(fe, ob)) { hfdl_b200_destroy(fe); return -1; }
	}
	// host-fed ring: B + 1 blocks of capacity plus one batch that may still be read by the channeliser stage while the next
	// samples arrive (the H2D copies run on their own stream, beside the FFT of the previous batch)
	fe->ring_len = (long long)g.overlap_length + (long long)(2 * B + 1) * g.input_size;
	CKD(cudaMalloc(&fe->d_ring, (size_t)fe->ring_len * fe->bps));
	CKD(cudaMemset(fe->d_ring, 0, (size_t)fe->ring_len * fe->bps));
	CKD(cudaMalloc((void **)&fe->d_tapslice, sizeof(cf) * (size_t)C * M));
	CKD(cudaMalloc((void **)&fe->d_offsetbin, sizeof(int) * (size_t)C));
	CKD(cudaMalloc((void **)&fe->d_dsa_rate, sizeof(float) * (size_t)C));
	{
		std::vector<int> ob((size_t)C); std::vector<float> dr((size_t)C);
		for(int i = 0; i < C; i++) { ob[(size_t)i] = fe->chg[(size_t)i].offsetbin; dr[(size_t)i] = fe->chg[(size_t)i].dsa_rate; }
		CKD(cudaMemcpy(fe->d_offsetbin, ob.data(), sizeof(int) * (size_t)C, cudaMemcpyHostToDevice));
		CKD(cudaMemcpy(fe->d_dsa_rate, dr.data(), sizeof(float) * (size_t)C, cudaMemcpyHostToDevice));
	}
	fe->bb_stride = HFDL_RS_HIST + (long long)B * fe->out_per_block + 16;
	CKD(cudaMalloc((void **)&fe->d_bb, sizeof(cf) * (size_t)C * fe->bb_stride));
	CKD(cudaMemset(fe->d_bb, 0, sizeof(cf) * (size_t)C * fe->bb_stride));
	fe->rs_stride = (long long)B * fe->out_per_block + 16;
	for(int q = 0; q < HFDL_NSETS; q++) CKD(cudaMalloc((void **)&fe->d_rs[q], sizeof(cf) * (size_t)C * fe->rs_stride));
	{
		std::vector<float> h((size_t)HFDL_RS_NPFB * HFDL_RS_TAPS);
		hfdl_design::resamp_design(fe->resamp_rate, h.data(), &fe->rs_step);
		CKD(cudaMalloc((void **)&fe->d_rs_h, sizeof(float) * h.size()));
		CKD(cudaMemcpy(fe->d_rs_h, h.data(), sizeof(float) * h.size(), cudaMemcpyHostToDevice));
	}
	{
		DemodTables *T = new DemodTables();
		hfdl_design::demod_tables(*T);
		CKD(cudaMalloc((void **)&fe->d_tab, sizeof(DemodTables)));
		CKD(cudaMemcpy(fe->d_tab, T, sizeof(DemodTables), cudaMemcpyHostToDevice));
		std::vector<DemodState> st((size_t)C);
		for(int i = 0; i < C; i++) hfdl_design::demod_state_init(st[(size_t)i], *T);
		CKD(cudaMalloc((void **)&fe->d_state, sizeof(DemodState) * (size_t)C));
		CKD(cudaMemcpy(fe->d_state, st.data(), sizeof(DemodState) * (size_t)C, cudaMemcpyHostToDevice));
		std::vector<AgcState> ag((size_t)C);
		for(int i = 0; i < C; i++) { ag[(size_t)i].g = 1.0f; ag[(size_t)i].y2 = 1.0f; }      // agc_crcf_create / reset
		CKD(cudaMalloc((void **)&fe->d_agc_state, sizeof(AgcState) * (size_t)C));
		CKD(cudaMemcpy(fe->d_agc_state, ag.data(), sizeof(AgcState) * (size_t)C, cudaMemcpyHostToDevice));
		fe->agc_stride = HFDL_AGC_HIST + fe->rs_stride; fe->mfo_stride = HFDL_MFO_HIST + fe->rs_stride;
		for(int q = 0; q < HFDL_NSETS; q++) {
			CKD(cudaMalloc((void **)&fe->d_agc[q], sizeof(cf) * (size_t)C * fe->agc_stride));
			CKD(cudaMemset(fe->d_agc[q], 0, sizeof(cf) * (size_t)C * fe->agc_stride));
			CKD(cudaMalloc((void **)&fe->d_mfo[q], sizeof(cf) * (size_t)C * fe->mfo_stride));
			CKD(cudaMemset(fe->d_mfo[q], 0, sizeof(cf) * (size_t)C * fe->mfo_stride));
			CKD(cudaMalloc((void **)&fe->d_lvl[q], sizeof(float) * (size_t)C * fe->rs_stride));
			CKD(cudaMalloc((void **)&fe->d_bank[q], sizeof(cf) * (size_t)C * fe->rs_stride * 32));
		}
		delete T;
	}
	{
		// A frame lasts >= 2.34 s of signal (448 + 531 + 72 * 45 symbols at 1800 baud): a batch can end at most
		// fpb frames per channel.  The data symbols of a frame stay in their slot until fec of that batch has run,
		// which overlaps with the loop stage of the next batch -> twice as many slots.
		const double batch_s = (double)B * g.input_size / (double)cfg->sample_rate;
		const int fpb = (int)(batch_s / 2.3) + 2;
		fe->nslots = std::max(HFDL_FRAME_SLOTS_MIN, 2 * fpb);
		fe->max_frames = C * fpb;
	}
	CKD(cudaMalloc((void **)&fe->d_datasym, sizeof(cf) * (size_t)C * fe->nslots * HFDL_DATA_SYMS_MAX));
	for(int q = 0; q < HFDL_NSETS; q++) {
		CKD(cudaMalloc((void **)&fe->d_frames[q], sizeof(FrameRec) * (size_t)fe->max_frames));
		CKD(cudaMalloc((void **)&fe->d_nframes[q], sizeof(int)));
		CKD(cudaMemset(fe->d_nframes[q], 0, sizeof(int)));
		CKD(cudaMalloc((void **)&fe->d_pdus[q], sizeof(PduRec) * (size_t)fe->max_frames));
		CKD(cudaMallocHost((void **)&fe->h_pdus[q], sizeof(PduRec) * (size_t)fe->max_frames));
		CKD(cudaMallocHost((void **)&fe->h_nframes[q], sizeof(int)));
	}
	if(fe->cfg.capture_channel >= 0 && fe->cfg.capture_max > 0) {
		size_t n = (size_t)fe->cfg.capture_max;
		CKD(cudaMalloc((void **)&fe->d_cap_agc, sizeof(cf) * n));
		CKD(cudaMalloc((void **)&fe->d_cap_mf, sizeof(cf) * n));
		CKD(cudaMalloc((void **)&fe->d_cap_eq, sizeof(cf) * n));
	} else fe->cfg.capture_channel = -1;
	CKD(cudaMalloc((void **)&fe->d_cap_cnt, sizeof(int) * 2));
	CKD(cudaMemset(fe->d_cap_cnt, 0, sizeof(int) * 2));
	fe->tmp_len = std::max((long long)N, (long long)B * fe->out_per_block + 64);
	CKD(cudaMalloc((void **)&fe->d_tmp, sizeof(cf) * (size_t)fe->tmp_len));
	if(fe->debug_mode) { CKD(cudaMalloc((void **)&fe->d_dbg, sizeof(long long) * 32 * (size_t)C)); CKD(cudaMemset(fe->d_dbg, 0, sizeof(long long) * 32 * (size_t)C)); }
	CKD(cudaEventCreate(&fe->ev0));
	CKD(cudaEventCreate(&fe->ev1));
	if(compute_tapslices(fe)) { hfdl_b200_destroy(fe); return -1; }
#undef CKD
	*out = fe;
	return 0;
}

void hfdl_b200_destroy(hfdl_b200_frontend_t *fe) {
	if(!fe) return;
	cudaSetDevice(fe->cfg.device);
	cudaStream_t sts[6] = { fe->stream, fe->stream2, fe->st_loop, fe->st_fec, fe->st_stats, fe->st_h2d };
	for(cudaStream_t q : sts) if(q) cudaStreamSynchronize(q);
	cudaFree(fe->d_work); cudaFree(fe->d_spec); cudaFree(fe->d_mask); cudaFree(fe->d_ring); cudaFree(fe->d_tapslice); cudaFree(fe->d_offsetbin); cudaFree(fe->d_dsa_rate);
	cudaFree(fe->d_bb); cudaFree(fe->d_rs_h); cudaFree(fe->d_tab); cudaFree(fe->d_state); cudaFree(fe->d_datasym);
	cudaFree(fe->d_cap_agc); cudaFree(fe->d_cap_mf); cudaFree(fe->d_cap_eq); cudaFree(fe->d_cap_cnt); cudaFree(fe->d_tmp); cudaFree(fe->d_dbg);
	cudaFree(fe->d_agc_state); cudaFree(fe->d_all_offsetbin);
	for(int q = 0; q < HFDL_NSETS; q++) {
		cudaFree(fe->d_agc[q]); cudaFree(fe->d_mfo[q]); cudaFree(fe->d_lvl[q]); cudaFree(fe->d_bank[q]);
		cudaFree(fe->d_rs[q]); cudaFree(fe->d_frames[q]); cudaFree(fe->d_nframes[q]); cudaFree(fe->d_pdus[q]);
		if(fe->h_pdus[q]) cudaFreeHost(fe->h_pdus[q]);
		if(fe->h_nframes[q]) cudaFreeHost(fe->h_nframes[q]);
		if(fe->ev_front[q]) cudaEventDestroy(fe->ev_front[q]);
		if(fe->ev_bank[q]) cudaEventDestroy(fe->ev_bank[q]);
		if(fe->ev_loop[q]) cudaEventDestroy(fe->ev_loop[q]);
		if(fe->ev_fec_done[q]) cudaEventDestroy(fe->ev_fec_done[q]);
	}
	for(auto &r : fe->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
	if(fe->ev0) cudaEventDestroy(fe->ev0);
	if(fe->ev1) cudaEventDestroy(fe->ev1);
	if(fe->ev_h2d) cudaEventDestroy(fe->ev_h2d);
	if(fe->ev_ext) cudaEventDestroy(fe->ev_ext);
	fe->fft.destroy();
	for(cudaStream_t q : sts) if(q) cudaStreamDestroy(q);
	delete fe;
}

int32_t hfdl_b200_get_geometry(const hfdl_b200_frontend_t *fe, hfdl_b200_geometry_t *o) {
	if(!fe || !o) return -1;
	const auto &g = fe->g;
	memset(o, 0, sizeof(*o));
	o->decimation = g.decimation; o->pre_decimation = g.pre_decimation; o->post_decimation = g.post_decimation;
	o->taps_length = g.taps_length; o->overlap_length = g.overlap_length; o->fft_size = g.fft_size; o->fft_inv_size = g.fft_inv_size;
	o->input_size = g.input_size; o->post_input_size = g.post_input_size; o->scrap = g.scrap; o->out_per_block = fe->out_per_block;
	o->transition_bw = g.transition_bw; o->resamp_rate = fe->resamp_rate;
	o->fft_passes = fe->plan.P;
	for(int i = 0; i < fe->plan.P; i++) o->fft_len[i] = 1 << fe->plan.lgL[i];
	return 0;
}

static int process_pending(hfdl_b200_frontend *fe, bool all) {
	const auto &g = fe->g;
	int done = 0;
	for(;;) {
		long long pending = fe->fed - fe->blocks_done * (long long)g.input_size;
		int nb = (int)(pending / g.input_size);
		if(nb < 1 || (!all && nb < fe->Bmax)) break;
		if(nb > fe->Bmax) nb = fe->Bmax;
		RawSource src;
		src.base = fe->d_ring; src.ring_len = fe->ring_len; src.ring_origin = 0; src.block_stride = g.input_size; src.sfmt = fe->sfmt;
		src.pos0 = fe->blocks_done * (long long)g.input_size - g.overlap_length;
		if(fe->h2d_pending) { if(cudaStreamWaitEvent(fe->stream, fe->ev_h2d, 0) != cudaSuccess) return -1; fe->h2d_pending = false; }     // the samples of this batch have landed
		if(run_batch(fe, src, nb)) return -1;
		done += nb;
	}
	return done;
}

static int32_t push_samples_impl(hfdl_b200_frontend_t *fe, const void *samples, int64_t nsamples, bool wait_copy) {
	if(!fe || (!samples && nsamples > 0) || nsamples < 0) return -1;
	HFDL_API(fe, -1);
	const auto &g = fe->g;
	const unsigned char *p = (const unsigned char *)samples;
	int blocks = 0;
	while(nsamples > 0) {
		// oldest sample still needed: start of the overlap of the next unprocessed block.  One batch worth of ring below that
		// is left alone: the channeliser stage of the newest queued batch may still be reading it
		long long keep_from = fe->blocks_done * (long long)g.input_size - g.overlap_length;
		long long space = fe->ring_len - (long long)fe->Bmax * g.input_size - (fe->fed - keep_from);
		if(space <= 0) {
			int r = process_pending(fe, true);
			if(r < 0) return -1;
			blocks += r;
			continue;
		}
		long long n = std::min<long long>(nsamples, space);
		long long idx = fe->fed % fe->ring_len;
		long long first = std::min(n, fe->ring_len - idx);
		// what this copy overwrites was read by batches up to the one before the newest: their channeliser stage must be done
		if(fe->batch_seq >= 2) CK(cudaStreamWaitEvent(fe->st_h2d, fe->ev_front[(fe->batch_seq - 2) % HFDL_NSETS], 0));
		CK(cudaMemcpyAsync((unsigned char *)fe->d_ring + idx * fe->bps, p, (size_t)(first * fe->bps), cudaMemcpyHostToDevice, fe->st_h2d));
		if(n > first)
			CK(cudaMemcpyAsync(fe->d_ring, p + first * fe->bps, (size_t)((n - first) * fe->bps), cudaMemcpyHostToDevice, fe->st_h2d));
		fe->fed += n; p += n * fe->bps; nsamples -= n;
		CK(cudaEventRecord(fe->ev_h2d, fe->st_h2d));
		fe->h2d_pending = true;
		int r = process_pending(fe, false);
		if(r < 0) return -1;
		blocks += r;
	}
	// the caller may reuse its buffer when this returns: wait for the last H2D copy (not for the processing)
	if(wait_copy && fe->ev_h2d && blocks >= 0) CK(cudaEventSynchronize(fe->ev_h2d));
	return blocks;
}

int32_t hfdl_b200_push_samples(hfdl_b200_frontend_t *fe, const void *samples, int64_t nsamples) { return push_samples_impl(fe, samples, nsamples, true); }
int32_t hfdl_b200_push_samples_nowait(hfdl_b200_frontend_t *fe, const void *samples, int64_t nsamples) { return push_samples_impl(fe, samples, nsamples, false); }
int32_t hfdl_b200_wait_host_buffer(hfdl_b200_frontend_t *fe) {
	HFDL_API(fe, -1);
	if(fe->ev_h2d) CK(cudaEventSynchronize(fe->ev_h2d));
	return 0;
}

// Multi-GPU: dst (another device) takes the samples src has received and dst has not, ring to ring over NVLink
// (cudaMemcpyPeerAsync with peer access enabled); both frontends have the same geometry, so the rings are congruent.
int32_t hfdl_b200_push_peer(hfdl_b200_frontend_t *dst, hfdl_b200_frontend_t *src) {
	if(!dst || !src || dst == src) return -1;
	// lock order by address: two threads pushing in opposite directions cannot deadlock
	std::unique_lock<std::recursive_mutex> l1(dst < src ? dst->mtx : src->mtx), l2(dst < src ? src->mtx : dst->mtx);
	if(dst->failed || src->failed || dst->ring_len != src->ring_len || dst->sfmt != src->sfmt || dst->g.input_size != src->g.input_size) return -1;
	int prev = -1;
	cudaGetDevice(&prev);
	struct Restore { int d; ~Restore() { if(d >= 0) cudaSetDevice(d); } } restore{ prev };
	CK(cudaSetDevice(dst->cfg.device));
	if(!dst->peer_enabled) {
		int can = 0;
		if(dst->cfg.device != src->cfg.device && cudaDeviceCanAccessPeer(&can, dst->cfg.device, src->cfg.device) == cudaSuccess && can) {
			cudaError_t e = cudaDeviceEnablePeerAccess(src->cfg.device, 0);
			if(e != cudaSuccess) cudaGetLastError();            // already enabled by someone else is fine
		}
		dst->peer_enabled = true;
	}
	int blocks = 0;
	while(dst->fed < src->fed) {
		// same bookkeeping as hfdl_b200_push_samples, the source being the peer's ring
		long long keep_from = dst->blocks_done * (long long)dst->g.input_size - dst->g.overlap_length;
		long long space = dst->ring_len - (long long)dst->Bmax * dst->g.input_size - (dst->fed - keep_from);
		if(space <= 0) {
			int r = process_pending(dst, true);
			if(r < 0) return -1;
			blocks += r;
			continue;
		}
		if(src->fed - dst->fed > src->ring_len) { fprintf(stderr, "hfdl_b200_push_peer: the peer's ring has been overwritten\n"); return -1; }
		long long n = std::min<long long>(src->fed - dst->fed, space);
		long long idx = dst->fed % dst->ring_len;
		long long first = std::min(n, dst->ring_len - idx);
		CK(cudaStreamWaitEvent(dst->stream, src->ev_h2d, 0));     // the peer's H2D copy of these samples
		CK(cudaMemcpyPeerAsync((unsigned char *)dst->d_ring + idx * dst->bps, dst->cfg.device, (const unsigned char *)src->d_ring + idx * src->bps, src->cfg.device, (size_t)(first * dst->bps), dst->stream));
		if(n > first)
			CK(cudaMemcpyPeerAsync(dst->d_ring, dst->cfg.device, src->d_ring, src->cfg.device, (size_t)((n - first) * dst->bps), dst->stream));
		dst->fed += n;
		CK(cudaEventRecord(dst->ev_h2d, dst->stream));
		// the peer must not overwrite this part of its ring before the copy has read it
		CK(cudaStreamWaitEvent(src->st_h2d, dst->ev_h2d, 0));
		int r = process_pending(dst, false);
		if(r < 0) return -1;
		blocks += r;
	}
	return blocks;
}

int32_t hfdl_b200_flush(hfdl_b200_frontend_t *fe) {
	HFDL_API(fe, -1);
	int r = process_pending(fe, true);
	if(r < 0 || drain(fe)) return -1;
	return r;
}

int32_t hfdl_b200_wait_input(hfdl_b200_frontend_t *fe, int32_t keep) {
	HFDL_API(fe, -1);
	if(keep < 0) keep = 0;
	if(keep > HFDL_NSETS - 1) keep = HFDL_NSETS - 1;
	const long long last = fe->batch_seq - 1 - keep;          // newest batch that must have read its input
	if(last < 0) return 0;
	CK(cudaEventSynchronize(fe->ev_front[last % HFDL_NSETS]));
	return 0;
}

int32_t hfdl_b200_submit(hfdl_b200_frontend_t *fe) {
	HFDL_API(fe, -1);
	return process_pending(fe, true);
}

int32_t hfdl_b200_poll(hfdl_b200_frontend_t *fe) {
	HFDL_API(fe, -1);
	for(int k = 0; k < HFDL_NSETS; k++) {                    // oldest batch first: PDU order per channel is kept
		const int q = (int)((fe->batch_seq + k) % HFDL_NSETS);
		if(!fe->flight[q].busy) continue;
		if(collect_batch(fe, q, false)) return -1;
		if(fe->flight[q].busy) break;
	}
	return (int32_t)fe->pduq.size();
}

int32_t hfdl_b200_busy(hfdl_b200_frontend_t *fe) {
	HFDL_API(fe, -1);
	for(int q = 0; q < HFDL_NSETS; q++) if(fe->flight[q].busy) return 1;
	return 0;
}

int32_t hfdl_b200_process_device(hfdl_b200_frontend_t *fe, const void *d_samples, int64_t ring_samples, int64_t start_sample, int32_t nblocks) {
	if(!fe || !d_samples || ring_samples < fe->g.fft_size || nblocks < 0) return -1;
	HFDL_API(fe, -1);
	const auto &g = fe->g;
	int done = 0;
	while(done < nblocks) {
		int nb = std::min(fe->Bmax, nblocks - done);
		RawSource src;
		src.base = d_samples; src.ring_len = ring_samples; src.ring_origin = 0; src.block_stride = g.input_size; src.sfmt = fe->sfmt;
		src.pos0 = start_sample + (long long)done * g.input_size - g.overlap_length;
		if(run_batch(fe, src, nb)) return -1;
		done += nb;
	}
	// keep the host-fed bookkeeping consistent if the two paths are mixed
	fe->fed = fe->blocks_done * (long long)g.input_size;
	return done;
}

// ---- sharded spectrum (multi-GPU): every rank transforms a share of the overlap-save blocks for ALL channels of the job
// and demodulates its own channels for ALL blocks; in between, the pass-band slices change hands (all-to-all over NVLink)
int32_t hfdl_b200_set_exchange(hfdl_b200_frontend_t *fe, const int32_t *all_freqs_hz, int32_t n_all, int32_t nranks) {
	if(!fe || !all_freqs_hz || nranks < 1 || nranks > HFDL_MAX_RANKS || n_all < nranks || n_all % nranks != 0) return -1;
	HFDL_API(fe, -1);
	if(n_all / nranks != fe->C) { fprintf(stderr, "hfdl_b200_set_exchange: this frontend has %d channels, the job gives every rank %d\n", fe->C, n_all / nranks); return -1; }
	if(drain(fe)) return -1;
	std::vector<int> ob((size_t)n_all);
	for(int j = 0; j < n_all; j++) {
		if(abs(fe->cfg.centerfreq_hz - all_freqs_hz[j]) >= fe->cfg.sample_rate / 2) return -1;
		ob[(size_t)j] = hfdl_design::channel_geom(fe->g, fe->cfg.sample_rate, fe->cfg.centerfreq_hz, all_freqs_hz[j]).offsetbin;
	}
	if(fe->d_all_offsetbin) { cudaFree(fe->d_all_offsetbin); fe->d_all_offsetbin = nullptr; }
	CK(cudaMalloc((void **)&fe->d_all_offsetbin, sizeof(int) * (size_t)n_all));
	CK(cudaMemcpy(fe->d_all_offsetbin, ob.data(), sizeof(int) * (size_t)n_all, cudaMemcpyHostToDevice));
	if(fe->d_mask && build_spec_mask(fe, ob)) return -1;          // the last FFT pass now stores what ANY rank's channels read
	fe->xr_ranks = nranks; fe->xr_nall = n_all;
	if(fe->loop_auto && nranks > 1) {
#ifndef HFDL_CUSIM
		fe->loop_layout = LK_ROLE2;
#endif
		fe->loop_smem_pad_to = (size_t)216 * 1024;
	}
	return 0;
}

static int32_t spectrum_slices_impl(hfdl_b200_frontend_t *fe, const void *d_samples, int64_t first_block, int32_t nblocks, const SliceDst &dst, void *cuda_stream) {
	const auto &g = fe->g;
	cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : fe->stream;
	// the buffer holds stream positions [first_block * input_size - overlap, (first_block + nblocks) * input_size): as a "ring"
	// that the windows never wrap in
	RawSource src0;
	src0.base = d_samples; src0.ring_len = (long long)g.overlap_length + (long long)nblocks * g.input_size;
	src0.pos0 = first_block * (long long)g.input_size - g.overlap_length;
	src0.ring_origin = ((-src0.pos0) % src0.ring_len + src0.ring_len) % src0.ring_len;
	src0.block_stride = g.input_size; src0.sfmt = fe->sfmt;
	ProfRec pr;
	for(int b0 = 0; b0 < nblocks; b0 += fe->Bsub) {
		const int nsb = std::min(fe->Bsub, nblocks - b0);
		RawSource src = src0;
		src.pos0 = src0.pos0 + (long long)b0 * src0.block_stride;
		cf *spec = fe->plan.natural ? fe->d_spec + (long long)b0 * g.fft_size : nullptr;
		cf *work = fe->plan.natural ? fe->d_work : fe->d_work + (long long)b0 * g.fft_size;
		if(run_fft(fe, fe->fft, fe->plan, src, work, spec, nsb, st, fe->d_mask)) return -1;
	}
	prof_begin2(fe, KC_PACK, pr, st);
	HFDL_LAUNCH(slice_pack, dim3((unsigned)fe->xr_nall, (unsigned)nblocks), dim3(256), 0, st,
		fe->plan.natural ? fe->d_spec : fe->d_work, fe->plan, g.fft_inv_size, fe->d_all_offsetbin, fe->xr_ranks, dst);
	prof_end2(fe, pr, st);
	fe->launches++;
	CK(cudaGetLastError());
	return nblocks;
}

int32_t hfdl_b200_spectrum_slices(hfdl_b200_frontend_t *fe, const void *d_samples, int64_t first_block, int32_t nblocks, void *d_send, void *cuda_stream) {
	if(!fe || !d_samples || !d_send || nblocks < 1 || first_block < 0) return -1;
	HFDL_API(fe, -1);
	if(fe->xr_ranks < 1 || nblocks > fe->Bmax) return -1;
	SliceDst dst;
	memset(&dst, 0, sizeof(dst));
	const long long part = (long long)nblocks * (fe->xr_nall / fe->xr_ranks) * fe->g.fft_inv_size;
	for(int q = 0; q < fe->xr_ranks; q++) dst.base[q] = (cf *)d_send + (long long)q * part;
	dst.blk0 = 0;
	return spectrum_slices_impl(fe, d_samples, first_block, nblocks, dst, cuda_stream);
}

int32_t hfdl_b200_spectrum_slices_to(hfdl_b200_frontend_t *fe, const void *d_samples, int64_t first_block, int32_t nblocks,
		void *const *d_recv_of_rank, int32_t batch_block0, void *cuda_stream) {
	if(!fe || !d_samples || !d_recv_of_rank || nblocks < 1 || first_block < 0 || batch_block0 < 0) return -1;
	HFDL_API(fe, -1);
	if(fe->xr_ranks < 1 || nblocks > fe->Bmax) return -1;
	SliceDst dst;
	memset(&dst, 0, sizeof(dst));
	for(int q = 0; q < fe->xr_ranks; q++) { if(!d_recv_of_rank[q]) return -1; dst.base[q] = (cf *)d_recv_of_rank[q]; }
	dst.blk0 = batch_block0;
	return spectrum_slices_impl(fe, d_samples, first_block, nblocks, dst, cuda_stream);
}

int32_t hfdl_b200_process_slices(hfdl_b200_frontend_t *fe, const void *d_slices, int32_t nblocks, void *cuda_stream) {
	if(!fe || !d_slices || nblocks < 1) return -1;
	HFDL_API(fe, -1);
	if(nblocks > fe->Bmax) return -1;
	// whatever the caller's stream has queued so far (the exchange that fills d_slices) comes before the channeliser stage
	CK(cudaEventRecord(fe->ev_ext, cuda_stream ? (cudaStream_t)cuda_stream : (cudaStream_t)0));
	RawSource none;
	memset(&none, 0, sizeof(none));
	if(run_batch(fe, none, nblocks, (const cf *)d_slices)) return -1;
	fe->fed = fe->blocks_done * (long long)fe->g.input_size;
	return nblocks;
}

int64_t hfdl_b200_slice_elems(const hfdl_b200_frontend_t *fe) { return fe ? (int64_t)fe->g.fft_inv_size : -1; }

int32_t hfdl_b200_sync(hfdl_b200_frontend_t *fe) {
	HFDL_API(fe, -1);
	return drain(fe);
}

int32_t hfdl_b200_pdu_count(hfdl_b200_frontend_t *fe) {
	if(!fe) return -1;
	std::lock_guard<std::recursive_mutex> lk(fe->mtx);
	return (int32_t)fe->pduq.size();
}

int32_t hfdl_b200_pop_pdu(hfdl_b200_frontend_t *fe, hfdl_b200_pdu_t *pdu) {
	if(!fe || !pdu) return -1;
	std::lock_guard<std::recursive_mutex> lk(fe->mtx);
	if(fe->pduq.empty()) return 0;
	*pdu = fe->pduq.front();
	fe->pduq.pop_front();
	return 1;
}

int32_t hfdl_b200_pop_pdus(hfdl_b200_frontend_t *fe, hfdl_b200_pdu_t *pdus, int32_t max) {
	if(!fe || !pdus || max < 0) return -1;
	std::lock_guard<std::recursive_mutex> lk(fe->mtx);
	int32_t n = 0;
	while(n < max && !fe->pduq.empty()) { pdus[n++] = fe->pduq.front(); fe->pduq.pop_front(); }
	return n;
}

// Snapshot of one channel's demodulator state.  It does NOT drain the pipeline (the reference's stats thread reads
// c->noise_floor while the channel thread runs, hfdl.c:1093): the copy goes through its own stream and sees the state
// as the last finished loop launch left it; after hfdl_b200_flush / _sync it is the state after all pushed samples.
static int read_state(hfdl_b200_frontend *fe, int ch, DemodState *S) {
	if(ch < 0 || ch >= fe->C) return -1;
	CK(cudaMemcpyAsync(S, fe->d_state + ch, sizeof(DemodState), cudaMemcpyDeviceToHost, fe->st_stats));
	CK(cudaStreamSynchronize(fe->st_stats));
	return 0;
}

int32_t hfdl_b200_channel_noise_floor(hfdl_b200_frontend_t *fe, int32_t channel, float *level) {
	DemodState S;
	if(!level) return -1;
	HFDL_API(fe, -1);
	if(read_state(fe, channel, &S)) return -1;
	*level = S.noise_floor;
	return 0;
}

int32_t hfdl_b200_channel_stats(hfdl_b200_frontend_t *fe, int32_t channel, int32_t out[4]) {
	DemodState S;
	if(!out) return -1;
	HFDL_API(fe, -1);
	if(read_state(fe, channel, &S)) return -1;
	out[0] = S.st_a1; out[1] = S.st_a2; out[2] = S.st_m1; out[3] = S.st_frames;
	return 0;
}

int32_t hfdl_b200_channel_counters(hfdl_b200_frontend_t *fe, int32_t channel, hfdl_b200_counters_t *out) {
	DemodState S;
	if(!out) return -1;
	HFDL_API(fe, -1);
	if(read_state(fe, channel, &S)) return -1;
	memset(out, 0, sizeof(*out));
	out->freq = fe->freqs[(size_t)channel];
	out->A1_found = S.st_a1; out->A2_found = S.st_a2; out->M1_found = S.st_m1; out->M1_not_found = S.st_m1_fail;
	out->noise_floor = S.noise_floor;
	const hfdl_b200_frontend::ChanTally &t = fe->tally[(size_t)channel];
	out->frames_processed = t.processed; out->frames_good = t.good; out->frames_bad_fcs = t.bad_fcs; out->frames_too_short = t.too_short;
	out->frames_air2gnd = t.air2gnd; out->frames_gnd2air = t.gnd2air;
	out->lpdus_processed = t.lpdus_processed; out->lpdus_good = t.lpdus_good; out->lpdus_bad_fcs = t.lpdus_bad; out->lpdus_too_short = t.lpdus_short;
	return 0;
}

void hfdl_b200_print_summary(hfdl_b200_frontend_t *fe) {
	if(!fe) return;
	ApiGuard guard_(fe);
	if(fe->d_dbg) {
		std::vector<long long> v((size_t)fe->C * 32);
		drain(fe);
		cudaMemcpy(v.data(), fe->d_dbg, sizeof(long long) * v.size(), cudaMemcpyDeviceToHost);
		for(int c = 0; c < fe->C && c < 4; c++) {
			const long long *q = &v[(size_t)c * 32];
			fprintf(stderr, "loop_kernel ch%d cycles: timing warp %lld (blocked %lld; polls: ring full %lld, loader %lld; outputs %lld, of them generic %lld; fast-loop entries %lld, left at loader limit %lld, at ring limit %lld)  demod warp %lld (waiting %lld)\n",
				c, q[0], q[1], q[4], q[5], q[6], q[8], q[7], q[9], q[10], q[2], q[3]);
			const char *mn[8] = { "bits", "train", "-", "skip", "A1", "data-bpsk", "data-psk4", "data-psk8" };
			for(int m = 0; m < 8; m++) if(q[12 + 2 * m + 1] > 0)
				fprintf(stderr, "    demod run %-9s: %lld symbols, %.0f cycles/symbol (waiting excluded)\n", mn[m], q[12 + 2 * m + 1], (double)q[12 + 2 * m] / (double)q[12 + 2 * m + 1]);
		}
	}
	long long t[4] = { 0, 0, 0, 0 };
	for(int c = 0; c < fe->C; c++) { int32_t s[4]; if(hfdl_b200_channel_stats(fe, c, s) == 0) for(int i = 0; i < 4; i++) t[i] += s[i]; }
	fprintf(stderr, "A1_found:\t\t%lld\nA2_found:\t\t%lld\nM1_found:\t\t%lld\nframes:\t\t\t%lld\n", t[0], t[1], t[2], t[3]);
}

int32_t hfdl_b200_timer_start(hfdl_b200_frontend_t *fe) {
	HFDL_API(fe, -1);
	if(drain(fe)) return -1;
	CK(cudaEventRecord(fe->ev0, fe->stream));
	return 0;
}
int32_t hfdl_b200_timer_stop(hfdl_b200_frontend_t *fe, float *ms) {
	if(!ms) return -1;
	HFDL_API(fe, -1);
	if(drain(fe)) return -1;                      // every batch of the timed region has finished on all four streams
	CK(cudaEventRecord(fe->ev1, fe->stream));
	CK(cudaEventSynchronize(fe->ev1));
	CK(cudaEventElapsedTime(ms, fe->ev0, fe->ev1));
	return 0;
}
int32_t hfdl_b200_profile_enable(hfdl_b200_frontend_t *fe, int32_t on) {
	HFDL_API(fe, -1);
	fe->profiling = on != 0;
	return 0;
}
int32_t hfdl_b200_profile_read(hfdl_b200_frontend_t *fe, int32_t max, char names[][32], float *ms, int32_t *launches) {
	HFDL_API(fe, -1);
	if(drain(fe)) return -1;
	for(auto &r : fe->prof) {
		float t = 0;
		if(cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) { fe->prof_ms[r.cls] += t; fe->prof_n[r.cls]++; }
		cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
	}
	fe->prof.clear();
	int n = std::min<int>(max, KC_COUNT);
	for(int i = 0; i < n; i++) {
		strncpy(names[i], kc_names[i], 31); names[i][31] = 0;
		ms[i] = fe->prof_ms[i]; launches[i] = fe->prof_n[i];
		fe->prof_ms[i] = 0; fe->prof_n[i] = 0;
	}
	return n;
}
int64_t hfdl_b200_kernel_launches(hfdl_b200_frontend_t *fe) { return fe ? fe->launches : -1; }
int64_t hfdl_b200_result_bytes_per_batch(hfdl_b200_frontend_t *fe) { return fe ? (int64_t)(sizeof(int) + sizeof(PduRec) * (size_t)fe->max_frames) : -1; }

int64_t hfdl_b200_read_checkpoint(hfdl_b200_frontend_t *fe, int32_t what, int32_t index, void *dst, int64_t max) {
	if(max < 0) return -1;
	HFDL_API(fe, -1);
	if(drain(fe)) return -1;
	const auto &g = fe->g;
	long long avail = 0;
	const cf *srcp = nullptr;
	switch(what) {
	case HFDL_B200_CP_SPECTRUM: {
		if(index < 0) index = fe->last_nblocks - 1;       // -1: last block processed
		if(index < 0 || index >= fe->last_nblocks) return -1;
		if(fe->d_mask) return -1;          // pruned spectrum: only the channels' granules exist (create with a capture channel for the full one)
		avail = g.fft_size;
		HFDL_LAUNCH(fft_gather_bins, dim3((unsigned)((g.fft_size + 255) / 256)), dim3(256), 0, fe->stream, fe->plan.natural ? fe->d_spec : fe->d_work, fe->plan, index, 0, g.fft_size, fe->d_tmp);
		CK(cudaGetLastError());
		CK(cudaStreamSynchronize(fe->stream));
		srcp = fe->d_tmp;
		break; }
	case HFDL_B200_CP_DDC:
		if(index < 0 || index >= fe->C) return -1;
		avail = (long long)fe->last_nblocks * fe->out_per_block;
		srcp = fe->d_bb + (long long)index * fe->bb_stride + HFDL_RS_HIST;
		break;
	case HFDL_B200_CP_CHAN:
		if(index < 0 || index >= fe->C) return -1;
		avail = fe->last_nout;
		srcp = fe->d_rs[(fe->batch_seq + HFDL_NSETS - 1) % HFDL_NSETS] + (long long)index * fe->rs_stride;      // set of the last batch
		break;
	case HFDL_B200_CP_AGC: case HFDL_B200_CP_MF: case HFDL_B200_CP_EQ: {
		if(fe->cfg.capture_channel < 0) return -1;
		int cnt[2];
		CK(cudaMemcpy(cnt, fe->d_cap_cnt, sizeof(cnt), cudaMemcpyDeviceToHost));
		avail = std::min<long long>(what == HFDL_B200_CP_EQ ? cnt[1] : fe->cap_n, fe->cfg.capture_max);
		srcp = what == HFDL_B200_CP_AGC ? fe->d_cap_agc : (what == HFDL_B200_CP_MF ? fe->d_cap_mf : fe->d_cap_eq);
		break; }
	case HFDL_B200_CP_TAPSLICE:
		if(index < 0 || index >= fe->C) return -1;
		avail = g.fft_inv_size;
		srcp = fe->d_tapslice + (long long)index * g.fft_inv_size;
		break;
	default: return -1;
	}
	long long n = std::min<long long>(avail, max);
	if(dst && n > 0) CK(cudaMemcpy(dst, srcp, sizeof(cf) * (size_t)n, cudaMemcpyDeviceToHost));
	return avail;
}

// ---------------- stage entry points ----------------
}  // extern "C"

namespace {
// Device allocations, stream, FFT engine and tables of a stage entry point: released on EVERY return path (the CK macro
// returns early on a CUDA error).
struct Scratch {
	std::vector<void *> dev;
	cudaStream_t st = nullptr;
	FftEngine eng; bool eng_up = false;
	DemodTables *T = nullptr;
	template <typename P> cudaError_t alloc(P **p, size_t bytes) {
		cudaError_t e = cudaMalloc((void **)p, bytes);
		if(e == cudaSuccess) dev.push_back((void *)*p);
		return e;
	}
	~Scratch() {
		for(void *p : dev) cudaFree(p);
		if(st) cudaStreamDestroy(st);
		if(eng_up) eng.destroy();
		delete T;
	}
};
}  // namespace

extern "C" {

int32_t hfdl_b200_fft_forward(int32_t device, const void *in, void *outp, int32_t n, int32_t batch) {
	if(!in || !outp || n < 2 || (n & (n - 1)) || batch < 1 || hfdl_b200_device_count() < 1) return -1;
	CK(cudaSetDevice(device));
	Scratch S;
	FftEngine &eng = S.eng;
	S.eng_up = true;
	if(eng.init()) return -1;
	FftPlan pl = make_plan(n);
	cf *d_in = nullptr, *d_work = nullptr, *d_spec = nullptr, *d_out = nullptr;
	size_t bytes = sizeof(cf) * (size_t)n * (size_t)batch;
	CK(S.alloc(&d_in, bytes)); CK(S.alloc(&d_work, bytes)); CK(S.alloc(&d_out, sizeof(cf) * (size_t)n));
	if(pl.natural) CK(S.alloc(&d_spec, bytes));
	CK(cudaMemcpy(d_in, in, bytes, cudaMemcpyHostToDevice));
	RawSource src;
	src.base = d_in; src.ring_len = (long long)n * batch; src.pos0 = 0; src.ring_origin = 0; src.block_stride = n; src.sfmt = HFDL_SFMT_CF32;
	CK(cudaStreamCreateWithFlags(&S.st, cudaStreamNonBlocking));
	cudaStream_t st = S.st;
	int rc = run_fft(nullptr, eng, pl, src, d_work, d_spec, batch, st);
	for(int b = 0; b < batch && rc == 0; b++) {
		HFDL_LAUNCH(fft_gather_bins, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, pl.natural ? d_spec : d_work, pl, b, 0, n, d_out);
		if(cudaStreamSynchronize(st) != cudaSuccess) { rc = -1; break; }
		if(cudaMemcpy((cf *)outp + (size_t)b * n, d_out, sizeof(cf) * (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = -1;
	}
	if(cudaGetLastError() != cudaSuccess) rc = -1;
	return rc;
}

static int fec_run(int device, const void *symbols, const uint8_t *vin, int nframes, int M1, uint32_t bitmask, int nbits,
		uint8_t *pdu_out, int stride_out, uint8_t *soft_out, int32_t *crc_out) {
	if(nframes < 1 || hfdl_b200_device_count() < 1) return -1;
	CK(cudaSetDevice(device));
	CK(cudaFuncSetAttribute(fec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HFDL_FEC_SMEM));
	Scratch S;
	DemodTables *T = S.T = new DemodTables();
	hfdl_design::demod_tables(*T);
	DemodTables *d_tab = nullptr; FrameRec *d_fr = nullptr; int *d_n = nullptr; PduRec *d_p = nullptr; cf *d_sym = nullptr;
	unsigned char *d_soft = nullptr, *d_vin = nullptr;
	CK(S.alloc(&d_tab, sizeof(DemodTables)));
	CK(cudaMemcpy(d_tab, T, sizeof(DemodTables), cudaMemcpyHostToDevice));
	std::vector<FrameRec> fr((size_t)nframes);
	for(int q = 0; q < nframes; q++) {
		memset(&fr[(size_t)q], 0, sizeof(FrameRec));
		fr[(size_t)q].channel = q / HFDL_FRAME_SLOTS_MIN; fr[(size_t)q].slot = q % HFDL_FRAME_SLOTS_MIN; fr[(size_t)q].M1 = M1; fr[(size_t)q].bitmask = bitmask;
	}
	CK(S.alloc(&d_fr, sizeof(FrameRec) * (size_t)nframes));
	CK(cudaMemcpy(d_fr, fr.data(), sizeof(FrameRec) * (size_t)nframes, cudaMemcpyHostToDevice));
	CK(S.alloc(&d_n, sizeof(int)));
	CK(cudaMemcpy(d_n, &nframes, sizeof(int), cudaMemcpyHostToDevice));
	CK(S.alloc(&d_p, sizeof(PduRec) * (size_t)nframes));
	int nsym = T->mode_segments[M1] * 30;
	if(symbols) {
		CK(S.alloc(&d_sym, sizeof(cf) * (size_t)nframes * HFDL_DATA_SYMS_MAX));
		for(int q = 0; q < nframes; q++)
			CK(cudaMemcpy(d_sym + (size_t)q * HFDL_DATA_SYMS_MAX, (const cf *)symbols + (size_t)q * nsym, sizeof(cf) * (size_t)nsym, cudaMemcpyHostToDevice));
	}
	if(vin) {
		CK(S.alloc(&d_vin, (size_t)nframes * 2 * nbits));
		CK(cudaMemcpy(d_vin, vin, (size_t)nframes * 2 * nbits, cudaMemcpyHostToDevice));
	}
	if(soft_out) CK(S.alloc(&d_soft, (size_t)nframes * HFDL_FEC_VIN_MAX));
	FecArgs a;
	a.frames = d_fr; a.nframes = d_n; a.max_frames = nframes; a.datasym = d_sym; a.nslots = HFDL_FRAME_SLOTS_MIN; a.tab = d_tab; a.pdus = d_p; a.soft_out = d_soft;
	a.vin_direct = d_vin; a.vin_nbits = nbits;
	HFDL_LAUNCH(fec_kernel, dim3((unsigned)nframes), dim3(32), HFDL_FEC_SMEM, 0, a);
	CK(cudaGetLastError());
	CK(cudaDeviceSynchronize());
	std::vector<PduRec> out((size_t)nframes);
	CK(cudaMemcpy(out.data(), d_p, sizeof(PduRec) * (size_t)nframes, cudaMemcpyDeviceToHost));
	for(int q = 0; q < nframes; q++) {
		int len = out[(size_t)q].len;
		memcpy(pdu_out + (size_t)q * stride_out, out[(size_t)q].octets, (size_t)std::min(len, stride_out));
		if(crc_out) crc_out[q] = out[(size_t)q].crc_good;
	}
	if(soft_out) CK(cudaMemcpy(soft_out, d_soft, (size_t)nframes * HFDL_FEC_VIN_MAX, cudaMemcpyDeviceToHost));
	return 0;
}

int32_t hfdl_b200_fec_decode(int32_t device, const void *symbols, int32_t nframes, int32_t M1, uint32_t bitmask,
		uint8_t *pdu_out, int32_t stride_out, uint8_t *soft_out, int32_t *crc_good_out) {
	if(!symbols || !pdu_out || M1 < 0 || M1 > 7 || stride_out < hfdl_design::pdu_len(M1)) return -1;
	return fec_run(device, symbols, nullptr, nframes, M1, bitmask, 0, pdu_out, stride_out, soft_out, crc_good_out);
}

int32_t hfdl_b200_pdu_front_parse(int32_t device, const uint8_t *pdus, int32_t stride, const int32_t *lens, int32_t n, hfdl_b200_pdu_t *out) {
	if(!pdus || !lens || !out || n < 1 || stride < 1 || hfdl_b200_device_count() < 1) return -1;
	CK(cudaSetDevice(device));
	std::vector<PduRec> recs((size_t)n);
	for(int i = 0; i < n; i++) {
		memset(&recs[(size_t)i], 0, sizeof(PduRec));
		if(lens[i] < 0 || lens[i] > HFDL_MAX_PDU || lens[i] > stride) return -1;
		recs[(size_t)i].len = lens[i];
		memcpy(recs[(size_t)i].octets, pdus + (size_t)i * stride, (size_t)lens[i]);
	}
	PduRec *d = nullptr;
	CK(cudaMalloc((void **)&d, sizeof(PduRec) * (size_t)n));
	int rc = 0;
	if(cudaMemcpy(d, recs.data(), sizeof(PduRec) * (size_t)n, cudaMemcpyHostToDevice) != cudaSuccess) rc = -1;
	if(rc == 0) {
		HFDL_LAUNCH(front_kernel, dim3((unsigned)n), dim3(32), 0, 0, d, n);
		if(cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) rc = -1;
	}
	if(rc == 0 && cudaMemcpy(recs.data(), d, sizeof(PduRec) * (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = -1;
	cudaFree(d);
	if(rc) return -1;
	for(int i = 0; i < n; i++) {
		const PduRec &r = recs[(size_t)i];
		memset(&out[i], 0, sizeof(out[i]));
		out[i].len = r.len; memcpy(out[i].octets, r.octets, (size_t)r.len);
		out[i].crc_good = r.crc_good; out[i].frame_status = r.frame_status; out[i].direction = r.direction;
		out[i].lpdus_processed = r.lpdus_processed; out[i].lpdus_good = r.lpdus_good; out[i].lpdus_bad_fcs = r.lpdus_bad_fcs;
		out[i].lpdus_too_short = r.lpdus_too_short; out[i].lpdu_good_mask = r.lpdu_good_mask;
	}
	return 0;
}

int32_t hfdl_b200_viterbi27(int32_t device, const uint8_t *syms, int32_t nframes, int32_t nbits, uint8_t *out) {
	if(!syms || !out || nbits < 8 || nbits > 7560) return -1;
	return fec_run(device, nullptr, syms, nframes, 1, 0, nbits, out, (nbits + 7) / 8, nullptr, nullptr);
}

}  // extern "C"
