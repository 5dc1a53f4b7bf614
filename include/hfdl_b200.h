/*
 * include/hfdl_b200.h -- C ABI of libhfdl_b200.so, the B200 (sm_100a) replacement for dumphfdl's
 * multichannel front-end hot path:  fft.c (overlap-save forward FFT)  ->  fastddc.c/libcsdr*.c
 * (channeliser)  ->  hfdl.c (demodulator, framer, FEC driver)  ->  libfec/viterbi27_port.c.
 *
 * Plain C, plain pointers and sizes; every entry point returns 0 on success and -1 on error
 * (message on stderr, like the reference).  There is NO CPU implementation behind this ABI:
 * creation fails when no CUDA device is present.
 *
 * Reference interfaces replaced (paths relative to dumphfdl's src/):
 *   hfdl_b200_create            <- csdr_fft_init (main.c:697, fft.h:23), compute_fft_decimation_rate /
 *                                  compute_filter_relative_transition_bw (main.c:699-704), fft_create (fft.h:31),
 *                                  hfdl_init_globals + hfdl_channel_create x C (hfdl.h:10-12, main.c:739-750),
 *                                  block_connect_one2many (main.c:752-755)
 *   hfdl_b200_destroy           <- fft_destroy (fft.h:32), hfdl_channel_destroy (hfdl.h:13), csdr_fft_destroy
 *   hfdl_b200_push_samples      <- the consumer side of fft_thread (fft.c:38-61) fed by
 *                                  complex_samples_produce (input-helpers.c:80-92), plus the sample conversion
 *                                  of input-helpers.c:10-78 when raw CU8/CS16 is pushed
 *   hfdl_b200_process_device    <- same, for a capture already resident in HBM (replay / multi-GPU broadcast)
 *   hfdl_b200_flush             <- the drain-then-exit rule of fft.c:38-47 (only whole blocks are processed)
 *   hfdl_b200_pop_pdu           <- dispatch_pdu -> pdu_decoder_queue_push (hfdl.c:1058-1080, pdu.h:39)
 *   hfdl_b200_push_peer         <- block_connect_one2many's shared spectrum buffer (block.c:90-120): every GPU sees the capture
 *   hfdl_b200_submit / _poll    <- the per-block hand-over of fft_thread / hfdl_decoder_thread (fft.c:57-61, hfdl.c:663-664)
 *                                  without its barrier pair: work is queued, finished batches are picked up later
 *   hfdl_b200_channel_noise_floor <- the c->noise_floor read of noise_floor_stats_thread (hfdl.c:1082-1105)
 *   hfdl_b200_channel_counters  <- the per-channel statsd metrics (doc/STATSD_METRICS.md; hook points hfdl.c:818,828,840,
 *                                  pdu.c:123, mpdu.c:68-102, spdu.c:56-70, lpdu.c:129-150)
 *   pdu.frame_status, lpdus_*   <- the front of pdu_decoder_thread: IS_MPDU split, header-length rule, MPDU / SPDU /
 *                                  LPDU frame check sequences (pdu.c:104, mpdu.c:56-159, spdu.c:55-64, lpdu.c:124-150)
 *   hfdl_b200_print_summary     <- hfdl_print_summary (hfdl.h:14)
 *   hfdl_b200_fft_forward       <- csdr_make_fft_c2c + csdr_fft_execute (fft.h:25-28) [stage entry for parity tests]
 *   hfdl_b200_fec_decode        <- decode_user_data (hfdl.c:993-1056)              [stage entry for parity tests]
 *   hfdl_b200_viterbi27         <- init/update_blk/chainback_viterbi27 (libfec/fec.h:18-21) [stage entry]
 * The block.c-facing wrapper (struct block with a thread routine, hfdl_gpu_frontend_create) is declared in
 * include/hfdl_b200_block.h.
 */
#ifndef HFDL_B200_H
#define HFDL_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct hfdl_b200_frontend hfdl_b200_frontend_t;

/* sample formats: values of the reference's enum sample_format (input-common.h) */
#define HFDL_B200_SFMT_CU8  1
#define HFDL_B200_SFMT_CS16 2
#define HFDL_B200_SFMT_CF32 3

#define HFDL_B200_MAX_PDU_OCTETS 945

typedef struct {
	int32_t sample_rate;
	int32_t centerfreq_hz;
	const int32_t *freqs_hz;      /* dial frequencies, one channel each (carrier = dial + 1440 Hz, hfdl.c:46) */
	int32_t nfreq;
	int32_t sample_format;        /* HFDL_B200_SFMT_* of the samples that will be pushed */
	int32_t device;               /* CUDA device ordinal */
	int32_t max_blocks_per_batch; /* overlap-save blocks processed per launch group (0 = default) */
	int32_t capture_channel;      /* >=0: keep DATADUMPS-style checkpoints of this channel (debug/parity); -1 off */
	int32_t capture_max;          /* samples kept per checkpoint */
} hfdl_b200_config_t;

/* block geometry = fastddc_t of the reference (fastddc.h:8-27) + the resampler rate (hfdl.c:471) */
typedef struct {
	int32_t decimation, pre_decimation, post_decimation;
	int32_t taps_length, overlap_length, fft_size, fft_inv_size, input_size, post_input_size, scrap;
	int32_t out_per_block;        /* post_input_size / post_decimation */
	float   transition_bw, resamp_rate;
	int32_t fft_passes, fft_len[3];
} hfdl_b200_geometry_t;

/* one decoded frame = struct hfdl_pdu_metadata (pdu.h:8-17) + the octet string (hfdl.c:1077-1079) */
typedef struct {
	int32_t version;              /* 1 (hfdl.c:1061) */
	int32_t freq;                 /* channel dial frequency, Hz */
	int32_t bit_rate;             /* hfdl.c:1072-1073 */
	float   freq_err_hz;          /* hfdl.c:812 */
	float   rssi;                 /* 20*log10(signal_level), hfdl.c:1064 */
	float   noise_floor;          /* 20*log10(noise_floor), hfdl.c:1065 */
	char    slot;                 /* 'S' / 'D' */
	int32_t M1;                   /* mode index 0..7 (hfdl.c:81-138) */
	int32_t crc_good;             /* 1 when the MPDU header / SPDU FCS is good (mpdu.c:83, spdu.c:62) */
	int32_t train_bits_bad, train_bits_total;
	uint64_t sample_cnt_a2;       /* 5400 Hz sample clock at A2 detection (stands in for the wall clock of hfdl.c:808) */
	uint64_t sample_cnt_end;      /* ... at the end of the frame */
	double  rx_time_s;            /* stream time of the frame start: (sample_cnt_a2/5400) - (448+254)/1800, hfdl.c:657-660 */
	float   signal_level, noise_floor_lin;
	int32_t len;
	uint8_t octets[HFDL_B200_MAX_PDU_OCTETS + 3];
	/* front parser, computed on the device (what pdu_decoder_thread finds out first, pdu.c:104-123) */
	int32_t frame_status;         /* HFDL_B200_FRAME_*: frames.good / frame.errors.bad_fcs / frame.errors.too_short */
	int32_t direction;            /* 1 air2gnd (downlink MPDU), 0 gnd2air (uplink MPDU or SPDU); valid when frame_status == GOOD */
	int32_t lpdus_processed, lpdus_good, lpdus_bad_fcs, lpdus_too_short;      /* lpdu.c:129-150 */
	uint64_t lpdu_good_mask;      /* bit j: the j-th LPDU of the PDU has a good FCS */
} hfdl_b200_pdu_t;
#define HFDL_B200_FRAME_GOOD 0
#define HFDL_B200_FRAME_BAD_FCS 1
#define HFDL_B200_FRAME_TOO_SHORT 2

/* per-channel counters = the reference's per-channel statsd metrics (doc/STATSD_METRICS.md:21-49) */
typedef struct {
	int32_t freq;
	int64_t A1_found, A2_found, M1_found, M1_not_found;                        /* demod.preamble.* (hfdl.c:818,828,840) */
	int64_t frames_processed, frames_good, frames_bad_fcs, frames_too_short;   /* pdu.c:123, mpdu.c:68-88, spdu.c:56-66 */
	int64_t frames_air2gnd, frames_gnd2air;                                    /* frame.dir.* (mpdu.c:92,100, spdu.c:70) */
	int64_t lpdus_processed, lpdus_good, lpdus_bad_fcs, lpdus_too_short;       /* lpdu.c:129-150 */
	float noise_floor;                                                         /* linear; the gauge is 10*|20log10| (hfdl.c:1093-1099) */
} hfdl_b200_counters_t;

int32_t hfdl_b200_device_count(void);
int32_t hfdl_b200_create(hfdl_b200_frontend_t **out, const hfdl_b200_config_t *cfg);
void    hfdl_b200_destroy(hfdl_b200_frontend_t *fe);
int32_t hfdl_b200_get_geometry(const hfdl_b200_frontend_t *fe, hfdl_b200_geometry_t *g);

/* Host samples in the configured sample format.  Whole batches are queued on the GPU as they fill; the call
 * returns the number of overlap-save blocks queued (>=0) or -1.  Batches are pipelined: when the call returns,
 * the samples have been copied to the device (the caller's buffer is free), the newest batch may still be
 * running and its PDUs appear in the queue with the next push / flush / sync (the reference's fft and channel
 * threads hand over through a barrier per block, block.c:90-120; the order of PDUs per channel is the same). */
int32_t hfdl_b200_push_samples(hfdl_b200_frontend_t *fe, const void *samples, int64_t nsamples);
/* The same without waiting for the last host-to-device copy: the caller's buffer (pinned memory, or the copy is not
 * asynchronous) must stay untouched until hfdl_b200_wait_host_buffer returns.  A producer with two staging buffers
 * fills one while the other is still being copied (complex_samples_produce into a ring does the same for the
 * reference's consumers, input-helpers.c:80-95). */
int32_t hfdl_b200_push_samples_nowait(hfdl_b200_frontend_t *fe, const void *samples, int64_t nsamples);
int32_t hfdl_b200_wait_host_buffer(hfdl_b200_frontend_t *fe);
/* Process every whole block still buffered (a final partial block is dropped, as in fft.c:41-46) and wait
 * until all PDUs are in the queue. */
int32_t hfdl_b200_flush(hfdl_b200_frontend_t *fe);
/* Device-resident cyclic capture: 'd_samples' holds ring_samples samples of the configured format in
 * HBM; processes nblocks blocks starting at stream position start_sample (stream position p lives at
 * ring index p % ring_samples; positions < 0 read as zeros).  Consecutive calls must continue the stream. */
int32_t hfdl_b200_process_device(hfdl_b200_frontend_t *fe, const void *d_samples, int64_t ring_samples,
		int64_t start_sample, int32_t nblocks);
/* hfdl_b200_process_device returns while its batches are still queued: this call blocks until all queued batches but
 * the newest `keep` (0..3) have finished READING the caller's device buffer (the channeliser stage; the demodulator
 * stages keep running), i.e. until the buffers of the older batches may be overwritten.  A caller that rotates over
 * k buffers, one process_device call (of at most max_blocks_per_batch blocks) each, passes keep = k - 1. */
int32_t hfdl_b200_wait_input(hfdl_b200_frontend_t *fe, int32_t keep);
/* Blocks until all queued GPU work of this frontend is complete and PDUs are collected. */
int32_t hfdl_b200_sync(hfdl_b200_frontend_t *fe);
/* Streaming use: queue every whole block buffered so far on the GPU and return at once (no waiting; returns the
 * number of blocks queued), and pick up the PDUs of batches that have finished meanwhile without blocking (returns
 * the number of PDUs waiting in the queue).  hfdl_b200_busy: 1 while a batch is still in flight. */
int32_t hfdl_b200_submit(hfdl_b200_frontend_t *fe);
/* Multi-GPU, one process: 'dst' (a frontend with the same sample rate / format / batch size on ANOTHER device) takes
 * the samples 'src' has been pushed and 'dst' has not seen yet, device ring to device ring over NVLink.  The capture
 * crosses PCIe once; this is the broadcast of the reference's one2many connection (block.c:90-120, fft.c:60-61). */
int32_t hfdl_b200_push_peer(hfdl_b200_frontend_t *dst, hfdl_b200_frontend_t *src);
int32_t hfdl_b200_poll(hfdl_b200_frontend_t *fe);
int32_t hfdl_b200_busy(hfdl_b200_frontend_t *fe);

/* ---- Multi-GPU, sharded spectrum (one process per GPU or one process for all).  The reference hands every block's
 * spectrum to every channel (fft.c:60-61, block.c:90-120: one producer, many consumers); across GPUs the cheap thing to
 * hand over is not the capture but the M-bin pass-band slice each channel reads (fastddc.c:152-167 touches nothing
 * else of the spectrum once the alias fold is restricted to the pass-band): C x M x 8 bytes per block instead of
 * input_size x sample bytes.  So every rank runs the forward FFT of 1/R of a batch's blocks for ALL channels of the
 * job, the slices change hands (all-to-all over NVLink: NCCL, or peer copies in one process), and every rank
 * demodulates ITS channels for all blocks.  FFT work and PCIe ingest are then divided by R instead of replicated.
 *
 * hfdl_b200_set_exchange: declares the channel list of the whole job.  Channel j of the list belongs to rank j % nranks,
 *   where it is local channel j / nranks (the frontend was created with exactly those n_all / nranks frequencies, in
 *   that order).  From now on the last FFT pass stores the spectrum granules ANY listed channel reads.
 * hfdl_b200_spectrum_slices: forward FFT of `nblocks` consecutive overlap-save windows, the first one being block
 *   `first_block` of the stream; d_samples (device, configured sample format) holds the stream positions
 *   [first_block * input_size - overlap_length, (first_block + nblocks) * input_size) (positions < 0 read as zero).
 *   Writes d_send[nranks][nblocks][n_all / nranks][fft_inv_size] complex64: destination rank, block, that rank's local
 *   channel, bins offsetbin - M/2 .. offsetbin + M/2 - 1 in inverse-FFT input order.  Runs on `cuda_stream` (a
 *   cudaStream_t; NULL = the frontend's own front stream), so the caller's exchange can simply be queued behind it.
 * hfdl_b200_process_slices: one batch of `nblocks` blocks from d_slices[nblocks][channels of this frontend][fft_inv_size]
 *   (what the exchange delivers: the d_send parts addressed to this rank, in block order).  Everything queued on
 *   `cuda_stream` so far (the exchange) is waited for on the device; the call itself does not block.  Consecutive
 *   calls continue the stream, exactly like hfdl_b200_process_device; hfdl_b200_wait_input tells when d_slices may be
 *   overwritten. */
int32_t hfdl_b200_set_exchange(hfdl_b200_frontend_t *fe, const int32_t *all_freqs_hz, int32_t n_all, int32_t nranks);
int32_t hfdl_b200_spectrum_slices(hfdl_b200_frontend_t *fe, const void *d_samples, int64_t first_block, int32_t nblocks,
		void *d_send, void *cuda_stream);
/* The same with the exchange fused into the kernel (one process, peer access enabled between the GPUs): the slices of
 * rank q's channels are stored straight into d_recv_of_rank[q] -- rank q's [batch blocks][channels][fft_inv_size] array,
 * peer memory for q != this rank -- at block position batch_block0 + (0 .. nblocks-1).  The stores cross NVLink from
 * inside the pack kernel; no copy follows, an event on cuda_stream tells the receivers when to go on. */
int32_t hfdl_b200_spectrum_slices_to(hfdl_b200_frontend_t *fe, const void *d_samples, int64_t first_block, int32_t nblocks,
		void *const *d_recv_of_rank, int32_t batch_block0, void *cuda_stream);
int32_t hfdl_b200_process_slices(hfdl_b200_frontend_t *fe, const void *d_slices, int32_t nblocks, void *cuda_stream);
int64_t hfdl_b200_slice_elems(const hfdl_b200_frontend_t *fe);

int32_t hfdl_b200_pdu_count(hfdl_b200_frontend_t *fe);
/* returns 1 and fills *pdu when one is available, 0 when the queue is empty */
int32_t hfdl_b200_pop_pdu(hfdl_b200_frontend_t *fe, hfdl_b200_pdu_t *pdu);
/* up to max PDUs at once (oldest first); returns how many were written */
int32_t hfdl_b200_pop_pdus(hfdl_b200_frontend_t *fe, hfdl_b200_pdu_t *pdus, int32_t max);
int32_t hfdl_b200_channel_noise_floor(hfdl_b200_frontend_t *fe, int32_t channel, float *level_linear);
/* counters: A1 found, A2 found, M1 found, frames (hfdl.c:162-179) */
int32_t hfdl_b200_channel_stats(hfdl_b200_frontend_t *fe, int32_t channel, int32_t out[4]);
/* The demodulator counters reflect the batches finished so far, the frame / LPDU counters the PDUs collected so far;
 * none of the three calls waits for queued work (the reference's stats thread does not stop the channels either). */
int32_t hfdl_b200_channel_counters(hfdl_b200_frontend_t *fe, int32_t channel, hfdl_b200_counters_t *out);
void    hfdl_b200_print_summary(hfdl_b200_frontend_t *fe);

/* timing of the device work issued since the previous call (CUDA events on the frontend's stream) */
int32_t hfdl_b200_timer_start(hfdl_b200_frontend_t *fe);
int32_t hfdl_b200_timer_stop(hfdl_b200_frontend_t *fe, float *ms);
/* per-kernel-class device time accumulated while profiling is on: fills names/ms/launches for up to max classes */
int32_t hfdl_b200_profile_enable(hfdl_b200_frontend_t *fe, int32_t on);
int32_t hfdl_b200_profile_read(hfdl_b200_frontend_t *fe, int32_t max, char names[][32], float *ms, int32_t *launches);
int64_t hfdl_b200_kernel_launches(hfdl_b200_frontend_t *fe);
/* bytes copied device -> host per batch for the results (frame counter + PDU records) */
int64_t hfdl_b200_result_bytes_per_batch(hfdl_b200_frontend_t *fe);

/* ---- checkpoints for parity tests (the reference's DATADUMPS taps, hfdl.c:616-655) ---- */
#define HFDL_B200_CP_SPECTRUM 0   /* last batch: forward spectrum of block 'index' (-1 = last), natural FFTW order */
#define HFDL_B200_CP_DDC      1   /* last batch: channel 'index' fastddc_inv_cc output, out_per_block*blocks samples */
#define HFDL_B200_CP_CHAN     2   /* last batch: channel 'index' resampler output (f_chan_out) */
#define HFDL_B200_CP_AGC      3   /* capture channel: AGC output since create (f_agc_out) */
#define HFDL_B200_CP_MF       4   /* capture channel: matched filter output (f_mf_out) */
#define HFDL_B200_CP_EQ       5   /* capture channel: equaliser output per symbol (f_eq_out) */
#define HFDL_B200_CP_TAPSLICE 6   /* channel 'index': M tap-spectrum bins in inverse-FFT input order */
/* copies up to max complex64 values into dst (host); returns the number available or -1 */
int64_t hfdl_b200_read_checkpoint(hfdl_b200_frontend_t *fe, int32_t what, int32_t index, void *dst, int64_t max);

/* ---- stage entry points (host buffers in, host buffers out; run on the device) ---- */
/* forward unnormalised DFT of 'batch' windows of n complex64 each, FFTW_FORWARD sign, natural order out */
int32_t hfdl_b200_fft_forward(int32_t device, const void *in_cf32, void *out_cf32, int32_t n, int32_t batch);
/* decode_user_data for 'nframes' frames of mode M1: symbols [nframes][nsym(M1)] complex64, bitmask per call;
 * pdu_out [nframes][stride_out] octets (stride_out >= pdu length), soft_out optional [nframes][15120] */
int32_t hfdl_b200_fec_decode(int32_t device, const void *symbols_cf32, int32_t nframes, int32_t M1, uint32_t bitmask,
		uint8_t *pdu_out, int32_t stride_out, uint8_t *soft_out, int32_t *crc_good_out);
/* Viterbi only: syms [nframes][2*nbits] soft bytes -> out [nframes][(nbits+7)/8]; nbits must be one of the HFDL sizes */
int32_t hfdl_b200_viterbi27(int32_t device, const uint8_t *syms, int32_t nframes, int32_t nbits, uint8_t *out);
int32_t hfdl_b200_pdu_len(int32_t M1);
/* front parser only: n PDUs of lens[i] octets at pdus + i*stride -> frame_status, direction, lpdus_*, lpdu_good_mask, crc_good */
int32_t hfdl_b200_pdu_front_parse(int32_t device, const uint8_t *pdus, int32_t stride, const int32_t *lens, int32_t n, hfdl_b200_pdu_t *out);

#ifdef __cplusplus
}
#endif
#endif
