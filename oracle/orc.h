/*
 * oracle/orc.h -- CPU ORACLE for the dumphfdl multichannel HFDL front-end hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may call anything in oracle/.
 * The product (dumphfdl_b200/libhfdl_b200.so) never links or calls it.
 *
 * It is a plain-C restatement (no code copied) of the reference's algorithm for the path
 *   fft.c (overlap-save forward FFT)  ->  fastddc.c / libcsdr*.c (channeliser)
 *   ->  hfdl.c (demodulator + framer + FEC driver)  ->  libfec/viterbi27_port.c  ->  crc.c
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference/src).
 *
 * PARITY STATUS
 *   - PINNED against the reference's own sources compiled where they lie into oracle/_ref/ (recipe:
 *     oracle/Makefile): geometry, tap design, all-bin channeliser, shift/decimate (fastddc.c, libcsdr.c,
 *     libcsdr_gpl.c), Viterbi (libfec/viterbi27_port.c), CRC (crc.c) -- tests/test_oracle_ref.py -- and the
 *     WHOLE of hfdl.c (sample loop, Costas loop, sampler, framer FSM, descrambler, deinterleaver,
 *     decode_user_data, dispatch_pdu, statsd hook points) plus block.c / fft.c / input-helpers.c running as
 *     the reference's own threads: every DATADUMPS tap bit-identical, every PDU and its metadata identical --
 *     tests/test_oracle_hfdl_ref.py.  The PDU front (orc_pdu_front_parse, orc_fcs_check) is pinned against the
 *     reference's own pdu.c / mpdu.c / spdu.c / lpdu.c / util.c / crc.c running their real pdu_decoder_thread
 *     (oracle/_ref/libref_front.so) -- tests/test_oracle_front_ref.py.
 *   - NOT PINNED: the inside of the liquid-dsp objects (agc, firfilt, msresamp, symsync, eqlms, modem,
 *     bsequence, msequence; oracle/orc_liquid.c).  liquid-dsp is NOT in /root/reference (un-vendored
 *     dependency jgaeddert/liquid-dsp, any 1.3.0 <= v < 2.0 accepted by src/CMakeLists.txt:71-101), is not
 *     installed here and the reference ships no test vectors: in oracle/_ref/ the reference's hfdl.c is linked
 *     against these same restated objects (ref_shim/liquid_shim.c).  They restate the published liquid-dsp
 *     1.3.2 algorithms; ground truth for them is the HFDL transmitter in orc_tx.c (decoded PDU octets ==
 *     transmitted octets, FCS good, all 8 modes, Es/N0 3..30 dB).  For these objects: "parity unpinned".
 */
#ifndef ORC_H
#define ORC_H
#include <stdint.h>
#include <stddef.h>
#include <complex.h>

#ifdef __cplusplus
#error "oracle is C11"
#endif

typedef float complex cf32;

/* ---------------- protocol constants (hfdl.c:29-46, hfdl.h:6-8) ---------------- */
#define ORC_SYMBOL_RATE 1800
#define ORC_SPS 3
#define ORC_PREKEY_LEN 448
#define ORC_A_LEN 127
#define ORC_M1_LEN 127
#define ORC_M2_LEN 15
#define ORC_T_LEN 15
#define ORC_EQ_LEN 15
#define ORC_DATA_FRAME_LEN 30
#define ORC_SEG_SINGLE 72
#define ORC_SEG_DOUBLE 168
#define ORC_DATA_SYMS_MAX (ORC_SEG_DOUBLE * ORC_DATA_FRAME_LEN)
#define ORC_PREAMBLE_LEN (2 * ORC_A_LEN + ORC_M1_LEN + ORC_M2_LEN + 9 * ORC_T_LEN)
#define ORC_SINGLE_SLOT_FRAME_LEN (ORC_PREKEY_LEN + ORC_PREAMBLE_LEN + ORC_SEG_SINGLE * (ORC_DATA_FRAME_LEN + ORC_T_LEN))
#define ORC_SSB_CARRIER_OFFSET_HZ 1440
#define ORC_TRANSITION_BW_HZ 250
#define ORC_MAX_PDU_OCTETS 945
#define ORC_MF_TAPS 19

/* mode table (hfdl.c:81-138) */
typedef struct { int arity, segments, code_rate, col_shift; } orc_mode_t;
extern const orc_mode_t orc_modes[8];
extern const float orc_mf_taps[ORC_MF_TAPS];        /* hfdl.c:148-154 */
extern const uint8_t orc_A_octets[16];              /* hfdl.c:420-437 */
extern const uint8_t orc_M1_bits[127];              /* hfdl.c:441-447 */
extern const int orc_M_shifts[8];                   /* hfdl.c:449 */
#define ORC_T_WORD 0x9AFu                           /* hfdl.c:181 */

/* ---------------- FFT (stand-in for fftw3f, fft_fftw.c:22-41) ---------------- */
/* unnormalised DFT; dir=+1: FFTW_FORWARD (e^-j), dir=-1: FFTW_BACKWARD (e^+j). n power of 2. */
void orc_fft(const cf32 *in, cf32 *out, int n, int dir);
void orc_fft_set_threads(int nthreads);

/* ---------------- geometry: fastddc_init (fastddc.c:46-80) ---------------- */
typedef struct {
	int32_t pre_decimation, post_decimation;
	int32_t taps_length, taps_min_length, overlap_length;
	int32_t fft_size, fft_inv_size, input_size, post_input_size;
	float pre_shift;
	int32_t startbin, v, offsetbin;
	float post_shift;
	int32_t scrap;
	float dsa_sindelta, dsa_cosdelta, dsa_rate;     /* libcsdr_gpl.c:26-39 */
} orc_ddc_t;
int32_t orc_next_pow2(int32_t x);                                         /* libcsdr.c:35-44 */
int32_t orc_fft_decimation_rate(int32_t sample_rate, int32_t target);     /* libcsdr.c:140-144 */
float   orc_relative_transition_bw(int32_t sample_rate, int32_t bw_hz);   /* libcsdr.c:135-138 */
int     orc_ddc_init(orc_ddc_t *d, float transition_bw, int32_t decimation, float shift_rate);
float   orc_channel_shift_rate(int32_t sample_rate, int32_t centerfreq, int32_t freq); /* hfdl.c:476 */
void    orc_bandpass_taps(cf32 *out, int32_t length, float lowcut, float highcut);     /* libcsdr.c:84-133 */

/* shift/decimate state (libcsdr_gpl.h:35-40) */
typedef struct { int32_t decimation_remain; float starting_phase; int32_t output_size; } orc_dsa_status_t;

/* ---------------- channeliser (fastddc.c:152-252) ---------------- */
#define ORC_FOLD_FULL  0   /* reference-exact: all N bins folded (fastddc.c:123-150) */
#define ORC_FOLD_SLICE 1   /* pass-band slice of M bins (what the GPU computes)      */
typedef struct {
	orc_ddc_t ddc;
	int fold_mode;
	cf32 *taps_fft;        /* FULL: N swapped bins.  SLICE: M bins, index i <-> swapped bin (startbin-M/2+i) mod N */
	cf32 *inv_in, *inv_out;
	orc_dsa_status_t shift_status;
} orc_channelizer_t;
orc_channelizer_t *orc_channelizer_create(int32_t decimation, float transition_bw, float freq_shift, int fold_mode);
void orc_channelizer_destroy(orc_channelizer_t *c);
const cf32 *orc_channelizer_taps(const orc_channelizer_t *c);      /* taps_fft: N (FULL) or M (SLICE) bins */
const orc_ddc_t *orc_channelizer_ddc(const orc_channelizer_t *c);
/* spectrum_swapped: N bins after fft_swap_sides (DC at N/2).  out: >= post_input_size. returns #outputs */
int  orc_channelizer_execute(orc_channelizer_t *c, const cf32 *spectrum_swapped, cf32 *out);
void orc_swap_sides(cf32 *io, int32_t n);                                 /* fastddc.c:102-112 */

/* ---------------- liquid-dsp restatements (parity unpinned, see header comment) ---------------- */
void  orc_firdes_kaiser(int n, float fc, float As, float mu, float *h);
typedef struct orc_resamp orc_resamp_t;
orc_resamp_t *orc_resamp_create(float rate, float As);
void orc_resamp_destroy(orc_resamp_t *q);
void orc_resamp_execute(orc_resamp_t *q, const cf32 *x, int nx, cf32 *y, uint32_t *ny);
/* expose the resampler design so the product can be checked against it */
int   orc_resamp_design(float rate, float As, float *h_out /*npfb*sublen, [filter][tap]*/, int *npfb, int *sublen, uint32_t *step);

/* ---------------- FEC / bits ---------------- */
uint16_t orc_crc16(const uint8_t *data, uint32_t len, uint16_t init);            /* crc.c:4-47 */
int      orc_fcs_check(const uint8_t *buf, uint32_t hdr_len);                    /* pdu.c:68-79 */
/* returns 1 if the PDU's frame check is good by the rules of pdu.c:104 + mpdu.c:56-85 / spdu.c:55-64 */
int      orc_pdu_crc_good(const uint8_t *buf, uint32_t len);
/* front of pdu_decoder_thread (pdu.c:104-123, mpdu.c:56-159, spdu.c:55-70, lpdu.c:129-150):
 * out = { status 0 good / 1 bad_fcs / 2 too_short, direction 1 air2gnd, lpdus processed, good, bad_fcs, too_short } */
void     orc_pdu_front_parse(const uint8_t *buf, uint32_t len, int32_t out[6], uint64_t *good_mask);
/* K=7 r=1/2 Viterbi restated from libfec/viterbi27_port.c: syms 2*nbits soft bytes -> ceil(nbits/8) octets (MSB first) */
void     orc_viterbi27(const uint8_t *syms, int nbits, uint8_t *out);
void     orc_conv_encode27(const uint8_t *bits, int nbits, uint8_t *chips /*2*nbits, values 0/1*/);
uint32_t orc_scrambler_bits(uint8_t *out, int n);                                /* hfdl.c:300-347 */
/* decode_user_data (hfdl.c:993-1056): data symbols -> PDU octets.  returns octet count */
int      orc_decode_user_data(const cf32 *symbols, int M1, uint32_t bitmask, uint8_t *pdu_out, uint8_t *softbits_out /*optional*/);
/* inverse for the transmitter */
int      orc_encode_user_data(const uint8_t *pdu, int M1, cf32 *symbols_out);
int      orc_pdu_len_octets(int M1);

/* ---------------- per-channel demodulator (hfdl.c:468-534, 593-935) ---------------- */
typedef struct {
	int32_t freq;
	int32_t M1;
	int32_t len;
	float freq_err_hz, signal_level, noise_floor;   /* hfdl.c:1061-1067 (levels linear; dB = 20log10) */
	int32_t bit_rate; char slot;
	uint64_t sample_cnt_end;    /* 5400 Hz sample clock when the frame completed */
	uint64_t sample_cnt_a2;     /* ... when A2 was found (replaces wall-clock timestamp hfdl.c:808) */
	int32_t train_bits_bad, train_bits_total;
	int32_t crc_good;
	uint8_t octets[ORC_MAX_PDU_OCTETS + 3];
} orc_pdu_t;

/* capture taps = the reference's DATADUMPS checkpoints (hfdl.c:616-655) */
enum { ORC_CAP_CHAN = 0, ORC_CAP_AGC, ORC_CAP_MF, ORC_CAP_SYMSYNC, ORC_CAP_COSTAS, ORC_CAP_EQ, ORC_CAP_DATASYM, ORC_CAP_DDC, ORC_CAP_COUNT };

typedef struct orc_channel orc_channel_t;
orc_channel_t *orc_channel_create(int32_t sample_rate, int32_t pre_decimation_rate, float transition_bw,
		int32_t centerfreq, int32_t frequency, int fold_mode);
void orc_channel_destroy(orc_channel_t *c);
void orc_channel_set_capture(orc_channel_t *c, uint32_t mask, size_t max_per_tap);
size_t orc_channel_get_capture(orc_channel_t *c, int tap, cf32 *dst, size_t max);
/* one overlap-save block: swapped N-bin spectrum in, PDUs appended to the channel's list */
void orc_channel_process_block(orc_channel_t *c, const cf32 *spectrum_swapped);
/* demod only: feed samples at the post-DDC rate (before msresamp), for stage tests */
void orc_channel_process_baseband(orc_channel_t *c, const cf32 *x, int n);
int  orc_channel_pdu_count(orc_channel_t *c);
int  orc_channel_get_pdu(orc_channel_t *c, int idx, orc_pdu_t *out);
const orc_ddc_t *orc_channel_ddc(orc_channel_t *c);
float orc_channel_resamp_rate(orc_channel_t *c);
void orc_channel_stats(orc_channel_t *c, int32_t *a1, int32_t *a2, int32_t *m1, int32_t *frames);
int32_t orc_channel_m1_not_found(orc_channel_t *c);      /* statsd demod.preamble.errors.M1_not_found, hfdl.c:840 */
float orc_channel_noise_floor(orc_channel_t *c);          /* c->noise_floor as noise_floor_stats_thread reads it, hfdl.c:1093 */

/* ---------------- whole pipeline: fft.c thread + C channel threads ---------------- */
typedef struct orc_pipeline orc_pipeline_t;
enum { ORC_SFMT_CU8 = 1, ORC_SFMT_CS16 = 2, ORC_SFMT_CF32 = 3 };               /* input-common.h sample_format */
orc_pipeline_t *orc_pipeline_create(int32_t sample_rate, int32_t centerfreq, const int32_t *freqs, int32_t nfreq,
		int fold_mode, int nthreads);
void orc_pipeline_destroy(orc_pipeline_t *p);
/* run on the spectrum ring (include/hfdl_b200_ring.h) instead of the reference's barrier pair: `depth` spectra in flight;
 * must be called before the first feed.  orc_pipeline_sync waits until every fed block has been demodulated. */
int  orc_pipeline_use_ring(orc_pipeline_t *p, int depth);
void orc_pipeline_sync(orc_pipeline_t *p);
orc_channel_t *orc_pipeline_channel(orc_pipeline_t *p, int idx);
const orc_ddc_t *orc_pipeline_ddc(orc_pipeline_t *p);
/* raw samples in 'sfmt' (input-helpers.c:10-78 scaling); whole blocks are processed, the rest is kept. returns blocks run */
int  orc_pipeline_feed(orc_pipeline_t *p, const void *raw, int64_t nsamples, int sfmt);
int  orc_pipeline_pdu_count(orc_pipeline_t *p);
int  orc_pipeline_get_pdu(orc_pipeline_t *p, int idx, orc_pdu_t *out);   /* canonical order: (sample_cnt_end, freq) */
/* copy of the last forward spectrum (swapped) for K1 parity */
int  orc_pipeline_last_spectrum(orc_pipeline_t *p, cf32 *dst, int n);
void orc_convert_samples(const void *raw, int64_t n, int sfmt, cf32 *out);     /* input-helpers.c:10-78,108-125 */

/* ---------------- transmitter / synthetic capture generator (no reference counterpart) ---------------- */
typedef struct {
	int32_t freq_hz;          /* dial frequency; carrier = freq + 1440 Hz */
	int32_t M1;               /* mode 0..7 */
	double start_s;           /* start time of the prekey within the capture (s) */
	double cfo_hz;            /* carrier frequency offset */
	double phase0;            /* initial carrier phase (rad) */
	double amplitude;         /* linear amplitude (RMS of the modulated part) */
	int32_t pdu_len;
	uint8_t pdu[ORC_MAX_PDU_OCTETS + 3];
} orc_tx_frame_t;
/* builds a PDU of the mode's size: kind 0 = downlink MPDU with 2 LPDUs, 1 = SPDU, 2 = raw random, 3 = uplink MPDU for two
 * aircraft with 3 LPDUs (one with a broken FCS), 4 = downlink MPDU with a too-short and a truncated LPDU */
int  orc_tx_make_pdu(int M1, int kind, uint64_t seed, uint8_t *out);
/* 3 samples/symbol complex baseband of one frame (prekey..last T), shaped with orc_mf_taps; returns sample count */
int  orc_tx_frame_baseband(const orc_tx_frame_t *f, cf32 *out, int max);
int  orc_tx_frame_symbols(const orc_tx_frame_t *f, cf32 *out, int max);
/* render frames into a wideband capture (adds to 'out').  cyclic!=0: capture is treated as periodic (looped slab) */
void orc_tx_render(cf32 *out, int64_t nsamples, int32_t sample_rate, int32_t centerfreq,
		const orc_tx_frame_t *frames, int nframes, int cyclic, int nthreads);
void orc_tx_render_range(cf32 *out, int64_t first, int64_t count, int64_t nsamples, int32_t sample_rate, int32_t centerfreq,
		const orc_tx_frame_t *frames, int nframes, int cyclic, int nthreads);
void orc_tx_add_noise(cf32 *out, int64_t nsamples, double sigma, uint64_t seed, int nthreads);
void orc_quantize_cs16(const cf32 *in, int64_t n, int16_t *out);
void orc_quantize_cu8(const cf32 *in, int64_t n, uint8_t *out);

#endif
