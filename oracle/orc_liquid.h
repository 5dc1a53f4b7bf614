/*
 * oracle/orc_liquid.h -- ORACLE (test infrastructure only, see orc.h).
 * Object-form restatements of the liquid-dsp objects src/hfdl.c calls (un-vendored dependency
 * jgaeddert/liquid-dsp, any 1.3.0 <= v < 2.0 accepted by src/CMakeLists.txt:71-101; restated from the
 * published liquid-dsp 1.3.2 algorithms -- parity with real liquid-dsp is UNPINNED, see orc.h).
 * Two users:
 *   orc_hfdl.c                 the oracle's restatement of hfdl.c
 *   ref_shim/liquid_shim.c     the liquid C API the REFERENCE's own hfdl.c / block.c / fft.c are linked
 *                              against when they are compiled in place into oracle/_ref/ (so that the
 *                              reference's framer, descrambler, deinterleaver and FEC driver run unmodified
 *                              on exactly the same object arithmetic as the oracle).
 */
#ifndef ORC_LIQUID_H
#define ORC_LIQUID_H
#include "orc.h"

/* ---- modem (modem_psk.c / modem_bpsk.c / modem_demod_soft.c): arity m = 1 BPSK, 2 PSK4, 3 PSK8 ---- */
typedef struct { cf32 r, x_hat; } orc_modem_t;
cf32     orc_psk_point(int m, uint32_t sym);
uint32_t orc_modem_demod(int m, cf32 x, orc_modem_t *st);
float    orc_modem_phase_error(const orc_modem_t *st);
void     orc_modem_demod_soft(int m, cf32 x, orc_modem_t *st, uint8_t *soft);

/* ---- agc_crcf (agc.c): create -> g = 1, bandwidth, y2 = 1 ---- */
typedef struct { float g, alpha, y2; } orc_agc_t;
void orc_agc_init(orc_agc_t *q, float bandwidth);
cf32 orc_agc_execute(orc_agc_t *q, cf32 x);
static inline float orc_agc_signal_level(const orc_agc_t *q) { return 1.0f / q->g; }

/* ---- firfilt_crcf: y[n] = sum h[k] x[n-k], scale 1 ---- */
#define ORC_FIRFILT_MAX 64
typedef struct { int n; float h[ORC_FIRFILT_MAX]; cf32 win[ORC_FIRFILT_MAX]; /* win[0] newest */ } orc_firfilt_t;
void orc_firfilt_init(orc_firfilt_t *q, const float *h, int n);
void orc_firfilt_push(orc_firfilt_t *q, cf32 x);
cf32 orc_firfilt_execute(const orc_firfilt_t *q);

/* ---- symsync_crcf_create_kaiser(k=3, m=3, beta, M=16), hfdl.c:503-505 ---- */
#define ORC_SS_NPFB 16
#define ORC_SS_K 3
#define ORC_SS_M 3
#define ORC_SS_SUB 18                      /* (2*npfb*k*m+1)/npfb */
typedef struct {
	float mf[ORC_SS_NPFB][ORC_SS_SUB], dmf[ORC_SS_NPFB][ORC_SS_SUB];   /* [filter][n] = h[filter + n*npfb] */
	cf32 win_mf[ORC_SS_SUB], win_dmf[ORC_SS_SUB];                      /* [0] = newest */
	uint32_t k, k_out, decim_counter;
	float rate, del, tau, bf, q, q_hat;
	int b;
	float b0, a1, a2;                                  /* normalised loop-filter SOS */
	float v[3];
	float rate_adjustment;
} orc_symsync_t;
void orc_symsync_init_kaiser(orc_symsync_t *q);        /* create_kaiser(3, 3, *, 16): k_out = 1, lf_bw 0.01 */
void orc_symsync_set_lf_bw(orc_symsync_t *q, float bt);
void orc_symsync_set_output_rate(orc_symsync_t *q, uint32_t k_out);
void orc_symsync_reset(orc_symsync_t *q);
int  orc_symsync_step(orc_symsync_t *q, cf32 x, cf32 *y);

/* ---- eqlms_cccf_create_lowpass(15, 0.45), hfdl.c:495-496 ---- */
typedef struct {
	cf32 h0[ORC_EQ_LEN], w[ORC_EQ_LEN], win[ORC_EQ_LEN];   /* win[0] = oldest */
	float x2[ORC_EQ_LEN];                                   /* delay line, x2[0] = oldest */
	float x2_sum, mu;
	uint32_t count; int buf_full;
} orc_eqlms_t;
void orc_eqlms_init_lowpass(orc_eqlms_t *q, float fc);  /* mu = 0.5 as liquid's create(); hfdl.c sets 0.1 */
void orc_eqlms_reset(orc_eqlms_t *q);
void orc_eqlms_push(orc_eqlms_t *q, cf32 x);
cf32 orc_eqlms_execute(const orc_eqlms_t *q);
void orc_eqlms_step(orc_eqlms_t *q, cf32 d, cf32 d_hat);

/* ---- msequence (sequence/src/msequence.c).  convention 0: liquid < 1.6 (state shifts left, g = genpoly >> 1);
 *      convention 1: liquid >= 1.6 (state shifts right, new bit enters at bit m-1) -- hfdl.c:333-345 picks the
 *      arguments by library version so that both emit the same scrambler sequence ---- */
typedef struct { uint32_t m, g, a, v; int convention; } orc_msequence_t;
void     orc_msequence_init(orc_msequence_t *q, uint32_t m, uint32_t genpoly, uint32_t a, int convention);
void     orc_msequence_reset(orc_msequence_t *q);
uint32_t orc_msequence_advance(orc_msequence_t *q);

#endif
