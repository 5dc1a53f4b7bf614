/*
 * oracle/orc_liquid.c -- ORACLE (test infrastructure only, see orc.h / orc_liquid.h).
 * liquid-dsp objects used by src/hfdl.c, restated from the published liquid-dsp 1.3.2 algorithms
 * (parity with real liquid-dsp unpinned: the library is an un-vendored dependency, absent here).
 * Call sites in the reference: hfdl.c:472-509 (create), 686-738 (per sample / symbol), 937-966, 968-991, 1013.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "orc_liquid.h"

/* ====================================================================================
 * modem (liquid modem_psk.c / modem_bpsk.c / modem_demod_soft.c), schemes BPSK, PSK4, PSK8
 * ==================================================================================== */
static uint32_t gray_enc(uint32_t s) { return s ^ (s >> 1); }
static uint32_t gray_dec(uint32_t s) { uint32_t m = s >> 1; while(m) { s ^= m; m >>= 1; } return s; }

cf32 orc_psk_point(int m, uint32_t sym) {
	int M = 1 << m;
	float alpha = (float)(M_PI / (float)M);
	float ang = (float)gray_dec(sym) * 2 * alpha;
	return cosf(ang) + I * sinf(ang);
}

uint32_t orc_modem_demod(int m, cf32 x, orc_modem_t *st) {
	uint32_t sym;
	if(m == 1) {
		sym = (crealf(x) > 0) ? 0 : 1;
		st->x_hat = sym ? -1.0f : 1.0f;
	} else {
		int M = 1 << m;
		float alpha = (float)(M_PI / (float)M);
		float d_phi = (float)(M_PI * (1.0f - 1.0f / (float)M));
		float theta = cargf(x);
		theta -= d_phi;
		if(theta < -M_PI) theta += 2 * M_PI;
		uint32_t s = 0;
		float v = theta;
		for(int i = 0; i < m; i++) {
			float ref = (float)(1 << (m - i - 1)) * alpha;
			s <<= 1;
			if(v > 0) { s |= 1; v -= ref; } else { v += ref; }
		}
		sym = gray_enc(s);
		st->x_hat = orc_psk_point(m, sym);
	}
	st->r = x;
	return sym;
}
float orc_modem_phase_error(const orc_modem_t *st) { return cimagf(st->r * conjf(st->x_hat)); }

void orc_modem_demod_soft(int m, cf32 x, orc_modem_t *st, uint8_t *soft) {
	if(m == 1) {
		float gamma = 4.0f;
		float LLR = -2.0f * crealf(x) * gamma;
		int sb = (int)(LLR * 16 + 127);
		if(sb > 255) sb = 255;
		if(sb < 0) sb = 0;
		soft[0] = (uint8_t)sb;
		orc_modem_demod(1, x, st);
		return;
	}
	uint32_t s = orc_modem_demod(m, x, st);
	if(m == 2) {   /* LIQUID_MODEM_PSK4: no soft table (built for m>=3 only) -> hard bits 0/255, MSB first */
		for(int i = 0; i < m; i++) soft[i] = ((s >> (m - i - 1)) & 1) ? 255 : 0;
		return;
	}
	/* PSK8: nearest-neighbour table, p = 2 (the two adjacent constellation points) */
	int M = 1 << m;
	float gamma = 1.2f * M;
	float dmin0[3], dmin1[3];
	for(int k = 0; k < m; k++) dmin0[k] = dmin1[k] = 4.0f;
	cf32 e = x - st->x_hat;
	float d = crealf(e * conjf(e));
	for(int k = 0; k < m; k++) {
		if((s >> (m - k - 1)) & 1) dmin1[k] = d; else dmin0[k] = d;
	}
	uint32_t g = gray_dec(s);
	for(int i = 0; i < 2; i++) {
		uint32_t nb = gray_enc((g + (i == 0 ? 1 : (uint32_t)(M - 1))) % (uint32_t)M);
		cf32 xh = orc_psk_point(m, nb);
		cf32 ee = x - xh;
		d = crealf(ee * conjf(ee));
		for(int k = 0; k < m; k++) {
			if((nb >> (m - k - 1)) & 1) { if(d < dmin1[k]) dmin1[k] = d; }
			else { if(d < dmin0[k]) dmin0[k] = d; }
		}
	}
	for(int k = 0; k < m; k++) {
		int sb = (int)(((dmin0[k] - dmin1[k]) * gamma) * 16 + 127);
		if(sb > 255) sb = 255;
		if(sb < 0) sb = 0;
		soft[k] = (uint8_t)sb;
	}
}

/* ====================================================================================
 * agc_crcf (agc.c): y = x*g; y2' = (1-a) y2' + a |y|^2; if(y2' > 1e-6) g *= exp(-0.5 a ln y2'); g <= 1e6
 * ==================================================================================== */
void orc_agc_init(orc_agc_t *q, float bandwidth) { q->g = 1.0f; q->y2 = 1.0f; q->alpha = bandwidth; }
cf32 orc_agc_execute(orc_agc_t *q, cf32 x) {
	cf32 r = x * q->g;
	float y2 = crealf(r * conjf(r));
	q->y2 = (1.0 - q->alpha) * q->y2 + q->alpha * y2;
	if(q->y2 > 1e-6f) q->g *= expf(-0.5f * q->alpha * logf(q->y2));
	if(q->g > 1e6f) q->g = 1e6f;
	return r;
}

/* ====================================================================================
 * firfilt_crcf
 * ==================================================================================== */
void orc_firfilt_init(orc_firfilt_t *q, const float *h, int n) {
	memset(q, 0, sizeof(*q));
	q->n = n > ORC_FIRFILT_MAX ? ORC_FIRFILT_MAX : n;
	memcpy(q->h, h, sizeof(float) * (size_t)q->n);
}
void orc_firfilt_push(orc_firfilt_t *q, cf32 x) {
	memmove(q->win + 1, q->win, sizeof(cf32) * (size_t)(q->n - 1));
	q->win[0] = x;
}
cf32 orc_firfilt_execute(const orc_firfilt_t *q) {
	cf32 s = 0;
	for(int k = q->n - 1; k >= 0; k--) s += q->h[k] * q->win[k];      /* oldest first */
	return s;
}

/* ====================================================================================
 * symsync_crcf (symsync.c), Kaiser prototype
 * ==================================================================================== */
void orc_symsync_reset(orc_symsync_t *q) {             /* SYMSYNC(_reset): clears the mf window only */
	memset(q->win_mf, 0, sizeof(q->win_mf));
	q->rate = (float)q->k / (float)q->k_out;
	q->del = q->rate;
	q->b = 0; q->bf = 0; q->tau = 0; q->q = 0; q->q_hat = 0; q->decim_counter = 0;
	q->v[0] = q->v[1] = q->v[2] = 0;
}

void orc_symsync_set_lf_bw(orc_symsync_t *q, float bt) {
	float alpha = 1.000f - bt, beta = 0.220f * bt, a = 0.500f, b = 0.495f;
	float B0 = beta, A0 = 1.0f - a * alpha, A1 = -b * alpha, A2 = 0;
	q->b0 = B0 / A0; q->a1 = A1 / A0; q->a2 = A2 / A0;
	q->rate_adjustment = 0.5 * bt;
}

void orc_symsync_set_output_rate(orc_symsync_t *q, uint32_t k_out) {
	q->k_out = k_out;
	q->rate = (float)q->k / (float)q->k_out;
	q->del = q->rate;
}

void orc_symsync_init_kaiser(orc_symsync_t *q) {
	memset(q, 0, sizeof(*q));
	enum { HL = 2 * ORC_SS_NPFB * ORC_SS_K * ORC_SS_M + 1 };
	float Hf[HL], H[HL], dH[HL];
	float fc = 0.75f, As = 40.0f;
	orc_firdes_kaiser(HL, fc / (float)(ORC_SS_K * ORC_SS_NPFB), As, 0.0f, Hf);
	for(int i = 0; i < HL; i++) H[i] = Hf[i] * 2.0f * fc;
	float hdh_max = 0;
	for(int i = 0; i < HL; i++) {
		if(i == 0) dH[i] = H[i + 1] - H[HL - 1];
		else if(i == HL - 1) dH[i] = H[0] - H[i - 1];
		else dH[i] = H[i + 1] - H[i - 1];
		if(fabsf(H[i] * dH[i]) > hdh_max || i == 0) hdh_max = fabsf(H[i] * dH[i]);
	}
	for(int i = 0; i < HL; i++) dH[i] *= 0.06f / hdh_max;
	for(int f = 0; f < ORC_SS_NPFB; f++)
		for(int n = 0; n < ORC_SS_SUB; n++) {
			q->mf[f][n] = H[f + n * ORC_SS_NPFB];
			q->dmf[f][n] = dH[f + n * ORC_SS_NPFB];
		}
	q->k = ORC_SS_K;
	q->k_out = 1;
	orc_symsync_reset(q);
	orc_symsync_set_lf_bw(q, 0.01f);
}

static cf32 pfb_exec(const float *h, const cf32 *win) {
	cf32 acc = 0;
	for(int n = ORC_SS_SUB - 1; n >= 0; n--) acc += h[n] * win[n];   /* oldest first */
	return acc;
}

int orc_symsync_step(orc_symsync_t *q, cf32 x, cf32 *y) {
	memmove(q->win_mf + 1, q->win_mf, sizeof(cf32) * (ORC_SS_SUB - 1));
	q->win_mf[0] = x;
	memmove(q->win_dmf + 1, q->win_dmf, sizeof(cf32) * (ORC_SS_SUB - 1));
	q->win_dmf[0] = x;
	int n = 0;
	while(q->b < ORC_SS_NPFB) {
		cf32 mf = pfb_exec(q->mf[q->b], q->win_mf);
		y[n] = mf / (float)q->k;
		if(q->decim_counter == q->k_out) {
			q->decim_counter = 0;
			cf32 dmf = pfb_exec(q->dmf[q->b], q->win_dmf);
			/* advance_internal_loop */
			q->q = crealf(conjf(mf) * dmf);
			if(q->q > 1.0f) q->q = 1.0f; else if(q->q < -1.0f) q->q = -1.0f;
			q->v[2] = q->v[1]; q->v[1] = q->v[0];
			q->v[0] = q->q - q->a1 * q->v[1] - q->a2 * q->v[2];
			q->q_hat = q->b0 * q->v[0];
			q->rate += q->rate_adjustment * q->q_hat;
			q->del = q->rate + q->q_hat;
		}
		q->decim_counter++;
		q->tau += q->del;
		q->bf = q->tau * (float)ORC_SS_NPFB;
		q->b = (int)roundf(q->bf);
		n++;
	}
	q->tau -= 1.0f;
	q->bf -= (float)ORC_SS_NPFB;
	q->b -= ORC_SS_NPFB;
	return n;
}

/* ====================================================================================
 * eqlms_cccf (eqlms.c)
 * ==================================================================================== */
void orc_eqlms_reset(orc_eqlms_t *q) {
	memcpy(q->w, q->h0, sizeof(q->w));
	memset(q->win, 0, sizeof(q->win));
	memset(q->x2, 0, sizeof(q->x2));
	q->count = 0; q->buf_full = 0; q->x2_sum = 0;
}
void orc_eqlms_init_lowpass(orc_eqlms_t *q, float fc) {
	float h[ORC_EQ_LEN];
	orc_firdes_kaiser(ORC_EQ_LEN, fc, 40.0f, 0.0f, h);
	for(int i = 0; i < ORC_EQ_LEN; i++) q->h0[i] = conjf((cf32)(h[ORC_EQ_LEN - 1 - i] * 2 * fc));
	q->mu = 0.5f;
	orc_eqlms_reset(q);
}
void orc_eqlms_push(orc_eqlms_t *q, cf32 x) {
	memmove(q->win, q->win + 1, sizeof(cf32) * (ORC_EQ_LEN - 1));
	q->win[ORC_EQ_LEN - 1] = x;
	float x2n = crealf(x * conjf(x));
	float x20 = q->x2[0];
	memmove(q->x2, q->x2 + 1, sizeof(float) * (ORC_EQ_LEN - 1));
	q->x2[ORC_EQ_LEN - 1] = x2n;
	q->x2_sum = q->x2_sum + x2n - x20;
	q->count++;
}
cf32 orc_eqlms_execute(const orc_eqlms_t *q) {
	cf32 y = 0;
	for(int i = 0; i < ORC_EQ_LEN; i++) y += conjf(q->w[i]) * q->win[i];
	return y;
}
void orc_eqlms_step(orc_eqlms_t *q, cf32 d, cf32 d_hat) {
	if(!q->buf_full) {
		if(q->count < ORC_EQ_LEN) return;
		q->buf_full = 1;
	}
	cf32 alpha = d - d_hat;
	for(int i = 0; i < ORC_EQ_LEN; i++) q->w[i] = q->w[i] + q->mu * conjf(alpha) * q->win[i] / q->x2_sum;
}

/* ====================================================================================
 * msequence
 * ==================================================================================== */
void orc_msequence_init(orc_msequence_t *q, uint32_t m, uint32_t genpoly, uint32_t a, int convention) {
	q->m = m; q->a = a; q->convention = convention;
	q->g = convention == 0 ? genpoly >> 1 : genpoly;
	q->v = a;
}
void orc_msequence_reset(orc_msequence_t *q) { q->v = q->a; }
uint32_t orc_msequence_advance(orc_msequence_t *q) {
	uint32_t b = (uint32_t)__builtin_parity(q->v & q->g);
	if(q->convention == 0) q->v = ((q->v << 1) | b) & ((1u << q->m) - 1u);
	else q->v = (q->v >> 1) | (b << (q->m - 1));
	return b;
}
