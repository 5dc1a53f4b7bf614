/*
 * oracle/orc_tx.c -- ORACLE support (test infrastructure only, see orc.h).
 * HFDL transmitter / synthetic multichannel capture generator.  The reference has only the
 * receive side and no sample captures, so known-answer inputs are produced here as
 * the exact inverse of the receive conventions of src/hfdl.c (framing hfdl.c:29-46,779-891,
 * symbol mapping / scrambler / interleaver / FEC inverse of hfdl.c:993-1056), pulse-shaped with
 * the reference's own matched-filter taps (hfdl.c:148-154).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include "orc.h"

static uint64_t splitmix(uint64_t *s) {
	uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

static void put_fcs(uint8_t *buf, uint32_t len) {      /* FCS little-endian after the covered octets, pdu.c:68-70 */
	uint16_t f = (uint16_t)(orc_crc16(buf, len, 0xFFFFu) ^ 0xFFFFu);
	buf[len] = (uint8_t)(f & 0xff);
	buf[len + 1] = (uint8_t)(f >> 8);
}

int orc_tx_make_pdu(int M1, int kind, uint64_t seed, uint8_t *out) {
	int L = orc_pdu_len_octets(M1);
	const orc_mode_t *p = &orc_modes[M1];
	int nbits = p->segments * ORC_DATA_FRAME_LEN * p->arity / p->code_rate;
	uint64_t s = seed * 0x2545F4914F6CDD1Dull + 12345;
	for(int i = 0; i < L; i++) out[i] = (uint8_t)splitmix(&s);
	if(kind == 0) {
		/* downlink MPDU: bit0 = MPDU (pdu.c:104), bit1 = downlink, bits 2-5 = LPDU count (mpdu.c:56-59) */
		int lpdu_cnt = 2, lpdu_len = 20;
		out[0] = (uint8_t)(0x03 | (lpdu_cnt << 2));
		out[1] &= 0x7f;
		int hdr_len = 6 + lpdu_cnt;
		for(int j = 0; j < lpdu_cnt; j++) out[6 + j] = (uint8_t)(lpdu_len - 1);      /* mpdu.c:142 */
		put_fcs(out, (uint32_t)hdr_len);
		uint8_t *d = out + hdr_len + 2;
		for(int j = 0; j < lpdu_cnt; j++) {
			d[0] = 0x0D;                               /* unnumbered data */
			put_fcs(d, (uint32_t)(lpdu_len - 2));      /* lpdu.c:143-144 */
			d += lpdu_len;
		}
	} else if(kind == 1) {
		out[0] &= (uint8_t)~1u;                        /* SPDU: bit0 clear; FCS over 64 octets (spdu.c:62) */
		put_fcs(out, 64);
	} else if(kind == 3) {
		/* uplink MPDU (mpdu.c:60-75,100-119): two aircraft with 2 and 1 LPDUs; the second LPDU's FCS is broken */
		int cnt[2] = { 2, 1 }, lens[3] = { 12, 9, 15 };
		out[0] = (uint8_t)(0x01 | (1 << 4));           /* MPDU, uplink, aircraft_cnt - 1 = 1 */
		int h = 2, k = 0;
		for(int a = 0; a < 2; a++) {
			out[h] = (uint8_t)(0x10 + a);              /* aircraft id */
			out[h + 1] = (uint8_t)(cnt[a] << 4);
			for(int j = 0; j < cnt[a]; j++) out[h + 2 + j] = (uint8_t)(lens[k++] - 1);
			h += 2 + cnt[a];
		}
		put_fcs(out, (uint32_t)h);
		uint8_t *d = out + h + 2;
		for(int j = 0; j < 3; j++) {
			d[0] = 0x0D;
			put_fcs(d, (uint32_t)(lens[j] - 2));
			if(j == 1) d[lens[j] - 1] ^= 0x5A;
			d += lens[j];
		}
	} else if(kind == 4) {
		/* downlink MPDU with a too-short LPDU (2 octets, lpdu.c:137) and a last LPDU that runs past the PDU (mpdu.c:152) */
		int lpdu_cnt = 3;
		out[0] = (uint8_t)(0x03 | (lpdu_cnt << 2));
		int hdr_len = 6 + lpdu_cnt;
		out[6] = 10 - 1; out[7] = 2 - 1; out[8] = 255;
		put_fcs(out, (uint32_t)hdr_len);
		uint8_t *d = out + hdr_len + 2;
		d[0] = 0x0D;
		put_fcs(d, 8);
	}
	/* last 6 information bits are the convolutional tail (decoder forces them to 0); bits are LSB-first */
	for(int i = nbits - 6; i < L * 8; i++) out[i >> 3] &= (uint8_t)~(1u << (i & 7));
	return L;
}

int orc_tx_frame_symbols(const orc_tx_frame_t *f, cf32 *out, int max) {
	const orc_mode_t *p = &orc_modes[f->M1];
	int n = 0;
#define EMIT(v) do { if(n < max) out[n] = (v); n++; } while(0)
	for(int i = 0; i < ORC_PREKEY_LEN; i++) EMIT(1.0f);
	for(int rep = 0; rep < 2; rep++)
		for(int i = 0; i < ORC_A_LEN; i++) EMIT(((orc_A_octets[i >> 3] >> (7 - (i & 7))) & 1) ? -1.0f : 1.0f);
	for(int j = 0; j < ORC_M1_LEN; j++) EMIT(orc_M1_bits[(orc_M_shifts[f->M1] + j) % ORC_M1_LEN] ? -1.0f : 1.0f);
	for(int j = 0; j < ORC_M2_LEN; j++) EMIT(orc_M1_bits[(orc_M_shifts[f->M1] + j) % ORC_M1_LEN] ? -1.0f : 1.0f);
	float T[ORC_T_LEN];
	for(int i = 0; i < ORC_T_LEN; i++) T[i] = ((ORC_T_WORD >> (ORC_T_LEN - 1 - i)) & 1) ? -1.0f : 1.0f;
	for(int rep = 0; rep < 9; rep++) for(int i = 0; i < ORC_T_LEN; i++) EMIT(T[i]);
	cf32 *data = malloc(sizeof(cf32) * ORC_DATA_SYMS_MAX);
	orc_encode_user_data(f->pdu, f->M1, data);
	for(int seg = 0; seg < p->segments; seg++) {
		for(int i = 0; i < ORC_DATA_FRAME_LEN; i++) EMIT(data[seg * ORC_DATA_FRAME_LEN + i]);
		for(int i = 0; i < ORC_T_LEN; i++) EMIT(T[i]);
	}
	free(data);
#undef EMIT
	return n;
}

int orc_tx_frame_baseband(const orc_tx_frame_t *f, cf32 *out, int max) {
	int maxsym = ORC_PREKEY_LEN + ORC_PREAMBLE_LEN + ORC_SEG_DOUBLE * 45;
	cf32 *sym = malloc(sizeof(cf32) * (size_t)maxsym);
	int ns = orc_tx_frame_symbols(f, sym, maxsym);
	int n = ORC_SPS * ns + ORC_MF_TAPS - 1;
	if(n > max) { free(sym); return n; }
	for(int i = 0; i < n; i++) out[i] = 0;
	for(int k = 0; k < ns; k++)
		for(int t = 0; t < ORC_MF_TAPS; t++) out[ORC_SPS * k + t] += sym[k] * (orc_mf_taps[t] * ORC_SPS);
	free(sym);
	return n;
}

/* ---------------- band-limited interpolation 5400 Hz -> capture rate ---------------- */
#define IK_HALF 8
#define IK_OS 1024
static float ik_tab[IK_OS + 1][2 * IK_HALF];
static int ik_ok;
static double bessel_i0(double x) {
	double s = 1, t = 1;
	for(int k = 1; k < 40; k++) { t *= (x / (2.0 * k)) * (x / (2.0 * k)); s += t; }
	return s;
}
static void ik_init(void) {
	double beta = 8.0, fc = 0.5;
	for(int f = 0; f <= IK_OS; f++) {
		double frac = (double)f / IK_OS;
		for(int j = 0; j < 2 * IK_HALF; j++) {
			double t = (double)(j - (IK_HALF - 1)) - frac;      /* tap position relative to the point */
			double w = 0;
			double r = t / IK_HALF;
			if(fabs(r) < 1.0) w = bessel_i0(beta * sqrt(1 - r * r)) / bessel_i0(beta);
			double sx = (fabs(t) < 1e-12) ? 1.0 : sin(M_PI * 2 * fc * t) / (M_PI * 2 * fc * t);
			ik_tab[f][j] = (float)(2 * fc * sx * w);
		}
	}
	ik_ok = 1;
}

struct render_job {
	cf32 *out; int64_t n0, n1, nsamples; int32_t sr, centerfreq;
	const orc_tx_frame_t *frames; int nframes; int cyclic;
	cf32 **bb; int *bbn;
};

static void *render_worker(void *arg) {
	struct render_job *J = arg;
	double T = (double)J->nsamples / J->sr;
	for(int fi = 0; fi < J->nframes; fi++) {
		const orc_tx_frame_t *f = &J->frames[fi];
		const cf32 *bb = J->bb[fi];
		int bbn = J->bbn[fi];
		double dur = (double)bbn / 5400.0;
		double foff = (double)f->freq_hz + ORC_SSB_CARRIER_OFFSET_HZ + f->cfo_hz - (double)J->centerfreq;
		int nwrap = J->cyclic ? 2 : 1;
		for(int w = 0; w < nwrap; w++) {
			/* frame occupies [start, start+dur) (+ w*T shift backwards for the wrapped part) */
			double start = f->start_s - w * T;
			int64_t a = (int64_t)ceil(start * J->sr), b = (int64_t)floor((start + dur) * J->sr);
			if(a < J->n0) a = J->n0;
			if(b >= J->n1) b = J->n1 - 1;
			for(int64_t n = a; n <= b; n++) {
				double t = (double)n / J->sr;
				double u = (t - start) * 5400.0;
				int i0 = (int)floor(u);
				double frac = u - i0;
				const float *k = ik_tab[(int)(frac * IK_OS + 0.5)];
				float re = 0, im = 0;
				for(int j = 0; j < 2 * IK_HALF; j++) {
					int idx = i0 + j - (IK_HALF - 1);
					if(idx < 0 || idx >= bbn) continue;
					re += k[j] * crealf(bb[idx]);
					im += k[j] * cimagf(bb[idx]);
				}
				/* carrier phase runs on frame-relative time: continuous across the wrap of a cyclic slab */
				double ph = 2 * M_PI * fmod(foff * (t - start), 1.0) + f->phase0;
				float c = (float)cos(ph), s = (float)sin(ph);
				float amp = (float)f->amplitude;
				J->out[n] += CMPLXF(amp * (re * c - im * s), amp * (re * s + im * c));
			}
		}
	}
	return NULL;
}

void orc_tx_render(cf32 *out, int64_t nsamples, int32_t sr, int32_t centerfreq,
		const orc_tx_frame_t *frames, int nframes, int cyclic, int nthreads) {
	orc_tx_render_range(out, 0, nsamples, nsamples, sr, centerfreq, frames, nframes, cyclic, nthreads);
}

/* samples [first, first + count) of a capture of nsamples samples; out[0] is sample 'first' */
void orc_tx_render_range(cf32 *out, int64_t first, int64_t count, int64_t nsamples, int32_t sr, int32_t centerfreq,
		const orc_tx_frame_t *frames, int nframes, int cyclic, int nthreads) {
	if(!ik_ok) ik_init();
	cf32 **bb = malloc(sizeof(cf32 *) * (size_t)nframes);
	int *bbn = malloc(sizeof(int) * (size_t)nframes);
	int maxbb = ORC_SPS * (ORC_PREKEY_LEN + ORC_PREAMBLE_LEN + ORC_SEG_DOUBLE * 45) + ORC_MF_TAPS;
	for(int i = 0; i < nframes; i++) {
		bb[i] = malloc(sizeof(cf32) * (size_t)maxbb);
		bbn[i] = orc_tx_frame_baseband(&frames[i], bb[i], maxbb);
	}
	if(nthreads < 1) nthreads = 1;
	if(nthreads > 64) nthreads = 64;
	pthread_t th[64];
	struct render_job jobs[64];
	for(int t = 0; t < nthreads; t++) {
		jobs[t] = (struct render_job){ out - first, first + count * t / nthreads, first + count * (t + 1) / nthreads, nsamples, sr, centerfreq,
			frames, nframes, cyclic, bb, bbn };
		pthread_create(&th[t], NULL, render_worker, &jobs[t]);
	}
	for(int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
	for(int i = 0; i < nframes; i++) free(bb[i]);
	free(bb); free(bbn);
}

/* ---------------- AWGN: deterministic per 65536-sample chunk, independent of thread count ---------------- */
struct noise_job { cf32 *out; int64_t c0, c1, nsamples; double sigma; uint64_t seed; };
static void *noise_worker(void *arg) {
	struct noise_job *J = arg;
	for(int64_t c = J->c0; c < J->c1; c++) {
		uint64_t s = J->seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(c + 1));
		int64_t a = c * 65536, b = a + 65536;
		if(b > J->nsamples) b = J->nsamples;
		for(int64_t n = a; n < b; n++) {
			double u1 = ((splitmix(&s) >> 11) + 1.0) * (1.0 / 9007199254740993.0);
			double u2 = (splitmix(&s) >> 11) * (1.0 / 9007199254740992.0);
			double r = J->sigma * sqrt(-2.0 * log(u1));
			J->out[n] += CMPLXF((float)(r * cos(2 * M_PI * u2)), (float)(r * sin(2 * M_PI * u2)));
		}
	}
	return NULL;
}
/* sigma = standard deviation per real component */
void orc_tx_add_noise(cf32 *out, int64_t nsamples, double sigma, uint64_t seed, int nthreads) {
	if(sigma <= 0) return;
	int64_t nch = (nsamples + 65535) / 65536;
	if(nthreads < 1) nthreads = 1;
	if(nthreads > 64) nthreads = 64;
	pthread_t th[64];
	struct noise_job jobs[64];
	for(int t = 0; t < nthreads; t++) {
		jobs[t] = (struct noise_job){ out, nch * t / nthreads, nch * (t + 1) / nthreads, nsamples, sigma, seed };
		pthread_create(&th[t], NULL, noise_worker, &jobs[t]);
	}
	for(int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

/* inverse of the input scalings of input-helpers.c:108-125 */
void orc_quantize_cs16(const cf32 *in, int64_t n, int16_t *out) {
	for(int64_t i = 0; i < n; i++) {
		float re = roundf(crealf(in[i]) * 32767.5f), im = roundf(cimagf(in[i]) * 32767.5f);
		re = fminf(fmaxf(re, -32768.f), 32767.f);
		im = fminf(fmaxf(im, -32768.f), 32767.f);
		out[2 * i] = (int16_t)re; out[2 * i + 1] = (int16_t)im;
	}
}
void orc_quantize_cu8(const cf32 *in, int64_t n, uint8_t *out) {
	for(int64_t i = 0; i < n; i++) {
		float re = roundf(crealf(in[i]) * 127.0f + 63.5f), im = roundf(cimagf(in[i]) * 127.0f + 63.5f);
		re = fminf(fmaxf(re, 0.f), 255.f);
		im = fminf(fmaxf(im, 0.f), 255.f);
		out[2 * i] = (uint8_t)re; out[2 * i + 1] = (uint8_t)im;
	}
}
