/*
 * oracle/orc_dsp.c -- ORACLE (test infrastructure only, see orc.h).
 * FFT stand-in, fastddc geometry, tap design, channeliser, shift/decimate.
 * Restated from src/fastddc.c, src/libcsdr.c, src/libcsdr_gpl.c, src/fft_fftw.c (cited per function).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include "orc.h"

/* ====================================================================================
 * FFT: stand-in for fftwf_plan_dft_1d/fftwf_execute (fft_fftw.c:22-41).  Plain unnormalised
 * DFT, power-of-two sizes, Stockham autosort radix-4 (+ one radix-2 stage), twiddles computed
 * in double and rounded once.  Any correct FFT is equivalent to ~1e-6 relative.
 * ==================================================================================== */
typedef struct { int n; cf32 *w; } orc_twid_t;
static orc_twid_t g_twid[32];
static pthread_mutex_t g_twid_lock = PTHREAD_MUTEX_INITIALIZER;
static int g_fft_threads = 1;

void orc_fft_set_threads(int n) { g_fft_threads = n < 1 ? 1 : n; }

static const cf32 *twiddles_for(int n) {
	int lg = 0;
	while((1 << lg) < n) lg++;
	pthread_mutex_lock(&g_twid_lock);
	if(g_twid[lg].w == NULL) {
		cf32 *w = malloc(sizeof(cf32) * (size_t)n);
		for(int i = 0; i < n; i++) {
			double a = -2.0 * M_PI * (double)i / (double)n;
			w[i] = (float)cos(a) + I * (float)sin(a);
		}
		g_twid[lg].n = n;
		g_twid[lg].w = w;
	}
	pthread_mutex_unlock(&g_twid_lock);
	return g_twid[lg].w;
}

/* one radix-4 DIF Stockham stage: n = current length, s = stride; tw stride ts = N/n */
static void stage4(int n, int s, int ts, const cf32 *restrict W, int conj, const cf32 *restrict x, cf32 *restrict y, int p0, int p1) {
	int n1 = n / 4, n2 = n / 2, n3 = n1 + n2;
	for(int p = p0; p < p1; p++) {
		cf32 w1 = W[(size_t)p * ts], w2 = W[(size_t)2 * p * ts], w3 = W[(size_t)3 * p * ts];
		if(conj) { w1 = conjf(w1); w2 = conjf(w2); w3 = conjf(w3); }
		const cf32 *xa = x + (size_t)s * p, *xb = x + (size_t)s * (p + n1), *xc = x + (size_t)s * (p + n2), *xd = x + (size_t)s * (p + n3);
		cf32 *y0 = y + (size_t)s * (4 * p);
		for(int q = 0; q < s; q++) {
			cf32 a = xa[q], b = xb[q], c = xc[q], d = xd[q];
			cf32 apc = a + c, amc = a - c, bpd = b + d, bmd = b - d;
			/* forward: -j*(b-d); inverse: +j*(b-d) */
			cf32 jbmd = conj ? (-cimagf(bmd) + I * crealf(bmd)) : (cimagf(bmd) - I * crealf(bmd));
			y0[q] = apc + bpd;
			y0[q + s] = w1 * (amc + jbmd);
			y0[q + 2 * s] = w2 * (apc - bpd);
			y0[q + 3 * s] = w3 * (amc - jbmd);
		}
	}
}

static void stage2(int n, int s, int ts, const cf32 *restrict W, int conj, const cf32 *restrict x, cf32 *restrict y, int p0, int p1) {
	int m = n / 2;
	for(int p = p0; p < p1; p++) {
		cf32 w = W[(size_t)p * ts];
		if(conj) w = conjf(w);
		const cf32 *xa = x + (size_t)s * p, *xb = x + (size_t)s * (p + m);
		cf32 *y0 = y + (size_t)s * (2 * p);
		for(int q = 0; q < s; q++) {
			cf32 a = xa[q], b = xb[q];
			y0[q] = a + b;
			y0[q + s] = (a - b) * w;
		}
	}
}

struct stage_job { int radix, n, s, ts, conj, p0, p1; const cf32 *W, *x; cf32 *y; };
static void stage_run(struct stage_job *j) {
	if(j->radix == 4) stage4(j->n, j->s, j->ts, j->W, j->conj, j->x, j->y, j->p0, j->p1);
	else stage2(j->n, j->s, j->ts, j->W, j->conj, j->x, j->y, j->p0, j->p1);
}

/* Persistent worker pool for the stages of large transforms (fftw3f_threads keeps its workers too): created once,
 * woken per stage.  One transform uses it at a time; a concurrent caller (the channel threads' small inverse FFTs
 * never get here, n*s < 2^19) simply runs its stage on its own thread. */
#define POOL_MAX 64
static struct {
	pthread_mutex_t busy, m;
	pthread_cond_t cv_go, cv_done;
	pthread_t th[POOL_MAX];
	int nth;                       /* workers started */
	struct stage_job jobs[POOL_MAX + 1];
	int njobs, next, pending;
} P = { PTHREAD_MUTEX_INITIALIZER, PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, {0}, 0, {{0}}, 0, 0, 0 };

static void *pool_worker(void *arg) {
	(void)arg;
	pthread_mutex_lock(&P.m);
	for(;;) {
		while(P.next >= P.njobs) pthread_cond_wait(&P.cv_go, &P.m);
		struct stage_job *j = &P.jobs[P.next++];
		pthread_mutex_unlock(&P.m);
		stage_run(j);
		pthread_mutex_lock(&P.m);
		if(--P.pending == 0) pthread_cond_signal(&P.cv_done);
	}
	return NULL;
}

static void run_stage(int radix, int n, int s, int ts, const cf32 *W, int conj, const cf32 *x, cf32 *y) {
	int np = n / radix;
	int nt = g_fft_threads;
	if(nt > np) nt = np;
	if(nt > POOL_MAX) nt = POOL_MAX;
	if(nt <= 1 || (size_t)n * s < (1u << 19) || pthread_mutex_trylock(&P.busy) != 0) {
		struct stage_job j = { radix, n, s, ts, conj, 0, np, W, x, y };
		stage_run(&j);
		return;
	}
	pthread_mutex_lock(&P.m);
	while(P.nth < nt - 1) {                       /* the calling thread is worker number nt */
		pthread_create(&P.th[P.nth], NULL, pool_worker, NULL);
		pthread_detach(P.th[P.nth]);
		P.nth++;
	}
	for(int t = 0; t < nt; t++)
		P.jobs[t] = (struct stage_job){ radix, n, s, ts, conj, (int)((int64_t)np * t / nt), (int)((int64_t)np * (t + 1) / nt), W, x, y };
	P.njobs = nt; P.next = 0; P.pending = nt;
	pthread_cond_broadcast(&P.cv_go);
	while(P.next < P.njobs) {                     /* the caller works too */
		struct stage_job *j = &P.jobs[P.next++];
		pthread_mutex_unlock(&P.m);
		stage_run(j);
		pthread_mutex_lock(&P.m);
		P.pending--;
	}
	while(P.pending > 0) pthread_cond_wait(&P.cv_done, &P.m);
	pthread_mutex_unlock(&P.m);
	pthread_mutex_unlock(&P.busy);
}

void orc_fft(const cf32 *in, cf32 *out, int N, int dir) {
	if(N == 1) { out[0] = in[0]; return; }
	const cf32 *W = twiddles_for(N);
	int conj = dir < 0;
	cf32 *bufA = malloc(sizeof(cf32) * (size_t)N);
	cf32 *bufB = malloc(sizeof(cf32) * (size_t)N);
	int lg = 0;
	while((1 << lg) < N) lg++;
	int nstages = lg / 2 + (lg & 1);
	const cf32 *src = in;
	if(in == out) {            /* in-place call: work from a copy */
		memcpy(bufB, in, sizeof(cf32) * (size_t)N);
		src = bufB;
	}
	int n = N, s = 1;
	for(int k = 0; k < nstages; k++) {
		int last = (k == nstages - 1);
		cf32 *dst = last ? out : ((src == bufA) ? bufB : bufA);
		int radix = ((lg & 1) && last) ? 2 : 4;
		run_stage(radix, n, s, N / n, W, conj, src, dst);
		src = dst; n /= radix; s *= radix;
	}
	free(bufA);
	free(bufB);
}

/* ====================================================================================
 * geometry
 * ==================================================================================== */
/* libcsdr.c:35-44 : smallest power of two strictly greater than x */
int32_t orc_next_pow2(int32_t x) {
	for(int32_t i = 0; i < 31; i++) {
		int32_t p = (int32_t)1 << i;
		if(x < p) return p;
	}
	return -1;
}

/* libcsdr.c:140-144 */
int32_t orc_fft_decimation_rate(int32_t sample_rate, int32_t target_rate) {
	int32_t r = (int32_t)floorf((float)sample_rate / (float)target_rate);
	return orc_next_pow2(r) / 2;
}

/* libcsdr.c:135-138 */
float orc_relative_transition_bw(int32_t sample_rate, int32_t bw_hz) {
	return (float)bw_hz / (float)sample_rate;
}

/* hfdl.c:476 */
float orc_channel_shift_rate(int32_t sample_rate, int32_t centerfreq, int32_t freq) {
	return (float)(centerfreq - (freq + ORC_SSB_CARRIER_OFFSET_HZ)) / (float)sample_rate;
}

/* libcsdr.c:46-51 : odd tap count for a relative transition bandwidth; the 4.0 is a double */
static int32_t filter_len(float transition_bw) {
	int32_t r = (int32_t)(4.0 / transition_bw);
	if((r % 2) == 0) r++;
	return r;
}

/* fastddc.c:46-80 (float/double/int conversions follow the C expression types of the reference) */
int orc_ddc_init(orc_ddc_t *d, float transition_bw, int32_t decimation, float shift_rate) {
	memset(d, 0, sizeof(*d));
	d->pre_decimation = 1;
	d->post_decimation = decimation;
	/* move factors of two into the frequency-domain decimation while post/2 is an integer != 1 */
	for(;;) {
		float half = (float)d->post_decimation / 2;
		if(!(floorf(half) == half) || d->post_decimation / 2 == 1) break;
		d->post_decimation /= 2;
		d->pre_decimation *= 2;
	}
	d->taps_min_length = filter_len(transition_bw);
	double tl = ceil(d->taps_min_length / (float)d->pre_decimation) * d->pre_decimation;
	d->taps_length = orc_next_pow2((int32_t)tl) + 1;
	d->fft_size = orc_next_pow2(d->taps_length * 4);
	while(d->fft_size < d->pre_decimation) d->fft_size *= 2;
	d->overlap_length = d->taps_length - 1;
	d->input_size = d->fft_size - d->overlap_length;
	d->fft_inv_size = d->fft_size / d->pre_decimation;

	d->v = d->fft_size / d->overlap_length;
	int32_t middlebin = d->fft_size / 2;
	/* int + (int*float*int -> float) -> float -> truncated on assignment */
	d->startbin = (int32_t)(middlebin + middlebin * (-shift_rate) * 2);
	d->startbin = (int32_t)(d->v * round(d->startbin / (float)d->v));
	d->offsetbin = d->startbin - middlebin;
	d->post_shift = (d->pre_decimation) * (shift_rate + ((float)d->offsetbin / d->fft_size));
	d->pre_shift = d->offsetbin / (float)d->fft_size;
	/* decimating_shift_addition_init -> shift_addition_init (libcsdr_gpl.c:26-39) */
	float rate = d->post_shift * d->post_decimation;
	rate *= 2;
	d->dsa_sindelta = (float)sin(rate * M_PI);
	d->dsa_cosdelta = (float)cos(rate * M_PI);
	d->dsa_rate = rate;

	d->scrap = d->overlap_length / d->pre_decimation;
	d->post_input_size = d->fft_inv_size - d->scrap;
	return d->fft_size <= 2;
}

/* ====================================================================================
 * taps: firdes_bandpass_c -> firdes_lowpass_f (Hamming) -> normalize_fir_f   (libcsdr.c:62-68,84-133)
 * ==================================================================================== */
static float hamming_kernel(float rate) {           /* libcsdr.c:62-68 */
	rate = 0.5 + rate / 2;                          /* double expression narrowed to float */
	return 0.54 - 0.46 * cos(2 * M_PI * rate);
}

static void lowpass_taps(float *out, int32_t length, float cutoff_rate) {   /* libcsdr.c:84-99 */
	int32_t middle = length / 2;
	out[middle] = 2 * M_PI * cutoff_rate * hamming_kernel(0);
	for(int32_t i = 1; i <= middle; i++) {
		out[middle - i] = out[middle + i] = (sin(2 * M_PI * cutoff_rate * i) / i) * hamming_kernel((float)i / middle);
	}
	/* normalize_fir_f (libcsdr.c:74-82): float accumulation in index order */
	float sum = 0;
	for(int32_t i = 0; i < length; i++) sum += out[i];
	for(int32_t i = 0; i < length; i++) out[i] = out[i] / sum;
}

void orc_bandpass_taps(cf32 *out, int32_t length, float lowcut, float highcut) {   /* libcsdr.c:101-124 */
	float *real = calloc((size_t)length, sizeof(float));
	lowpass_taps(real, length, (highcut - lowcut) / 2);
	float center = (highcut + lowcut) / 2;
	float phase = 0, sinval, cosval;
	for(int32_t i = 0; i < length; i++) {
		cosval = cos(phase);
		sinval = sin(phase);
		phase += 2 * M_PI * center;
		while(phase > 2 * M_PI) phase -= 2 * M_PI;
		while(phase < 0) phase += 2 * M_PI;
		out[i] = CMPLXF(cosval * real[i], sinval * real[i]);
	}
	free(real);
}

/* fastddc.c:102-112 */
void orc_swap_sides(cf32 *io, int32_t n) {
	int32_t h = n / 2;
	for(int32_t i = 0; i < h; i++) {
		cf32 t = io[i];
		io[i] = io[i + h];
		io[i + h] = t;
	}
}

/* ====================================================================================
 * channeliser: fft_channelizer_create (fastddc.c:217-252), fastddc_inv_cc (fastddc.c:152-215)
 * ==================================================================================== */
orc_channelizer_t *orc_channelizer_create(int32_t decimation, float transition_bw, float freq_shift, int fold_mode) {
	orc_channelizer_t *c = calloc(1, sizeof(*c));
	if(orc_ddc_init(&c->ddc, transition_bw, decimation, freq_shift)) { free(c); return NULL; }
	c->fold_mode = fold_mode;
	int32_t N = c->ddc.fft_size, M = c->ddc.fft_inv_size;
	cf32 *taps = calloc((size_t)N, sizeof(cf32));
	cf32 *tf = calloc((size_t)N, sizeof(cf32));
	float half_bw = 0.5f / decimation;
	orc_bandpass_taps(taps, c->ddc.taps_length, (-freq_shift) - half_bw, (-freq_shift) + half_bw);
	orc_fft(taps, tf, N, +1);
	orc_swap_sides(tf, N);
	free(taps);
	if(fold_mode == ORC_FOLD_FULL) {
		c->taps_fft = tf;
	} else {
		c->taps_fft = calloc((size_t)M, sizeof(cf32));
		for(int32_t i = 0; i < M; i++) {
			int32_t k = ((c->ddc.startbin - M / 2 + i) % N + N) % N;
			c->taps_fft[i] = tf[k];
		}
		free(tf);
	}
	c->inv_in = calloc((size_t)M, sizeof(cf32));
	c->inv_out = calloc((size_t)M, sizeof(cf32));
	return c;
}

const cf32 *orc_channelizer_taps(const orc_channelizer_t *c) { return c->taps_fft; }
const orc_ddc_t *orc_channelizer_ddc(const orc_channelizer_t *c) { return &c->ddc; }

void orc_channelizer_destroy(orc_channelizer_t *c) {
	if(!c) return;
	free(c->taps_fft); free(c->inv_in); free(c->inv_out); free(c);
}

/* fastddc.c:123-150: out[(h+k) mod M] += X[k]*H[k] over all N swapped bins, h=(N-offset+M/2) mod M */
static void fold_full(const cf32 *X, const cf32 *H, int32_t N, cf32 *out, int32_t M, int32_t offset) {
	int32_t head = (N - offset + M / 2) % M;
	memset(out, 0, sizeof(cf32) * (size_t)M);
	int32_t k = 0, o = head;
	/* same summation order as the reference: head, whole blocks, tail */
	for(; o < M; o++, k++) out[o] += H[k] * X[k];
	int32_t whole = N / M - 1;
	for(int32_t b = 0; b < whole; b++)
		for(o = 0; o < M; o++, k++) out[o] += H[k] * X[k];
	for(o = 0; o < head; o++, k++) out[o] += H[k] * X[k];
}

/* pass-band slice: the single alias nearest the channel centre for every output index */
static void fold_slice(const cf32 *X, const cf32 *Hs, int32_t N, cf32 *out, int32_t M, int32_t startbin) {
	for(int32_t i = 0; i < M; i++) {
		int32_t k = ((startbin - M / 2 + i) % N + N) % N;
		out[i] = Hs[i] * X[k];
	}
}

/* libcsdr_gpl.c:41-74 */
static orc_dsa_status_t shift_decimate(const cf32 *in, cf32 *out, int32_t n, const orc_ddc_t *d, orc_dsa_status_t s) {
	float cosphi = cos(s.starting_phase);
	float sinphi = sin(s.starting_phase);
	int32_t i, k = 0;
	int32_t dec = d->post_decimation;
	for(i = s.decimation_remain; i < n; i += dec) {
		float re = crealf(in[i]), im = cimagf(in[i]);
		out[k++] = CMPLXF(cosphi * re - sinphi * im, sinphi * re + cosphi * im);
		float cl = cosphi, sl = sinphi;
		cosphi = cl * d->dsa_cosdelta - sl * d->dsa_sindelta;
		sinphi = sl * d->dsa_cosdelta + cl * d->dsa_sindelta;
	}
	s.decimation_remain = i - n;
	s.starting_phase += d->dsa_rate * M_PI * k;
	s.output_size = k;
	while(s.starting_phase > M_PI) s.starting_phase -= 2 * M_PI;
	while(s.starting_phase < -M_PI) s.starting_phase += 2 * M_PI;
	return s;
}

int orc_channelizer_execute(orc_channelizer_t *c, const cf32 *X, cf32 *out) {
	const orc_ddc_t *d = &c->ddc;
	int32_t N = d->fft_size, M = d->fft_inv_size;
	if(c->fold_mode == ORC_FOLD_FULL) fold_full(X, c->taps_fft, N, c->inv_in, M, d->offsetbin);
	else fold_slice(X, c->taps_fft, N, c->inv_in, M, d->startbin);
	orc_swap_sides(c->inv_in, M);
	orc_fft(c->inv_in, c->inv_out, M, -1);
	/* fastddc.c:193-197: divide by pre_decimation*M (= N) as a complex division by a real */
	float norm = (float)(d->pre_decimation * M);
	for(int32_t i = 0; i < M; i++) c->inv_out[i] /= norm;
	c->shift_status = shift_decimate(c->inv_out + d->scrap, out, d->post_input_size, d, c->shift_status);
	return c->shift_status.output_size;
}

/* ====================================================================================
 * liquid-dsp restatements used by several files (parity unpinned: see orc.h)
 * ==================================================================================== */
static float besseli0(float z) {                    /* liquid: math.bessel.c besseli0f, series in log domain */
	if(z == 0.0f) return 1.0f;
	float y = 0.0f;
	for(int k = 0; k < 32; k++) {
		float t = k * logf(0.5f * z) - lgammaf((float)k + 1.0f);
		y += expf(2 * t);
	}
	return y;
}

static float kaiser_win(int n, int N, float beta, float mu) {   /* liquid 1.3.x: math.windows.c kaiser() */
	float t = (float)n - (float)(N - 1) / 2 + mu;
	float r = 2.0f * t / (float)(N);
	float a = besseli0(beta * sqrtf(1 - r * r));
	float b = besseli0(beta);
	return a / b;
}

static float sincf_(float x) {                      /* liquid: math.c sincf() */
	if(fabsf(x) < 0.01f)
		return cosf(M_PI * x / 2.0f) * cosf(M_PI * x / 4.0f) * cosf(M_PI * x / 8.0f);
	return sinf(M_PI * x) / (M_PI * x);
}

static float kaiser_beta_As(float As) {             /* liquid: filter/src/firdes.c kaiser_beta_As */
	As = fabsf(As);
	if(As > 50.0f) return 0.1102f * (As - 8.7f);
	if(As > 21.0f) return 0.5842f * powf(As - 21, 0.4f) + 0.07886f * (As - 21);
	return 0.0f;
}

void orc_firdes_kaiser(int n, float fc, float As, float mu, float *h) {   /* liquid_firdes_kaiser */
	float beta = kaiser_beta_As(As);
	for(int i = 0; i < n; i++) {
		float t = (float)i - (float)(n - 1) / 2 + mu;
		float h1 = sincf_(2.0f * fc * t);
		float h2 = kaiser_win(i, n, beta, mu);
		h[i] = h1 * h2;
	}
}

/* ------------------------------------------------------------------------------------
 * msresamp_crcf(rate, As) for 0.5 <= rate <= 1 (the only range hfdl.c:471-472 can produce:
 * rate = 5400/(sr/dec) with dec = next_pow2(floor(sr/5400))/2): zero half-band stages and one
 * resamp_crcf(rate, m=7, fc=min(0.515*rate,0.49), As, npfb=256), liquid >= 1.3.2 fixed-point phase.
 * ------------------------------------------------------------------------------------ */
#define RS_M 7
#define RS_NPFB 256
#define RS_SUBLEN (2 * RS_M)
struct orc_resamp {
	float rate;
	uint32_t step, phase;
	float h[RS_NPFB][RS_SUBLEN];   /* h[i][n] = proto[i + n*npfb] */
	cf32 win[RS_SUBLEN];           /* win[0] = newest */
};

int orc_resamp_design(float rate, float As, float *h_out, int *npfb, int *sublen, uint32_t *step) {
	int n = 2 * RS_M * RS_NPFB + 1;
	float *hf = malloc(sizeof(float) * (size_t)n);
	float fc = 0.515f * rate;
	if(fc > 0.49f) fc = 0.49f;
	orc_firdes_kaiser(n, fc / (float)RS_NPFB, As, 0.0f, hf);
	float gain = 0.0f;
	for(int i = 0; i < n; i++) gain += hf[i];
	gain = (float)RS_NPFB / gain;
	for(int i = 0; i < RS_NPFB; i++)
		for(int k = 0; k < RS_SUBLEN; k++)
			h_out[i * RS_SUBLEN + k] = hf[i + k * RS_NPFB] * gain;
	free(hf);
	*npfb = RS_NPFB; *sublen = RS_SUBLEN;
	*step = (uint32_t)roundf((float)(1 << 24) / rate);
	return 0;
}

orc_resamp_t *orc_resamp_create(float rate, float As) {
	if(!(rate >= 0.5f && rate <= 1.0f)) {
		fprintf(stderr, "orc_resamp_create: rate %f outside [0.5,1] not restated\n", rate);
		return NULL;
	}
	orc_resamp_t *q = calloc(1, sizeof(*q));
	int a, b;
	q->rate = rate;
	orc_resamp_design(rate, As, &q->h[0][0], &a, &b, &q->step);
	return q;
}

void orc_resamp_destroy(orc_resamp_t *q) { free(q); }

void orc_resamp_execute(orc_resamp_t *q, const cf32 *x, int nx, cf32 *y, uint32_t *ny) {
	uint32_t n = 0;
	for(int i = 0; i < nx; i++) {
		memmove(q->win + 1, q->win, sizeof(cf32) * (RS_SUBLEN - 1));
		q->win[0] = x[i];
		while(q->phase < (1u << 24)) {
			uint32_t idx = q->phase >> (24 - 8);
			const float *h = q->h[idx];
			cf32 acc = 0;
			/* dotprod over the window oldest..newest with reversed sub-filter == sum h[k]*x[newest-k] */
			for(int k = RS_SUBLEN - 1; k >= 0; k--) acc += h[k] * q->win[k];
			y[n++] = acc;
			q->phase += q->step;
		}
		q->phase -= (1u << 24);
	}
	*ny = n;
}
