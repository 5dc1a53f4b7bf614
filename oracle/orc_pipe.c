/*
 * oracle/orc_pipe.c -- ORACLE (test infrastructure only, see orc.h).
 * Whole-path driver: input conversion (input-helpers.c:10-78), overlap-save framing + forward FFT
 * + swap (fft.c:34-61), then every channel on the shared spectrum (hfdl.c:662-675).  Threading
 * shape mirrors the reference: the FFT is multi-threaded (fftw3f_threads, fft_fftw.c:8-14) and
 * channels run concurrently between two synchronisation points per block (fft.c:60-61).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include "orc.h"
#include "../include/hfdl_b200_ring.h"      /* the spectrum ring runtime under test (product host code, dumphfdl_b200/csrc/spectrum_ring.c) */

struct orc_pipeline {
	int32_t sample_rate, centerfreq, nch;
	int fold_mode, nthreads;
	orc_ddc_t ddc;                 /* fft_create (fft.c:70-86): fastddc_init with shift 0 */
	orc_channel_t **ch;
	cf32 *window;                  /* N samples: [overlap | input_size] (fft.c:49-54) */
	cf32 *spectrum;
	cf32 *pending; int64_t npending, cap_pending;
	int64_t blocks_done;
	orc_pdu_t *sorted; int nsorted;
	/* optional: spectrum ring instead of the reference's barrier pair: FFT of block k+1 beside the channels of block k */
	hfdl_spectrum_ring_t *ring; int nworkers; pthread_t *workers; struct ring_worker *wctx;
};
struct ring_worker { orc_pipeline_t *p; int id; };

void orc_convert_samples(const void *raw, int64_t n, int sfmt, cf32 *out) {
	if(sfmt == ORC_SFMT_CF32) {                         /* full_scale 1.0 */
		const float *f = raw;
		for(int64_t i = 0; i < n; i++) out[i] = CMPLXF(f[2 * i] / 1.0f, f[2 * i + 1] / 1.0f);
	} else if(sfmt == ORC_SFMT_CS16) {                  /* full_scale SHRT_MAX + 0.5 */
		const int16_t *s = raw;
		const float fs = 32767.5f;
		for(int64_t i = 0; i < n; i++) out[i] = CMPLXF((float)s[2 * i] / fs, (float)s[2 * i + 1] / fs);
	} else {                                            /* CU8: full_scale SCHAR_MAX, shift = fs/2 */
		const uint8_t *b = raw;
		const float fs = 127.0f, shift = 127.0f / 2.0f;
		for(int64_t i = 0; i < n; i++) out[i] = CMPLXF((b[2 * i] - shift) / fs, (b[2 * i + 1] - shift) / fs);
	}
}

orc_pipeline_t *orc_pipeline_create(int32_t sample_rate, int32_t centerfreq, const int32_t *freqs, int32_t nfreq,
		int fold_mode, int nthreads) {
	orc_pipeline_t *p = calloc(1, sizeof(*p));
	p->sample_rate = sample_rate; p->centerfreq = centerfreq; p->nch = nfreq;
	p->fold_mode = fold_mode; p->nthreads = nthreads < 1 ? 1 : nthreads;
	int32_t dec = orc_fft_decimation_rate(sample_rate, ORC_SYMBOL_RATE * ORC_SPS);   /* main.c:699 */
	float tbw = orc_relative_transition_bw(sample_rate, ORC_TRANSITION_BW_HZ);       /* main.c:704 */
	if(orc_ddc_init(&p->ddc, tbw, dec, 0)) { free(p); return NULL; }
	p->ch = calloc((size_t)nfreq, sizeof(*p->ch));
	for(int i = 0; i < nfreq; i++) {
		p->ch[i] = orc_channel_create(sample_rate, dec, tbw, centerfreq, freqs[i], fold_mode);
		if(!p->ch[i]) return NULL;
	}
	p->window = calloc((size_t)p->ddc.fft_size, sizeof(cf32));
	p->spectrum = calloc((size_t)p->ddc.fft_size, sizeof(cf32));
	return p;
}

static void *ring_worker_main(void *arg) {
	struct ring_worker *w = arg;
	orc_pipeline_t *p = w->p;
	for(;;) {
		const float *spec = hfdl_spectrum_ring_consume_begin(p->ring, w->id);
		if(!spec) break;
		for(int c = w->id; c < p->nch; c += p->nworkers) orc_channel_process_block(p->ch[c], (const cf32 *)spec);
		hfdl_spectrum_ring_consume_end(p->ring, w->id);
	}
	return NULL;
}

/* switch the pipeline to the ring runtime: `depth` spectra in flight, min(nthreads, channels) channel workers */
int orc_pipeline_use_ring(orc_pipeline_t *p, int depth) {
	if(p->ring || p->blocks_done > 0) return -1;
	p->nworkers = p->nthreads < p->nch ? p->nthreads : p->nch;
	if(p->nworkers < 1) p->nworkers = 1;
	p->ring = hfdl_spectrum_ring_create((size_t)p->ddc.fft_size, depth, p->nworkers);
	if(!p->ring) return -1;
	p->workers = calloc((size_t)p->nworkers, sizeof(pthread_t));
	p->wctx = calloc((size_t)p->nworkers, sizeof(struct ring_worker));
	for(int i = 0; i < p->nworkers; i++) {
		p->wctx[i] = (struct ring_worker){ p, i };
		pthread_create(&p->workers[i], NULL, ring_worker_main, &p->wctx[i]);
	}
	return 0;
}

void orc_pipeline_destroy(orc_pipeline_t *p) {
	if(!p) return;
	if(p->ring) {
		hfdl_spectrum_ring_shutdown(p->ring);
		for(int i = 0; i < p->nworkers; i++) pthread_join(p->workers[i], NULL);
		hfdl_spectrum_ring_destroy(p->ring);
		free(p->workers); free(p->wctx);
	}
	for(int i = 0; i < p->nch; i++) orc_channel_destroy(p->ch[i]);
	free(p->ch); free(p->window); free(p->spectrum); free(p->pending); free(p->sorted); free(p);
}

orc_channel_t *orc_pipeline_channel(orc_pipeline_t *p, int idx) { return (idx >= 0 && idx < p->nch) ? p->ch[idx] : NULL; }
const orc_ddc_t *orc_pipeline_ddc(orc_pipeline_t *p) { return &p->ddc; }

struct chan_job { orc_pipeline_t *p; int c0, c1; };
static void *chan_worker(void *arg) {
	struct chan_job *j = arg;
	for(int c = j->c0; c < j->c1; c++) orc_channel_process_block(j->p->ch[c], j->p->spectrum);
	return NULL;
}

static void run_block(orc_pipeline_t *p, const cf32 *newsamples) {
	const orc_ddc_t *d = &p->ddc;
	memmove(p->window, p->window + d->input_size, sizeof(cf32) * (size_t)d->overlap_length);
	memcpy(p->window + d->overlap_length, newsamples, sizeof(cf32) * (size_t)d->input_size);
	orc_fft_set_threads(p->nthreads);
	if(p->ring) {
		/* producer side of the ring: this thread is the fft thread (fft.c:49-61 without the barrier pair) */
		cf32 *slot = (cf32 *)hfdl_spectrum_ring_produce_begin(p->ring);
		orc_fft(p->window, slot, d->fft_size, +1);
		orc_swap_sides(slot, d->fft_size);
		hfdl_spectrum_ring_produce_end(p->ring);
		p->blocks_done++;
		return;
	}
	orc_fft(p->window, p->spectrum, d->fft_size, +1);
	orc_swap_sides(p->spectrum, d->fft_size);
	int nt = p->nthreads < p->nch ? p->nthreads : p->nch;
	if(nt <= 1) {
		struct chan_job j = { p, 0, p->nch };
		chan_worker(&j);
	} else {
		pthread_t th[256];
		struct chan_job jobs[256];
		if(nt > 256) nt = 256;
		for(int t = 0; t < nt; t++) {
			jobs[t] = (struct chan_job){ p, p->nch * t / nt, p->nch * (t + 1) / nt };
			pthread_create(&th[t], NULL, chan_worker, &jobs[t]);
		}
		for(int t = 0; t < nt; t++) pthread_join(th[t], NULL);
	}
	p->blocks_done++;
}

int orc_pipeline_feed(orc_pipeline_t *p, const void *raw, int64_t nsamples, int sfmt) {
	if(p->npending + nsamples > p->cap_pending) {
		p->cap_pending = p->npending + nsamples;
		p->pending = realloc(p->pending, sizeof(cf32) * (size_t)p->cap_pending);
	}
	orc_convert_samples(raw, nsamples, sfmt, p->pending + p->npending);
	p->npending += nsamples;
	int64_t isz = p->ddc.input_size, off = 0;
	int blocks = 0;
	while(p->npending - off >= isz) {
		run_block(p, p->pending + off);
		off += isz;
		blocks++;
	}
	memmove(p->pending, p->pending + off, sizeof(cf32) * (size_t)(p->npending - off));
	p->npending -= off;
	return blocks;
}

static int pdu_cmp(const void *a, const void *b) {
	const orc_pdu_t *x = a, *y = b;
	if(x->sample_cnt_end != y->sample_cnt_end) return x->sample_cnt_end < y->sample_cnt_end ? -1 : 1;
	return (x->freq > y->freq) - (x->freq < y->freq);
}
void orc_pipeline_sync(orc_pipeline_t *p) { if(p->ring) hfdl_spectrum_ring_drain(p->ring); }
static void collect(orc_pipeline_t *p) {
	int n = 0;
	orc_pipeline_sync(p);
	for(int i = 0; i < p->nch; i++) n += orc_channel_pdu_count(p->ch[i]);
	p->sorted = realloc(p->sorted, sizeof(orc_pdu_t) * (size_t)(n ? n : 1));
	int k = 0;
	for(int i = 0; i < p->nch; i++)
		for(int j = 0; j < orc_channel_pdu_count(p->ch[i]); j++) orc_channel_get_pdu(p->ch[i], j, &p->sorted[k++]);
	qsort(p->sorted, (size_t)n, sizeof(orc_pdu_t), pdu_cmp);
	p->nsorted = n;
}
int orc_pipeline_pdu_count(orc_pipeline_t *p) { collect(p); return p->nsorted; }
int orc_pipeline_get_pdu(orc_pipeline_t *p, int idx, orc_pdu_t *out) {
	if(idx < 0 || idx >= p->nsorted) return -1;
	*out = p->sorted[idx];
	return 0;
}
int orc_pipeline_last_spectrum(orc_pipeline_t *p, cf32 *dst, int n) {
	if(n > p->ddc.fft_size) n = p->ddc.fft_size;
	memcpy(dst, p->spectrum, sizeof(cf32) * (size_t)n);
	return n;
}
