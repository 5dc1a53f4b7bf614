/* dumphfdl_b200/csrc/spectrum_ring.c -- see include/hfdl_b200_ring.h.  Host C, pthreads only. */
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "../../include/hfdl_b200_ring.h"

struct hfdl_spectrum_ring {
	size_t bins;
	int32_t depth, consumers;
	float *buf;                     /* depth x bins x 2 floats */
	int64_t produced;               /* spectra published so far */
	int64_t *consumed;              /* per consumer: spectra released so far */
	int shutdown;
	pthread_mutex_t m;
	pthread_cond_t cv_space, cv_data;
};

hfdl_spectrum_ring_t *hfdl_spectrum_ring_create(size_t bins, int32_t depth, int32_t consumers) {
	if(bins == 0 || depth < 1 || consumers < 1) return NULL;
	hfdl_spectrum_ring_t *r = calloc(1, sizeof(*r));
	if(!r) return NULL;
	r->bins = bins; r->depth = depth; r->consumers = consumers;
	r->buf = calloc((size_t)depth * bins * 2, sizeof(float));
	r->consumed = calloc((size_t)consumers, sizeof(int64_t));
	if(!r->buf || !r->consumed) { free(r->buf); free(r->consumed); free(r); return NULL; }
	pthread_mutex_init(&r->m, NULL);
	pthread_cond_init(&r->cv_space, NULL);
	pthread_cond_init(&r->cv_data, NULL);
	return r;
}

void hfdl_spectrum_ring_destroy(hfdl_spectrum_ring_t *r) {
	if(!r) return;
	pthread_mutex_destroy(&r->m);
	pthread_cond_destroy(&r->cv_space);
	pthread_cond_destroy(&r->cv_data);
	free(r->buf); free(r->consumed); free(r);
}

static int64_t min_consumed(const hfdl_spectrum_ring_t *r) {
	int64_t m = r->consumed[0];
	for(int32_t i = 1; i < r->consumers; i++) if(r->consumed[i] < m) m = r->consumed[i];
	return m;
}

float *hfdl_spectrum_ring_produce_begin(hfdl_spectrum_ring_t *r) {
	pthread_mutex_lock(&r->m);
	while(r->produced - min_consumed(r) >= r->depth) pthread_cond_wait(&r->cv_space, &r->m);
	float *slot = r->buf + (size_t)(r->produced % r->depth) * r->bins * 2;
	pthread_mutex_unlock(&r->m);
	return slot;
}

void hfdl_spectrum_ring_produce_end(hfdl_spectrum_ring_t *r) {
	pthread_mutex_lock(&r->m);
	r->produced++;
	pthread_cond_broadcast(&r->cv_data);
	pthread_mutex_unlock(&r->m);
}

const float *hfdl_spectrum_ring_consume_begin(hfdl_spectrum_ring_t *r, int32_t c) {
	if(c < 0 || c >= r->consumers) return NULL;
	pthread_mutex_lock(&r->m);
	while(r->consumed[c] >= r->produced && !r->shutdown) pthread_cond_wait(&r->cv_data, &r->m);
	const float *slot = NULL;
	if(r->consumed[c] < r->produced) slot = r->buf + (size_t)(r->consumed[c] % r->depth) * r->bins * 2;
	pthread_mutex_unlock(&r->m);
	return slot;
}

void hfdl_spectrum_ring_consume_end(hfdl_spectrum_ring_t *r, int32_t c) {
	if(c < 0 || c >= r->consumers) return;
	pthread_mutex_lock(&r->m);
	r->consumed[c]++;
	pthread_cond_broadcast(&r->cv_space);
	pthread_mutex_unlock(&r->m);
}

void hfdl_spectrum_ring_shutdown(hfdl_spectrum_ring_t *r) {
	pthread_mutex_lock(&r->m);
	r->shutdown = 1;
	pthread_cond_broadcast(&r->cv_data);
	pthread_mutex_unlock(&r->m);
}

void hfdl_spectrum_ring_drain(hfdl_spectrum_ring_t *r) {
	pthread_mutex_lock(&r->m);
	while(min_consumed(r) < r->produced) pthread_cond_wait(&r->cv_space, &r->m);
	pthread_mutex_unlock(&r->m);
}

int64_t hfdl_spectrum_ring_produced(hfdl_spectrum_ring_t *r) {
	pthread_mutex_lock(&r->m);
	int64_t v = r->produced;
	pthread_mutex_unlock(&r->m);
	return v;
}
