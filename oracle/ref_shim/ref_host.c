/*
 * oracle/ref_shim/ref_host.c -- host program stand-in around the REFERENCE's own hot-path sources compiled where
 * they lie (block.c, fft.c, fastddc.c, libcsdr*.c, hfdl.c, input-helpers.c, libfec/viterbi27_port.c, crc.c):
 *   - the util.c helpers those files call (start_thread, pthread_*_initialize, octet_string_new; util.c:55-105),
 *   - the downstream callee pdu_decoder_queue_push / hfdl_pdu_metadata_create (pdu.c:37-43,81-85) as a capture list,
 *   - the statsd hook points (statsd.h:21-26, hfdl.c:818,828,840,1099) as per-channel counters,
 *   - the DATADUMPS taps (dumpfile.h:13-22, hfdl.c:616-655) as in-memory captures,
 *   - a driver that wires the blocks exactly as main.c:697-711,739-755,770-774 does and feeds samples the way
 *     input-file.c:50-63 does (convert_* + complex_samples_produce).
 * Nothing here restates arithmetic of the path.  TEST INFRASTRUCTURE ONLY (oracle/_ref/, never the product).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdbool.h>
#include <string.h>
#include <errno.h>
#include <unistd.h>
#include <complex.h>
#include <pthread.h>
#include <liquid/liquid.h>
#include "config.h"
#include "block.h"
#include "fft.h"
#include "fastddc.h"
#include "libcsdr.h"
#include "hfdl.h"
#include "input-common.h"
#include "input-helpers.h"
#include "metadata.h"
#include "pdu.h"
#include "dumpfile.h"
#include "util.h"

struct dumphfdl_config Config = { .nf_stats_interval = 3600, .datadumps = false };
int32_t do_exit = 0;

/* ---------------- util.c:55-105 ---------------- */
int32_t start_thread(pthread_t *pth, void *(*start_routine)(void *), void *thread_ctx) {
	int ret = pthread_create(pth, NULL, start_routine, thread_ctx);
	if(ret != 0) { errno = ret; perror("pthread_create() failed"); return ret; }
	return pthread_detach(*pth);
}
int32_t pthread_barrier_create(pthread_barrier_t *barrier, unsigned count) { return pthread_barrier_init(barrier, NULL, count); }
int32_t pthread_cond_initialize(pthread_cond_t *cond) { return pthread_cond_init(cond, NULL); }
int32_t pthread_mutex_initialize(pthread_mutex_t *mutex) { return pthread_mutex_init(mutex, NULL); }
struct octet_string *octet_string_new(void *buf, size_t len) {
	NEW(struct octet_string, o);
	o->buf = buf; o->len = len;
	return o;
}
void octet_string_destroy(struct octet_string *o) { if(o) { free(o->buf); free(o); } }

/* ---------------- PDU capture (pdu.c:37-43,81-85) ---------------- */
typedef struct {
	int32_t version, freq, bit_rate;
	float freq_err_hz, rssi, noise_floor;
	char slot;
	int32_t len;
	uint32_t flags;
	uint8_t octets[948];
} ref_pdu_t;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static ref_pdu_t *g_pdus; static int g_npdu, g_cappdu;

struct metadata *hfdl_pdu_metadata_create() {
	NEW(struct hfdl_pdu_metadata, m);
	return &m->metadata;
}
void pdu_decoder_queue_push(struct metadata *metadata, struct octet_string *pdu, uint32_t flags) {
	struct hfdl_pdu_metadata *hm = container_of(metadata, struct hfdl_pdu_metadata, metadata);
	pthread_mutex_lock(&g_lock);
	if(g_npdu == g_cappdu) { g_cappdu = g_cappdu ? 2 * g_cappdu : 64; g_pdus = realloc(g_pdus, sizeof(ref_pdu_t) * (size_t)g_cappdu); }
	ref_pdu_t *p = &g_pdus[g_npdu++];
	memset(p, 0, sizeof(*p));
	p->version = hm->version; p->freq = hm->freq; p->bit_rate = hm->bit_rate;
	p->freq_err_hz = hm->freq_err_hz; p->rssi = hm->rssi; p->noise_floor = hm->noise_floor; p->slot = hm->slot;
	p->len = (int32_t)(pdu->len > sizeof(p->octets) ? sizeof(p->octets) : pdu->len);
	p->flags = flags;
	memcpy(p->octets, pdu->buf, (size_t)p->len);
	pthread_mutex_unlock(&g_lock);
	octet_string_destroy(pdu);          /* what pdu_decoder_thread does with them (pdu.c:171-172) */
	free(hm);
}
int ref_pdu_count(void) { return g_npdu; }
int ref_pdu_get(int i, ref_pdu_t *out) { if(i < 0 || i >= g_npdu) return -1; *out = g_pdus[i]; return 0; }
void ref_pdu_clear(void) { pthread_mutex_lock(&g_lock); g_npdu = 0; pthread_mutex_unlock(&g_lock); }

/* ---------------- statsd hooks (statsd.h:21-26): counters per (channel, name) ---------------- */
typedef struct { int32_t freq; char name[48]; long count; long gauge; } ref_stat_t;
static ref_stat_t g_stats[8192]; static int g_nstats;
static ref_stat_t *stat_slot(int32_t freq, const char *name) {
	for(int i = 0; i < g_nstats; i++) if(g_stats[i].freq == freq && strcmp(g_stats[i].name, name) == 0) return &g_stats[i];
	if(g_nstats == 8192) return NULL;
	ref_stat_t *s = &g_stats[g_nstats++];
	s->freq = freq; strncpy(s->name, name, sizeof(s->name) - 1); s->count = 0; s->gauge = 0;
	return s;
}
void statsd_counter_per_channel_increment(int32_t freq, char *counter) {
	pthread_mutex_lock(&g_lock);
	ref_stat_t *s = stat_slot(freq, counter);
	if(s) s->count++;
	pthread_mutex_unlock(&g_lock);
}
void statsd_gauge_per_channel_set(int32_t freq, char *gauge, size_t value) {
	pthread_mutex_lock(&g_lock);
	ref_stat_t *s = stat_slot(freq, gauge);
	if(s) s->gauge = (long)value;
	pthread_mutex_unlock(&g_lock);
}
long ref_stat_count(int32_t freq, const char *name) {
	long v = 0;
	pthread_mutex_lock(&g_lock);
	for(int i = 0; i < g_nstats; i++) if(g_stats[i].freq == freq && strcmp(g_stats[i].name, name) == 0) v = g_stats[i].count;
	pthread_mutex_unlock(&g_lock);
	return v;
}
void ref_stat_clear(void) { pthread_mutex_lock(&g_lock); g_nstats = 0; pthread_mutex_unlock(&g_lock); }

/* ---------------- DATADUMPS taps (dumpfile.h:13-22): (time, value) pairs kept in memory ---------------- */
#ifdef DATADUMPS
typedef struct ref_dump { char name[32]; int is_complex; uint64_t *t; float complex *v; size_t n, cap; } ref_dump_t;
static ref_dump_t *g_dumps[4096]; static int g_ndumps;
static int g_dumps_enabled;
void ref_dumps_enable(int on) { g_dumps_enabled = on; }
static ref_dump_t *dump_open(const char *name, int is_complex) {
	if(!g_dumps_enabled) return NULL;
	ref_dump_t *d = calloc(1, sizeof(*d));
	strncpy(d->name, name, sizeof(d->name) - 1);
	d->is_complex = is_complex;
	pthread_mutex_lock(&g_lock);
	if(g_ndumps < 4096) g_dumps[g_ndumps++] = d;
	pthread_mutex_unlock(&g_lock);
	return d;
}
static void dump_put(ref_dump_t *d, uint64_t time, float complex v) {
	if(!d) return;
	if(d->n == d->cap) {
		d->cap = d->cap ? 2 * d->cap : 4096;
		d->t = realloc(d->t, sizeof(uint64_t) * d->cap);
		d->v = realloc(d->v, sizeof(float complex) * d->cap);
	}
	d->t[d->n] = time; d->v[d->n] = v; d->n++;
}
dumpfile_rf32 do_dumpfile_rf32_open(char const *name, float fillval) { (void)fillval; return (dumpfile_rf32)dump_open(name, 0); }
void do_dumpfile_rf32_write_value(dumpfile_rf32 f, uint64_t time, float val) { dump_put((ref_dump_t *)f, time, val); }
void do_dumpfile_rf32_destroy(dumpfile_rf32 f) { (void)f; }
dumpfile_cf32 do_dumpfile_cf32_open(char const *name, float complex fillval) { (void)fillval; return (dumpfile_cf32)dump_open(name, 1); }
void do_dumpfile_cf32_write_value(dumpfile_cf32 f, uint64_t time, float complex val) { dump_put((ref_dump_t *)f, time, val); }
void do_dumpfile_cf32_write_block(dumpfile_cf32 f, uint64_t time, float complex *buf, size_t len) {
	for(size_t i = 0; i < len; i++) dump_put((ref_dump_t *)f, time + i, buf[i]);
}
void do_dumpfile_cf32_destroy(dumpfile_cf32 f) { (void)f; }
/* the idx-th dump opened under 'name' (one per channel thread): returns its length, copies up to max entries */
long ref_dump_read(const char *name, int idx, uint64_t *t, float complex *v, long max) {
	for(int i = 0; i < g_ndumps; i++) {
		if(strcmp(g_dumps[i]->name, name) != 0) continue;
		if(idx-- > 0) continue;
		ref_dump_t *d = g_dumps[i];
		long n = (long)d->n < max ? (long)d->n : max;
		if(t) memcpy(t, d->t, sizeof(uint64_t) * (size_t)n);
		if(v) memcpy(v, d->v, sizeof(float complex) * (size_t)n);
		return (long)d->n;
	}
	return -1;
}
void ref_dumps_clear(void) {
	for(int i = 0; i < g_ndumps; i++) { free(g_dumps[i]->t); free(g_dumps[i]->v); free(g_dumps[i]); }
	g_ndumps = 0;
}
#else
void ref_dumps_enable(int on) { (void)on; }
long ref_dump_read(const char *name, int idx, uint64_t *t, float complex *v, long max) { (void)name; (void)idx; (void)t; (void)v; (void)max; return -1; }
void ref_dumps_clear(void) {}
#endif

/* ---------------- the wiring of main.c:697-711,739-755,770-774 ---------------- */
typedef struct ref_pipeline_s {
	struct input input;            /* stands for the input block (input-common.c:43-63, input-file.c:76-110) */
	struct input_cfg cfg;
	struct block *fft;
	struct block **channels;
	int32_t nch;
	void *inbuf; float complex *outbuf; size_t max_tu;
	fastddc_t ddc;
} ref_pipeline_t;

static int g_globals_done;

struct create_job { struct ref_pipeline_s *p; int32_t sample_rate, dec; float tbw; int32_t centerfreq; const int32_t *freqs; int t, nt, fail; };
static void *create_worker(void *arg);

ref_pipeline_t *ref_pipeline_create(int32_t sample_rate, int32_t centerfreq, const int32_t *freqs, int32_t nfreq, int32_t sfmt, int32_t fft_threads) {
	ref_pipeline_t *p = calloc(1, sizeof(*p));
	csdr_fft_init(fft_threads);                                                              /* main.c:697 */
	int32_t dec = compute_fft_decimation_rate(sample_rate, HFDL_SYMBOL_RATE * SPS);          /* main.c:699 */
	float tbw = compute_filter_relative_transition_bw(sample_rate, HFDL_CHANNEL_TRANSITION_BW_HZ);   /* main.c:704 */
	p->fft = fft_create(dec, tbw);                                                           /* main.c:708 */
	if(!p->fft) { free(p); return NULL; }
	fastddc_init(&p->ddc, tbw, dec, 0);
	if(!g_globals_done) { hfdl_init_globals(); g_globals_done = 1; }                        /* main.c:739 */
	p->nch = nfreq;
	p->channels = calloc((size_t)nfreq, sizeof(struct block *));
	{
		/* main.c:741-750 creates the channels one after the other; each creation is independent (taps design + an
		 * N-point FFT), so with hundreds of channels the set-up is spread over a few threads -- set-up only, not timed */
		struct create_job jobs[16];
		pthread_t th[16];
		int nt = nfreq < 16 ? nfreq : 16;
		long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
		if(ncpu > 0 && nt > ncpu) nt = (int)ncpu;
		if(nt < 1) nt = 1;
		for(int t = 0; t < nt; t++) {
			jobs[t] = (struct create_job){ p, sample_rate, dec, tbw, centerfreq, freqs, t, nt, 0 };
			pthread_create(&th[t], NULL, create_worker, &jobs[t]);
		}
		int fail = 0;
		for(int t = 0; t < nt; t++) { pthread_join(th[t], NULL); fail |= jobs[t].fail; }
		if(fail) return NULL;
	}
	/* input block as input_create + file_input_init leave it (input-common.c:55, input-file.c:98-107) */
	p->cfg.sfmt = (sample_format)sfmt; p->cfg.sample_rate = sample_rate; p->cfg.centerfreq = centerfreq;
	p->input.config = &p->cfg;
	p->input.full_scale = get_sample_full_scale_value(p->cfg.sfmt);
	p->input.bytes_per_sample = (int32_t)get_sample_size(p->cfg.sfmt);
	p->input.convert_sample_buffer = get_sample_converter(p->cfg.sfmt);
	p->max_tu = (size_t)p->ddc.input_size;
	p->input.block.producer = (struct producer){ .type = PRODUCER_SINGLE, .max_tu = p->max_tu };
	p->input.block.consumer = (struct consumer){ .type = CONSUMER_NONE };
	p->inbuf = malloc(p->max_tu * 8);
	p->outbuf = malloc(p->max_tu * sizeof(float complex));
	if(block_connect_one2one(&p->input.block, p->fft) != 1 ||
			block_connect_one2many(p->fft, (size_t)nfreq, p->channels) != nfreq) return NULL;       /* main.c:752-753 */
	if(block_set_start((size_t)nfreq, p->channels) != nfreq || block_start(p->fft) != 1) return NULL;   /* main.c:770-771 */
	return p;
}

static void *create_worker(void *arg) {
	struct create_job *j = arg;
	for(int32_t i = j->t; i < j->p->nch; i += j->nt) {
		j->p->channels[i] = hfdl_channel_create(j->sample_rate, j->dec, j->tbw, j->centerfreq, j->freqs[i]);   /* main.c:743-744 */
		if(!j->p->channels[i]) j->fail = 1;
	}
	return NULL;
}

/* what file_input_thread does per read (input-file.c:50-63): wait for ring space, convert, produce */
int64_t ref_pipeline_feed(ref_pipeline_t *p, const void *raw, int64_t nsamples) {
	struct circ_buffer *cb = &p->input.block.producer.out->circ_buffer;
	const uint8_t *src = raw;
	int64_t done = 0;
	const size_t bps = (size_t)p->input.bytes_per_sample;
	while(done < nsamples) {
		size_t n = (size_t)(nsamples - done) < p->max_tu ? (size_t)(nsamples - done) : p->max_tu;
		for(;;) {
			pthread_mutex_lock(cb->mutex);
			size_t space = cbuffercf_space_available(cb->buf);
			pthread_mutex_unlock(cb->mutex);
			if(space >= n) break;
			usleep(50);
		}
		memcpy(p->inbuf, src + (size_t)done * bps, n * bps);
		p->input.convert_sample_buffer(&p->input, p->inbuf, n * bps, p->outbuf);
		complex_samples_produce(cb, p->outbuf, n);
		done += (int64_t)n;
	}
	return done;
}

/* ordered shutdown: input-file.c:68 -> fft.c:41-46,63-66 -> hfdl.c:665-668; returns when every thread has left */
void ref_pipeline_finish(ref_pipeline_t *p) {
	block_connection_one2one_shutdown(p->input.block.producer.out);
	for(;;) {
		bool any = *(volatile bool *)&p->fft->running;
		for(int32_t i = 0; i < p->nch; i++) any |= *(volatile bool *)&p->channels[i]->running;
		if(!any) break;
		usleep(200);
	}
	__sync_synchronize();
}

/* waits until the ring holds less than one block and the fft thread is parked again (for timing runs) */
void ref_pipeline_drain(ref_pipeline_t *p) {
	struct circ_buffer *cb = &p->input.block.producer.out->circ_buffer;
	for(;;) {
		pthread_mutex_lock(cb->mutex);
		size_t have = cbuffercf_size(cb->buf);
		pthread_mutex_unlock(cb->mutex);
		if(have < (size_t)p->ddc.input_size) break;
		usleep(50);
	}
}

void ref_pipeline_destroy(ref_pipeline_t *p) {
	if(!p) return;
	block_disconnect_one2many(p->fft, (size_t)p->nch, p->channels);
	block_disconnect_one2one(&p->input.block, p->fft);
	for(int32_t i = 0; i < p->nch; i++) hfdl_channel_destroy(p->channels[i]);
	fft_destroy(p->fft);
	free(p->channels); free(p->inbuf); free(p->outbuf); free(p);
}
int32_t ref_pipeline_input_size(ref_pipeline_t *p) { return p->ddc.input_size; }
int32_t ref_pipeline_fft_size(ref_pipeline_t *p) { return p->ddc.fft_size; }
int32_t ref_struct_sizes(int which) {
	switch(which) {
	case 0: return (int32_t)sizeof(struct block);
	case 1: return (int32_t)sizeof(struct block_connection);
	case 2: return (int32_t)sizeof(struct hfdl_pdu_metadata);
	case 3: return (int32_t)offsetof(struct block, thread_routine);
	case 4: return (int32_t)offsetof(struct hfdl_pdu_metadata, slot);
	case 5: return (int32_t)offsetof(struct block, running);
	default: return -1;
	}
}
