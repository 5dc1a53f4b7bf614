// dumphfdl_b200/csrc/block_shim.cu -- block.c-facing wrapper (include/hfdl_b200_block.h): a `struct block` whose
// thread routine follows the consumer protocol of fft_thread (fft.c:38-55), feeds the GPU front-end and hands
// decoded frames to the reference's pdu_decoder_queue_push (hfdl.c:1058-1080).  Host code only.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <sys/time.h>
#include <vector>
#include "../../include/hfdl_b200_block.h"
#ifdef HFDL_CUSIM
#include "cusim.h"
#else
#include <cuda_runtime.h>
#endif

// ---- symbols of the host program (dumphfdl), bound at load time when present --------------------------------
extern "C" {
struct metadata_vtable;
struct metadata { struct metadata_vtable *vtable; struct timeval rx_timestamp; };          // metadata.h:5-8
struct hfdl_pdu_metadata {                                                                   // pdu.h:8-17
	struct metadata metadata;
	int32_t version, freq, bit_rate;
	float freq_err_hz, rssi, noise_floor;
	char slot;
};
struct octet_string;
struct metadata *hfdl_pdu_metadata_create(void) __attribute__((weak));
struct octet_string *octet_string_new(void *buf, size_t len) __attribute__((weak));
void pdu_decoder_queue_push(struct metadata *metadata, struct octet_string *pdu, uint32_t flags) __attribute__((weak));
}

// ---- cbuffercf stand-in: mirrored storage so that any read of <= max_size elements is contiguous ---------------
struct hfdl_cbuffercf_s {
	unsigned int max_size, num_elements, read_index, write_index;
	float *v;                      // 2 * max_size complex elements (re, im)
};

extern "C" {

cbuffercf cbuffercf_create(unsigned int max_size) {
	if(max_size == 0) return NULL;
	cbuffercf q = (cbuffercf)calloc(1, sizeof(*q));
	q->max_size = max_size;
	q->v = (float *)calloc((size_t)max_size * 2 * 2, sizeof(float));
	return q;
}
void cbuffercf_destroy(cbuffercf q) { if(q) { free(q->v); free(q); } }
void cbuffercf_reset(cbuffercf q) { q->num_elements = q->read_index = q->write_index = 0; }
unsigned int cbuffercf_size(cbuffercf q) { return q->num_elements; }
unsigned int cbuffercf_max_size(cbuffercf q) { return q->max_size; }
unsigned int cbuffercf_space_available(cbuffercf q) { return q->max_size - q->num_elements; }
int cbuffercf_write(cbuffercf q, void *v, unsigned int n) {
	if(n > q->max_size - q->num_elements) return -1;
	const float *src = (const float *)v;
	for(unsigned int i = 0; i < n; i++) {
		unsigned int w = q->write_index;
		q->v[2 * w] = src[2 * i]; q->v[2 * w + 1] = src[2 * i + 1];
		q->v[2 * (w + q->max_size)] = src[2 * i]; q->v[2 * (w + q->max_size) + 1] = src[2 * i + 1];
		q->write_index = (w + 1 == q->max_size) ? 0 : w + 1;
	}
	q->num_elements += n;
	return 0;
}
int cbuffercf_read(cbuffercf q, unsigned int n, void **v, unsigned int *num_read) {
	if(n > q->num_elements) n = q->num_elements;
	*v = q->v + 2 * (size_t)q->read_index;
	*num_read = n;
	return 0;
}
int cbuffercf_release(cbuffercf q, unsigned int n) {
	if(n > q->num_elements) return -1;
	q->read_index = (q->read_index + n) % q->max_size;
	q->num_elements -= n;
	return 0;
}

}  // extern "C"

// ---- the block --------------------------------------------------------------------------------------------------
struct gpu_frontend {
	struct block block;            // must stay first: container_of idiom of fft.c:24 / hfdl.c:596
	hfdl_b200_frontend_t *fe;
	hfdl_b200_geometry_t geom;
	float *staging; size_t staging_samples;
	hfdl_gpu_pdu_callback cb; void *cb_user;
	struct timeval t_start;
	int64_t delivered;
};

static void deliver(gpu_frontend *g) {
	hfdl_b200_pdu_t p;
	while(hfdl_b200_pop_pdu(g->fe, &p) == 1) {
		g->delivered++;
		if(pdu_decoder_queue_push && hfdl_pdu_metadata_create && octet_string_new) {
			struct metadata *m = hfdl_pdu_metadata_create();
			struct hfdl_pdu_metadata *hm = (struct hfdl_pdu_metadata *)m;
			hm->version = p.version; hm->freq = p.freq; hm->freq_err_hz = p.freq_err_hz;
			hm->rssi = p.rssi; hm->noise_floor = p.noise_floor; hm->bit_rate = p.bit_rate; hm->slot = p.slot;
			// the reference stamps wall clock at A2 minus prekey+2A (hfdl.c:657-660,808-809); here stream time is known
			double t = (double)g->t_start.tv_sec + 1e-6 * g->t_start.tv_usec + p.rx_time_s;
			m->rx_timestamp.tv_sec = (time_t)floor(t);
			m->rx_timestamp.tv_usec = (suseconds_t)((t - floor(t)) * 1e6);
			uint8_t *copy = (uint8_t *)calloc((size_t)p.len, 1);
			memcpy(copy, p.octets, (size_t)p.len);
			pdu_decoder_queue_push(m, octet_string_new(copy, (size_t)p.len), 0);
		} else if(g->cb) {
			g->cb(&p, g->cb_user);
		}
	}
}

static void *gpu_frontend_thread(void *ctx) {
	struct block *block = (struct block *)ctx;
	gpu_frontend *g = (gpu_frontend *)block;
	struct circ_buffer *cb = &block->consumer.in->circ_buffer;
	const unsigned int isz = (unsigned int)g->geom.input_size;
	gettimeofday(&g->t_start, NULL);
	for(;;) {
		pthread_mutex_lock(cb->mutex);
		// the shutdown flag is honoured only when less than one block is buffered (drain, then exit: fft.c:38-47)
		bool stop = false;
		while(cbuffercf_size(cb->buf) < isz) {
			if(block->consumer.in->flags & BLOCK_CONNECTION_SHUTDOWN) { stop = true; break; }
			pthread_cond_wait(cb->cond, cb->mutex);
		}
		if(stop) { pthread_mutex_unlock(cb->mutex); break; }
		unsigned int avail = cbuffercf_size(cb->buf);
		unsigned int take = (avail / isz) * isz;
		if(take > g->staging_samples) take = (unsigned int)(g->staging_samples / isz) * isz;
		void *rp; unsigned int nr;
		cbuffercf_read(cb->buf, take, &rp, &nr);
		memcpy(g->staging, rp, (size_t)nr * 2 * sizeof(float));
		cbuffercf_release(cb->buf, nr);
		pthread_mutex_unlock(cb->mutex);
		if(hfdl_b200_push_samples(g->fe, g->staging, nr) < 0) { fprintf(stderr, "hfdl_gpu_frontend: GPU processing failed\n"); break; }
		if(hfdl_b200_flush(g->fe) < 0) break;
		deliver(g);
	}
	hfdl_b200_flush(g->fe);
	deliver(g);
	block->running = false;
	return NULL;
}

extern "C" {

struct block *hfdl_gpu_frontend_create(int32_t sample_rate, int32_t centerfreq_hz, const int32_t *freqs_hz, int32_t nfreq, int32_t device) {
	gpu_frontend *g = (gpu_frontend *)calloc(1, sizeof(*g));
	hfdl_b200_config_t cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.sample_rate = sample_rate; cfg.centerfreq_hz = centerfreq_hz; cfg.freqs_hz = freqs_hz; cfg.nfreq = nfreq;
	cfg.sample_format = HFDL_B200_SFMT_CF32; cfg.device = device; cfg.capture_channel = -1;
	if(hfdl_b200_create(&g->fe, &cfg) != 0) { fprintf(stderr, "Error in hfdl_gpu_frontend_create()\n"); free(g); return NULL; }
	hfdl_b200_get_geometry(g->fe, &g->geom);
	g->staging_samples = (size_t)g->geom.input_size * 8;
	if(cudaMallocHost((void **)&g->staging, g->staging_samples * 2 * sizeof(float)) != cudaSuccess) { hfdl_b200_destroy(g->fe); free(g); return NULL; }
	g->block.consumer.type = CONSUMER_SINGLE;
	g->block.consumer.min_ru = (size_t)g->geom.fft_size;
	g->block.producer.type = PRODUCER_NONE;
	g->block.thread_routine = gpu_frontend_thread;
	return &g->block;
}

void hfdl_gpu_frontend_destroy(struct block *b) {
	if(!b) return;
	gpu_frontend *g = (gpu_frontend *)b;
	hfdl_b200_destroy(g->fe);
	cudaFreeHost(g->staging);
	free(g);
}

void hfdl_gpu_frontend_print_summary(struct block *b) { if(b) hfdl_b200_print_summary(((gpu_frontend *)b)->fe); }

int32_t hfdl_gpu_frontend_noise_floor_db(struct block *b, int32_t channel, float *db) {
	float lvl;
	if(!b || !db || hfdl_b200_channel_noise_floor(((gpu_frontend *)b)->fe, channel, &lvl)) return -1;
	*db = 20.0f * log10f(lvl);
	return 0;
}

void hfdl_gpu_frontend_set_pdu_callback(struct block *b, hfdl_gpu_pdu_callback cb, void *user) {
	if(!b) return;
	((gpu_frontend *)b)->cb = cb; ((gpu_frontend *)b)->cb_user = user;
}

}  // extern "C"
