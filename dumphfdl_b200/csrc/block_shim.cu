// dumphfdl_b200/csrc/block_shim.cu -- block.c-facing wrapper (include/hfdl_b200_block.h): a `struct block` whose
// thread routine follows the consumer protocol of fft_thread (fft.c:38-55), feeds the GPU front-end(s) and hands
// decoded frames to the reference's pdu_decoder_queue_push (hfdl.c:1058-1080).  Host code only.
//
// Everything the wrapper needs from the host program is a WEAK reference resolved when the library is loaded into
// dumphfdl: liquid-dsp's cbuffercf_* (the ring block_connect_one2one creates, block.c:20), hfdl_pdu_metadata_create,
// octet_string_new, pdu_decoder_queue_push and (optional) the statsd hook statsd_counter_per_channel_increment.
// The library itself defines none of them (it must not interpose libliquid's cbuffercf).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <errno.h>
#include <time.h>
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
void statsd_counter_per_channel_increment(int32_t freq, char *counter) __attribute__((weak));       // statsd.h:15 (WITH_STATSD builds)
// liquid-dsp cbuffercf (liquid.h; used by block.c:20,28, fft.c:41-54, input-helpers.c:83-89): the four calls the consumer makes
unsigned int cbuffercf_size(cbuffercf q) __attribute__((weak));
// (liquid-dsp 1.3.x returns void from the next two, >= 1.4 an int status that is not used here)
void cbuffercf_read(cbuffercf q, unsigned int num_requested, void **v, unsigned int *num_read) __attribute__((weak));
void cbuffercf_release(cbuffercf q, unsigned int n) __attribute__((weak));
}

// ---- the block --------------------------------------------------------------------------------------------------
#define HFDL_SHIM_NRECV 3
struct gpu_frontend {
	struct block block;            // must stay first: container_of idiom of fft.c:24 / hfdl.c:596
	int ngpus;
	std::vector<hfdl_b200_frontend_t *> fe;        // fe[g] owns channels g, g + ngpus, g + 2*ngpus, ...
	hfdl_b200_geometry_t geom;
	float *staging; size_t staging_samples;
	hfdl_gpu_pdu_callback cb; void *cb_user;
	struct timeval t_start;
	int64_t delivered;
	std::vector<int64_t> statsd_sent;              // per channel x 3: demod.preamble.* increments already forwarded to the host's statsd hook
	struct timespec statsd_last;
	// ngpus > 1, sharded spectrum (include/hfdl_b200.h): GPU d transforms its share of every batch's blocks for all
	// channels, the pass-band slices change hands over NVLink (peer copies), GPU q demodulates channels q, q + ngpus, ...
	bool sharded;
	bool direct;                                   // every GPU can store into every other GPU's memory: the pack kernel writes the receive buffers itself
	int device0, nfreq, bmax, recv_i;
	int64_t blocks_done;
	std::vector<void *> d_in, d_send;              // per GPU: its samples of a batch / the slices it computed
	std::vector<void *> d_recv[HFDL_SHIM_NRECV];   // per GPU, rotated: the slices of all blocks for its channels
	std::vector<cudaStream_t> st, xs;              // per GPU: upload + FFT + sends / hand-over to the demodulator
	std::vector<cudaEvent_t> ev_up, ev_sent;
};

static void deliver(gpu_frontend *g) {
	hfdl_b200_pdu_t p;
	for(int d = 0; d < g->ngpus; d++)
	while(hfdl_b200_pop_pdu(g->fe[(size_t)d], &p) == 1) {
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

// The reference fires statsd_increment_per_channel(freq, "demod.preamble.A2_found" / "M1_found" / "errors.M1_not_found") from its
// channel threads (hfdl.c:818,828,840).  Here those events happen on the GPU; when the host program exports the hook (a
// WITH_STATSD build of dumphfdl) the block thread forwards the increments about once a second and when it exits, so the
// metrics of doc/STATSD_METRICS.md keep flowing.  The frame / LPDU counters are fired downstream by the host's own
// pdu_decoder_thread, which is not replaced.
static void forward_statsd(gpu_frontend *g, bool force) {
	if(!statsd_counter_per_channel_increment) return;
	struct timespec now;
	clock_gettime(CLOCK_MONOTONIC, &now);
	if(!force && (now.tv_sec - g->statsd_last.tv_sec) * 1000 + (now.tv_nsec - g->statsd_last.tv_nsec) / 1000000 < 1000) return;
	g->statsd_last = now;
	static char n_a2[] = "demod.preamble.A2_found", n_m1[] = "demod.preamble.M1_found", n_fail[] = "demod.preamble.errors.M1_not_found";
	char *names[3] = { n_a2, n_m1, n_fail };      // (statsd.h:12-15: not const, the client may modify them)
	if(g->statsd_sent.size() != (size_t)g->nfreq * 3) g->statsd_sent.assign((size_t)g->nfreq * 3, 0);
	for(int k = 0; k < g->nfreq; k++) {
		hfdl_b200_counters_t c;
		if(hfdl_b200_channel_counters(g->fe[(size_t)(k % g->ngpus)], k / g->ngpus, &c) != 0) continue;
		const int64_t cur[3] = { c.A2_found, c.M1_found, c.M1_not_found };
		for(int i = 0; i < 3; i++) {
			int64_t &sent = g->statsd_sent[(size_t)k * 3 + (size_t)i];
			for(; sent < cur[i]; sent++) statsd_counter_per_channel_increment(c.freq, names[i]);
		}
	}
}

// One batch of nb blocks in sharded-spectrum mode.  staging = [overlap | nb * input_size] CF32 samples (pinned).
static bool sharded_batch(gpu_frontend *g, int nb) {
	const int G = g->ngpus, cper = g->nfreq / G;
	const size_t isz = (size_t)g->geom.input_size, ovl = (size_t)g->geom.overlap_length;
	const size_t slice = (size_t)g->geom.fft_inv_size * 2 * sizeof(float);
	std::vector<int> n((size_t)G), first((size_t)G);
	for(int d = 0, at = 0; d < G; d++) { n[(size_t)d] = nb / G + (d < nb % G ? 1 : 0); first[(size_t)d] = at; at += n[(size_t)d]; }
	// the receive buffers rotate; the batch that read this one HFDL_SHIM_NRECV batches ago must be through the channeliser
	const int slot = g->recv_i++ % HFDL_SHIM_NRECV;
	for(int q = 0; q < G; q++) if(hfdl_b200_wait_input(g->fe[(size_t)q], HFDL_SHIM_NRECV - 1) < 0) return false;
	// every GPU uploads ITS blocks (plus the overlap in front of them) over its own PCIe link and transforms them; the
	// slices of GPU q's channels go to q's [nb][channels][M] array at the place of the sender's blocks -- stored there by
	// the pack kernel itself over NVLink (peer access), or packed locally and copied device to device
	for(int d = 0; d < G; d++) {
		if(n[(size_t)d] == 0) continue;
		cudaSetDevice(g->device0 + d);
		const float *src = g->staging + (size_t)first[(size_t)d] * isz * 2;
		if(cudaMemcpyAsync(g->d_in[(size_t)d], src, (ovl + (size_t)n[(size_t)d] * isz) * 2 * sizeof(float), cudaMemcpyHostToDevice, g->st[(size_t)d]) != cudaSuccess) return false;
		cudaEventRecord(g->ev_up[(size_t)d], g->st[(size_t)d]);
		if(g->direct) {
			if(hfdl_b200_spectrum_slices_to(g->fe[(size_t)d], g->d_in[(size_t)d], g->blocks_done + first[(size_t)d], n[(size_t)d],
					g->d_recv[slot].data(), first[(size_t)d], (void *)g->st[(size_t)d]) < 0) return false;
		} else {
			if(hfdl_b200_spectrum_slices(g->fe[(size_t)d], g->d_in[(size_t)d], g->blocks_done + first[(size_t)d], n[(size_t)d], g->d_send[(size_t)d], (void *)g->st[(size_t)d]) < 0) return false;
			const size_t part = (size_t)n[(size_t)d] * cper * slice;
			for(int q = 0; q < G; q++) {
				unsigned char *dst = (unsigned char *)g->d_recv[slot][(size_t)q] + (size_t)first[(size_t)d] * cper * slice;
				const unsigned char *srcp = (const unsigned char *)g->d_send[(size_t)d] + (size_t)q * part;
				if(cudaMemcpyPeerAsync(dst, g->device0 + q, srcp, g->device0 + d, part, g->st[(size_t)d]) != cudaSuccess) return false;
			}
		}
		cudaEventRecord(g->ev_sent[(size_t)d], g->st[(size_t)d]);
	}
	for(int q = 0; q < G; q++) {
		cudaSetDevice(g->device0 + q);
		for(int d = 0; d < G; d++) if(n[(size_t)d] > 0) cudaStreamWaitEvent(g->xs[(size_t)q], g->ev_sent[(size_t)d], 0);
		if(hfdl_b200_process_slices(g->fe[(size_t)q], g->d_recv[slot][(size_t)q], nb, (void *)g->xs[(size_t)q]) < 0) return false;
	}
	// the staging buffer is refilled next: the uploads must have left it
	for(int d = 0; d < G; d++) if(n[(size_t)d] > 0 && cudaEventSynchronize(g->ev_up[(size_t)d]) != cudaSuccess) return false;
	g->blocks_done += nb;
	memmove(g->staging, g->staging + (size_t)nb * isz * 2, ovl * 2 * sizeof(float));      // overlap of the next batch
	cudaSetDevice(g->device0);
	return true;
}

// Consumer side of the one2one ring, as fft_thread runs it (fft.c:38-55) -- but nothing here waits for the GPU: the
// blocks read from the ring are queued (hfdl_b200_submit) and the PDUs of batches that have finished meanwhile are
// delivered (hfdl_b200_poll); the ring is drained again while the GPU works.  When the producer pauses, a timed wait
// picks up the stragglers.  With several GPUs the samples cross PCIe once (to the first GPU) and reach the others
// over NVLink (hfdl_b200_push_peer).
static void *gpu_frontend_thread(void *ctx) {
	struct block *block = (struct block *)ctx;
	gpu_frontend *g = (gpu_frontend *)block;
	struct circ_buffer *cb = &block->consumer.in->circ_buffer;
	const unsigned int isz = (unsigned int)g->geom.input_size;
	gettimeofday(&g->t_start, NULL);
	bool ok = true;
	while(ok) {
		pthread_mutex_lock(cb->mutex);
		// the shutdown flag is honoured only when less than one block is buffered (drain, then exit: fft.c:38-47)
		bool stop = false, idle = false;
		while(cbuffercf_size(cb->buf) < isz) {
			if(block->consumer.in->flags & BLOCK_CONNECTION_SHUTDOWN) { stop = true; break; }
			struct timespec ts;
			clock_gettime(CLOCK_REALTIME, &ts);
			ts.tv_nsec += 20 * 1000 * 1000;
			if(ts.tv_nsec >= 1000000000L) { ts.tv_sec++; ts.tv_nsec -= 1000000000L; }
			if(pthread_cond_timedwait(cb->cond, cb->mutex, &ts) == ETIMEDOUT) { idle = true; break; }
		}
		if(stop) { pthread_mutex_unlock(cb->mutex); break; }
		unsigned int nr = 0;
		if(!idle || cbuffercf_size(cb->buf) >= isz) {
			unsigned int avail = cbuffercf_size(cb->buf);
			unsigned int take = (avail / isz) * isz;
			if(take > g->staging_samples) take = (unsigned int)(g->staging_samples / isz) * isz;
			void *rp;
			cbuffercf_read(cb->buf, take, &rp, &nr);
			// (sharded mode keeps the previous batch's last overlap_length samples in front of the new ones)
			memcpy(g->staging + (g->sharded ? (size_t)g->geom.overlap_length * 2 : 0), rp, (size_t)nr * 2 * sizeof(float));
			cbuffercf_release(cb->buf, nr);
		}
		pthread_mutex_unlock(cb->mutex);
		if(nr > 0 && g->sharded) {
			if(!sharded_batch(g, (int)(nr / isz))) ok = false;
		} else if(nr > 0) {
			if(hfdl_b200_push_samples(g->fe[0], g->staging, nr) < 0) ok = false;
			for(int d = 1; d < g->ngpus && ok; d++) if(hfdl_b200_push_peer(g->fe[(size_t)d], g->fe[0]) < 0) ok = false;
			for(int d = 0; d < g->ngpus && ok; d++) if(hfdl_b200_submit(g->fe[(size_t)d]) < 0) ok = false;
		}
		for(int d = 0; d < g->ngpus && ok; d++) if(hfdl_b200_poll(g->fe[(size_t)d]) < 0) ok = false;
		if(!ok) { fprintf(stderr, "hfdl_gpu_frontend: GPU processing failed\n"); break; }
		deliver(g);
		forward_statsd(g, false);
	}
	for(int d = 0; d < g->ngpus; d++) hfdl_b200_flush(g->fe[(size_t)d]);      // (sharded mode: nothing is buffered, this drains the pipelines)
	deliver(g);
	forward_statsd(g, true);
	block->running = false;
	return NULL;
}

extern "C" {

struct block *hfdl_gpu_frontend_create(int32_t sample_rate, int32_t centerfreq_hz, const int32_t *freqs_hz, int32_t nfreq, int32_t device, int32_t ngpus) {
	if(!cbuffercf_size || !cbuffercf_read || !cbuffercf_release) {
		fprintf(stderr, "hfdl_gpu_frontend_create: the host program does not export liquid-dsp's cbuffercf_size/read/release\n");
		return NULL;
	}
	if(!freqs_hz || nfreq < 1) { fprintf(stderr, "hfdl_gpu_frontend_create: no channel frequencies\n"); return NULL; }      // (main.c:687-695 refuses to start without one)
	if(ngpus < 1) ngpus = 1;
	if(ngpus > nfreq) ngpus = nfreq;
	if(device < 0 || device + ngpus > hfdl_b200_device_count()) {
		fprintf(stderr, "hfdl_gpu_frontend_create: devices %d..%d requested, %d present\n", device, device + ngpus - 1, hfdl_b200_device_count());
		return NULL;
	}
	gpu_frontend *g = new gpu_frontend();
	memset(&g->block, 0, sizeof(g->block));
	g->ngpus = ngpus; g->staging = NULL; g->cb = NULL; g->cb_user = NULL; g->delivered = 0;
	g->statsd_last.tv_sec = 0; g->statsd_last.tv_nsec = 0;
	g->device0 = device; g->nfreq = nfreq; g->recv_i = 0; g->blocks_done = 0; g->bmax = 0;
	g->sharded = ngpus > 1 && nfreq % ngpus == 0 && !getenv("HFDL_B200_SHIM_BROADCAST");
	g->direct = g->sharded && !getenv("HFDL_B200_SHIM_COPY");
	for(int d = 0; d < ngpus; d++) {
		std::vector<int32_t> mine;
		for(int k = d; k < nfreq; k += ngpus) mine.push_back(freqs_hz[k]);
		hfdl_b200_config_t cfg;
		memset(&cfg, 0, sizeof(cfg));
		cfg.sample_rate = sample_rate; cfg.centerfreq_hz = centerfreq_hz; cfg.freqs_hz = mine.data(); cfg.nfreq = (int32_t)mine.size();
		cfg.sample_format = HFDL_B200_SFMT_CF32; cfg.device = device + d; cfg.capture_channel = -1;
		hfdl_b200_frontend_t *fe = NULL;
		if(hfdl_b200_create(&fe, &cfg) != 0) {
			fprintf(stderr, "Error in hfdl_gpu_frontend_create()\n");
			for(auto q : g->fe) hfdl_b200_destroy(q);
			delete g;
			return NULL;
		}
		g->fe.push_back(fe);
	}
	hfdl_b200_get_geometry(g->fe[0], &g->geom);
	g->staging_samples = (size_t)g->geom.input_size * 8;
	cudaSetDevice(device);
	bool ok = true;
	if(g->sharded) {
		// a batch = what one ring read delivers, at most 8 blocks (and at least ngpus, so that every GPU can get one)
		g->bmax = ngpus > 8 ? ngpus : 8;
		g->staging_samples = (size_t)g->geom.input_size * (size_t)g->bmax;
		const size_t isz = (size_t)g->geom.input_size, ovl = (size_t)g->geom.overlap_length, cper = (size_t)(nfreq / ngpus);
		const size_t slice = (size_t)g->geom.fft_inv_size * 2 * sizeof(float);
		const size_t share = (size_t)(g->bmax + ngpus - 1) / (size_t)ngpus;           // blocks one GPU transforms per batch
		g->d_in.assign((size_t)ngpus, nullptr); g->d_send.assign((size_t)ngpus, nullptr);
		for(int k = 0; k < HFDL_SHIM_NRECV; k++) g->d_recv[k].assign((size_t)ngpus, nullptr);
		g->st.assign((size_t)ngpus, nullptr); g->xs.assign((size_t)ngpus, nullptr);
		g->ev_up.assign((size_t)ngpus, nullptr); g->ev_sent.assign((size_t)ngpus, nullptr);
		for(int d = 0; d < ngpus && ok; d++) {
			cudaSetDevice(device + d);
			for(int q = 0; q < ngpus; q++) if(q != d) {
				int can = 0;
				if(cudaDeviceCanAccessPeer(&can, device + d, device + q) != cudaSuccess || !can) { g->direct = false; continue; }
				if(cudaDeviceEnablePeerAccess(device + q, 0) != cudaSuccess) cudaGetLastError();      // "already enabled" is fine
			}
			ok = ok && hfdl_b200_set_exchange(g->fe[(size_t)d], freqs_hz, nfreq, ngpus) == 0;
			ok = ok && cudaMalloc(&g->d_in[(size_t)d], (ovl + share * isz) * 2 * sizeof(float)) == cudaSuccess;
			ok = ok && cudaMalloc(&g->d_send[(size_t)d], (size_t)ngpus * share * cper * slice) == cudaSuccess;
			for(int k = 0; k < HFDL_SHIM_NRECV; k++) ok = ok && cudaMalloc(&g->d_recv[k][(size_t)d], (size_t)g->bmax * cper * slice) == cudaSuccess;
			ok = ok && cudaStreamCreateWithFlags(&g->st[(size_t)d], cudaStreamNonBlocking) == cudaSuccess;
			ok = ok && cudaStreamCreateWithFlags(&g->xs[(size_t)d], cudaStreamNonBlocking) == cudaSuccess;
			ok = ok && cudaEventCreateWithFlags(&g->ev_up[(size_t)d], cudaEventDisableTiming) == cudaSuccess;
			ok = ok && cudaEventCreateWithFlags(&g->ev_sent[(size_t)d], cudaEventDisableTiming) == cudaSuccess;
		}
		cudaSetDevice(device);
	}
	const size_t staging_total = g->staging_samples + (g->sharded ? (size_t)g->geom.overlap_length : 0);
	if(!ok || cudaMallocHost((void **)&g->staging, staging_total * 2 * sizeof(float)) != cudaSuccess) {
		hfdl_gpu_frontend_destroy(&g->block);
		return NULL;
	}
	memset(g->staging, 0, staging_total * 2 * sizeof(float));
	g->block.consumer.type = CONSUMER_SINGLE;
	g->block.consumer.min_ru = (size_t)g->geom.fft_size;
	g->block.producer.type = PRODUCER_NONE;
	g->block.thread_routine = gpu_frontend_thread;
	return &g->block;
}

void hfdl_gpu_frontend_destroy(struct block *b) {
	if(!b) return;
	gpu_frontend *g = (gpu_frontend *)b;
	for(auto q : g->fe) hfdl_b200_destroy(q);
	for(size_t d = 0; d < g->d_in.size(); d++) {
		cudaSetDevice(g->device0 + (int)d);
		cudaFree(g->d_in[d]); cudaFree(g->d_send[d]);
		for(int k = 0; k < HFDL_SHIM_NRECV; k++) cudaFree(g->d_recv[k][d]);
		if(g->st[d]) cudaStreamDestroy(g->st[d]);
		if(g->xs[d]) cudaStreamDestroy(g->xs[d]);
		if(g->ev_up[d]) cudaEventDestroy(g->ev_up[d]);
		if(g->ev_sent[d]) cudaEventDestroy(g->ev_sent[d]);
	}
	if(g->staging) cudaFreeHost(g->staging);
	delete g;
}

void hfdl_gpu_frontend_print_summary(struct block *b) {
	if(!b) return;
	for(auto q : ((gpu_frontend *)b)->fe) hfdl_b200_print_summary(q);
}

int32_t hfdl_gpu_frontend_noise_floor_db(struct block *b, int32_t channel, float *db) {
	float lvl;
	if(!b || !db || channel < 0) return -1;
	gpu_frontend *g = (gpu_frontend *)b;
	if(hfdl_b200_channel_noise_floor(g->fe[(size_t)(channel % g->ngpus)], channel / g->ngpus, &lvl)) return -1;
	*db = 20.0f * log10f(lvl);
	return 0;
}

int32_t hfdl_gpu_frontend_counters(struct block *b, int32_t channel, hfdl_b200_counters_t *out) {
	if(!b || !out || channel < 0) return -1;
	gpu_frontend *g = (gpu_frontend *)b;
	return hfdl_b200_channel_counters(g->fe[(size_t)(channel % g->ngpus)], channel / g->ngpus, out);
}

void hfdl_gpu_frontend_set_pdu_callback(struct block *b, hfdl_gpu_pdu_callback cb, void *user) {
	if(!b) return;
	((gpu_frontend *)b)->cb = cb; ((gpu_frontend *)b)->cb_user = user;
}

}  // extern "C"
