/* tests/block_driver.c -- drives libhfdl_b200.so (or the host-emulation build) through the REFERENCE'S OWN block.c
 * the way dumphfdl's main.c does (main.c:687-755,770-774,789-802).  oracle/_ref/libref.so holds the reference's
 * block.c and input-helpers.c compiled where they lie, a cbuffercf with liquid-dsp's API and the downstream callee
 * pdu_decoder_queue_push as a capture list (oracle/ref_shim/ref_host.c); it is loaded RTLD_GLOBAL first, so the
 * front-end library's weak references (cbuffercf_*, hfdl_pdu_metadata_create, octet_string_new,
 * pdu_decoder_queue_push) bind to it exactly as they would bind to dumphfdl + libliquid.
 *   input thread  = file_input_thread (input-file.c:35-74): read, wait for ring space, complex_samples_produce
 *   wiring        = block_connect_one2one(input, gpu) / block_start / block_connection_one2one_shutdown / block_is_running
 * Prints one line per PDU as the reference's decoder thread would receive it (struct hfdl_pdu_metadata, pdu.h:8-17).
 * usage: block_driver <libref.so> <lib.so> <capture.cf32 | badargs> <sample_rate> <centerfreq_hz> <ngpus> <freq_hz>... */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "../include/hfdl_b200_block.h"

typedef struct {       /* ref_pdu_t of oracle/ref_shim/ref_host.c */
	int32_t version, freq, bit_rate;
	float freq_err_hz, rssi, noise_floor;
	char slot;
	int32_t len;
	uint32_t flags;
	uint8_t octets[948];
} ref_pdu_t;

static struct {
	int32_t (*connect)(struct block *, struct block *);
	void (*disconnect)(struct block *, struct block *);
	int32_t (*start)(struct block *);
	void (*shutdown)(struct block_connection *);
	bool (*is_running)(struct block *);
	void (*produce)(struct circ_buffer *, float *, size_t);
	unsigned int (*cb_space)(cbuffercf);
	int (*pdu_count)(void);
	long (*stat_count)(int32_t, const char *);
	int (*pdu_get)(int, ref_pdu_t *);
	struct block *(*create)(int32_t, int32_t, const int32_t *, int32_t, int32_t, int32_t);
	void (*destroy)(struct block *);
	int32_t (*counters)(struct block *, int32_t, hfdl_b200_counters_t *);
	int32_t (*nf_db)(struct block *, int32_t, float *);
} A;

struct input { struct block block; FILE *fh; };

static void *input_thread(void *ctx) {               /* file_input_thread, input-file.c:35-74 */
	struct input *in = (struct input *)ctx;
	struct circ_buffer *cb = &in->block.producer.out->circ_buffer;
	const size_t batch = in->block.producer.max_tu;
	float *buf = malloc(batch * 8);
	size_t n;
	do {
		n = fread(buf, 8, batch, in->fh);
		for(;;) {
			pthread_mutex_lock(cb->mutex);
			size_t space = A.cb_space(cb->buf);
			pthread_mutex_unlock(cb->mutex);
			if(space >= n) break;
			usleep(500);
		}
		A.produce(cb, buf, n);                       /* complex_samples_produce, input-helpers.c:80-92 */
	} while(n > 0);
	A.shutdown(in->block.producer.out);              /* block_connection_one2one_shutdown, input-file.c:68 */
	in->block.running = false;
	free(buf);
	return NULL;
}

#define SYM(h, field, name) do { *(void **)&A.field = dlsym(h, name); if(!A.field) { fprintf(stderr, "missing symbol %s\n", name); return 2; } } while(0)

int main(int argc, char **argv) {
	if(argc < 8) { fprintf(stderr, "usage: block_driver libref.so lib.so capture.cf32 sample_rate centerfreq ngpus freq...\n"); return 2; }
	void *hr = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
	if(!hr) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
	void *h = dlopen(argv[2], RTLD_NOW | RTLD_GLOBAL);
	if(!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
	SYM(hr, connect, "block_connect_one2one"); SYM(hr, disconnect, "block_disconnect_one2one"); SYM(hr, start, "block_start");
	SYM(hr, shutdown, "block_connection_one2one_shutdown"); SYM(hr, is_running, "block_is_running");
	SYM(hr, produce, "complex_samples_produce"); SYM(hr, cb_space, "cbuffercf_space_available");
	SYM(hr, pdu_count, "ref_pdu_count"); SYM(hr, pdu_get, "ref_pdu_get"); SYM(hr, stat_count, "ref_stat_count");
	SYM(h, create, "hfdl_gpu_frontend_create"); SYM(h, destroy, "hfdl_gpu_frontend_destroy");
	SYM(h, counters, "hfdl_gpu_frontend_counters"); SYM(h, nf_db, "hfdl_gpu_frontend_noise_floor_db");
	int32_t sr = atoi(argv[4]), cf = atoi(argv[5]), ngpus = atoi(argv[6]);
	int nf = argc - 7;
	int32_t freqs[512];
	for(int i = 0; i < nf; i++) freqs[i] = atoi(argv[7 + i]);
	if(strcmp(argv[3], "badargs") == 0) {
		/* what main.c refuses before it builds its blocks (no frequency: main.c:687-695; check_frequency_span: main.c:214-226)
		 * and what cannot be served (devices that are not there) must come back as NULL, and the accessors must take a NULL block */
		int32_t far_away[2] = { freqs[0], cf + sr / 2 };
		hfdl_b200_counters_t c;
		float db;
		int bad = 0;
		bad += A.create(sr, cf, NULL, 0, 0, 1) != NULL;
		bad += A.create(sr, cf, freqs, 0, 0, 1) != NULL;
		bad += A.create(sr, cf, NULL, nf, 0, 1) != NULL;
		bad += A.create(sr, cf, far_away, 2, 0, 1) != NULL;
		bad += A.create(sr, cf, freqs, nf, 99, 1) != NULL;
		bad += A.create(sr, cf, freqs, nf, -1, 1) != NULL;
		bad += A.create(5000, cf, freqs, nf, 0, 1) != NULL;
		bad += A.counters(NULL, 0, &c) != -1;
		bad += A.nf_db(NULL, 0, &db) != -1;
		A.destroy(NULL);
		struct block *ok = A.create(sr, cf, freqs, nf, 0, 0);      /* ngpus < 1 means one */
		bad += ok == NULL;
		if(ok) {
			bad += ok->consumer.type != CONSUMER_SINGLE || ok->producer.type != PRODUCER_NONE || ok->thread_routine == NULL || ok->running;
			bad += A.counters(ok, nf, &c) != -1 || A.counters(ok, -1, &c) != -1 || A.counters(ok, 0, NULL) != -1;
			bad += A.counters(ok, 0, &c) != 0 || c.freq != freqs[0] || c.frames_processed != 0;
			A.destroy(ok);                                         /* never connected, never started */
		}
		printf("BADARGS %d\n", bad);
		return bad ? 1 : 0;
	}
	struct input in;
	memset(&in, 0, sizeof(in));
	in.fh = fopen(argv[3], "rb");
	if(!in.fh) { perror("capture"); return 2; }
	in.block.producer.type = PRODUCER_SINGLE;
	in.block.producer.max_tu = 40000;                /* 320000-byte default read buffer / 8 (input-file.c:16,107) */
	in.block.thread_routine = input_thread;
	struct block *fe = A.create(sr, cf, freqs, nf, 0, ngpus);
	if(!fe) return 1;
	if(A.connect(&in.block, fe) != 1) { fprintf(stderr, "block_connect_one2one failed\n"); return 1; }       /* main.c:752 */
	if(A.start(fe) != 1 || A.start(&in.block) != 1) { fprintf(stderr, "block_start failed\n"); return 1; }   /* main.c:771-772 */
	/* a stats thread would poll while the blocks run (noise_floor_stats_thread, hfdl.c:1082-1105): do so here */
	int polls = 0;
	while(A.is_running(&in.block) || A.is_running(fe)) {     /* main.c:794-802 */
		float db; hfdl_b200_counters_t c;
		for(int i = 0; i < nf; i++) { A.nf_db(fe, i, &db); A.counters(fe, i, &c); }
		polls++;
		usleep(2000);
	}
	for(int i = 0; i < A.pdu_count(); i++) {
		ref_pdu_t p;
		A.pdu_get(i, &p);
		printf("PDU %d %d %c %d %.6g %.6g %.6g ", p.freq, p.bit_rate, p.slot, p.version, p.freq_err_hz, p.rssi, p.noise_floor);
		for(int k = 0; k < p.len; k++) printf("%02x", p.octets[k]);
		printf("\n");
	}
	for(int i = 0; i < nf; i++) {
		hfdl_b200_counters_t c;
		if(A.counters(fe, i, &c) == 0)
			printf("CNT %d %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld\n", c.freq, (long long)c.A2_found, (long long)c.M1_found, (long long)c.M1_not_found,
				(long long)c.frames_processed, (long long)c.frames_good, (long long)c.frames_bad_fcs, (long long)c.frames_air2gnd, (long long)c.frames_gnd2air,
				(long long)c.lpdus_processed, (long long)c.lpdus_good);
	}
	/* what the host program's statsd hook (statsd_counter_per_channel_increment, captured in libref.so) received from the block */
	for(int i = 0; i < nf; i++)
		printf("STATSD %d %lld %lld %lld\n", freqs[i], (long long)A.stat_count(freqs[i], "demod.preamble.A2_found"),
			(long long)A.stat_count(freqs[i], "demod.preamble.M1_found"), (long long)A.stat_count(freqs[i], "demod.preamble.errors.M1_not_found"));
	printf("POLLS %d\n", polls);
	fflush(stdout);
	A.disconnect(&in.block, fe);
	A.destroy(fe);
	fclose(in.fh);
	return 0;
}
