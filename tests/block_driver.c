/* tests/block_driver.c -- drives libhfdl_b200.so (or the host-emulation build) through the block.c contract the
 * way dumphfdl's main.c does (main.c:687-755,770-774): an input block produces CF32 into the one2one ring
 * (file_input_thread + complex_samples_produce, input-file.c:35-74, input-helpers.c:80-92), the GPU front-end
 * block consumes it, EOF triggers the ordered shutdown of block.c:137-143.  Prints one line per PDU.
 * usage: block_driver <lib.so> <capture.cf32> <sample_rate> <centerfreq_hz> <freq_hz>... */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "../include/hfdl_b200_block.h"

static pthread_mutex_t out_lock = PTHREAD_MUTEX_INITIALIZER;
static void on_pdu(const hfdl_b200_pdu_t *p, void *user) {
	(void)user;
	pthread_mutex_lock(&out_lock);
	printf("PDU %d %d %d %llu ", p->freq, p->M1, p->crc_good, (unsigned long long)p->sample_cnt_a2);
	for(int i = 0; i < p->len; i++) printf("%02x", p->octets[i]);
	printf("\n");
	pthread_mutex_unlock(&out_lock);
}

struct api {
	struct block *(*create)(int32_t, int32_t, const int32_t *, int32_t, int32_t);
	void (*destroy)(struct block *);
	void (*set_cb)(struct block *, hfdl_gpu_pdu_callback, void *);
	cbuffercf (*cb_create)(unsigned int);
	void (*cb_destroy)(cbuffercf);
	unsigned int (*cb_space)(cbuffercf);
	int (*cb_write)(cbuffercf, void *, unsigned int);
} A;

struct input { struct block block; FILE *fh; };

static void *input_thread(void *ctx) {               /* file_input_thread, input-file.c:35-74 */
	struct input *in = ctx;
	struct circ_buffer *cb = &in->block.producer.out->circ_buffer;
	const size_t batch = 40000;                      /* 320000-byte default read buffer / 8 */
	float *buf = malloc(batch * 8);
	size_t n;
	do {
		n = fread(buf, 8, batch, in->fh);
		for(;;) {
			pthread_mutex_lock(cb->mutex);
			size_t space = A.cb_space(cb->buf);
			pthread_mutex_unlock(cb->mutex);
			if(space >= n) break;
			usleep(1000);
		}
		pthread_mutex_lock(cb->mutex);              /* complex_samples_produce */
		A.cb_write(cb->buf, buf, (unsigned int)n);
		pthread_mutex_unlock(cb->mutex);
		pthread_cond_signal(cb->cond);
	} while(n > 0);
	pthread_mutex_lock(cb->mutex);                  /* block_connection_one2one_shutdown */
	in->block.producer.out->flags |= BLOCK_CONNECTION_SHUTDOWN;
	pthread_mutex_unlock(cb->mutex);
	pthread_cond_signal(cb->cond);
	in->block.running = false;
	free(buf);
	return NULL;
}

int main(int argc, char **argv) {
	if(argc < 6) { fprintf(stderr, "usage\n"); return 2; }
	void *h = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
	if(!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
	A.create = dlsym(h, "hfdl_gpu_frontend_create"); A.destroy = dlsym(h, "hfdl_gpu_frontend_destroy");
	A.set_cb = dlsym(h, "hfdl_gpu_frontend_set_pdu_callback");
	A.cb_create = dlsym(h, "cbuffercf_create"); A.cb_destroy = dlsym(h, "cbuffercf_destroy");
	A.cb_space = dlsym(h, "cbuffercf_space_available"); A.cb_write = dlsym(h, "cbuffercf_write");
	if(!A.create || !A.destroy || !A.set_cb || !A.cb_create || !A.cb_space || !A.cb_write) { fprintf(stderr, "missing symbols\n"); return 2; }
	int32_t sr = atoi(argv[3]), cf = atoi(argv[4]);
	int nf = argc - 5;
	int32_t freqs[512];
	for(int i = 0; i < nf; i++) freqs[i] = atoi(argv[5 + i]);
	struct input in;
	memset(&in, 0, sizeof(in));
	in.fh = fopen(argv[2], "rb");
	if(!in.fh) { perror("capture"); return 2; }
	in.block.producer.type = PRODUCER_SINGLE;
	in.block.producer.max_tu = 40000;
	struct block *fe = A.create(sr, cf, freqs, nf, 0);
	if(!fe) return 1;
	A.set_cb(fe, on_pdu, NULL);
	/* block_connect_one2one (block.c:55-76) */
	size_t bs = 8 * in.block.producer.max_tu;
	if(2 * fe->consumer.min_ru > bs) bs = 2 * fe->consumer.min_ru;
	struct block_connection *conn = calloc(1, sizeof(*conn));
	conn->circ_buffer.buf = A.cb_create((unsigned int)bs);
	conn->circ_buffer.cond = calloc(1, sizeof(pthread_cond_t));
	conn->circ_buffer.mutex = calloc(1, sizeof(pthread_mutex_t));
	pthread_cond_init(conn->circ_buffer.cond, NULL);
	pthread_mutex_init(conn->circ_buffer.mutex, NULL);
	in.block.producer.out = fe->consumer.in = conn;
	/* block_start (block.c:157-166) */
	fe->running = true;
	pthread_create(&fe->thread, NULL, fe->thread_routine, fe);
	in.block.running = true;
	pthread_create(&in.block.thread, NULL, input_thread, &in);
	while(in.block.running || fe->running) usleep(2000);     /* main.c:789-802 */
	pthread_join(in.block.thread, NULL);
	pthread_join(fe->thread, NULL);
	fflush(stdout);
	A.destroy(fe);
	A.cb_destroy(conn->circ_buffer.buf);
	fclose(in.fh);
	return 0;
}
