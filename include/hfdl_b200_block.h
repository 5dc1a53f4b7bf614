/*
 * include/hfdl_b200_block.h -- the block.c-facing side of libhfdl_b200.so: one `struct block` that replaces
 * the reference's fft block and all of its hfdl channel blocks, so that input-file.c / input-soapysdr.c keep
 * producing into the same ring and pdu.c / fmtr-*.c / output-*.c keep consuming PDUs, untouched.
 *
 * Replaces (dumphfdl src/):
 *   hfdl_gpu_frontend_create   <- fft_create (fft.h:31) + hfdl_init_globals / hfdl_channel_create x C (hfdl.h:10-12)
 *                                 + block_connect_one2many(fft, channels) (main.c:752-755) + csdr_fft_init (main.c:697)
 *   hfdl_gpu_frontend_destroy  <- fft_destroy (fft.h:32) + hfdl_channel_destroy x C (hfdl.h:13) + block_disconnect_one2many
 *   block->thread_routine      <- fft_thread (fft.c:22-68) + hfdl_decoder_thread x C (hfdl.c:593-935); started by the
 *                                 reference's own block_start() (block.c:157-166) after block_connect_one2one(input, this)
 *   PDU delivery               <- dispatch_pdu (hfdl.c:1058-1080): hfdl_pdu_metadata_create + octet_string_new +
 *                                 pdu_decoder_queue_push (pdu.h:39-40), resolved against the host program at load time
 *   hfdl_gpu_frontend_print_summary <- hfdl_print_summary (hfdl.h:14)
 * The ring of block_connect_one2one is liquid-dsp's cbuffercf (block.c:20): the library CALLS the host program's
 * cbuffercf_size / cbuffercf_read / cbuffercf_release (weak references, resolved at load time) and defines none.
 *
 * The struct layouts below restate the ABI of src/block.h:27-68 (field order and types are the interface).
 */
#ifndef HFDL_B200_BLOCK_H
#define HFDL_B200_BLOCK_H
#include <stdint.h>
#include <stdbool.h>
#include <stddef.h>
#include <pthread.h>
#include "hfdl_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

#ifndef HFDL_B200_NO_BLOCK_STRUCTS      /* define when compiling inside dumphfdl, which has its own block.h */
enum producer_type { PRODUCER_NONE = 0, PRODUCER_SINGLE, PRODUCER_MULTI, PRODUCER_MAX };
enum consumer_type { CONSUMER_NONE = 0, CONSUMER_SINGLE, CONSUMER_MULTI, CONSUMER_MAX };
typedef struct cbuffercf_s *cbuffercf;        /* liquid.h */
struct circ_buffer { cbuffercf buf; pthread_cond_t *cond; pthread_mutex_t *mutex; };
struct shared_buffer { void *buf; pthread_barrier_t *data_ready; pthread_barrier_t *consumers_ready; };
struct block_connection {
	union { struct circ_buffer circ_buffer; struct shared_buffer shared_buffer; };
	uint32_t flags;
};
#define BLOCK_CONNECTION_SHUTDOWN (1 << 0)
struct block;
struct producer { struct block_connection *out; size_t max_tu; enum producer_type type; };
struct consumer { struct block_connection *in; size_t min_ru; enum consumer_type type; };
struct block {
	struct consumer consumer;
	struct producer producer;
	pthread_t thread;
	void *(*thread_routine)(void *);
	bool running;
};
#endif

/* Returns &obj->block with consumer = { CONSUMER_SINGLE, min_ru = fft_size }, producer = { PRODUCER_NONE } and
 * thread_routine set; NULL on error (no channel, a channel outside the capture's span as main.c:214-226 checks it, devices
 * that are not there, a host program without liquid-dsp's cbuffercf).  The ring carries CF32 (what complex_samples_produce
 * writes).  When the host program exports statsd_counter_per_channel_increment (statsd.h:15, WITH_STATSD builds) the block
 * thread forwards the demod.preamble.A2_found / M1_found / errors.M1_not_found increments of hfdl.c:818,828,840 to it.
 * ngpus >= 1: CUDA devices device .. device+ngpus-1 share the work; channel k runs on GPU k mod ngpus end to end, the
 * samples cross PCIe once and reach the other GPUs over NVLink (the reference's one2many broadcast, block.c:90-120). */
struct block *hfdl_gpu_frontend_create(int32_t sample_rate, int32_t centerfreq_hz, const int32_t *freqs_hz,
		int32_t nfreq, int32_t device, int32_t ngpus);
void hfdl_gpu_frontend_destroy(struct block *frontend_block);
void hfdl_gpu_frontend_print_summary(struct block *frontend_block);
/* channel noise floor in dBFS, as noise_floor_stats_thread reports it (hfdl.c:1093); does not stop the processing */
int32_t hfdl_gpu_frontend_noise_floor_db(struct block *frontend_block, int32_t channel, float *db);
/* the per-channel statsd metrics (doc/STATSD_METRICS.md) as counters; channel = index into freqs_hz */
int32_t hfdl_gpu_frontend_counters(struct block *frontend_block, int32_t channel, hfdl_b200_counters_t *out);
/* When the host program does not export pdu_decoder_queue_push (e.g. the tests), PDUs go to this callback. */
typedef void (*hfdl_gpu_pdu_callback)(const hfdl_b200_pdu_t *pdu, void *user);
void hfdl_gpu_frontend_set_pdu_callback(struct block *frontend_block, hfdl_gpu_pdu_callback cb, void *user);

#ifdef __cplusplus
}
#endif
#endif
