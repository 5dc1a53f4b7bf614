/*
 * include/hfdl_b200_ring.h -- spectrum ring: a drop-in for the reference's one2many block connection on the CPU path.
 *
 * The reference hands ONE spectrum buffer from the fft thread to the channel threads through a pair of barriers
 * (block_shared_buffer_init / block_connect_one2many, block.c:35-43,90-120; fft.c:57-61; hfdl.c:663-664): the forward
 * FFT of block k+1 cannot start before every channel has finished block k.  This ring keeps `depth`
 * spectra: the producer fills slot k mod depth while the consumers still read older slots, every consumer sees every
 * spectrum exactly once and in order, and the producer only blocks when all slots are still in use.  depth = 1 is the
 * reference's behaviour.  Host C only (pthreads); used for hybrid CPU/GPU operation and for a fairer CPU baseline.
 *
 * Replaces:  block_shared_buffer_init / _destroy (block.c:35-53)      -> hfdl_spectrum_ring_create / _destroy
 *            pthread_barrier_wait(data_ready) by the producer (fft.c:60) -> hfdl_spectrum_ring_produce_end
 *            pthread_barrier_wait(consumers_ready) + (data_ready) by a consumer (hfdl.c:663-664)
 *                                                                        -> hfdl_spectrum_ring_consume_begin / _end
 *            block_connection_one2many_shutdown (block.c:145-149)        -> hfdl_spectrum_ring_shutdown
 */
#ifndef HFDL_B200_RING_H
#define HFDL_B200_RING_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct hfdl_spectrum_ring hfdl_spectrum_ring_t;

/* bins complex64 values per spectrum, depth >= 1 slots, consumers >= 1 readers; NULL on error */
hfdl_spectrum_ring_t *hfdl_spectrum_ring_create(size_t bins, int32_t depth, int32_t consumers);
void hfdl_spectrum_ring_destroy(hfdl_spectrum_ring_t *r);
/* producer: the slot to fill next (blocks while the oldest slot is still being read); then publish it */
float *hfdl_spectrum_ring_produce_begin(hfdl_spectrum_ring_t *r);
void   hfdl_spectrum_ring_produce_end(hfdl_spectrum_ring_t *r);
/* consumer c: the next spectrum it has not seen (blocks until published); NULL once the ring is shut down AND drained
 * (the reference's consumers also finish the blocks already handed over, hfdl.c:665-668) */
const float *hfdl_spectrum_ring_consume_begin(hfdl_spectrum_ring_t *r, int32_t consumer);
void   hfdl_spectrum_ring_consume_end(hfdl_spectrum_ring_t *r, int32_t consumer);
void   hfdl_spectrum_ring_shutdown(hfdl_spectrum_ring_t *r);
/* blocks until every published spectrum has been consumed by every consumer */
void   hfdl_spectrum_ring_drain(hfdl_spectrum_ring_t *r);
int64_t hfdl_spectrum_ring_produced(hfdl_spectrum_ring_t *r);

#ifdef __cplusplus
}
#endif
#endif
