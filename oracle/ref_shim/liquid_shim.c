/*
 * oracle/ref_shim/liquid_shim.c -- the liquid-dsp C API subset declared in ref_shim/liquid/liquid.h, served by the
 * oracle's restated objects (oracle/orc_liquid.c, oracle/orc_dsp.c).  Linked with the reference's own hfdl.c /
 * block.c / fft.c / input-helpers.c in oracle/_ref/ so that those files run UNMODIFIED on the same object
 * arithmetic as the oracle: what the comparison then pins is everything hfdl.c itself does (sample loop, Costas
 * loop, sampler, framer FSM, descrambler, deinterleaver, decode_user_data, dispatch_pdu).  The objects themselves
 * stay "parity unpinned" against real liquid-dsp (absent here).  TEST INFRASTRUCTURE ONLY.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <liquid/liquid.h>
#include "orc_liquid.h"

int liquid_libversion_number(void) { return ORC_LIQUID_VERSION; }

/* ---- msresamp_crcf ---- */
struct msresamp_crcf_s { orc_resamp_t *rs; };
msresamp_crcf msresamp_crcf_create(float r, float As) {
	msresamp_crcf q = calloc(1, sizeof(*q));
	q->rs = orc_resamp_create(r, As);
	if(!q->rs) { free(q); return NULL; }
	return q;
}
void msresamp_crcf_destroy(msresamp_crcf q) { if(q) { orc_resamp_destroy(q->rs); free(q); } }
float msresamp_crcf_get_delay(msresamp_crcf q) { (void)q; return 7.0f; }      /* resamp_crcf semi-length m; sizes a buffer only (hfdl.c:473,600) */
void msresamp_crcf_execute(msresamp_crcf q, float complex *x, unsigned int nx, float complex *y, unsigned int *ny) {
	uint32_t n = 0;
	orc_resamp_execute(q->rs, x, (int)nx, y, &n);
	*ny = n;
}

/* ---- agc_crcf ---- */
struct agc_crcf_s { orc_agc_t a; };
agc_crcf agc_crcf_create(void) { agc_crcf q = calloc(1, sizeof(*q)); orc_agc_init(&q->a, 1e-2f); return q; }
void agc_crcf_destroy(agc_crcf q) { free(q); }
void agc_crcf_set_bandwidth(agc_crcf q, float bt) { q->a.alpha = bt; }
void agc_crcf_execute(agc_crcf q, float complex x, float complex *y) { *y = orc_agc_execute(&q->a, x); }
float agc_crcf_get_signal_level(agc_crcf q) { return orc_agc_signal_level(&q->a); }
float agc_crcf_get_gain(agc_crcf q) { return q->a.g; }
float agc_crcf_get_rssi(agc_crcf q) { return -20.0f * log10f(q->a.g); }
void agc_crcf_unlock(agc_crcf q) { (void)q; }      /* the reference never locks the AGC */

/* ---- firfilt_crcf ---- */
struct firfilt_crcf_s { orc_firfilt_t f; };
firfilt_crcf firfilt_crcf_create(float *h, unsigned int n) { firfilt_crcf q = calloc(1, sizeof(*q)); orc_firfilt_init(&q->f, h, (int)n); return q; }
void firfilt_crcf_destroy(firfilt_crcf q) { free(q); }
void firfilt_crcf_push(firfilt_crcf q, float complex x) { orc_firfilt_push(&q->f, x); }
void firfilt_crcf_execute(firfilt_crcf q, float complex *y) { *y = orc_firfilt_execute(&q->f); }

/* ---- eqlms_cccf ---- */
struct eqlms_cccf_s { orc_eqlms_t e; };
eqlms_cccf eqlms_cccf_create_lowpass(unsigned int n, float fc) {
	if(n != ORC_EQ_LEN) return NULL;                /* only the length hfdl.c:495 asks for is restated */
	eqlms_cccf q = calloc(1, sizeof(*q));
	orc_eqlms_init_lowpass(&q->e, fc);
	return q;
}
void eqlms_cccf_destroy(eqlms_cccf q) { free(q); }
void eqlms_cccf_reset(eqlms_cccf q) { orc_eqlms_reset(&q->e); }
void eqlms_cccf_set_bw(eqlms_cccf q, float mu) { q->e.mu = mu; }
void eqlms_cccf_push(eqlms_cccf q, float complex x) { orc_eqlms_push(&q->e, x); }
void eqlms_cccf_execute(eqlms_cccf q, float complex *y) { *y = orc_eqlms_execute(&q->e); }
void eqlms_cccf_step(eqlms_cccf q, float complex d, float complex d_hat) { orc_eqlms_step(&q->e, d, d_hat); }

/* ---- modem ---- */
struct modem_s { int m; orc_modem_t st; };
modem modem_create(modulation_scheme scheme) {
	int m = scheme == LIQUID_MODEM_BPSK ? 1 : scheme == LIQUID_MODEM_PSK4 ? 2 : scheme == LIQUID_MODEM_PSK8 ? 3 : 0;
	if(!m) return NULL;
	modem q = calloc(1, sizeof(*q));
	q->m = m;
	return q;
}
void modem_destroy(modem q) { free(q); }
void modem_demodulate(modem q, float complex x, unsigned int *sym) { *sym = orc_modem_demod(q->m, x, &q->st); }
float modem_get_demodulator_phase_error(modem q) { return orc_modem_phase_error(&q->st); }
void modem_demodulate_soft(modem q, float complex x, unsigned int *sym, unsigned char *soft_bits) {
	orc_modem_demod_soft(q->m, x, &q->st, soft_bits);
	*sym = 0;                                       /* hfdl.c:1013 ignores the hard symbol */
}

/* ---- symsync_crcf ---- */
struct symsync_crcf_s { orc_symsync_t s; };
symsync_crcf symsync_crcf_create_kaiser(unsigned int k, unsigned int m, float beta, unsigned int M) {
	(void)beta;
	if(k != ORC_SS_K || m != ORC_SS_M || M != ORC_SS_NPFB) return NULL;     /* only hfdl.c:503's shape is restated */
	symsync_crcf q = calloc(1, sizeof(*q));
	orc_symsync_init_kaiser(&q->s);
	return q;
}
void symsync_crcf_destroy(symsync_crcf q) { free(q); }
void symsync_crcf_reset(symsync_crcf q) { orc_symsync_reset(&q->s); }
void symsync_crcf_set_lf_bw(symsync_crcf q, float bt) { orc_symsync_set_lf_bw(&q->s, bt); }
void symsync_crcf_set_output_rate(symsync_crcf q, unsigned int k_out) { orc_symsync_set_output_rate(&q->s, k_out); }
void symsync_crcf_execute(symsync_crcf q, float complex *x, unsigned int nx, float complex *y, unsigned int *ny) {
	unsigned int n = 0;
	for(unsigned int i = 0; i < nx; i++) n += (unsigned int)orc_symsync_step(&q->s, x[i], y + n);
	*ny = n;
}

/* ---- bsequence (sequence/src/bsequence.c): bit array, push shifts in at the LSB end ---- */
struct bsequence_s { unsigned int num_bits, nwords; uint32_t *s; uint32_t msb_mask; };
bsequence bsequence_create(unsigned int num_bits) {
	bsequence q = calloc(1, sizeof(*q));
	q->num_bits = num_bits;
	q->nwords = (num_bits + 31) / 32;
	q->s = calloc(q->nwords ? q->nwords : 1, sizeof(uint32_t));      /* s[0] = most significant word */
	unsigned int r = num_bits % 32;
	q->msb_mask = r ? ((1u << r) - 1u) : 0xFFFFFFFFu;
	return q;
}
void bsequence_destroy(bsequence q) { if(q) { free(q->s); free(q); } }
void bsequence_reset(bsequence q) { memset(q->s, 0, sizeof(uint32_t) * q->nwords); }
void bsequence_push(bsequence q, unsigned int bit) {
	for(unsigned int i = 0; i + 1 < q->nwords; i++) q->s[i] = (q->s[i] << 1) | (q->s[i + 1] >> 31);
	q->s[q->nwords - 1] = (q->s[q->nwords - 1] << 1) | (bit & 1u);
	q->s[0] &= q->msb_mask;
}
void bsequence_init(bsequence q, unsigned char *v) {      /* MSB-first from bytes */
	unsigned int k = 0;
	unsigned char byte = 0, mask = 0x80;
	for(unsigned int i = 0; i < q->num_bits; i++) {
		if((i % 8) == 0) { byte = v[k++]; mask = 0x80; }
		bsequence_push(q, (byte & mask) ? 1 : 0);
		mask >>= 1;
	}
}
int bsequence_correlate(bsequence a, bsequence b) {       /* number of equal bit positions */
	int diff = 0;
	for(unsigned int i = 0; i < a->nwords; i++) diff += __builtin_popcount(a->s[i] ^ b->s[i]);
	return (int)a->num_bits - diff;
}
unsigned int bsequence_get_length(bsequence q) { return q->num_bits; }

/* ---- msequence ---- */
struct msequence_s { orc_msequence_t m; };
msequence msequence_create(unsigned int m, unsigned int g, unsigned int a) {
	msequence q = calloc(1, sizeof(*q));
	orc_msequence_init(&q->m, m, g, a, ORC_LIQUID_VERSION < 1006000 ? 0 : 1);
	return q;
}
void msequence_destroy(msequence q) { free(q); }
void msequence_reset(msequence q) { orc_msequence_reset(&q->m); }
unsigned int msequence_advance(msequence q) { return orc_msequence_advance(&q->m); }

unsigned int count_bit_errors(unsigned int s1, unsigned int s2) { return (unsigned int)__builtin_popcount(s1 ^ s2); }

/* ---- cbuffercf (buffer/src/cbuffer.c): circular buffer whose read() hands out a LINEAR view, so the storage
 *      keeps a mirror of the first max_read elements behind the end ---- */
struct cbuffercf_s { unsigned int max_size, max_read, num_allocated, num_elements, read_index, write_index; float complex *v; };
cbuffercf cbuffercf_create(unsigned int max_size) {
	cbuffercf q = calloc(1, sizeof(*q));
	q->max_size = max_size; q->max_read = max_size;
	q->num_allocated = q->max_size + q->max_read - 1;
	q->v = calloc(q->num_allocated ? q->num_allocated : 1, sizeof(float complex));
	return q;
}
void cbuffercf_destroy(cbuffercf q) { if(q) { free(q->v); free(q); } }
void cbuffercf_reset(cbuffercf q) { q->read_index = q->write_index = q->num_elements = 0; }
unsigned int cbuffercf_size(cbuffercf q) { return q->num_elements; }
unsigned int cbuffercf_max_size(cbuffercf q) { return q->max_size; }
unsigned int cbuffercf_space_available(cbuffercf q) { return q->max_size - q->num_elements; }
void cbuffercf_push(cbuffercf q, float complex v) {
	if(q->num_elements == q->max_size) return;
	q->v[q->write_index] = v;
	q->write_index = (q->write_index + 1) % q->max_size;
	q->num_elements++;
}
void cbuffercf_write(cbuffercf q, float complex *v, unsigned int n) {
	if(n > q->max_size - q->num_elements) return;
	q->num_elements += n;
	unsigned int k = q->max_size - q->write_index;      /* room before the wrap */
	if(n > k) {
		memmove(q->v + q->write_index, v, k * sizeof(float complex));
		memmove(q->v, v + k, (n - k) * sizeof(float complex));
		q->write_index = n - k;
	} else {
		memmove(q->v + q->write_index, v, n * sizeof(float complex));
		q->write_index = (q->write_index + n) % q->max_size;
	}
}
void cbuffercf_pop(cbuffercf q, float complex *v) {
	if(q->num_elements == 0) return;
	if(v) *v = q->v[q->read_index];
	q->read_index = (q->read_index + 1) % q->max_size;
	q->num_elements--;
}
void cbuffercf_read(cbuffercf q, unsigned int num_requested, float complex **v, unsigned int *num_read) {
	if(num_requested > q->num_elements) num_requested = q->num_elements;
	if(num_requested > q->max_read) num_requested = q->max_read;
	/* linearise: mirror the head of the storage behind its end when the requested run wraps */
	if(num_requested > q->max_size - q->read_index)
		memmove(q->v + q->max_size, q->v, (q->max_read - 1) * sizeof(float complex));
	*v = q->v + q->read_index;
	*num_read = num_requested;
}
void cbuffercf_release(cbuffercf q, unsigned int n) {
	if(n > q->num_elements) return;
	q->read_index = (q->read_index + n) % q->max_size;
	q->num_elements -= n;
}
