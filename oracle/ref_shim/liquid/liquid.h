/*
 * oracle/ref_shim/liquid/liquid.h -- DECLARATION-ONLY stand-in for <liquid/liquid.h> (jgaeddert/liquid-dsp,
 * un-vendored and not installed here) so that the reference's own src/hfdl.c, src/block.c, src/fft.c and
 * src/input-helpers.c compile where they lie into oracle/_ref/ (recipe: oracle/Makefile).  Only the subset of the
 * liquid-dsp 1.3.x C API those files call is declared; the symbols are served by ref_shim/liquid_shim.c on top of
 * the oracle's restated objects (oracle/orc_liquid.c).  TEST INFRASTRUCTURE ONLY -- never part of the product.
 */
#ifndef ORC_REF_SHIM_LIQUID_H
#define ORC_REF_SHIM_LIQUID_H
#include <complex.h>

/* reported library version: 1.3.2 -> hfdl.c:339-341 takes its "< 1.6.0" msequence branch.
 * Build with -DORC_LIQUID_VERSION=1006000 to exercise the other branch (same scrambler sequence). */
#ifndef ORC_LIQUID_VERSION
#define ORC_LIQUID_VERSION 1003002
#endif
#define LIQUID_VERSION_NUMBER ORC_LIQUID_VERSION
int liquid_libversion_number(void);

typedef enum { LIQUID_MODEM_UNKNOWN = 0, LIQUID_MODEM_PSK2, LIQUID_MODEM_PSK4, LIQUID_MODEM_PSK8, LIQUID_MODEM_BPSK = 39 } modulation_scheme;

typedef struct msresamp_crcf_s *msresamp_crcf;
msresamp_crcf msresamp_crcf_create(float r, float As);
void  msresamp_crcf_destroy(msresamp_crcf q);
float msresamp_crcf_get_delay(msresamp_crcf q);
void  msresamp_crcf_execute(msresamp_crcf q, float complex *x, unsigned int nx, float complex *y, unsigned int *ny);

typedef struct agc_crcf_s *agc_crcf;
agc_crcf agc_crcf_create(void);
void  agc_crcf_destroy(agc_crcf q);
void  agc_crcf_set_bandwidth(agc_crcf q, float bt);
void  agc_crcf_execute(agc_crcf q, float complex x, float complex *y);
float agc_crcf_get_signal_level(agc_crcf q);
float agc_crcf_get_gain(agc_crcf q);
float agc_crcf_get_rssi(agc_crcf q);
void  agc_crcf_unlock(agc_crcf q);

typedef struct firfilt_crcf_s *firfilt_crcf;
firfilt_crcf firfilt_crcf_create(float *h, unsigned int n);
void firfilt_crcf_destroy(firfilt_crcf q);
void firfilt_crcf_push(firfilt_crcf q, float complex x);
void firfilt_crcf_execute(firfilt_crcf q, float complex *y);

typedef struct eqlms_cccf_s *eqlms_cccf;
eqlms_cccf eqlms_cccf_create_lowpass(unsigned int n, float fc);
void eqlms_cccf_destroy(eqlms_cccf q);
void eqlms_cccf_reset(eqlms_cccf q);
void eqlms_cccf_set_bw(eqlms_cccf q, float mu);
void eqlms_cccf_push(eqlms_cccf q, float complex x);
void eqlms_cccf_execute(eqlms_cccf q, float complex *y);
void eqlms_cccf_step(eqlms_cccf q, float complex d, float complex d_hat);

typedef struct modem_s *modem;
modem modem_create(modulation_scheme scheme);
void  modem_destroy(modem q);
void  modem_demodulate(modem q, float complex x, unsigned int *sym);
float modem_get_demodulator_phase_error(modem q);
void  modem_demodulate_soft(modem q, float complex x, unsigned int *sym, unsigned char *soft_bits);

typedef struct symsync_crcf_s *symsync_crcf;
symsync_crcf symsync_crcf_create_kaiser(unsigned int k, unsigned int m, float beta, unsigned int M);
void symsync_crcf_destroy(symsync_crcf q);
void symsync_crcf_reset(symsync_crcf q);
void symsync_crcf_set_lf_bw(symsync_crcf q, float bt);
void symsync_crcf_set_output_rate(symsync_crcf q, unsigned int k_out);
void symsync_crcf_execute(symsync_crcf q, float complex *x, unsigned int nx, float complex *y, unsigned int *ny);

typedef struct bsequence_s *bsequence;
bsequence bsequence_create(unsigned int num_bits);
void bsequence_destroy(bsequence bs);
void bsequence_reset(bsequence bs);
void bsequence_init(bsequence bs, unsigned char *v);
void bsequence_push(bsequence bs, unsigned int bit);
int  bsequence_correlate(bsequence a, bsequence b);
unsigned int bsequence_get_length(bsequence bs);

typedef struct msequence_s *msequence;
msequence msequence_create(unsigned int m, unsigned int g, unsigned int a);
void msequence_destroy(msequence ms);
void msequence_reset(msequence ms);
unsigned int msequence_advance(msequence ms);

unsigned int count_bit_errors(unsigned int s1, unsigned int s2);

typedef struct cbuffercf_s *cbuffercf;
cbuffercf cbuffercf_create(unsigned int max_size);
void cbuffercf_destroy(cbuffercf q);
void cbuffercf_reset(cbuffercf q);
unsigned int cbuffercf_size(cbuffercf q);
unsigned int cbuffercf_max_size(cbuffercf q);
unsigned int cbuffercf_space_available(cbuffercf q);
void cbuffercf_write(cbuffercf q, float complex *v, unsigned int n);
void cbuffercf_read(cbuffercf q, unsigned int num_requested, float complex **v, unsigned int *num_read);
void cbuffercf_release(cbuffercf q, unsigned int n);
void cbuffercf_push(cbuffercf q, float complex v);
void cbuffercf_pop(cbuffercf q, float complex *v);

#endif
