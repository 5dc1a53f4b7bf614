/* oracle/ref_shim/ref_glue.c -- glue so the reference's own fastddc.c / libcsdr.c link without fftw3f:
 * the five csdr_* FFT entry points of src/fft.h:23-28 (implemented in the reference by fft_fftw.c on
 * fftw3f, which is not installed here) are served by the oracle's FFT.  Test infrastructure only. */
#include <math.h>
#include <complex.h>
#include <stdlib.h>
#include "fft.h"
#include "fastddc.h"

void orc_fft(const float complex *in, float complex *out, int n, int dir);

void csdr_fft_init(int32_t thread_cnt) { (void)thread_cnt; }
void csdr_fft_destroy() {}
FFT_PLAN_T *csdr_make_fft_c2c(int32_t size, float complex *input, float complex *output, int32_t forward, int32_t benchmark) {
	(void)benchmark;
	FFT_PLAN_T *p = calloc(1, sizeof(*p));
	p->size = size; p->input = input; p->output = output;
	p->plan = forward ? (void *)1 : (void *)2;
	return p;
}
void csdr_destroy_fft_c2c(FFT_PLAN_T *plan) { free(plan); }
void csdr_fft_execute(FFT_PLAN_T *plan) {
	orc_fft(plan->input, plan->output, plan->size, plan->plan == (void *)1 ? +1 : -1);
}
/* fastddc.c declares is_integer as a C99 'inline' without an external definition */
int32_t is_integer(float a) { return floorf(a) == a; }

/* small accessors so Python can drive the reference structs without knowing their layout */
int ref_sizeof_fastddc(void) { return (int)sizeof(fastddc_t); }
void ref_fastddc_fields(fastddc_t *d, int32_t *iv, float *fv) {
	iv[0] = d->pre_decimation; iv[1] = d->post_decimation; iv[2] = d->taps_length; iv[3] = d->taps_min_length;
	iv[4] = d->overlap_length; iv[5] = d->fft_size; iv[6] = d->fft_inv_size; iv[7] = d->input_size;
	iv[8] = d->post_input_size; iv[9] = d->startbin; iv[10] = d->v; iv[11] = d->offsetbin; iv[12] = d->scrap;
	fv[0] = d->pre_shift; fv[1] = d->post_shift; fv[2] = d->dsadata.sindelta; fv[3] = d->dsadata.cosdelta; fv[4] = d->dsadata.rate;
}
float complex *ref_channelizer_taps_fft(fft_channelizer c) { return c->filtertaps_fft; }
fastddc_t *ref_channelizer_ddc(fft_channelizer c) { return c->ddc; }
/* one block through the reference's own fastddc_inv_cc; returns output count */
int ref_channelizer_execute(fft_channelizer c, float complex *spectrum, float complex *out) {
	c->shift_status = fastddc_inv_cc(spectrum, out, c->ddc, c->inv_plan, c->filtertaps_fft, c->shift_status);
	return c->shift_status.output_size;
}
