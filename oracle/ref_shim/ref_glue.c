/* oracle/ref_shim/ref_glue.c -- glue so the reference's own fastddc.c / libcsdr.c / fft.c link without fftw3f:
 * the five csdr_* FFT entry points of src/fft.h:23-28 (implemented in the reference by fft_fftw.c on fftw3f).
 * When libfftw3f.so.3 can be dlopen'ed on the machine (it is not installed in the build container) the calls go
 * to fftwf_plan_dft_1d / fftwf_execute exactly as fft_fftw.c:8-41 makes them (FFTW_ESTIMATE, nthreads); otherwise
 * they are served by the oracle's FFT (orc_fft, persistent worker pool).  Test infrastructure only. */
#define _GNU_SOURCE
#include <math.h>
#include <complex.h>
#include <stdlib.h>
#include <stdio.h>
#include <dlfcn.h>
#include "fft.h"
#include "fastddc.h"

void orc_fft(const float complex *in, float complex *out, int n, int dir);
void orc_fft_set_threads(int nthreads);
/* timing build only (REF_TIMING_BUILD, libref_fast.so): cache-blocked four-step FFT with a worker pool for the large
 * transforms, ref_shim/fft4step.c; the parity build keeps orc_fft so that every tap stays bit-identical with the oracle */
int fft4step(const float complex *in, float complex *out, int n, int dir);      /* 0: busy, nothing done */
void fft4step_set_threads(int nthreads);
int fft4step_usable(int n);

typedef void *(*fftwf_plan_dft_1d_t)(int, void *, void *, int, unsigned);
typedef void (*fftwf_execute_t)(void *);
typedef void (*fftwf_destroy_plan_t)(void *);
typedef int (*fftwf_init_threads_t)(void);
typedef void (*fftwf_plan_with_nthreads_t)(int);
static struct {
	int probed, ok, threads;
	fftwf_plan_dft_1d_t plan; fftwf_execute_t exec; fftwf_destroy_plan_t destroy;
	fftwf_init_threads_t init_threads; fftwf_plan_with_nthreads_t with_nthreads;
} F;

static void probe_fftw(void) {
	if(F.probed) return;
	F.probed = 1;
	if(getenv("ORC_NO_FFTW")) return;
	void *h = dlopen("libfftw3f.so.3", RTLD_NOW | RTLD_GLOBAL);
	if(!h) return;
	F.plan = (fftwf_plan_dft_1d_t)dlsym(h, "fftwf_plan_dft_1d");
	F.exec = (fftwf_execute_t)dlsym(h, "fftwf_execute");
	F.destroy = (fftwf_destroy_plan_t)dlsym(h, "fftwf_destroy_plan");
	if(!F.plan || !F.exec || !F.destroy) return;
	F.ok = 1;
	void *ht = dlopen("libfftw3f_threads.so.3", RTLD_NOW | RTLD_GLOBAL);
	if(ht) {
		F.init_threads = (fftwf_init_threads_t)dlsym(ht, "fftwf_init_threads");
		F.with_nthreads = (fftwf_plan_with_nthreads_t)dlsym(ht, "fftwf_plan_with_nthreads");
		if(F.init_threads && F.with_nthreads) F.threads = 1;
	}
}
/* 1: fftw3f (+2: with fftw3f_threads); 0: the oracle's FFT; 4: the four-step FFT of the timing build */
int ref_fft_backend(void) {
	probe_fftw();
	if(F.ok) return F.threads ? 3 : 1;
#ifdef REF_TIMING_BUILD
	return getenv("REF_NO_FFT4STEP") ? 0 : 4;           /* (A/B switch for the stand-in) */
#else
	return 0;
#endif
}

void csdr_fft_init(int32_t thread_cnt) {                 /* fft_fftw.c:8-14 */
	probe_fftw();
	if(F.ok && F.threads) { F.init_threads(); F.with_nthreads(thread_cnt); }
	orc_fft_set_threads(thread_cnt);
#ifdef REF_TIMING_BUILD
	fft4step_set_threads(thread_cnt);
#endif
}
void csdr_fft_destroy() {}
FFT_PLAN_T *csdr_make_fft_c2c(int32_t size, float complex *input, float complex *output, int32_t forward, int32_t benchmark) {
	(void)benchmark;
	probe_fftw();
	FFT_PLAN_T *p = calloc(1, sizeof(*p));
	p->size = size; p->input = input; p->output = output;
	if(F.ok) p->plan = F.plan(size, input, output, forward ? -1 : +1, 1u << 6);      /* FFTW_FORWARD / BACKWARD, FFTW_ESTIMATE (fft_fftw.c:25) */
	else p->plan = forward ? (void *)1 : (void *)2;
	return p;
}
void csdr_destroy_fft_c2c(FFT_PLAN_T *plan) {
	if(F.ok && plan) F.destroy(plan->plan);
	free(plan);
}
void csdr_fft_execute(FFT_PLAN_T *plan) {
	const int dir = plan->plan == (void *)1 ? +1 : -1;
	if(F.ok) F.exec(plan->plan);
#ifdef REF_TIMING_BUILD
	else if(fft4step_usable(plan->size) && ref_fft_backend() == 4 && fft4step(plan->input, plan->output, plan->size, dir)) return;
#endif
	else orc_fft(plan->input, plan->output, plan->size, dir);
}
/* direct entry for tests: one transform through whatever backend this build uses */
void ref_fft_run(float complex *in, float complex *out, int32_t n, int32_t forward) {
	FFT_PLAN_T *p = csdr_make_fft_c2c(n, in, out, forward, 0);
	csdr_fft_execute(p);
	csdr_destroy_fft_c2c(p);
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
