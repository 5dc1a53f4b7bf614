/* oracle/ref_shim/fft4step.c -- TEST INFRASTRUCTURE: the forward / inverse FFT of the TIMING build of the reference arm
 * (oracle/_ref/libref_fast.so) when libfftw3f is not on the machine.
 *
 * The reference calls fftwf_execute on plans of 2^18 .. 2^23 points with fftw3f_threads (fft_fftw.c:8-41, --fft-threads).
 * A stand-in that makes log4(N) streaming passes over a 32-64 MB array is memory-bound and does not scale with threads,
 * which would flatter the GPU arm; FFTW itself works in cache-sized pieces.  So does this one -- the classic four-step
 * decomposition N = N1 * N2:
 *     X[k1 + N1*k2] = sum_n2 W_N2^(n2*k2) * ( W_N^(n2*k1) * sum_n1 W_N1^(n1*k1) * x[N2*n1 + n2] )
 *   step 1  blocks of 16 columns n2: gather (128-byte runs), N1-point transforms in cache, twiddle, store as B[k1][n2]
 *   step 2  blocks of 16 rows k1: N2-point transforms in cache, store X[k1 + N1*k2] (128-byte runs)
 * i.e. two reads and two writes of the array, every sub-transform in L1/L2, the blocks spread over a persistent pool of
 * worker threads (the stand-in for fftw3f_threads).  Plain unnormalised DFT, sign -1 forward / +1 inverse; table entries
 * computed in double and rounded once, the inter-step twiddle is the product of two of them.  Only the timed CPU baseline uses it: the parity build (libref.so) and the oracle
 * keep orc_fft, so every bit-exact comparison is untouched.
 */
#define _GNU_SOURCE
#include <complex.h>
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <unistd.h>
#include <stdlib.h>
#include <string.h>

typedef float complex cf32;
#define COLS 16
#define MAXSUB 4096                /* N1, N2 <= 4096: N <= 2^24 */
#define MAXT 64

static struct {
	pthread_mutex_t m;
	pthread_cond_t go, done;
	pthread_barrier_t mid;
	pthread_t th[MAXT];
	int nth, started;              /* workers wanted / running */
	unsigned long gen;
	int pending;
	/* the transform in flight */
	const cf32 *in; cf32 *out, *B; const cf32 *Whi, *Wlo, *W1, *W2;
	int sh;
	int N, N1, N2, lg1, lg2, conj;
	int next1, next2;
	/* tables */
	cf32 *tw[25];                  /* W_(2^lg)^i, i < 2^lg (sub-transform sizes only) */
	cf32 *hi[25], *lo[25];         /* W_N^(n2*k1) = hi[idx >> sh] * lo[idx & (2^sh - 1)], sh = ceil(lg / 2): two cache-resident tables instead of one of N entries */
	cf32 *scratch; size_t scratch_n;
} G = { PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER };

static const cf32 *table(int lg) {
	if(!G.tw[lg]) {
		int n = 1 << lg;
		cf32 *w = malloc(sizeof(cf32) * (size_t)n);
		for(int i = 0; i < n; i++) { double a = -2.0 * M_PI * (double)i / (double)n; w[i] = (float)cos(a) + I * (float)sin(a); }
		G.tw[lg] = w;
	}
	return G.tw[lg];
}

static void split_tables(int lg) {
	if(G.hi[lg]) return;
	const int sh = (lg + 1) / 2, nlo = 1 << sh, nhi = 1 << (lg - sh);
	const double N = (double)(1 << lg);
	cf32 *lo = malloc(sizeof(cf32) * (size_t)nlo), *hi = malloc(sizeof(cf32) * (size_t)nhi);
	for(int i = 0; i < nlo; i++) { double a = -2.0 * M_PI * (double)i / N; lo[i] = (float)cos(a) + I * (float)sin(a); }
	for(int i = 0; i < nhi; i++) { double a = -2.0 * M_PI * (double)i * (double)nlo / N; hi[i] = (float)cos(a) + I * (float)sin(a); }
	G.lo[lg] = lo; G.hi[lg] = hi;
}

/* in-cache transform of n = 2^lg points: Stockham autosort, radix 4 (+ one radix-2 stage); result in a (b is scratch) */
static void small_fft(cf32 *a, cf32 *b, int lg, const cf32 *W, int conj) {
	int n = 1 << lg, len = n, s = 1;
	cf32 *x = a, *y = b;
	for(int rem = lg; rem > 0;) {
		int radix = rem >= 2 ? 4 : 2, ts = n / len;
		if(radix == 4) {
			int q1 = len / 4;
			for(int p = 0; p < q1; p++) {
				cf32 w1 = W[p * ts], w2 = W[2 * p * ts], w3 = W[3 * p * ts];
				if(conj) { w1 = conjf(w1); w2 = conjf(w2); w3 = conjf(w3); }
				const cf32 *xa = x + s * p, *xb = xa + s * q1, *xc = xb + s * q1, *xd = xc + s * q1;
				cf32 *y0 = y + s * 4 * p;
				for(int q = 0; q < s; q++) {
					cf32 A = xa[q], Bv = xb[q], Cv = xc[q], D = xd[q];
					cf32 apc = A + Cv, amc = A - Cv, bpd = Bv + D, bmd = Bv - D;
					cf32 j = conj ? (-cimagf(bmd) + I * crealf(bmd)) : (cimagf(bmd) - I * crealf(bmd));
					y0[q] = apc + bpd; y0[q + s] = w1 * (amc + j); y0[q + 2 * s] = w2 * (apc - bpd); y0[q + 3 * s] = w3 * (amc - j);
				}
			}
			rem -= 2;
		} else {
			int h = len / 2;
			for(int p = 0; p < h; p++) {
				cf32 w = W[p * ts];
				if(conj) w = conjf(w);
				const cf32 *xa = x + s * p, *xb = xa + s * h;
				cf32 *y0 = y + s * 2 * p;
				for(int q = 0; q < s; q++) { cf32 A = xa[q], Bv = xb[q]; y0[q] = A + Bv; y0[q + s] = (A - Bv) * w; }
			}
			rem -= 1;
		}
		len /= radix; s *= radix;
		cf32 *t = x; x = y; y = t;
	}
	if(x != a) memcpy(a, x, sizeof(cf32) * (size_t)n);
}

static void work(void) {
	const int N1 = G.N1, N2 = G.N2, conj = G.conj;
	static __thread cf32 *buf;                         /* per thread, kept: (COLS + 1) sub-transforms */
	if(!buf) buf = malloc(sizeof(cf32) * (size_t)(COLS + 1) * MAXSUB);
	cf32 *tmp = buf + (size_t)COLS * MAXSUB;
	for(;;) {                                          /* step 1: column blocks */
		int c0 = __atomic_fetch_add(&G.next1, COLS, __ATOMIC_RELAXED);
		if(c0 >= N2) break;
		for(int n1 = 0; n1 < N1; n1++) {
			const cf32 *src = G.in + (size_t)n1 * N2 + c0;
			for(int c = 0; c < COLS; c++) buf[(size_t)c * N1 + n1] = src[c];
		}
		for(int c = 0; c < COLS; c++) {
			cf32 *col = buf + (size_t)c * N1;
			small_fft(col, tmp, G.lg1, G.W1, conj);
			const size_t n2 = (size_t)(c0 + c);
			const unsigned mlo = (1u << G.sh) - 1u;
			for(int k1 = 0; k1 < N1; k1++) {
				const unsigned idx = (unsigned)(n2 * (size_t)k1);         /* < N */
				cf32 w = G.Whi[idx >> G.sh] * G.Wlo[idx & mlo];
				col[k1] *= conj ? conjf(w) : w;
			}
		}
		for(int k1 = 0; k1 < N1; k1++) {
			cf32 *dst = G.B + (size_t)k1 * N2 + c0;
			for(int c = 0; c < COLS; c++) dst[c] = buf[(size_t)c * N1 + k1];
		}
	}
	pthread_barrier_wait(&G.mid);
	for(;;) {                                          /* step 2: row blocks */
		int r0 = __atomic_fetch_add(&G.next2, COLS, __ATOMIC_RELAXED);
		if(r0 >= N1) break;
		for(int r = 0; r < COLS; r++) {
			cf32 *row = buf + (size_t)r * N2;
			memcpy(row, G.B + (size_t)(r0 + r) * N2, sizeof(cf32) * (size_t)N2);
			small_fft(row, tmp, G.lg2, G.W2, conj);
		}
		for(int k2 = 0; k2 < N2; k2++) {
			cf32 *dst = G.out + (size_t)k2 * N1 + r0;
			for(int r = 0; r < COLS; r++) dst[r] = buf[(size_t)r * N2 + k2];
		}
	}
}

static void *worker(void *arg) {
	/* one core per worker: woken threads otherwise start on the waker's core and wait for the load balancer, which costs
	 * more than the whole transform */
	const long t = (long)arg;
	cpu_set_t allowed;
	if(!getenv("REF_FFT_NO_PIN") && sched_getaffinity(0, sizeof(allowed), &allowed) == 0 && CPU_COUNT(&allowed) > 1) {
		int want = (int)((t + 1) % CPU_COUNT(&allowed));           /* the (t+1)-th CPU this process may run on (cgroup cpusets) */
		for(int c = 0; c < CPU_SETSIZE; c++) {
			if(!CPU_ISSET(c, &allowed)) continue;
			if(want-- == 0) {
				cpu_set_t cs;
				CPU_ZERO(&cs);
				CPU_SET(c, &cs);
				pthread_setaffinity_np(pthread_self(), sizeof(cs), &cs);
				break;
			}
		}
	}
	unsigned long seen = 0;
	for(;;) {
		pthread_mutex_lock(&G.m);
		while(G.gen == seen) pthread_cond_wait(&G.go, &G.m);
		seen = G.gen;
		pthread_mutex_unlock(&G.m);
		work();
		pthread_mutex_lock(&G.m);
		if(--G.pending == 0) pthread_cond_signal(&G.done);
		pthread_mutex_unlock(&G.m);
	}
	return NULL;
}

/* nthreads as --fft-threads; fixed by the first large transform (the pool's barrier is sized once) */
void fft4step_set_threads(int n) { if(!G.started) G.nth = n < 1 ? 1 : (n > MAXT ? MAXT : n); }

/* below 2^19 points (4 MB) the array sits in the last-level cache and the streaming FFT is as fast */
int fft4step_usable(int n) { return n >= (1 << 19) && n <= (1 << 24) && (n & (n - 1)) == 0; }

/* (the reference has one fft thread; fft_channelizer_create runs before the threads start) */
/* returns 0 without doing anything when another transform is in flight (the channel blocks are created by several threads
 * at once, each with its own tap-spectrum FFT: those callers run the streaming FFT on their own thread instead of queueing) */
int fft4step(const cf32 *in, cf32 *out, int N, int dir) {
	static pthread_mutex_t busy = PTHREAD_MUTEX_INITIALIZER;
	int lg = 0;
	while((1 << lg) < N) lg++;
	if(pthread_mutex_trylock(&busy) != 0) return 0;
	pthread_mutex_lock(&G.m);
	if(!G.started) {                                    /* nth - 1 workers: the calling thread is the nth */
		if(G.nth < 1) G.nth = 1;
		pthread_barrier_init(&G.mid, NULL, (unsigned)G.nth);
		for(int t = 0; t < G.nth - 1; t++) { pthread_create(&G.th[t], NULL, worker, (void *)(long)t); pthread_detach(G.th[t]); }
		G.started = 1;
	}
	G.lg1 = (lg + 1) / 2; G.lg2 = lg - G.lg1; G.N1 = 1 << G.lg1; G.N2 = 1 << G.lg2; G.N = N;
	split_tables(lg);
	G.sh = (lg + 1) / 2; G.Whi = G.hi[lg]; G.Wlo = G.lo[lg]; G.W1 = table(G.lg1); G.W2 = table(G.lg2);
	if(G.scratch_n < (size_t)N) { free(G.scratch); G.scratch = malloc(sizeof(cf32) * (size_t)N); G.scratch_n = (size_t)N; }
	G.B = G.scratch; G.in = in; G.out = out; G.conj = dir < 0;
	G.next1 = 0; G.next2 = 0; G.pending = G.nth - 1;
	G.gen++;
	pthread_cond_broadcast(&G.go);
	pthread_mutex_unlock(&G.m);
	work();
	pthread_mutex_lock(&G.m);
	while(G.pending > 0) pthread_cond_wait(&G.done, &G.m);
	pthread_mutex_unlock(&G.m);
	pthread_mutex_unlock(&busy);
	return 1;
}
