// dumphfdl_b200/csrc/ddc_kernels.cuh -- sm_100a kernels for the overlap-save channeliser:
//   K0+K1  ingest (CU8/CS16/CF32 -> CF32, input-helpers.c:10-78) fused into the first pass of the
//          forward FFT of the overlap-save window (fft.c:49-59, fft_fftw.c:22-41)
//   K2+K3+K4  per-(channel, block) pass-band slice x tap spectrum, inverse FFT(M), 1/N, scrap,
//          phase rotation and post-decimation (fastddc.c:152-215, libcsdr_gpl.c:41-74)
//   K5     arbitrary resampler to 5400 Hz (msresamp_crcf, hfdl.c:472,676)
//
// Forward FFT layout.  N = L1*L2(*L3).  Pass p transforms axis p of the row-major view
// [L1][L2][L3] in place, so the spectrum ends up "digit-scrambled": natural bin
// k = k1 + L1*(k2 + L2*k3) lives at address (k1*L2 + k2)*L3 + k3.  No transposing pass is run:
// the only consumers are the channel extractor (gathers M bins per channel through
// fft_bin_addr()) and debug read-back.  fft_swap_sides (fastddc.c:102-112) is likewise folded
// into that gather.  Every global access of the passes is a >=128-byte contiguous run.
#pragma once
#include "common.cuh"

#define HFDL_TWN 4096          // twiddle table size: exp(-2*pi*i*k/4096); sub-FFT lengths <= 4096
#ifdef HFDL_FFT_THREADS_OVERRIDE
#define HFDL_FFT_THREADS HFDL_FFT_THREADS_OVERRIDE
#else
#define HFDL_FFT_THREADS 256
#endif
#define HFDL_RS_TAPS 14        // resamp_crcf sub-filter length (2*m, m = 7)
#define HFDL_RS_NPFB 256
#define HFDL_RS_HIST (HFDL_RS_TAPS - 1)

struct FftPlan {
	int N, lgN, P;
	int lgL[3];
	int natural;       // 1: the last pass writes the spectrum out of place in natural bin order (fft_last_pass_nat)
};

__host__ __device__ __forceinline__ int fft_bin_addr(const FftPlan &pl, int k) {
	if(pl.P == 1 || pl.natural) return k;
	int k1 = k & ((1 << pl.lgL[0]) - 1);
	int r = k >> pl.lgL[0];
	if(pl.P == 2) return (k1 << pl.lgL[1]) | r;
	int k2 = r & ((1 << pl.lgL[1]) - 1);
	int k3 = r >> pl.lgL[1];
	return (((k1 << pl.lgL[1]) | k2) << pl.lgL[2]) | k3;
}

__device__ __forceinline__ int brev_n(int x, int bits) { return (int)(__brev((unsigned)x) >> (32 - bits)); }

// In-place radix-2^2 DIT FFT over shared memory.  Element e of transform f lives at s[e*TP + f]
// and must have been stored in bit-reversed element order.  nf transforms of length L = 1<<lgL.
// dirsign = +1: forward (e^-j), -1: inverse (e^+j, unnormalised) -- FFTW_FORWARD/BACKWARD, fft_fftw.c:25.
__device__ __forceinline__ void smem_fft(cf *s, int lgL, int nf, int TP, const cf *__restrict__ tw, int inverse) {
	const int L = 1 << lgL;
	const int tid = threadIdx.x, nth = blockDim.x;
	int st = 0;
	for(; st + 1 < lgL; st += 2) {
		const int h = 1 << st;
		const int items = nf * (L >> 2);
		for(int w = tid; w < items; w += nth) {
			int f = w % nf, q = w / nf;
			int j = q & (h - 1);
			int i0 = ((q >> st) << (st + 2)) + j;
			cf w1 = __ldg(&tw[j * (HFDL_TWN >> (st + 1))]);
			cf w2 = __ldg(&tw[j * (HFDL_TWN >> (st + 2))]);
			if(inverse) { w1.y = -w1.y; w2.y = -w2.y; }
			cf *p = s + i0 * TP + f;
			const int hs = h * TP;
			cf x0 = p[0], x1 = p[hs], x2 = p[2 * hs], x3 = p[3 * hs];
			cf t1 = cmul(w1, x1), t3 = cmul(w1, x3);
			cf a = cadd(x0, t1), b = csub(x0, t1), c = cadd(x2, t3), d = csub(x2, t3);
			cf u = cmul(w2, c);
			cf wd = cmul(w2, d);
			// W_{4h}^{j+h} = -i * W_{4h}^j (forward), +i (inverse)
			cf v = inverse ? make_float2(-wd.y, wd.x) : make_float2(wd.y, -wd.x);
			p[0] = cadd(a, u);
			p[hs] = cadd(b, v);
			p[2 * hs] = csub(a, u);
			p[3 * hs] = csub(b, v);
		}
		__syncthreads();
	}
	if(st < lgL) {      // odd log2: one radix-2 stage, h = L/2
		const int h = 1 << st;
		const int items = nf * (L >> 1);
		for(int w = tid; w < items; w += nth) {
			int f = w % nf, q = w / nf;
			int j = q & (h - 1);
			int i0 = ((q >> st) << (st + 1)) + j;
			cf w1 = __ldg(&tw[j * (HFDL_TWN >> (st + 1))]);
			if(inverse) w1.y = -w1.y;
			cf *p = s + i0 * TP + f;
			cf x0 = p[0], x1 = p[h * TP];
			cf t = cmul(w1, x1);
			p[0] = cadd(x0, t);
			p[h * TP] = csub(x0, t);
		}
		__syncthreads();
	}
}

// Forward sub-FFT for the passes: radix-2^2 DIF over a shared-memory tile of `nf` = 1<<lgT transforms.
// Element e of transform f lives at s[tile_idx(e, f)], a row-major [L][nf] tile whose column is XOR-swizzled
// with the row so that natural-order rows, bit-reversed rows and whole-row accesses are all bank-conflict free.
// Input in natural order, output in bit-reversed order (callers read row brev(k) to get bin k).
__device__ __forceinline__ int tile_idx(int row, int col, int lgT) {
	const int mask = (1 << lgT) - 1;
	return (row << lgT) + ((col ^ row ^ (row >> 4)) & mask);
}

__device__ __forceinline__ cf cmul_w8(cf v) {       // v * W_8^1 = v * (1 - i) / sqrt(2)
	const float c = 0.70710678118654752f;
	return make_float2(c * (v.x + v.y), c * (v.y - v.x));
}
__device__ __forceinline__ cf cmul_w83(cf v) {      // v * W_8^3 = v * (-1 - i) / sqrt(2)
	const float c = 0.70710678118654752f;
	return make_float2(c * (v.y - v.x), -c * (v.x + v.y));
}
__device__ __forceinline__ cf cmul_mi(cf v) { return make_float2(v.y, -v.x); }      // v * (-i)

__device__ __forceinline__ void smem_fft_dif(cf *s, int lgL, int lgT, const cf *__restrict__ tw) {
	const int L = 1 << lgL, nf = 1 << lgT;
	const int tid = threadIdx.x, nth = blockDim.x;
	int st = lgL;                       // current sub-transform size is 1 << st
	const int rem = lgL % 3;            // radix-8 register steps; the 1 or 2 left-over bits go first
	if(rem == 1) {                      // one radix-2 stage of size L
		const int h = L >> 1;
		const int items = nf * h;
		for(int w = tid; w < items; w += nth) {
			const int f = w & (nf - 1), j = w >> lgT;
			const cf w1 = __ldg(&tw[j * (HFDL_TWN >> lgL)]);
			cf *p0 = s + tile_idx(j, f, lgT), *p1 = s + tile_idx(j + h, f, lgT);
			const cf x0 = *p0, x1 = *p1;
			*p0 = cadd(x0, x1);
			*p1 = cmul(w1, csub(x0, x1));
		}
		__syncthreads();
		st -= 1;
	} else if(rem == 2) {               // one radix-2^2 stage of size L = 4h
		const int lgh = st - 2, h = 1 << lgh;
		const int items = nf * (L >> 2);
		for(int w = tid; w < items; w += nth) {
			const int f = w & (nf - 1), q = w >> lgT;
			const int j = q & (h - 1);
			const int i0 = ((q >> lgh) << st) + j;
			const cf wa = __ldg(&tw[j * (HFDL_TWN >> st)]);            // W_{4h}^j
			const cf wb = __ldg(&tw[j * (HFDL_TWN >> (st - 1))]);      // W_{2h}^j
			cf *p0 = s + tile_idx(i0, f, lgT), *p1 = s + tile_idx(i0 + h, f, lgT);
			cf *p2 = s + tile_idx(i0 + 2 * h, f, lgT), *p3 = s + tile_idx(i0 + 3 * h, f, lgT);
			const cf x0 = *p0, x1 = *p1, x2 = *p2, x3 = *p3;
			// size-4h stage: (x0,x2) with W^j, (x1,x3) with W^(j+h) = -i W^j
			const cf a0 = cadd(x0, x2), d02 = csub(x0, x2), a1 = cadd(x1, x3), d13 = csub(x1, x3);
			const cf a2 = cmul(wa, d02);
			const cf t13 = cmul(wa, d13);
			const cf a3 = make_float2(t13.y, -t13.x);
			// size-2h stage: (a0,a1) and (a2,a3) with W_{2h}^j
			*p0 = cadd(a0, a1);
			*p1 = cmul(wb, csub(a0, a1));
			*p2 = cadd(a2, a3);
			*p3 = cmul(wb, csub(a2, a3));
		}
		__syncthreads();
		st -= 2;
	}
	// Radix-8 steps: a work item takes the 8 elements i0 + j*m (m = size/8) of one sub-transform into registers,
	// runs the three radix-2 DIF stages there (outputs in bit-reversed order: register j holds bin brev3(j)),
	// applies the twiddle W_size^(b*brev3(j)) and stores back in place.  One shared-memory round trip per 3 stages.
	for(; st >= 3; st -= 3) {
		const int lgm = st - 3, m = 1 << lgm;
		const int items = nf * (L >> 3);
		const int tws = HFDL_TWN >> st;
		for(int w = tid; w < items; w += nth) {
			const int f = w & (nf - 1), q = w >> lgT;
			const int b = q & (m - 1);
			const int i0 = ((q >> lgm) << st) + b;
			cf x[8];
#pragma unroll
			for(int j = 0; j < 8; j++) x[j] = s[tile_idx(i0 + (j << lgm), f, lgT)];
			const cf u0 = cadd(x[0], x[4]), u1 = cadd(x[1], x[5]), u2 = cadd(x[2], x[6]), u3 = cadd(x[3], x[7]);
			const cf v0 = csub(x[0], x[4]), v1 = cmul_w8(csub(x[1], x[5])), v2 = cmul_mi(csub(x[2], x[6])), v3 = cmul_w83(csub(x[3], x[7]));
			const cf p0 = cadd(u0, u2), p1 = cadd(u1, u3), p2 = csub(u0, u2), p3 = cmul_mi(csub(u1, u3));
			const cf q0 = cadd(v0, v2), q1 = cadd(v1, v3), q2 = csub(v0, v2), q3 = cmul_mi(csub(v1, v3));
			cf y[8];
			y[0] = cadd(p0, p1); y[1] = csub(p0, p1); y[2] = cadd(p2, p3); y[3] = csub(p2, p3);      // bins 0 4 2 6
			y[4] = cadd(q0, q1); y[5] = csub(q0, q1); y[6] = cadd(q2, q3); y[7] = csub(q2, q3);      // bins 1 5 3 7
			if(lgm > 0) {
				const int bt = b * tws;
				y[1] = cmul(__ldg(&tw[4 * bt]), y[1]); y[2] = cmul(__ldg(&tw[2 * bt]), y[2]); y[3] = cmul(__ldg(&tw[6 * bt]), y[3]);
				y[4] = cmul(__ldg(&tw[bt]), y[4]); y[5] = cmul(__ldg(&tw[5 * bt]), y[5]); y[6] = cmul(__ldg(&tw[3 * bt]), y[6]);
				y[7] = cmul(__ldg(&tw[7 * bt]), y[7]);
			}
#pragma unroll
			for(int j = 0; j < 8; j++) s[tile_idx(i0 + (j << lgm), f, lgT)] = y[j];
		}
		__syncthreads();
	}
}

// Source of the wideband stream for the first pass: a cyclic device buffer of raw samples.
// Window b starts at stream position pos0 + b*block_stride; positions < 0 read as zero (the
// reference's first window has a zeroed overlap, fft.c:79), others wrap modulo ring_len.
struct RawSource {
	const void *base;
	long long ring_len;        // samples
	long long pos0;            // stream position of window 0, element 0 (may be negative)
	long long ring_origin;     // ring index that holds stream position 0 (mod ring_len)
	long long block_stride;    // input_size
	int sfmt;
};

// per-CTA view of one window: ring index of its element 0 (64-bit modulo once per CTA, not per sample)
struct WindowView { long long start_ring; long long first_valid; long long ring_len; const void *base; int sfmt; };

__device__ __forceinline__ WindowView window_view(const RawSource &src, int b) {
	WindowView v;
	const long long pos = src.pos0 + (long long)b * src.block_stride;      // stream position of element 0
	v.first_valid = pos < 0 ? -pos : 0;                                     // elements before this index read as zero
	long long r = (pos + src.ring_origin) % src.ring_len;
	if(r < 0) r += src.ring_len;
	v.start_ring = r; v.ring_len = src.ring_len; v.base = src.base; v.sfmt = src.sfmt;
	return v;
}

// one raw sample -> CF32.  The division by full_scale (input-helpers.c:10-78) is a multiplication by its reciprocal, as
// in the reference's own build (-ffast-math implies -freciprocal-math, src/CMakeLists.txt:39-42).
__device__ __forceinline__ cf load_raw(const void *base, int sfmt, long long r) {
	if(sfmt == HFDL_SFMT_CF32) {
		return reinterpret_cast<const cf *>(base)[r];                     // full_scale 1.0
	} else if(sfmt == HFDL_SFMT_CS16) {
		short2 q = reinterpret_cast<const short2 *>(base)[r];
		const float rfs = 1.0f / 32767.5f;                                // SHRT_MAX + 0.5 (input-helpers.c:116)
		return make_float2((float)q.x * rfs, (float)q.y * rfs);
	} else {
		uchar2 q = reinterpret_cast<const uchar2 *>(base)[r];
		const float rfs = 1.0f / 127.0f, shift = 63.5f;                   // input-helpers.c:54,110
		return make_float2(((float)q.x - shift) * rfs, ((float)q.y - shift) * rfs);
	}
}
// all elements [n_lo, n_hi) of the window are real samples that lie in the ring without wrapping: they can be read
// through one base index (no per-element zero-fill test, no wrap test, 32-bit offsets)
__device__ __forceinline__ bool window_plain(const WindowView &v, long long n_lo, long long n_hi) {
	return n_lo >= v.first_valid && v.start_ring + n_hi <= v.ring_len;
}

__device__ __forceinline__ cf load_window(const WindowView &v, long long n) {      // n in [0, N), N <= ring_len
	if(n < v.first_valid) return make_float2(0.f, 0.f);
	long long r = v.start_ring + n;
	if(r >= v.ring_len) r -= v.ring_len;
	return load_raw(v.base, v.sfmt, r);
}

// Column pass: FFT along an axis of stride 'inner' for T adjacent inner positions, then twiddle
// W_{L*inner}^{k*n_rest}.  grid = (outer*inner/T, B).  FIRST: read the raw stream instead of 'work'.
struct ColPassArgs {
	RawSource src;
	cf *work;
	const cf *tw;
	int N, lgL, inner, lgInner, T, lgT, first;
};

__global__ void __launch_bounds__(HFDL_FFT_THREADS) fft_col_pass(ColPassArgs a) {
	HFDL_DYN_SMEM(cf, s);
	const int L = 1 << a.lgL, lgT = a.lgT, T = 1 << lgT;
	const int tiles_per_row = a.inner >> lgT;
	const int o = blockIdx.x / tiles_per_row;
	const int n0 = (blockIdx.x - o * tiles_per_row) << lgT;
	const int b = blockIdx.y;
	const long long base = (long long)o * L * a.inner + n0;
	const int total = L << lgT;
	cf *wk = a.work + (long long)b * a.N + base;
	if(a.first) {
		const WindowView wv = window_view(a.src, b);
		for(int idx = threadIdx.x; idx < total; idx += blockDim.x) {
			const int t = idx & (T - 1), e = idx >> lgT;
			s[tile_idx(e, t, lgT)] = load_window(wv, base + (long long)e * a.inner + t);
		}
	} else {
		for(int idx = threadIdx.x; idx < total; idx += blockDim.x) {
			const int t = idx & (T - 1), e = idx >> lgT;
			s[tile_idx(e, t, lgT)] = wk[(long long)e * a.inner + t];
		}
	}
	__syncthreads();
	smem_fft_dif(s, a.lgL, lgT, a.tw);
	// bin k sits in row brev(k); each thread walks k with a constant stride, so its inter-pass twiddle
	// W_{L*inner}^{k*(n0+t)} advances by a constant rotation (re-seeded from sincospif every 8 steps)
	const float inv_np = 1.0f / (float)(L * a.inner);    // power of two: exact
	const int t = threadIdx.x & (T - 1);
	const int kstep = blockDim.x >> lgT;
	const float c = (float)(n0 + t);
	float ssn, scs;
	sincospif(-2.0f * (float)kstep * c * inv_np, &ssn, &scs);          // kstep*c < L*inner <= 2^23: exact argument
	const cf rot = make_float2(scs, ssn);
	cf twd = make_float2(1.f, 0.f);
	int it = 0;
	for(int k = threadIdx.x >> lgT; k < L; k += kstep, it++) {
		if((it & 7) == 0) {
			float sn, cs;
			sincospif(-2.0f * (float)k * c * inv_np, &sn, &cs);
			twd = make_float2(cs, sn);
		}
		const cf v = cmul(s[tile_idx(brev_n(k, a.lgL), t, lgT)], twd);
		wk[(long long)k * a.inner + t] = v;
		twd = cmul(twd, rot);
	}
}

// Row pass (last pass): R adjacent contiguous rows of length L.  grid = (N/L/R, B).
struct RowPassArgs {
	RawSource src;
	cf *work;
	const cf *tw;
	int N, lgL, R, lgR, first;
	cf *out;           // fft_last_pass_nat: natural-order spectrum [B][N]
	int L1, mid;       // fft_last_pass_nat: length of the first axis; product of the middle axes (1 for two-pass plans)
	const unsigned *mask;   // fft_last_pass_nat: one bit per granule of HFDL_SPEC_GRAN natural bins, set when some channel's
	                        // pass-band slice reads the granule; results of other granules are not stored (nullptr: store all)
};
#define HFDL_SPEC_LG_GRAN 5
#define HFDL_SPEC_GRAN (1 << HFDL_SPEC_LG_GRAN)

__global__ void __launch_bounds__(HFDL_FFT_THREADS) fft_row_pass(RowPassArgs a) {
	HFDL_DYN_SMEM(cf, s);
	const int L = 1 << a.lgL, lgR = a.lgR;
	const int b = blockIdx.y;
	const long long row0 = (long long)blockIdx.x << lgR;
	const int total = L << lgR;
	cf *wk = a.work + (long long)b * a.N + row0 * L;
	if(a.first) {
		const WindowView wv = window_view(a.src, b);
		for(int idx = threadIdx.x; idx < total; idx += blockDim.x) {
			const int e = idx & (L - 1), rr = idx >> a.lgL;
			s[tile_idx(e, rr, lgR)] = load_window(wv, row0 * L + idx);
		}
	} else {
		for(int idx = threadIdx.x; idx < total; idx += blockDim.x) {
			const int e = idx & (L - 1), rr = idx >> a.lgL;
			s[tile_idx(e, rr, lgR)] = wk[idx];
		}
	}
	__syncthreads();
	smem_fft_dif(s, a.lgL, lgR, a.tw);
	for(int idx = threadIdx.x; idx < total; idx += blockDim.x) {
		const int k = idx & (L - 1), rr = idx >> a.lgL;
		wk[idx] = s[tile_idx(brev_n(k, a.lgL), rr, lgR)];
	}
}

// ======================================================================================
// Register-resident passes (L = 32 * B, B = 2..16; every plan of BASELINE configs 1-5).
// A pass transforms one axis of length L for a tile of T = 8192 / L columns (or rows).  Each of the 256 threads
//   1. loads 32 elements x[a] = X[a*B + b] of one (column, b) straight from global memory into registers,
//   2. runs a 32-point DIF there (five radix-2 stages with compile-time twiddles), applies W_L^(b*ka) and stores the
//      32 results to shared memory,                                                    -- the only __syncthreads --
//   3. reads B elements (fixed ka) back, runs the B-point DIF in registers and writes natural bin k = ka + 32*kb
//      to global memory (column pass: times the inter-pass twiddle W^(k * column)).
// Versus the shared-memory passes above: one barrier instead of four, ~2.5x fewer instructions per element, 32
// independent global loads in flight per thread.
// ======================================================================================
#ifndef HFDL_FFT_REG_MINB
#define HFDL_FFT_REG_MINB 3          // CTAs per SM the register-resident passes are compiled for
#endif
__host__ __device__ __forceinline__ constexpr float w32_cos(int m) {      // cos(2*pi*m/32), m = 0..15
	return m == 0 ? 1.0f : m == 1 ? 0.98078528040323043f : m == 2 ? 0.92387953251128674f : m == 3 ? 0.83146961230254524f :
	       m == 4 ? 0.70710678118654752f : m == 5 ? 0.55557023301960218f : m == 6 ? 0.38268343236508978f : m == 7 ? 0.19509032201612825f :
	       m == 8 ? 0.0f : m == 9 ? -0.19509032201612825f : m == 10 ? -0.38268343236508978f : m == 11 ? -0.55557023301960218f :
	       m == 12 ? -0.70710678118654752f : m == 13 ? -0.83146961230254524f : m == 14 ? -0.92387953251128674f : -0.98078528040323043f;
}
__host__ __device__ __forceinline__ constexpr float w32_sin(int m) { return m <= 8 ? w32_cos(8 - m) : w32_cos(m - 8); }      // sin(2*pi*m/32)
__host__ __device__ __forceinline__ constexpr int brev_ct(int x, int bits) {
	int r = 0;
	for(int i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
	return r;
}
// d * W_32^m, W_32 = exp(-2*pi*i/32); m is a compile-time constant after unrolling
__device__ __forceinline__ cf cmul_w32(cf d, int m) {
	if(m == 0) return d;
	if(m == 8) return make_float2(d.y, -d.x);
	const float c = w32_cos(m), sn = w32_sin(m);
	return make_float2(d.x * c + d.y * sn, d.y * c - d.x * sn);
}
// in-place DIF of N = 2..32 points in registers; x[j] ends up holding bin brev(j)
template <int N> __device__ __forceinline__ void fft_reg_dif(cf *x) {
#pragma unroll
	for(int S = N; S >= 2; S >>= 1) {
		const int h = S >> 1;
#pragma unroll
		for(int g = 0; g < N; g += S) {
#pragma unroll
			for(int j = 0; j < h; j++) {
				const cf a = x[g + j], b = x[g + j + h];
				x[g + j] = cadd(a, b);
				x[g + j + h] = cmul_w32(csub(a, b), j * (32 / S));
			}
		}
	}
}

template <int LGB>
__global__ void __launch_bounds__(256, HFDL_FFT_REG_MINB) fft_col_pass_reg(ColPassArgs a) {
	HFDL_DYN_SMEM(cf, s);
	constexpr int B = 1 << LGB, L = 32 << LGB, LGT = 8 - LGB, T = 1 << LGT;
	const int tiles_per_row = a.inner >> LGT;
	const int o = blockIdx.x / tiles_per_row;
	const int n0 = (blockIdx.x - o * tiles_per_row) << LGT;
	const int blk = blockIdx.y;
	const long long base = (long long)o * L * a.inner + n0;
	cf *wk = a.work + (long long)blk * a.N + base;
	const int tid = threadIdx.x;
	{
		const int t = tid & (T - 1), b = tid >> LGT;
		cf x[32];
		if(a.first) {
			const WindowView wv = window_view(a.src, blk);
			if(window_plain(wv, base, base + (long long)(L - 1) * a.inner + T)) {      // CTA-uniform: all but the windows at the stream start / ring wrap
				const long long r0 = wv.start_ring + base + (long long)b * a.inner + t;
				const int step = B * a.inner;
				if(wv.sfmt == HFDL_SFMT_CF32) {
					const cf *q = reinterpret_cast<const cf *>(wv.base) + r0;
#pragma unroll
					for(int i = 0; i < 32; i++) x[i] = q[i * step];
				} else if(wv.sfmt == HFDL_SFMT_CS16) {
					const short2 *q = reinterpret_cast<const short2 *>(wv.base) + r0;
					short2 raw[32];
#pragma unroll
					for(int i = 0; i < 32; i++) raw[i] = q[i * step];
					const float rfs = 1.0f / 32767.5f;
#pragma unroll
					for(int i = 0; i < 32; i++) x[i] = make_float2((float)raw[i].x * rfs, (float)raw[i].y * rfs);
				} else {
#pragma unroll
					for(int i = 0; i < 32; i++) x[i] = load_raw(wv.base, wv.sfmt, r0 + (long long)i * step);
				}
			} else {
#pragma unroll
				for(int i = 0; i < 32; i++) x[i] = load_window(wv, base + (long long)(i * B + b) * a.inner + t);
			}
		} else {
#pragma unroll
			for(int i = 0; i < 32; i++) x[i] = wk[(long long)(i * B + b) * a.inner + t];
		}
		fft_reg_dif<32>(x);
		const int bt = b * (HFDL_TWN / L);
#pragma unroll
		for(int j = 0; j < 32; j++) {
			const int ka = brev_ct(j, 5);
			cf v = x[j];
			if(ka != 0) v = cmul(__ldg(&a.tw[ka * bt]), v);
			s[((ka << LGB) + b) * T + t] = v;
		}
	}
	__syncthreads();
	// inter-pass twiddle W^(k*c), W = exp(-2*pi*i / (L*inner)), k = ka + 32*kb, c = column (fixed per thread: T divides
	// 256).  ka = ka0 + (256/T)*i walks with i, so W^(ka*c) = W^(ka0*c) * (W^((256/T)*c))^i: three sincospif per thread
	// (exactly representable arguments), then one complex multiply per step (<= 15 steps)
	const float inv_np = 1.0f / (float)((long long)L * a.inner);      // power of two: exact
	const int t = tid & (T - 1), ka0 = tid >> LGT;
	const float c = (float)(n0 + t);
	cf wka, wstep, pw[B];
	{
		float sn, cs;
		sincospif(-2.0f * (float)ka0 * c * inv_np, &sn, &cs);
		wka = make_float2(cs, sn);
		sincospif(-2.0f * (float)(256 >> LGT) * c * inv_np, &sn, &cs);
		wstep = make_float2(cs, sn);
		sincospif(-2.0f * 32.0f * c * inv_np, &sn, &cs);
		pw[0] = make_float2(1.f, 0.f);
		if(B > 1) pw[1] = make_float2(cs, sn);
#pragma unroll
		for(int q = 2; q < B; q++) pw[q] = cmul(pw[q >> 1], pw[q - (q >> 1)]);
	}
#pragma unroll
	for(int i = 0; i < 32 / B; i++) {
		const int ka = ka0 + (256 >> LGT) * i;
		cf z[B];
#pragma unroll
		for(int b = 0; b < B; b++) z[b] = s[((ka << LGB) + b) * T + t];
		fft_reg_dif<B>(z);
#pragma unroll
		for(int j = 0; j < B; j++) {
			const int kb = brev_ct(j, LGB);
			const cf w = cmul(wka, pw[kb]);
			wk[(long long)(ka + 32 * kb) * a.inner + t] = cmul(z[j], w);
		}
		wka = cmul(wka, wstep);
	}
}

// last pass: T contiguous rows of length L per CTA; shared-memory rows are padded to 33 elements
template <int LGB>
__global__ void __launch_bounds__(256, HFDL_FFT_REG_MINB) fft_row_pass_reg(RowPassArgs a) {
	HFDL_DYN_SMEM(cf, s);
	constexpr int B = 1 << LGB, L = 32 << LGB, LGT = 8 - LGB;
	const int blk = blockIdx.y;
	const long long row0 = (long long)blockIdx.x << LGT;
	cf *wk = a.work + (long long)blk * a.N + row0 * L;
	const int tid = threadIdx.x;
	{
		const int b = tid & (B - 1), r = tid >> LGB;
		cf x[32];
		if(a.first) {
			const WindowView wv = window_view(a.src, blk);
#pragma unroll
			for(int i = 0; i < 32; i++) x[i] = load_window(wv, (row0 + r) * L + i * B + b);
		} else {
#pragma unroll
			for(int i = 0; i < 32; i++) x[i] = wk[(long long)r * L + i * B + b];
		}
		fft_reg_dif<32>(x);
		const int bt = b * (HFDL_TWN / L);
#pragma unroll
		for(int j = 0; j < 32; j++) {
			const int ka = brev_ct(j, 5);
			cf v = x[j];
			if(ka != 0) v = cmul(__ldg(&a.tw[ka * bt]), v);
			s[((r << LGB) + b) * 33 + ka] = v;
		}
	}
	__syncthreads();
#pragma unroll
	for(int i = 0; i < 32 / B; i++) {
		const int p = tid + 256 * i;
		const int ka = p & 31, r = p >> 5;
		cf z[B];
#pragma unroll
		for(int b = 0; b < B; b++) z[b] = s[((r << LGB) + b) * 33 + ka];
		fft_reg_dif<B>(z);
#pragma unroll
		for(int j = 0; j < B; j++) wk[(long long)r * L + ka + 32 * brev_ct(j, LGB)] = z[j];
	}
}

// Last pass, out of place, natural bin order.  The T rows of a tile are the rows of T CONSECUTIVE values of the first
// (fastest natural) digit k1 for one value of the middle digits, so that for every output bin of the row transform
// the tile's T results are T adjacent natural bins: each row is read as one contiguous run, each result group is
// written as one contiguous run, and the consumers (chan_extract, tap-slice gather) read plain contiguous slices.
// Shared memory: element (r, b, ka) at r*(33*B + 1) + 33*b + ka -- both the (b, r)-major writes of phase 1 and the
// (r, ka)-major reads of phase 2 are bank-conflict free.
template <int LGB>
__global__ void __launch_bounds__(256, HFDL_FFT_REG_MINB) fft_last_pass_nat(RowPassArgs a) {
	HFDL_DYN_SMEM(cf, s);
	constexpr int B = 1 << LGB, L = 32 << LGB, LGT = 8 - LGB, T = 1 << LGT, RS = 33 * B + 1;
	const int blk = blockIdx.y;
	const int tiles1 = a.L1 >> LGT;
	const int kmid = blockIdx.x / tiles1;                     // value of the middle digit(s)
	const int k10 = (blockIdx.x - kmid * tiles1) << LGT;      // first k1 of the tile
	const cf *wk = a.work + (long long)blk * a.N;
	cf *out = a.out + (long long)blk * a.N;
	const int tid = threadIdx.x;
	const long long kstride = (long long)a.L1 * a.mid;          // natural-index stride of the last digit
	// Which of this thread's 32 results are stored: result (i, j) is bin bin0 + kstride * (ka + 32 * kb).  The 32 lanes of
	// a warp hold 32 adjacent bins (r = lane, T >= 32) or whole groups of T, so a granule of HFDL_SPEC_GRAN = 32 bins is
	// decided by one mask bit; the 32 bits are fetched here, ahead of the row loads, instead of one dependent load in
	// front of every store.
	unsigned keep = 0xFFFFFFFFu;
	if(a.mask) {
		keep = 0u;
		const int r2 = tid & (T - 1), kaq = tid >> LGT;
		const long long bin0 = (k10 + r2) + (long long)a.L1 * kmid;
#pragma unroll
		for(int i = 0; i < 32 / B; i++) {
#pragma unroll
			for(int j = 0; j < B; j++) {
				const int ka = kaq + (256 >> LGT) * i;
				const unsigned g = (unsigned)((bin0 + kstride * (ka + 32 * brev_ct(j, LGB))) >> HFDL_SPEC_LG_GRAN);
				keep |= ((__ldg(&a.mask[g >> 5]) >> (g & 31u)) & 1u) << (i * B + j);
			}
		}
	}
	{
		const int b = tid & (B - 1), r = tid >> LGB;
		const cf *row = wk + ((long long)(k10 + r) * a.mid + kmid) * L;
		cf x[32];
#pragma unroll
		for(int i = 0; i < 32; i++) x[i] = row[i * B + b];
		fft_reg_dif<32>(x);
		const int bt = b * (HFDL_TWN / L);
#pragma unroll
		for(int j = 0; j < 32; j++) {
			const int ka = brev_ct(j, 5);
			cf v = x[j];
			if(ka != 0) v = cmul(__ldg(&a.tw[ka * bt]), v);
			s[r * RS + b * 33 + ka] = v;
		}
	}
	__syncthreads();
#pragma unroll
	for(int i = 0; i < 32 / B; i++) {
		if(!((keep >> (i * B)) & ((1u << B) - 1u))) continue;      // none of the B results of this group is read by any channel
		const int p = tid + 256 * i;
		const int r = p & (T - 1), ka = p >> LGT;
		cf z[B];
#pragma unroll
		for(int b = 0; b < B; b++) z[b] = s[r * RS + b * 33 + ka];
		fft_reg_dif<B>(z);
		const long long bin0 = (k10 + r) + (long long)a.L1 * kmid;
		cf *o = out + bin0;
#pragma unroll
		for(int j = 0; j < B; j++) {
			if((keep >> (i * B + j)) & 1u) o[kstride * (ka + 32 * brev_ct(j, LGB))] = z[j];
		}
	}
}

// Debug / init helper: gather natural-order bins [k0, k0+n) of window b out of the scrambled layout.
__global__ void fft_gather_bins(const cf *work, FftPlan pl, int b, int k0, int n, cf *out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i < n) out[i] = work[(long long)b * pl.N + fft_bin_addr(pl, (k0 + i) & (pl.N - 1))];
}

// Tap-slice extraction at init: tapslice[c][i'] = H_c[natural bin offsetbin_c + sgn(i')], i' in IFFT input order
// (fastddc.c:229-230 computes the full N-bin tap spectrum; only the M bins the slice fold uses are kept).
__global__ void tapslice_gather(const cf *work, FftPlan pl, int M, const int *offsetbin, int c0, cf *tapslice) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	int cl = blockIdx.y;      // window index within this FFT batch
	if(i < M) {
		int sg = i < M / 2 ? i : i - M;
		int k = (offsetbin[c0 + cl] + sg) & (pl.N - 1);
		tapslice[(long long)(c0 + cl) * M + i] = work[(long long)cl * pl.N + fft_bin_addr(pl, k)];
	}
}

// K2+K3+K4.  grid = (C, B), block = HFDL_FFT_THREADS, smem = M*8.
struct ChanArgs {
	const cf *work;            // [B][N] scrambled spectra
	const cf *slices;          // sharded spectrum: [B][C][M] pass-band slices in inverse-FFT input order (slice_pack); nullptr: read `work`
	const cf *tapslice;        // [C][M]
	const int *offsetbin;      // [C]
	const float *dsa_rate;     // [C] phase increment per output sample / pi (libcsdr_gpl.c:26-39)
	cf *bb;                    // [C][bb_stride] baseband stream, HFDL_RS_HIST history samples first
	const cf *tw;
	FftPlan pl;
	int M, lgM, scrap, post_dec, out_per_block;
	long long bb_stride;
	long long out_index0;      // global output-sample index of (block 0 of the batch, output 0)
	int block0;                // first block of this launch within the batch (sub-batches)
	float inv_norm;            // 1/(pre_decimation*M) = 1/N (fastddc.c:193)
};

__global__ void __launch_bounds__(HFDL_FFT_THREADS) chan_extract(ChanArgs a) {
	HFDL_DYN_SMEM(cf, s);
	const int c = blockIdx.x, b = blockIdx.y;
	const int off = a.offsetbin[c];
	const cf *W = a.work + (long long)b * a.pl.N;
	const cf *H = a.tapslice + (long long)c * a.M;
	if(a.slices) {
		const cf *S = a.slices + ((long long)b * gridDim.x + c) * a.M;
		for(int i = threadIdx.x; i < a.M; i += blockDim.x) s[brev_n(i, a.lgM)] = cmul(__ldg(&H[i]), S[i]);
	} else {
		for(int i = threadIdx.x; i < a.M; i += blockDim.x) {
			int sg = i < a.M / 2 ? i : i - a.M;
			int k = (off + sg) & (a.pl.N - 1);
			cf v = cmul(__ldg(&H[i]), W[fft_bin_addr(a.pl, k)]);
			s[brev_n(i, a.lgM)] = v;
		}
	}
	__syncthreads();
	smem_fft(s, a.lgM, 1, 1, a.tw, 1);
	const double cyc = 0.5 * (double)a.dsa_rate[c];     // cycles of phase per output sample
	cf *out = a.bb + (long long)c * a.bb_stride + HFDL_RS_HIST + (long long)(b + a.block0) * a.out_per_block;
	for(int j = threadIdx.x; j < a.out_per_block; j += blockDim.x) {
		long long g = a.out_index0 + (long long)(b + a.block0) * a.out_per_block + j;
		double fr = cyc * (double)g;
		fr -= floor(fr);
		float sn, cs;
		sincospif((float)(2.0 * fr), &sn, &cs);
		cf v = cscale(s[a.scrap + j * a.post_dec], a.inv_norm);
		out[j] = make_float2(cs * v.x - sn * v.y, sn * v.x + cs * v.y);
	}
}

// Sharded spectrum (multi-GPU, one rank transforms a share of the blocks for every rank's channels): the pass-band slice
// of channel j of the WHOLE job's channel list (M bins around offsetbin[j], inverse-FFT input order -- exactly what
// chan_extract multiplies with the tap slice) goes to the rank that owns the channel:
//   dst.base[owner = j % R] + ((dst.blk0 + block b) * (C_all / R) + local channel j / R) * M,   grid = (n_all, nb).
// dst.base[q] is either part q of a local send buffer (an all-to-all then delivers it) or rank q's receive buffer itself
// (peer memory in a one-process multi-GPU setup: the stores cross NVLink straight from this kernel, no copy follows).
// Either way every rank ends up with [all blocks of the batch][its channels][M].
#define HFDL_MAX_RANKS 16
struct SliceDst { cf *base[HFDL_MAX_RANKS]; long long blk0; };
__global__ void slice_pack(const cf *spec, FftPlan pl, int M, const int *offsetbin, int nranks, SliceDst dst) {
	const int j = blockIdx.x, b = blockIdx.y;
	const int owner = j % nranks, local = j / nranks, cper = gridDim.x / nranks;
	const int off = offsetbin[j];
	const cf *W = spec + (long long)b * pl.N;
	cf *out = dst.base[owner] + ((dst.blk0 + b) * cper + local) * M;
	for(int i = threadIdx.x; i < M; i += blockDim.x) {
		const int sg = i < M / 2 ? i : i - M;
		out[i] = W[fft_bin_addr(pl, (off + sg) & (pl.N - 1))];
	}
}

// moves the last HFDL_RS_HIST baseband samples of a batch in front of the next one
__global__ void bb_carry(cf *bb, long long bb_stride, long long n_new) {
	int c = blockIdx.x, t = threadIdx.x;
	cf v = make_float2(0.f, 0.f);
	cf *row = bb + (long long)c * bb_stride;
	if(t < HFDL_RS_HIST) v = row[n_new + t];
	__syncthreads();
	if(t < HFDL_RS_HIST) row[t] = v;
}

// K5: resamp_crcf with 24-bit fixed-point phase.  Output m of this batch is taken at
// tau = phi0 + m*step (2^-24 input samples): input index tau>>24, filter (tau & 0xFFFFFF)>>16.
struct ResampArgs {
	const cf *bb; long long bb_stride;
	cf *rs; long long rs_stride;
	const float *h;            // [256][14]
	unsigned long long phi0; unsigned step;
	int n_out;
};

__global__ void resamp_kernel(ResampArgs a) {
	int m = blockIdx.x * blockDim.x + threadIdx.x;
	int c = blockIdx.y;
	if(m >= a.n_out) return;
	unsigned long long tau = a.phi0 + (unsigned long long)m * a.step;
	long long i = (long long)(tau >> 24);
	int idx = (int)((tau & 0xFFFFFFull) >> 16);
	const cf *x = a.bb + (long long)c * a.bb_stride + HFDL_RS_HIST + i;
	const float *h = a.h + idx * HFDL_RS_TAPS;
	float re = 0.f, im = 0.f;
#pragma unroll
	for(int k = HFDL_RS_TAPS - 1; k >= 0; k--) {      // oldest sample first
		cf v = x[-k];
		float hk = __ldg(&h[k]);
		re += hk * v.x;
		im += hk * v.y;
	}
	a.rs[(long long)c * a.rs_stride + m] = make_float2(re, im);
}
