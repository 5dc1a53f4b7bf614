// dumphfdl_b200/csrc/design.hpp -- host-side (create-time) design math of the front-end:
// block geometry (fastddc_init, fastddc.c:46-80), channel tap design (libcsdr.c:62-133) and the
// filter banks / tables of the liquid-dsp objects hfdl_channel_create builds (hfdl.c:468-521).
// Runs once per frontend on the CPU, exactly as the reference does its own init on the CPU; nothing
// here is on the per-sample path.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector>
#include <complex>
#include "demod_kernels.cuh"

namespace hfdl_design {

// ---------------- geometry (libcsdr.c:35-51,135-144; fastddc.c:46-80) ----------------
inline int32_t next_pow2(int32_t x) {                // smallest 2^i with x < 2^i
	for(int i = 0; i < 31; i++) { int32_t p = (int32_t)1 << i; if(x < p) return p; }
	return -1;
}
inline int32_t fft_decimation_rate(int32_t sample_rate, int32_t target) {
	return next_pow2((int32_t)floorf((float)sample_rate / (float)target)) / 2;
}
struct Geometry {
	int32_t decimation, pre_decimation, post_decimation;
	int32_t taps_length, taps_min_length, overlap_length, fft_size, fft_inv_size, input_size, post_input_size, scrap, v;
	float transition_bw;
};
struct ChannelGeom { int32_t startbin, offsetbin; float post_shift, dsa_rate, freq_shift; };

inline bool geometry_init(Geometry &g, int32_t sample_rate) {
	g.decimation = fft_decimation_rate(sample_rate, 1800 * 3);           // main.c:699
	g.transition_bw = (float)250 / (float)sample_rate;                   // main.c:704
	if(g.decimation < 1) return false;
	g.pre_decimation = 1; g.post_decimation = g.decimation;
	for(;;) {
		float half = (float)g.post_decimation / 2;
		if(floorf(half) != half || g.post_decimation / 2 == 1) break;
		g.post_decimation /= 2; g.pre_decimation *= 2;
	}
	int32_t tml = (int32_t)(4.0 / g.transition_bw);
	if(tml % 2 == 0) tml++;
	g.taps_min_length = tml;
	g.taps_length = next_pow2((int32_t)(ceil(tml / (float)g.pre_decimation) * g.pre_decimation)) + 1;
	g.fft_size = next_pow2(g.taps_length * 4);
	while(g.fft_size < g.pre_decimation) g.fft_size *= 2;
	g.overlap_length = g.taps_length - 1;
	g.input_size = g.fft_size - g.overlap_length;
	g.fft_inv_size = g.fft_size / g.pre_decimation;
	g.v = g.fft_size / g.overlap_length;
	g.scrap = g.overlap_length / g.pre_decimation;
	g.post_input_size = g.fft_inv_size - g.scrap;
	return g.fft_size > 2;
}
inline ChannelGeom channel_geom(const Geometry &g, int32_t sample_rate, int32_t centerfreq, int32_t freq) {
	ChannelGeom c;
	c.freq_shift = (float)(centerfreq - (freq + 1440)) / (float)sample_rate;          // hfdl.c:476
	int32_t middlebin = g.fft_size / 2;
	int32_t sb = (int32_t)(middlebin + middlebin * (-c.freq_shift) * 2);
	sb = (int32_t)(g.v * round(sb / (float)g.v));
	c.startbin = sb;
	c.offsetbin = sb - middlebin;
	c.post_shift = (g.pre_decimation) * (c.freq_shift + ((float)c.offsetbin / g.fft_size));
	float rate = c.post_shift * g.post_decimation;        // decimating_shift_addition_init, libcsdr_gpl.c:26-39
	rate *= 2;
	c.dsa_rate = rate;
	return c;
}

// ---------------- channel taps (firdes_bandpass_c with WINDOW_HAMMING, libcsdr.c:62-133) ----------------
inline void bandpass_taps(std::vector<std::complex<float>> &out, int32_t length, float lowcut, float highcut) {
	std::vector<float> real((size_t)length);
	float cutoff = (highcut - lowcut) / 2;
	int32_t middle = length / 2;
	auto hamming = [](float rate) -> float { rate = 0.5 + rate / 2; return 0.54 - 0.46 * cos(2 * M_PI * rate); };
	real[middle] = 2 * M_PI * cutoff * hamming(0);
	for(int32_t i = 1; i <= middle; i++)
		real[middle - i] = real[middle + i] = (sin(2 * M_PI * cutoff * i) / i) * hamming((float)i / middle);
	float sum = 0;
	for(int32_t i = 0; i < length; i++) sum += real[i];
	for(int32_t i = 0; i < length; i++) real[i] = real[i] / sum;
	float center = (highcut + lowcut) / 2;
	float phase = 0;
	out.resize((size_t)length);
	for(int32_t i = 0; i < length; i++) {
		float cv = cos(phase), sv = sin(phase);
		phase += 2 * M_PI * center;
		while(phase > 2 * M_PI) phase -= 2 * M_PI;
		while(phase < 0) phase += 2 * M_PI;
		out[i] = std::complex<float>(cv * real[i], sv * real[i]);
	}
}

// ---------------- liquid-dsp filter design (firdes.c liquid_firdes_kaiser, math.windows.c kaiser) ----------------
inline float besseli0f(float z) {
	if(z == 0.0f) return 1.0f;
	float y = 0.0f;
	for(int k = 0; k < 32; k++) { float t = k * logf(0.5f * z) - lgammaf((float)k + 1.0f); y += expf(2 * t); }
	return y;
}
inline void firdes_kaiser(int n, float fc, float As, float mu, float *h) {
	As = fabsf(As);
	float beta = As > 50.0f ? 0.1102f * (As - 8.7f) : (As > 21.0f ? 0.5842f * powf(As - 21, 0.4f) + 0.07886f * (As - 21) : 0.0f);
	for(int i = 0; i < n; i++) {
		float t = (float)i - (float)(n - 1) / 2 + mu;
		float x = 2.0f * fc * t;
		float h1 = fabsf(x) < 0.01f ? cosf(M_PI * x / 2.0f) * cosf(M_PI * x / 4.0f) * cosf(M_PI * x / 8.0f) : sinf(M_PI * x) / (M_PI * x);
		float r = 2.0f * t / (float)n;
		float h2 = besseli0f(beta * sqrtf(1 - r * r)) / besseli0f(beta);
		h[i] = h1 * h2;
	}
}

// msresamp_crcf(rate, 60 dB) -> resamp_crcf(rate, m=7, fc=min(0.515 rate, 0.49), As, npfb=256)
inline void resamp_design(float rate, float *h /*[256][14]*/, uint32_t *step) {
	const int npfb = 256, m = 7, sub = 2 * m, n = 2 * m * npfb + 1;
	std::vector<float> hf((size_t)n);
	float fc = 0.515f * rate;
	if(fc > 0.49f) fc = 0.49f;
	firdes_kaiser(n, fc / (float)npfb, 60.0f, 0.0f, hf.data());
	float gain = 0.0f;
	for(int i = 0; i < n; i++) gain += hf[i];
	gain = (float)npfb / gain;
	for(int i = 0; i < npfb; i++) for(int k = 0; k < sub; k++) h[i * sub + k] = hf[i + k * npfb] * gain;
	*step = (uint32_t)roundf((float)(1 << 24) / rate);
}

inline uint32_t gray_decode(uint32_t s) { uint32_t m = s >> 1; while(m) { s ^= m; m >>= 1; } return s; }

inline void demod_tables(DemodTables &T) {
	memset(&T, 0, sizeof(T));
	static const double mf[HFDL_MF_TAPS] = {            // hfdl.c:148-154
		-0.0170974647427123, 0.01148231492068473, 0.03138375667422348, 0.009454398851680437,
		-0.04161644170893816, -0.06451564801420356, -0.005495792933327306, 0.1316404671361545,
		0.2759693160697777, 0.3375901874933208, 0.2759693160697777, 0.1316404671361545,
		-0.005495792933327306, -0.06451564801420356, -0.04161644170893816, 0.009454398851680437,
		0.03138375667422348, 0.01148231492068473, -0.0170974647427123 };
	for(int i = 0; i < HFDL_MF_TAPS; i++) T.mf[i] = (float)mf[i];
	// symsync_crcf_create_kaiser(k=3, m=3, beta, npfb=16): prototype + derivative filter, hfdl.c:503
	{
		enum { HL = 2 * HFDL_SS_NPFB * 3 * 3 + 1 };
		float Hf[HL], H[HL], dH[HL];
		float fc = 0.75f;
		firdes_kaiser(HL, fc / (float)(3 * HFDL_SS_NPFB), 40.0f, 0.0f, Hf);
		for(int i = 0; i < HL; i++) H[i] = Hf[i] * 2.0f * fc;
		float hdh_max = 0;
		for(int i = 0; i < HL; i++) {
			if(i == 0) dH[i] = H[i + 1] - H[HL - 1];
			else if(i == HL - 1) dH[i] = H[0] - H[i - 1];
			else dH[i] = H[i + 1] - H[i - 1];
			if(fabsf(H[i] * dH[i]) > hdh_max || i == 0) hdh_max = fabsf(H[i] * dH[i]);
		}
		for(int i = 0; i < HL; i++) dH[i] *= 0.06f / hdh_max;
		for(int f = 0; f < HFDL_SS_NPFB; f++) for(int n = 0; n < HFDL_SS_SUB; n++) {
			T.ss_mf[f][n] = H[f + n * HFDL_SS_NPFB];
			T.ss_dmf[f][n] = dH[f + n * HFDL_SS_NPFB];
		}
		float bt = 0.001f;                               // symsync_crcf_set_lf_bw, hfdl.c:504
		float alpha = 1.000f - bt, beta = 0.220f * bt, a = 0.500f, b = 0.495f;
		float A0 = 1.0f - a * alpha, A1 = -b * alpha;
		T.ss_b0 = beta / A0; T.ss_a1 = A1 / A0; T.ss_a2 = 0.0f / A0;
		T.ss_rate_adj = 0.5 * bt;
	}
	{   // eqlms_cccf_create_lowpass(15, 0.45), hfdl.c:495
		float h[HFDL_EQ_LEN];
		firdes_kaiser(HFDL_EQ_LEN, 0.45f, 40.0f, 0.0f, h);
		for(int i = 0; i < HFDL_EQ_LEN; i++) T.eq_h0[i] = make_float2(h[HFDL_EQ_LEN - 1 - i] * 2 * 0.45f, -0.0f);
	}
	// preamble templates hfdl.c:420-459, pushed in time order (newest bit = bit 0 of word 0)
	static const uint8_t A_octets[16] = { 0x5B, 0xBC, 0x74, 0x57, 0x03, 0xD9, 0x89, 0x39, 0xF2, 0x08, 0xD5, 0x36, 0x94, 0x2C, 0x32, 0xFE };
	static const uint8_t M1b[127] = {
		0,1,1,1,0,1,1,0,1,1,1,1,0,1,0,0,0,1,0,1,1,0,0,
		1,0,1,1,1,1,1,0,0,0,1,0,0,0,0,0,0,1,1,0,0,1,1,0,1,1,
		0,0,0,1,1,1,0,0,1,1,1,0,1,0,1,1,1,0,0,0,0,1,0,0,1,1,
		0,0,0,0,0,1,0,1,0,1,0,1,1,0,1,0,0,1,0,0,1,0,1,0,0,1,
		1,1,1,0,0,1,0,0,0,1,1,0,1,0,1,0,0,0,0,1,1,1,1,1,1,1 };
	static const int shifts[8] = { 72, 82, 113, 123, 61, 103, 93, 9 };
	auto push = [](unsigned *b, unsigned bit) {
		b[3] = ((b[3] << 1) | (b[2] >> 31)) & 0x7FFFFFFFu; b[2] = (b[2] << 1) | (b[1] >> 31);
		b[1] = (b[1] << 1) | (b[0] >> 31); b[0] = (b[0] << 1) | (bit & 1u);
	};
	for(int i = 0; i < 127; i++) push(T.A_bits, (A_octets[i >> 3] >> (7 - (i & 7))) & 1u);
	for(int s = 0; s < 8; s++) for(int j = 0; j < 127; j++) push(T.M1_bits[s], M1b[(shifts[s] + j) % 127]);
	// PSK constellations (liquid modem_create_psk): exp(j*2*pi/M*gray_decode(sym))
	for(int m = 2; m <= 3; m++) {
		int M = 1 << m;
		float alpha = (float)(M_PI / (float)M);
		for(int sym = 0; sym < M; sym++) {
			float ang = (float)gray_decode((uint32_t)sym) * 2 * alpha;
			T.psk[m][sym] = make_float2(cosf(ang), sinf(ang));
		}
	}
	// scrambler: 15-bit LFSR x^15+x+1, preset 0x6959, 120-symbol period (hfdl.c:300-347)
	{
		uint32_t v = 0x6959u, g = 0x8002u >> 1;
		for(int i = 0; i < 120; i++) {
			uint32_t b = (uint32_t)__builtin_parity(v & g);
			v = ((v << 1) | b) & 0x7fffu;
			T.scr[i] = (unsigned char)b;
		}
	}
	static const int ar[8] = { 1, 1, 2, 3, 1, 1, 2, 3 }, seg[8] = { 72, 72, 72, 72, 168, 168, 168, 168 };
	static const int cr[8] = { 4, 2, 2, 2, 4, 2, 2, 2 }, cs[8] = { 17, 17, 17, 17, 23, 23, 23, 23 };
	for(int i = 0; i < 8; i++) { T.mode_arity[i] = ar[i]; T.mode_segments[i] = seg[i]; T.mode_code_rate[i] = cr[i]; T.mode_col_shift[i] = cs[i]; }
}

inline int pdu_len(int M1) {
	static const int ar[8] = { 1, 1, 2, 3, 1, 1, 2, 3 }, seg[8] = { 72, 72, 72, 72, 168, 168, 168, 168 }, cr[8] = { 4, 2, 2, 2, 4, 2, 2, 2 };
	int bits = seg[M1] * 30 * ar[M1] / cr[M1];
	return bits / 8 + (bits % 8 ? 1 : 0);
}

inline void demod_state_init(DemodState &S, const DemodTables &T) {      // hfdl_channel_create, hfdl.c:485-521
	memset(&S, 0, sizeof(S));
	S.noise_floor = 1.0f;                               // hfdl.c:490
	S.ss_since_reset = 0;
	// symsync created (reset), then k_out = 2
	S.ss_rate = 1.5f; S.ss_del = 1.5f;
	for(int i = 0; i < HFDL_EQ_LEN; i++) S.eq_w[i] = T.eq_h0[i];
	S.fr_state = HF_A1; S.symbols_wanted = 1; S.cur_arity = 1; S.s_state = HS_EMIT_BITS;
}

}  // namespace hfdl_design
