/*
 * oracle/orc_hfdl.c -- ORACLE (test infrastructure only, see orc.h).
 * Per-channel HFDL demodulator, framer and FEC restated from src/hfdl.c, plus the liquid-dsp
 * objects it calls (restated from the published liquid-dsp 1.3.2 algorithms; parity unpinned),
 * src/libfec/viterbi27_port.c and src/crc.c (both pinned against oracle/_ref in tests).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "orc.h"
#include "orc_liquid.h"

/* ---------------- tables from the reference text ---------------- */
const orc_mode_t orc_modes[8] = {           /* hfdl.c:81-138 */
	{ 1, ORC_SEG_SINGLE, 4, 17 }, { 1, ORC_SEG_SINGLE, 2, 17 }, { 2, ORC_SEG_SINGLE, 2, 17 }, { 3, ORC_SEG_SINGLE, 2, 17 },
	{ 1, ORC_SEG_DOUBLE, 4, 23 }, { 1, ORC_SEG_DOUBLE, 2, 23 }, { 2, ORC_SEG_DOUBLE, 2, 23 }, { 3, ORC_SEG_DOUBLE, 2, 23 },
};
const float orc_mf_taps[ORC_MF_TAPS] = {    /* hfdl.c:148-154 */
	-0.0170974647427123, 0.01148231492068473, 0.03138375667422348, 0.009454398851680437,
	-0.04161644170893816, -0.06451564801420356, -0.005495792933327306, 0.1316404671361545,
	0.2759693160697777, 0.3375901874933208, 0.2759693160697777, 0.1316404671361545,
	-0.005495792933327306, -0.06451564801420356, -0.04161644170893816, 0.009454398851680437,
	0.03138375667422348, 0.01148231492068473, -0.0170974647427123
};
const uint8_t orc_A_octets[16] = {          /* hfdl.c:420-437 */
	0x5B, 0xBC, 0x74, 0x57, 0x03, 0xD9, 0x89, 0x39, 0xF2, 0x08, 0xD5, 0x36, 0x94, 0x2C, 0x32, 0xFE
};
const uint8_t orc_M1_bits[127] = {          /* hfdl.c:441-447 */
	0,1,1,1,0,1,1,0,1,1,1,1,0,1,0,0,0,1,0,1,1,0,0,
	1,0,1,1,1,1,1,0,0,0,1,0,0,0,0,0,0,1,1,0,0,1,1,0,1,1,
	0,0,0,1,1,1,0,0,1,1,1,0,1,0,1,1,1,0,0,0,0,1,0,0,1,1,
	0,0,0,0,0,1,0,1,0,1,0,1,1,0,1,0,0,1,0,0,1,0,1,0,0,1,
	1,1,1,0,0,1,0,0,0,1,1,0,1,0,1,0,0,0,0,1,1,1,1,1,1,1
};
const int orc_M_shifts[8] = { 72, 82, 113, 123, 61, 103, 93, 9 };   /* hfdl.c:449 */

/* ====================================================================================
 * CRC (crc.c:4-47): reflected CRC-16, poly 0x8408; table regenerated, not copied
 * ==================================================================================== */
static uint16_t crc_tab[256];
static int crc_tab_ok;
static void crc_init(void) {
	for(int i = 0; i < 256; i++) {
		uint16_t c = (uint16_t)i;
		for(int b = 0; b < 8; b++) c = (c & 1) ? (uint16_t)((c >> 1) ^ 0x8408) : (uint16_t)(c >> 1);
		crc_tab[i] = c;
	}
	crc_tab_ok = 1;
}
uint16_t orc_crc16(const uint8_t *data, uint32_t len, uint16_t crc) {
	if(!crc_tab_ok) crc_init();
	while(len-- > 0) crc = (uint16_t)((crc >> 8) ^ crc_tab[(crc ^ *data++) & 0xff]);
	return crc;
}
int orc_fcs_check(const uint8_t *buf, uint32_t hdr_len) {   /* pdu.c:68-79 */
	uint16_t rx = (uint16_t)(buf[hdr_len] | (buf[hdr_len + 1] << 8));
	uint16_t calc = (uint16_t)(orc_crc16(buf, hdr_len, 0xFFFFu) ^ 0xFFFFu);
	return rx == calc;
}
/* "CRC-good PDU" = the frames.good events of mpdu.c:83-85 / spdu.c:62-64, dispatch rule pdu.c:104 */
int orc_pdu_crc_good(const uint8_t *buf, uint32_t len) {
	if(len < 1) return 0;
	if(buf[0] & 1) {                                   /* MPDU (pdu.c:104 IS_MPDU) */
		uint32_t hdr_len;
		if(buf[0] & 0x2) {                             /* downlink, mpdu.c:56-59 */
			uint32_t lpdu_cnt = (buf[0] >> 2) & 0xF;
			hdr_len = 6 + lpdu_cnt;
		} else {                                       /* uplink, mpdu.c:60-75 */
			uint32_t ac = ((buf[0] & 0x70) >> 4) + 1;
			hdr_len = 2;
			for(uint32_t i = 0; i < ac; i++) {
				if(len < hdr_len + 2) return 0;
				uint32_t lpdu_cnt = buf[hdr_len + 1] >> 4;
				hdr_len += 2 + lpdu_cnt;
			}
		}
		if(len < hdr_len + 2) return 0;                /* mpdu.c:77-81 */
		return orc_fcs_check(buf, hdr_len);
	}
	if(len < 66) return 0;                             /* spdu.c:12,55-59 */
	return orc_fcs_check(buf, 64u);
}

/* Front of pdu_decoder_thread restated: IS_MPDU split (pdu.c:104), MPDU header-length rule + FCS (mpdu.c:56-88),
 * LPDU walk (mpdu.c:90-119,136-159), LPDU length / FCS checks (lpdu.c:129-150), SPDU (spdu.c:55-70).
 * out: status (0 good, 1 bad_fcs, 2 too_short), direction (1 air2gnd), lpdus processed / good / bad_fcs / too_short */
void orc_pdu_front_parse(const uint8_t *buf, uint32_t len, int32_t out[6], uint64_t *good_mask) {
	int32_t status = 0, dir = 0, n = 0, ngood = 0, nbad = 0, nshort = 0;
	uint64_t mask = 0;
#define LPDU(ptr, ll) do { \
		if((ll) < 3) nshort++; \
		else if(orc_fcs_check((ptr), (ll) - 2)) { ngood++; if(n < 64) mask |= 1ull << n; } \
		else nbad++; \
		n++; } while(0)
	if(len < 1) status = 2;
	else if(buf[0] & 1) {
		uint32_t hdr_len, lpdu_cnt = 0, ac = 0;
		if(buf[0] & 0x2) { dir = 1; lpdu_cnt = (buf[0] >> 2) & 0xF; hdr_len = 6 + lpdu_cnt; }
		else {
			ac = ((buf[0] & 0x70) >> 4) + 1;
			hdr_len = 2;
			for(uint32_t i = 0; i < ac; i++) {
				if(len < hdr_len + 2) { status = 2; break; }
				lpdu_cnt = buf[hdr_len + 1] >> 4;
				hdr_len += 2 + lpdu_cnt;
			}
		}
		if(status == 0 && len < hdr_len + 2) status = 2;
		if(status == 0 && !orc_fcs_check(buf, hdr_len)) status = 1;
		if(status == 0) {
			const uint8_t *data = buf + hdr_len + 2, *end = buf + len;
			if(dir == 1) {
				const uint8_t *hp = buf + 6;
				for(uint32_t j = 0; j < lpdu_cnt; j++) {
					uint32_t ll = (uint32_t)*hp + 1;
					if(data + ll > end) break;
					LPDU(data, ll);
					data += ll; hp++;
				}
			} else {
				const uint8_t *hp = buf + 2;
				int stop = 0;
				for(uint32_t i = 0; i < ac && !stop; i++) {
					hp++;
					uint32_t cnt = (*hp++ >> 4) & 0xF;
					for(uint32_t j = 0; j < cnt; j++) {
						uint32_t ll = (uint32_t)hp[j] + 1;
						if(data + ll > end) { stop = 1; break; }
						LPDU(data, ll);
						data += ll;
					}
					hp += cnt;
				}
			}
		}
	} else {
		if(len < 66) status = 2;
		else if(!orc_fcs_check(buf, 64u)) status = 1;
	}
#undef LPDU
	out[0] = status; out[1] = dir; out[2] = n; out[3] = ngood; out[4] = nbad; out[5] = nshort;
	if(good_mask) *good_mask = mask;
}

/* ====================================================================================
 * Viterbi K=7 r=1/2 (libfec/viterbi27_port.c:65-79,92-134,147-221), polys 0x6d,0x4f (fec.h:13-14)
 * ==================================================================================== */
static int par32(uint32_t x) { return __builtin_parity(x); }

void orc_viterbi27(const uint8_t *syms, int nbits, uint8_t *out) {
	static uint8_t bt0[32], bt1[32];
	static int init;
	if(!init) {
		for(int s = 0; s < 32; s++) {
			bt0[s] = par32((2 * s) & 0x6d) ? 255 : 0;
			bt1[s] = par32((2 * s) & 0x4f) ? 255 : 0;
		}
		init = 1;
	}
	uint32_t m1[64], m2[64], *old = m1, *nw = m2;
	/* decisions: one 64-bit word per trellis step, plus 6 never-written (zero) tail steps */
	uint64_t *dec = calloc((size_t)nbits + 6, sizeof(uint64_t));
	for(int i = 0; i < 64; i++) old[i] = 63;
	old[0] = 0;
	for(int t = 0; t < nbits; t++) {
		uint32_t s0 = syms[2 * t], s1 = syms[2 * t + 1];
		uint64_t d = 0;
		for(int i = 0; i < 32; i++) {
			uint32_t metric = (bt0[i] ^ s0) + (bt1[i] ^ s1);
			uint32_t a = old[i] + metric, b = old[i + 32] + (510 - metric);
			uint32_t dd = (int32_t)(a - b) > 0;
			nw[2 * i] = dd ? b : a;
			d |= (uint64_t)dd << (2 * i);
			a = old[i] + (510 - metric);
			b = old[i + 32] + metric;
			dd = (int32_t)(a - b) > 0;
			nw[2 * i + 1] = dd ? b : a;
			d |= (uint64_t)dd << (2 * i + 1);
		}
		dec[t] = d;
		uint32_t *tmp = old; old = nw; nw = tmp;
	}
	/* chainback from state 0, reading the decision 6 steps ahead (viterbi27_port.c:105-134) */
	uint32_t endstate = 0;
	memset(out, 0, (size_t)((nbits + 7) / 8));
	for(int n = nbits - 1; n >= 0; n--) {
		uint32_t k = (uint32_t)((dec[n + 6] >> (endstate >> 2)) & 1);
		endstate = (endstate >> 1) | (k << 7);
		out[n >> 3] = (uint8_t)endstate;
	}
	free(dec);
}

/* encoder inverse of the above: sr=(sr<<1)|bit; c0=parity(sr&0x6d), c1=parity(sr&0x4f) */
void orc_conv_encode27(const uint8_t *bits, int nbits, uint8_t *chips) {
	uint32_t sr = 0;
	for(int i = 0; i < nbits; i++) {
		sr = ((sr << 1) | (bits[i] & 1)) & 0x7f;
		chips[2 * i] = (uint8_t)par32(sr & 0x6d);
		chips[2 * i + 1] = (uint8_t)par32(sr & 0x4f);
	}
}

/* ====================================================================================
 * scrambler: msequence (liquid < 1.6 convention selected by hfdl.c:339-341): m=15,
 * g = 0x8002 >> 1, state 0x6959; restarted every 120 symbols (hfdl.c:321-329).
 * ==================================================================================== */
uint32_t orc_scrambler_bits(uint8_t *out, int n) {
	uint32_t v = 0x6959u, g = 0x8002u >> 1;
	for(int i = 0; i < n; i++) {
		if(i % 120 == 0) v = 0x6959u;
		uint32_t b = (uint32_t)par32(v & g);
		v = ((v << 1) | b) & 0x7fffu;
		out[i] = (uint8_t)b;
	}
	return v;
}

/* ====================================================================================
 * decode_user_data (hfdl.c:993-1056) and its transmit-side inverse
 * ==================================================================================== */
int orc_pdu_len_octets(int M1) {
	const orc_mode_t *p = &orc_modes[M1];
	int bits = p->segments * ORC_DATA_FRAME_LEN * p->arity / p->code_rate;
	return bits / 8 + (bits % 8 != 0 ? 1 : 0);
}

int orc_decode_user_data(const cf32 *symbols, int M1, uint32_t bitmask, uint8_t *pdu_out, uint8_t *softbits_out) {
	const orc_mode_t *p = &orc_modes[M1];
	int num_symbols = p->segments * ORC_DATA_FRAME_LEN;
	int nenc = num_symbols * p->arity;
	int column_cnt = nenc / 40;
	uint8_t *table = calloc((size_t)nenc, 1);          /* [row][col], 40 rows */
	uint8_t *scr = malloc((size_t)num_symbols);
	orc_scrambler_bits(scr, num_symbols);
	orc_modem_t ms;
	int row = 0, col = 0;
	uint8_t soft[3];
	int si = 0;
	for(int i = 0; i < num_symbols; i++) {
		float flip = (scr[i] ? -1.0f : 1.0f) * ((bitmask & 1) ? -1.0f : 1.0f);
		orc_modem_demod_soft(p->arity, symbols[i] * flip, &ms, soft);
		for(int j = 0; j < p->arity; j++) {
			if(softbits_out) softbits_out[si] = soft[j];
			si++;
			/* deinterleaver_push (hfdl.c:387-399) */
			table[row * column_cnt + col] = soft[j];
			row++;
			if(row == 40) { row = 0; col++; }
			col -= p->col_shift;
			if(col < 0) col += column_cnt;
		}
	}
	int vin_len = nenc;
	if(p->code_rate == 4) vin_len /= 2;
	uint8_t *vin = malloc((size_t)vin_len);
	row = col = 0;
#define POP(dst) do { dst = table[row * column_cnt + col]; row = (row + 9) % 40; if(row == 0) col++; } while(0)
	if(p->code_rate == 4) {
		for(int i = 0; i < vin_len; i++) {
			uint8_t a, b;
			POP(a); POP(b);
			vin[i] = (uint8_t)((a & b) + ((a ^ b) >> 1));
		}
	} else {
		for(int i = 0; i < vin_len; i++) POP(vin[i]);
	}
#undef POP
	int out_bits = vin_len / 2;
	int out_octets = out_bits / 8 + (out_bits % 8 != 0 ? 1 : 0);
	orc_viterbi27(vin, out_bits, pdu_out);
	for(int i = 0; i < out_octets; i++) {              /* REVERSE_BYTE util.h:109 */
		uint8_t x = pdu_out[i], r = 0;
		for(int b = 0; b < 8; b++) r |= (uint8_t)(((x >> b) & 1) << (7 - b));
		pdu_out[i] = r;
	}
	free(table); free(scr); free(vin);
	return out_octets;
}

int orc_encode_user_data(const uint8_t *pdu, int M1, cf32 *symbols_out) {
	const orc_mode_t *p = &orc_modes[M1];
	int num_symbols = p->segments * ORC_DATA_FRAME_LEN;
	int nenc = num_symbols * p->arity;
	int column_cnt = nenc / 40;
	int vin_len = (p->code_rate == 4) ? nenc / 2 : nenc;
	int nbits = vin_len / 2;
	uint8_t *bits = calloc((size_t)nbits, 1);
	for(int i = 0; i < nbits; i++) bits[i] = (pdu[i >> 3] >> (i & 7)) & 1;     /* LSB first per octet */
	for(int i = nbits - 6; i < nbits; i++) bits[i] = 0;                         /* tail: decoder forces zeros */
	uint8_t *chips = malloc((size_t)vin_len);
	orc_conv_encode27(bits, nbits, chips);
	/* pop-order stream v[j]: r=1/4 repeats every chip twice */
	uint8_t *v = malloc((size_t)nenc);
	if(p->code_rate == 4) for(int j = 0; j < nenc; j++) v[j] = chips[j / 2];
	else memcpy(v, chips, (size_t)nenc);
	/* fill the table in pop order, read it in push order */
	uint8_t *table = malloc((size_t)nenc);
	int row = 0, col = 0;
	for(int j = 0; j < nenc; j++) {
		table[row * column_cnt + col] = v[j];
		row = (row + 9) % 40;
		if(row == 0) col++;
	}
	uint8_t *scr = malloc((size_t)num_symbols);
	orc_scrambler_bits(scr, num_symbols);
	row = col = 0;
	for(int i = 0; i < num_symbols; i++) {
		uint32_t sym = 0;
		for(int j = 0; j < p->arity; j++) {
			sym = (sym << 1) | table[row * column_cnt + col];
			row++;
			if(row == 40) { row = 0; col++; }
			col -= p->col_shift;
			if(col < 0) col += column_cnt;
		}
		cf32 x = (p->arity == 1) ? (sym ? -1.0f : 1.0f) : orc_psk_point(p->arity, sym);
		symbols_out[i] = scr[i] ? -x : x;
	}
	free(bits); free(chips); free(v); free(table); free(scr);
	return num_symbols;
}

/* ====================================================================================
 * the channel (struct hfdl_channel hfdl.c:203-244)
 * ==================================================================================== */
enum { S_EMIT_BITS = 1, S_EMIT_SYMBOLS = 2, S_SKIP = 3 };
enum { F_A1 = 1, F_A2, F_M1, F_M2_SKIP, F_EQ_TRAIN, F_DATA_1, F_DATA_2 };

typedef struct { uint64_t hi, lo; } bits128_t;     /* 127-bit window, newest bit = lo bit 0 */

struct orc_channel {
	orc_channelizer_t *chz;
	orc_resamp_t *rs;
	float resamp_rate;
	int32_t freq;
	/* agc_crcf */
	orc_agc_t agc;
	/* firfilt (matched filter) */
	orc_firfilt_t mf;
	orc_symsync_t ss;
	float c_alpha, c_beta, c_phi, c_dphi, c_err;   /* costas hfdl.c:250-294 */
	orc_eqlms_t eq;
	orc_modem_t modem;
	bits128_t bits, A_bs, M1_bs[8];
	cf32 training[ORC_T_LEN]; int training_n;
	cf32 *data_symbols; int data_n;
	int cur_buf;                   /* 0 = training, 1 = data */
	uint64_t symbol_cnt, sample_cnt;
	int s_state, fr_state, data_arity, cur_arity;
	int32_t symbols_wanted, search_retries, eq_train_seq_cnt, data_segment_cnt;
	int32_t train_bits_total, train_bits_bad, T_idx, M1;
	uint32_t bitmask, symsync_out_idx;
	float freq_err_hz, signal_level, noise_floor;
	uint64_t a2_sample_cnt;
	/* locals of hfdl_decoder_thread that persist across blocks (hfdl.c:603-612) */
	uint32_t nf_clk; float frame_symbol_cnt;
	/* outputs */
	orc_pdu_t *pdus; int npdu, cap_pdu;
	int32_t st_a1, st_a2, st_m1, st_frames, st_m1_fail;
	/* capture */
	uint32_t cap_mask; size_t cap_max; cf32 *cap[ORC_CAP_COUNT]; size_t cap_n[ORC_CAP_COUNT];
	cf32 *chan_out, *resampled;
};

static void cap_push(orc_channel_t *c, int tap, cf32 v) {
	if(!(c->cap_mask & (1u << tap))) return;
	if(c->cap_n[tap] < c->cap_max) c->cap[tap][c->cap_n[tap]] = v;
	c->cap_n[tap]++;
}

static void bits_push(bits128_t *b, uint32_t bit) {
	b->hi = ((b->hi << 1) | (b->lo >> 63)) & 0x7FFFFFFFFFFFFFFFull;   /* keep 127 bits */
	b->lo = (b->lo << 1) | (bit & 1);
}
static int bits_correlate(const bits128_t *a, const bits128_t *b) {   /* number of equal positions of 127 */
	uint64_t xh = (a->hi ^ b->hi) & 0x7FFFFFFFFFFFFFFFull, xl = a->lo ^ b->lo;
	return 127 - (__builtin_popcountll(xh) + __builtin_popcountll(xl));
}

static void sampler_reset(orc_channel_t *c) {          /* hfdl.c:968-972 */
	orc_symsync_reset(&c->ss);
	c->s_state = S_EMIT_BITS;
	c->bitmask = 0;
}
static void framer_reset(orc_channel_t *c) {           /* hfdl.c:974-991 */
	c->fr_state = F_A1;
	c->symbols_wanted = 1;
	c->search_retries = 0;
	c->cur_arity = 1;
	c->train_bits_total = c->train_bits_bad = 0;
	c->T_idx = 0;
	c->cur_buf = 0;
	/* agc_crcf_unlock: the AGC is never locked, no effect */
	orc_eqlms_reset(&c->eq);
	c->data_n = 0;
	c->training_n = 0;
	sampler_reset(c);
}

orc_channel_t *orc_channel_create(int32_t sample_rate, int32_t pre_dec, float tbw, int32_t centerfreq, int32_t frequency, int fold_mode) {
	orc_channel_t *c = calloc(1, sizeof(*c));
	c->resamp_rate = (float)(ORC_SYMBOL_RATE * ORC_SPS) / ((float)sample_rate / (float)pre_dec);   /* hfdl.c:471 */
	c->rs = orc_resamp_create(c->resamp_rate, 60.0f);
	c->freq = frequency;
	float freq_shift = orc_channel_shift_rate(sample_rate, centerfreq, frequency);
	c->chz = orc_channelizer_create(pre_dec, tbw, freq_shift, fold_mode);
	if(!c->chz || !c->rs) { free(c); return NULL; }
	orc_agc_init(&c->agc, 0.01f);                              /* hfdl.c:485-487 */
	orc_firfilt_init(&c->mf, orc_mf_taps, ORC_MF_TAPS);        /* hfdl.c:494 */
	c->noise_floor = 1.0f;                                     /* hfdl.c:490 */
	c->c_alpha = 0.1f;
	c->c_beta = 0.047f * c->c_alpha * c->c_alpha;              /* hfdl.c:254-259 */
	orc_eqlms_init_lowpass(&c->eq, 0.45f);                     /* hfdl.c:495 */
	c->eq.mu = 0.1f;                                           /* eqlms_cccf_set_bw, hfdl.c:496 */
	orc_symsync_init_kaiser(&c->ss);                           /* hfdl.c:503 */
	orc_symsync_set_lf_bw(&c->ss, 0.001f);                     /* hfdl.c:504 */
	orc_symsync_set_output_rate(&c->ss, 2);                    /* hfdl.c:505 */
	/* A and M1 templates, hfdl.c:438-459 (bits pushed in time order) */
	for(int i = 0; i < ORC_A_LEN; i++) bits_push(&c->A_bs, (orc_A_octets[i >> 3] >> (7 - (i & 7))) & 1);
	for(int s = 0; s < 8; s++)
		for(int j = 0; j < ORC_M1_LEN; j++) bits_push(&c->M1_bs[s], orc_M1_bits[(orc_M_shifts[s] + j) % ORC_M1_LEN]);
	c->data_symbols = calloc(ORC_DATA_SYMS_MAX, sizeof(cf32));
	framer_reset(c);
	c->chan_out = calloc((size_t)c->chz->ddc.post_input_size, sizeof(cf32));
	c->resampled = calloc((size_t)c->chz->ddc.post_input_size + 64, sizeof(cf32));
	return c;
}

void orc_channel_destroy(orc_channel_t *c) {
	if(!c) return;
	orc_channelizer_destroy(c->chz);
	orc_resamp_destroy(c->rs);
	for(int i = 0; i < ORC_CAP_COUNT; i++) free(c->cap[i]);
	free(c->data_symbols); free(c->chan_out); free(c->resampled); free(c->pdus); free(c);
}

void orc_channel_set_capture(orc_channel_t *c, uint32_t mask, size_t max_per_tap) {
	c->cap_mask = mask; c->cap_max = max_per_tap;
	for(int i = 0; i < ORC_CAP_COUNT; i++) {
		free(c->cap[i]); c->cap[i] = NULL; c->cap_n[i] = 0;
		if(mask & (1u << i)) c->cap[i] = calloc(max_per_tap, sizeof(cf32));
	}
}
size_t orc_channel_get_capture(orc_channel_t *c, int tap, cf32 *dst, size_t max) {
	size_t n = c->cap_n[tap] < c->cap_max ? c->cap_n[tap] : c->cap_max;
	if(n > max) n = max;
	if(dst && n) memcpy(dst, c->cap[tap], n * sizeof(cf32));
	return c->cap_n[tap];
}
const orc_ddc_t *orc_channel_ddc(orc_channel_t *c) { return &c->chz->ddc; }
float orc_channel_resamp_rate(orc_channel_t *c) { return c->resamp_rate; }
int orc_channel_pdu_count(orc_channel_t *c) { return c->npdu; }
int orc_channel_get_pdu(orc_channel_t *c, int idx, orc_pdu_t *out) {
	if(idx < 0 || idx >= c->npdu) return -1;
	*out = c->pdus[idx];
	return 0;
}
void orc_channel_stats(orc_channel_t *c, int32_t *a1, int32_t *a2, int32_t *m1, int32_t *frames) {
	*a1 = c->st_a1; *a2 = c->st_a2; *m1 = c->st_m1; *frames = c->st_frames;
}
int32_t orc_channel_m1_not_found(orc_channel_t *c) { return c->st_m1_fail; }
float orc_channel_noise_floor(orc_channel_t *c) { return c->noise_floor; }

static void dispatch_pdu(orc_channel_t *c, const uint8_t *buf, int len) {   /* hfdl.c:1058-1080 */
	if(c->npdu == c->cap_pdu) {
		c->cap_pdu = c->cap_pdu ? 2 * c->cap_pdu : 16;
		c->pdus = realloc(c->pdus, sizeof(orc_pdu_t) * (size_t)c->cap_pdu);
	}
	orc_pdu_t *p = &c->pdus[c->npdu++];
	memset(p, 0, sizeof(*p));
	const orc_mode_t *m = &orc_modes[c->M1];
	p->freq = c->freq; p->M1 = c->M1; p->len = len;
	p->freq_err_hz = c->freq_err_hz; p->signal_level = c->signal_level; p->noise_floor = c->noise_floor;
	p->bit_rate = ORC_SYMBOL_RATE * m->arity / m->code_rate * ORC_DATA_FRAME_LEN / (ORC_DATA_FRAME_LEN + ORC_T_LEN);
	p->slot = m->segments == ORC_SEG_SINGLE ? 'S' : 'D';
	p->sample_cnt_end = c->sample_cnt; p->sample_cnt_a2 = c->a2_sample_cnt;
	p->train_bits_bad = c->train_bits_bad; p->train_bits_total = c->train_bits_total;
	memcpy(p->octets, buf, (size_t)len);
	p->crc_good = orc_pdu_crc_good(buf, (uint32_t)len);
	c->st_frames++;
}

static void train_bit_errors(orc_channel_t *c) {       /* hfdl.c:952-966 */
	uint32_t T_seq = 0;
	orc_modem_t ms;
	for(int i = 0; i < ORC_T_LEN; i++) {
		uint32_t bit = orc_modem_demod(1, c->training[i], &ms);
		bit ^= (c->bitmask & 1);
		T_seq = (T_seq << 1) | bit;
	}
	c->train_bits_total += ORC_T_LEN;
	c->train_bits_bad += __builtin_popcount(ORC_T_WORD ^ T_seq);
	c->training_n = 0;
}

static const float T_sym[ORC_T_LEN] = { 1, 1, 1, -1, 1, 1, -1, -1, 1, -1, 1, -1, -1, -1, -1 };   /* hfdl.c:157-160 */

/* one sample at 5400 Hz: body of the k loop, hfdl.c:685-893 */
static void demod_sample(orc_channel_t *c, cf32 x) {
	/* agc_crcf_execute */
	cf32 r = orc_agc_execute(&c->agc, x);
	cap_push(c, ORC_CAP_AGC, r);
	/* matched filter */
	orc_firfilt_push(&c->mf, r);
	cf32 s = orc_firfilt_execute(&c->mf);
	cap_push(c, ORC_CAP_MF, s);
	/* noise floor hfdl.c:700-706 */
	if(c->fr_state == F_A1 && (++c->nf_clk & 0xFFu) == 0xFFu)
		c->noise_floor = 0.65f * c->noise_floor + 0.35f * fminf(c->noise_floor, orc_agc_signal_level(&c->agc)) + 1e-6f;
	cf32 symbols[8];
	int produced = orc_symsync_step(&c->ss, s, symbols);
	for(int i = 0; i < produced; i++, c->symsync_out_idx++) {
		/* costas_cccf_step + execute */
		c->c_phi += c->c_dphi;
		if(c->c_phi > M_PI) c->c_phi -= 2.0 * M_PI;
		else if(c->c_phi < -M_PI) c->c_phi += 2.0 * M_PI;
		r = symbols[i] * (cosf(c->c_phi) - I * sinf(c->c_phi));
		if(fabsf(c->c_dphi) > 0.25f && c->fr_state == F_A1) {
			c->c_phi = c->c_dphi = 0.f;
			orc_symsync_reset(&c->ss);
		}
		orc_eqlms_push(&c->eq, r);
		if(!(c->symsync_out_idx & 1)) continue;
		cap_push(c, ORC_CAP_SYMSYNC, symbols[i]);
		cap_push(c, ORC_CAP_COSTAS, r);
		s = orc_eqlms_execute(&c->eq);
		if(c->fr_state == F_EQ_TRAIN) {
			float d = T_sym[c->T_idx];
			if(c->bitmask & 1) d = -d;
			orc_eqlms_step(&c->eq, d, s);
			c->T_idx++;
		}
		cap_push(c, ORC_CAP_EQ, s);
		uint32_t bits = orc_modem_demod(c->cur_arity, s, &c->modem);
		/* costas_cccf_adjust */
		float err = orc_modem_phase_error(&c->modem);
		err = 0.5 * (fabsf(err + 1.0f) - fabsf(err - 1.0f));
		c->c_err = err;
		c->c_phi += c->c_alpha * err;
		c->c_dphi += c->c_beta * err;

		c->symbol_cnt++;
		if(c->symbol_cnt >= 13u * ORC_SINGLE_SLOT_FRAME_LEN && c->fr_state == F_A1) {
			c->symbol_cnt = 0;
			c->c_phi = c->c_dphi = 0.f;
			orc_symsync_reset(&c->ss);
		}
		if(c->s_state == S_EMIT_BITS) {
			bits ^= c->bitmask;
			for(int b = 0; b < c->cur_arity; b++, bits >>= 1) bits_push(&c->bits, bits);
		} else if(c->s_state == S_EMIT_SYMBOLS) {
			if(c->cur_buf == 0) { if(c->training_n < ORC_T_LEN) c->training[c->training_n++] = s; }
			else { if(c->data_n < ORC_DATA_SYMS_MAX) c->data_symbols[c->data_n++] = s; }
		}
		if(c->fr_state > F_A1) {
			c->signal_level = (c->signal_level * c->frame_symbol_cnt + orc_agc_signal_level(&c->agc)) / (c->frame_symbol_cnt + 1.0f);
			c->frame_symbol_cnt += 1.0f;
		}
		if(c->symbols_wanted > 1) { c->symbols_wanted--; continue; }

		switch(c->fr_state) {
		case F_A1: {
			float corr = 2.0f * (float)bits_correlate(&c->A_bs, &c->bits) / (float)ORC_A_LEN - 1.0f;
			if(fabsf(corr) > 0.36f) {
				c->st_a1++;
				c->bitmask = corr > 0.f ? 0 : ~0u;
				c->signal_level = orc_agc_signal_level(&c->agc);
				c->frame_symbol_cnt = 1.0f;
				c->symbols_wanted = ORC_A_LEN;
				c->search_retries = 0;
				c->fr_state = F_A2;
			}
			break; }
		case F_A2: {
			float corr = 2.0f * (float)bits_correlate(&c->A_bs, &c->bits) / (float)ORC_A_LEN - 1.0f;
			if(fabsf(corr) > 0.3f) {
				c->a2_sample_cnt = c->sample_cnt;
				c->freq_err_hz = c->c_dphi * ORC_SYMBOL_RATE / (2.0 * M_PI);
				c->st_a2++;
				c->symbols_wanted = ORC_M1_LEN;
				c->search_retries = 0;
				c->fr_state = F_M1;
			} else if(++c->search_retries >= 3) {
				framer_reset(c);
			}
			break; }
		case F_M1: {
			float max_corr = 0.f; int max_idx = -1;     /* match_sequence hfdl.c:937-950 */
			for(int idx = 0; idx < 8; idx++) {
				float corr = fabsf(2.0f * (float)bits_correlate(&c->M1_bs[idx], &c->bits) / (float)ORC_M1_LEN - 1.0f);
				if(corr > max_corr) { max_corr = corr; max_idx = idx; }
			}
			if(fabsf(max_corr) > 0.3f) {
				c->st_m1++;
				c->data_segment_cnt = orc_modes[max_idx].segments;
				c->data_arity = orc_modes[max_idx].arity;
				c->M1 = max_idx;
				c->symbols_wanted = ORC_M2_LEN;
				c->search_retries = 0;
				c->fr_state = F_M2_SKIP;
				c->s_state = S_SKIP;
			} else {
				c->st_m1_fail++;                        /* statsd demod.preamble.errors.M1_not_found, hfdl.c:840 */
				framer_reset(c);
			}
			break; }
		case F_M2_SKIP:
			c->training_n = 0;
			c->symbols_wanted = ORC_T_LEN;
			c->eq_train_seq_cnt = 9;
			c->fr_state = F_EQ_TRAIN;
			c->s_state = S_EMIT_SYMBOLS;
			break;
		case F_EQ_TRAIN:
			train_bit_errors(c);
			if(c->eq_train_seq_cnt > 1) {
				c->eq_train_seq_cnt--;
				c->symbols_wanted = ORC_T_LEN;
				c->T_idx = 0;
			} else if(c->data_segment_cnt > 0) {
				c->symbols_wanted = ORC_DATA_FRAME_LEN / 2;
				c->fr_state = F_DATA_1;
				c->cur_arity = c->data_arity;
				c->cur_buf = 1;
			} else {
				uint8_t pdu[ORC_MAX_PDU_OCTETS + 3];
				if(c->cap_mask & (1u << ORC_CAP_DATASYM))
					for(int q = 0; q < c->data_n; q++) cap_push(c, ORC_CAP_DATASYM, c->data_symbols[q]);
				int len = orc_decode_user_data(c->data_symbols, c->M1, c->bitmask, pdu, NULL);
				dispatch_pdu(c, pdu, len);
				framer_reset(c);
				c->symbol_cnt = 0;
			}
			break;
		case F_DATA_1:
			c->symbols_wanted = ORC_DATA_FRAME_LEN / 2;
			c->fr_state = F_DATA_2;
			break;
		case F_DATA_2:
			c->data_segment_cnt--;
			c->cur_arity = 1;
			c->cur_buf = 0;
			c->fr_state = F_EQ_TRAIN;
			c->eq_train_seq_cnt = 1;
			c->symbols_wanted = ORC_T_LEN;
			c->T_idx = 0;
			break;
		}
	}
}

void orc_channel_process_baseband(orc_channel_t *c, const cf32 *x, int n) {
	uint32_t cnt = 0;
	/* msresamp_crcf_execute in chunks of the scratch buffer */
	int done = 0;
	int chunk = c->chz->ddc.post_input_size;
	while(done < n) {
		int m = n - done < chunk ? n - done : chunk;
		orc_resamp_execute(c->rs, x + done, m, c->resampled, &cnt);
		for(uint32_t k = 0; k < cnt; k++, c->sample_cnt++) {
			cap_push(c, ORC_CAP_CHAN, c->resampled[k]);
			demod_sample(c, c->resampled[k]);
		}
		done += m;
	}
}

void orc_channel_process_block(orc_channel_t *c, const cf32 *spectrum_swapped) {   /* hfdl.c:674-893 */
	int n = orc_channelizer_execute(c->chz, spectrum_swapped, c->chan_out);
	if(c->cap_mask & (1u << ORC_CAP_DDC))
		for(int i = 0; i < n; i++) cap_push(c, ORC_CAP_DDC, c->chan_out[i]);
	orc_channel_process_baseband(c, c->chan_out, n);
}
