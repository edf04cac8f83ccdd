/* tfrec_oracle.c - plain-C restatement of the baycom/tfrec IQ->telegram decode path.
 *
 * TEST INFRASTRUCTURE ONLY (see tfrec_oracle.h).  Each function names the reference lines it
 * follows; file:line citations are relative to the reference tree (baycom/tfrec @ bf0803bd).
 * Written as flat C state machines over one struct, not as a copy of the C++ class layout.
 * Uninitialised reference members (tfa1/tfa2/whb demod last_i,last_q, whb avg_of) start at 0.
 *
 * Floating-point semantics.  The reference Makefile builds x86_64 with -O3 -ffast-math
 * (Makefile:11-13,20), and that binary - compiled unmodified into oracle/_ref - is the ground truth
 * on the box.  g++ 13.3 applies four value-changing rewrites on this path (read off the
 * disassembly of the objects in oracle/_ref, confirmed bit-for-bit by tests/test_oracle_vs_ref.py):
 *   1. fm_dev:     angle / M_PI * (1<<14)   ->  angle * fl(16384/pi)            (one multiply)
 *   2. iir2::step: b0*dn+b1*dn1+b2*dn2+a1*yn1+a2*yn2 ->
 *                  ((b2*dn2 + a1*yn1) + (b0*dn + b1*dn1)) + a2*yn2
 *   3. x / 10, x / 10.0                     ->  x * 0.1      (temperatures, TX22/WHB scalings)
 *   4. rssi / 4000                          ->  rssi * 0.00025 (whb.cpp:696)
 *   5. iir2::set (dsp_stuff.cpp:36-45) is reassociated too: two of the five coefficient sets
 *      come out 1 ulp away from the source-order values; the as-built bit patterns are tabulated.
 * This file restates those forms (ORC_AS_COMPILED=1, the default).  -DORC_AS_COMPILED=0 gives the
 * literal source order instead.  This file itself is built with strict IEEE double
 * (-fno-fast-math -ffp-contract=off) so that what is written here is what runs.
 */
#ifndef ORC_AS_COMPILED
#define ORC_AS_COMPILED 1
#endif
#include "tfrec_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define DEC_PER_BLOCK 8192          /* 65536 B -> 32768 raw IQ -> 8192 decimated IQ          */
#define IDX_PER_BLOCK 16384         /* fsk_demod::process len (int16 count), fm_demod.cpp:42 */
#define N_KINDS 5

/* ------------------------------------------------------------------ small growable arrays */
typedef struct { void *p; size_t n, cap, esz; } vec;
static void *vec_push(vec *v, const void *e)
{
	if (v->n == v->cap) {
		v->cap = v->cap ? v->cap * 2 : 64;
		v->p = realloc(v->p, v->cap * v->esz);
		if (!v->p) { perror("oracle realloc"); exit(-1); }
	}
	void *dst = (char *)v->p + v->n * v->esz;
	memcpy(dst, e, v->esz);
	v->n++;
	return dst;
}

/* ------------------------------------------------------------------ CRCs */
/* crc8.cpp:4-29: table-driven, poly 0x31, init 0, MSB first.  Restated bitwise. */
uint8_t orc_crc8(const uint8_t *d, int len)
{
	uint8_t c = 0;
	for (int n = 0; n < len; n++) {
		c ^= d[n];
		for (int m = 0; m < 8; m++)
			c = (c & 0x80) ? (uint8_t)((c << 1) ^ 0x31) : (uint8_t)(c << 1);
	}
	return c;
}
/* crc32.cpp:4-30: poly 0x04c11db7, caller-supplied init, MSB first, no reflection, no xorout. */
uint32_t orc_crc32(const uint8_t *d, int len, uint32_t init)
{
	uint32_t c = init;
	for (int n = 0; n < len; n++) {
		c ^= (uint32_t)d[n] << 24;
		for (int m = 0; m < 8; m++)
			c = (c & 0x80000000u) ? ((c << 1) ^ 0x04c11db7u) : (c << 1);
	}
	return c;
}

/* ------------------------------------------------------------------ decimator */
/* tap tables: dsp_stuff.cpp:61-88 (narrow), :91-117 (wide, -W), :119-130 (first stage) */
static const int16_t TAPS_S2_NARROW[20] = { -1087, -1082, -1065, -451, 912, 2997, 5556, 8157, 10285, 11484,
	11484, 10285, 8157, 5556, 2997, 912, -451, -1065, -1082, -1087 };
static const int16_t TAPS_S2_WIDE[20] = { 546, 451, -317, -1844, -3198, -2817, 494, 6469, 13074, 17421,
	17421, 13074, 6469, 494, -2817, -3198, -1844, -317, 451, 546 };
static const int16_t TAPS_S1[8] = { 2443, 6339, 11036, 14254, 14254, 11036, 6339, 2443 };

/* (a*b)>>16 on int: arithmetic shift = floor, per tap (dsp_stuff.cpp:194-195, 222-223) */
static inline int32_t tap_floor(int32_t a, int32_t b)
{
	int32_t p = a * b;
	return (p >= 0) ? (p >> 16) : -(int32_t)(((uint32_t)(-p) + 0xffffu) >> 16);
}

typedef struct {
	int16_t w1[8];   /* last 8 stage-1 inputs  (process2x1's t0, carried in hist0) */
	int16_t w2[20];  /* last 20 stage-2 inputs (process2x's t0)                    */
	int phase;       /* stage-1 outputs since the last stage-2 output (0/1)        */
} fir_chan;

/* one raw pair (two consecutive samples of one channel) -> optionally one 384 kS/s sample.
 * Stage 1 = decimate::process2x1 (dsp_stuff.cpp:204-230), stage 2 = process2x (:172-202);
 * downconvert::process_iq (:243-264) runs them block-wise in place, which is equivalent to
 * this sample-wise cascade because both carry their window in hist0. */
static int fir_push_pair(fir_chan *c, int16_t x0, int16_t x1, const int16_t *t2, int16_t *out)
{
	memmove(c->w1, c->w1 + 2, 6 * sizeof(int16_t));
	c->w1[6] = x0;
	c->w1[7] = x1;
	int32_t s = 0;
	for (int n = 0; n < 8; n++)
		s += tap_floor(c->w1[n], TAPS_S1[n]);
	int16_t y1 = (int16_t)s;
	if (c->phase == 0) {
		memmove(c->w2, c->w2 + 2, 18 * sizeof(int16_t));
		c->w2[18] = y1;
		c->phase = 1;
		return 0;
	}
	c->w2[19] = y1;
	c->phase = 0;
	s = 0;
	for (int n = 0; n < 20; n++)
		s += tap_floor(c->w2[n], t2[n]);
	*out = (int16_t)s;
	return 1;
}

typedef struct { fir_chan ch[2]; const int16_t *t2; } decim;

static void decim_init(decim *d, int filter)
{
	memset(d, 0, sizeof(*d));
	d->t2 = filter ? TAPS_S2_WIDE : TAPS_S2_NARROW;
}
/* u8 block -> int16 decimated; conversion (b-128)<<6 is engine.cpp:77-78 */
static size_t decim_block(decim *d, const uint8_t *iq, size_t nbytes, int16_t *out)
{
	size_t no = 0;
	for (size_t j = 0; j + 3 < nbytes; j += 4) {
		int16_t i0 = (int16_t)((iq[j] - 128) * 64), q0 = (int16_t)((iq[j + 1] - 128) * 64);
		int16_t i1 = (int16_t)((iq[j + 2] - 128) * 64), q1 = (int16_t)((iq[j + 3] - 128) * 64);
		int16_t yi, yq;
		int gi = fir_push_pair(&d->ch[0], i0, i1, d->t2, &yi);
		int gq = fir_push_pair(&d->ch[1], q0, q1, d->t2, &yq);
		if (gi && gq) {
			out[no++] = yi;
			out[no++] = yq;
		}
	}
	return no;
}
size_t orc_decimate(const uint8_t *iq, size_t nbytes, int filter, int16_t *out)
{
	decim d;
	decim_init(&d, filter);
	return decim_block(&d, iq, nbytes, out);
}

/* downconvert(passes)::process_iq, dsp_stuff.cpp:232-264, over a whole buffer with zero history (the reference
 * carries each stage's window in hist0, so whole-buffer == block-wise): passes-1 times process2x1 (8 taps,
 * :204-230) on I and on Q, then process2x (20 taps, narrow or wide, :172-202); every stage halves the rate
 * and stores int16.  Stage output k = sum_n floor(x[2k-(T-2)+n]*t[n] / 2^16), x[<0] = 0.
 * nbytes of u8 IQ -> (nbytes/2 >> passes) I,Q pairs; returns the number of int16 written. */
size_t orc_downconvert(const uint8_t *iq, size_t nbytes, int passes, int filter, int16_t *out)
{
	if (passes < 1 || passes > 8) return 0;
	size_t n = nbytes / 2;   /* IQ pairs at the current stage */
	int16_t *cur = (int16_t *)malloc((n ? n : 1) * 2 * sizeof(int16_t));
	if (!cur) return 0;
	for (size_t j = 0; j < 2 * n; j++) cur[j] = (int16_t)((iq[j] - 128) * 64);   /* engine.cpp:77-78 */
	for (int p = 0; p < passes; p++) {
		const int last = (p == passes - 1);
		const int T = last ? 20 : 8;
		const int16_t *t = last ? (filter ? TAPS_S2_WIDE : TAPS_S2_NARROW) : TAPS_S1;
		const size_t no = n / 2;
		int16_t *nx = (int16_t *)malloc((no ? no : 1) * 2 * sizeof(int16_t));
		if (!nx) { free(cur); return 0; }
		for (size_t k = 0; k < no; k++)
			for (int c = 0; c < 2; c++) {
				int32_t sum = 0;
				for (int m = 0; m < T; m++) {
					const long long j = 2 * (long long)k - (T - 2) + m;
					if (j >= 0) sum += tap_floor(cur[2 * j + c], t[m]);
				}
				nx[2 * k + c] = (int16_t)sum;
			}
		free(cur);
		cur = nx;
		n = no;
	}
	memcpy(out, cur, n * 2 * sizeof(int16_t));
	free(cur);
	return n * 2;
}

/* ------------------------------------------------------------------ discriminators, biquad */
/* dsp_stuff.cpp:269-279 */
int orc_fm_dev_nrzs(int ar, int aj, int br, int bj)
{
	int cr = ar * br + aj * bj;
	if (cr > 1000000000) cr = 1000000000;
	if (cr < -1000000000) cr = -1000000000;
	return cr;
}
/* dsp_stuff.cpp:284-292: double products in this operand order (signed zeros matter), atan2,
 * scale by 2^14/pi, truncate */
int orc_fm_dev(int ar, int aj, int br, int bj)
{
	double cr = ((double)ar) * br + ((double)aj) * bj;
	double cj = ((double)aj) * br - ((double)ar) * bj;
	double angle = atan2(cj, cr);
#if ORC_AS_COMPILED
	return (int)(angle * 5215.189175235227); /* 0x40b45f306dc9c883 = fl(16384/pi) */
#else
	return (int)(angle / M_PI * (1 << 14));
#endif
}

typedef struct { double b0, b1, b2, a1, a2, d1, d2, y0, y1, y2; } biquad;
/* coefficients {b0,b1,b2,a1,a2} exactly as the -ffast-math reference binary computes them for the
 * five cutoffs main.cpp registers: 0.5/spb for TFA_2, TFA_3, TX22 (tfa2.cpp:321), 2.0/64 and
 * 0.0025/64 for WHB (whb.cpp:610-611) */
static const struct { double cutoff; uint64_t c[5]; } BIQUAD_AS_BUILT[5] = {
	{ 0.5 / ((1536000 / 4.0) / 17240), { 0x3f727f98b1037a14ull, 0x3f827f98b1037a14ull, 0x3f727f98b1037a14ull, 0x3ffcd1527f4a26e2ull, 0xbfea36a1c41c6995ull } },
	{ 0.5 / ((1536000 / 4.0) / 9600), { 0x3f57ed02b18a270dull, 0x3f67ed02b18a270dull, 0x3f57ed02b18a270dull, 0x3ffe397ac010fc89ull, 0xbfeca2cf85850d62ull } },
	{ 0.5 / ((1536000 / 4.0) / 8842), { 0x3f5461fa1a309718ull, 0x3f6461fa1a309718ull, 0x3f5461fa1a309718ull, 0x3ffe5d4f47377e30ull, 0xbfece36282a35d90ull } },
	{ 2.0 / 64.0, { 0x3f814a67102a1ffdull, 0x3f914a67102a1ffdull, 0x3f814a67102a1ffdull, 0x3ffb949652fa3970ull, 0xbfe83dd316f714e0ull } },
	{ 0.0025 / 64.0, { 0x3e502ae4cfc8910aull, 0x3e602ae4cfc8910aull, 0x3e502ae4cfc8910aull, 0x3ffffe9409fe171bull, 0xbfeffd283451f7d3ull } },
};
/* iir2::set, dsp_stuff.cpp:36-45 */
static void biquad_init(biquad *f, double cutoff)
{
	memset(f, 0, sizeof(*f));
#if ORC_AS_COMPILED
	for (int k = 0; k < 5; k++)
		if (BIQUAD_AS_BUILT[k].cutoff == cutoff) {
			double v[5];
			memcpy(v, BIQUAD_AS_BUILT[k].c, sizeof(v));
			f->b0 = v[0]; f->b1 = v[1]; f->b2 = v[2]; f->a1 = v[3]; f->a2 = v[4];
			return;
		}
#endif
	double i = 1.0 / tan(M_PI * cutoff);
	double s = sqrt(2);
	f->b0 = 1 / (1 + s * i + i * i);
	f->b1 = 2 * f->b0;
	f->b2 = f->b0;
	f->a1 = 2 * (i * i - 1) * f->b0;
	f->a2 = -(1 - s * i + i * i) * f->b0;
}
/* test hook: the five doubles of biquad k as used by this oracle */
void orc_biquad_coeffs(int k, double *out5)
{
	biquad f;
	biquad_init(&f, BIQUAD_AS_BUILT[k].cutoff);
	out5[0] = f.b0; out5[1] = f.b1; out5[2] = f.b2; out5[3] = f.a1; out5[4] = f.a2;
}
/* iir2::step, dsp_stuff.cpp:47-55 (left-to-right sum) */
static double biquad_step(biquad *f, double dn)
{
	f->y2 = f->y1;
	f->y1 = f->y0;
#if ORC_AS_COMPILED
	f->y0 = ((f->b2 * f->d2 + f->a1 * f->y1) + (f->b0 * dn + f->b1 * f->d1)) + f->a2 * f->y2;
#else
	f->y0 = f->b0 * dn + f->b1 * f->d1 + f->b2 * f->d2 + f->a1 * f->y1 + f->a2 * f->y2;
#endif
	f->d2 = f->d1;
	f->d1 = dn;
	return f->y0;
}

#if ORC_AS_COMPILED
#define DIV10(x) ((x) * 0.1)
#define DIV4000(x) ((x) * 0.00025)
#else
#define DIV10(x) ((x) / 10.0)
#define DIV4000(x) ((x) / 4000)
#endif

/* double -> int the way x86 cvttsd2si does it (INT_MIN on overflow/NaN); the reference hits this
 * for 10*log10(0) = -inf (tfa1.cpp:180, tfa2.cpp:434) */
static int d2i(double v)
{
	if (!(v > -2147483649.0 && v < 2147483648.0))
		return (int)0x80000000;
	return (int)v;
}

/* ------------------------------------------------------------------ the whole receiver */
typedef struct { uint64_t id; int type; int sequence; } seen_t;

typedef struct {
	int kind;            /* index 0..4 in registration order */
	int type;            /* sensor_e */
	/* decoder base (decoder.h:33-59) */
	int synced, byte_cnt, bad;
	uint8_t rdata[256];
	uint32_t sr;
	int sr_cnt;
	int invert;                                 /* tfa2_decoder */
	int w_last_bit, w_psk, w_last_psk, w_nrzs;  /* whb_decoder  */
	uint32_t w_lfsr;
	vec seen;                                   /* decoder::data map, decoder.cpp:46-65 */
	/* demodulator base + per-type demod state */
	int last_bit_idx;
	int timeout_cnt, last_i, last_q;
	int mark_lvl, rssi_i;                       /* tfa1_demod */
	double spb, est_spb;                        /* tfa2_demod / whb_demod */
	int bitcnt, dmin, dmax, offset, last_bit;
	biquad lp, lp_avg;
	int last_dev, avg_of;                       /* whb_demod */
	uint64_t step, last_peak;
	double rssi_d;
} chan;

struct orc {
	int n_ch;
	chan ch[N_KINDS];
	decim dec;
	int thresh, thresh_mode, triggered_avg, runs;
	int64_t pos_base;       /* decimated index of the current block's first sample */
	int64_t cur_pos;
	int tap_mask;
	long inverted;
	vec frames, records, blocks, tap[3], tap_ch[3];
	int16_t dbuf[2 * DEC_PER_BLOCK];
};

static void tap_put(orc_t *o, const chan *c, int kind, const void *v)
{
	uint8_t k = (uint8_t)c->kind;
	vec_push(&o->tap[kind], v);
	vec_push(&o->tap_ch[kind], &k);
}

static orc_frame *emit_frame(orc_t *o, chan *c, int status, int rssi, int offset, int byte_cnt)
{
	orc_frame f;
	memset(&f, 0, sizeof(f));
	f.type = c->type;
	f.status = status;
	f.pos = o ? o->cur_pos : -1;
	f.byte_cnt = byte_cnt;
	f.rssi = rssi;
	f.offset = offset;
	memcpy(f.rdata, c->rdata, ORC_MAX_RDATA);
	return (orc_frame *)vec_push(&o->frames, &f);
}

/* decoder::store_data (decoder.cpp:46-65): record everything, flag what mode 0 would exec */
static void emit_record(orc_t *o, chan *c, orc_frame *f, uint64_t id, double temp, double hum,
			int sequence, int alarm, int rssi)
{
	orc_record r;
	memset(&r, 0, sizeof(r));
	r.type = c->type;
	r.id = id;
	r.temp = temp;
	r.humidity = hum;
	r.sequence = sequence;
	r.alarm = alarm;
	r.rssi = rssi;
	r.flags = 0;
	r.pos = f->pos;
	r.frame = (int32_t)(f - (orc_frame *)o->frames.p);
	int found = 0;
	seen_t *s = (seen_t *)c->seen.p;
	size_t k;
	for (k = 0; k < c->seen.n; k++)
		if (s[k].id == id)
			break;
	if (k == c->seen.n) {
		seen_t e = { id, c->type, sequence };
		vec_push(&c->seen, &e);
	} else if (s[k].type == ORC_TFA_WHB) {
		if (s[k].sequence == sequence)
			found = 1;
		else
			s[k].sequence = sequence;
	}
	r.flags |= found ? 0x100 : 0; /* bit 8: suppressed duplicate (never set by the reference's own flags=0) */
	vec_push(&o->records, &r);
	f->n_records++;
}

/* ---- TFA_1 --------------------------------------------------------------------------------- */
/* tfa1_decoder::flush, tfa1.cpp:47-118 */
static void tfa1_flush(orc_t *o, chan *c, int rssi)
{
	if (c->byte_cnt >= 10) {
		const uint8_t *r = c->rdata;
		int id = ((r[2] << 8) | r[3]) & 0x7fff;
		int batfail = (r[7] & 0x80) >> 7;
		double temp = ((r[4] & 0xf) * 100) + ((r[5] >> 4) * 10) + (r[5] & 0xf);
		temp = DIV10(temp) - 40;
		int hum = r[6];
		int seq = r[8] >> 4;
		uint8_t crc_val = r[10], crc_calc = orc_crc8(&r[2], 8);
		int sane = ((r[4] & 0xf0) == 0x80 || hum == 0x7f || hum == 0x6a) && hum <= 0x7f &&
			   (r[7] & 0x60) == 0x60 && (r[8] & 0xf) == 0 && r[9] == 0x56;
		if (crc_val == crc_calc && sane) {
			if (hum == 0x6a)
				hum = 0;
			if (r[5] == 0xff || r[5] == 0xaa || hum == 0x7f) {
				batfail = 2;
				hum = 0;
				temp = 0;
			}
			orc_frame *f = emit_frame(o, c, ORC_FRAME_OK, rssi, 0, c->byte_cnt);
			snprintf(f->line, sizeof(f->line), "TFA1 ID %04x %+.1f %i%% seq %x lowbat %i RSSI %i",
				 id, temp, hum, seq, batfail, rssi);
			emit_record(o, c, f, (uint64_t)id, temp, hum, seq, batfail, rssi);
		} else {
			c->bad++;
			emit_frame(o, c, crc_val != crc_calc ? ORC_FRAME_BAD_CRC : ORC_FRAME_BAD_SANITY, rssi, 0,
				   c->byte_cnt);
		}
	}
	c->sr_cnt = -1;
	c->byte_cnt = 0;
	c->rdata[10] = 0;
}
/* tfa1_decoder::store_bit, tfa1.cpp:120-134 */
static void tfa1_bit(chan *c, int bit)
{
	c->sr = (c->sr >> 1) | ((uint32_t)bit << 31);
	if ((c->sr & 0xffff) == 0xd42d) {
		c->sr_cnt = 0;
		c->byte_cnt = 0;
	}
	if (c->sr_cnt == 0) {
		if (c->byte_cnt < 256)
			c->rdata[c->byte_cnt] = c->sr & 0xff;
		c->byte_cnt++;
	}
	if (c->sr_cnt >= 0)
		c->sr_cnt = (c->sr_cnt + 1) & 7;
}
/* tfa1_demod::demod, tfa1.cpp:143-190; BITPERIOD = 10 (tfa1.cpp:34) */
static int tfa1_sample(orc_t *o, chan *c, int thresh, int pwr, int index, const int16_t *iq)
{
	int triggered = 0;
	if (pwr > thresh)
		c->timeout_cnt = 400;
	if (c->timeout_cnt) {
		triggered = 1;
		int dev = orc_fm_dev_nrzs(iq[0], iq[1], c->last_i, c->last_q);
		if (o->tap_mask & 2)
			tap_put(o, c, 1, &dev);
		if (dev > c->mark_lvl)
			c->mark_lvl = dev;
		else
			c->mark_lvl = (int)(c->mark_lvl * 0.95);
		if (c->mark_lvl > c->rssi_i)
			c->rssi_i = c->mark_lvl;
		c->timeout_cnt--;
		if (dev < c->mark_lvl / 2) {
			if (c->last_bit_idx) {
				if (index - c->last_bit_idx > 4) {
					for (int n = 22; n <= index - c->last_bit_idx; n += 20)
						tfa1_bit(c, 1);
					tfa1_bit(c, 0);
				}
			}
			if (index - c->last_bit_idx > 2)
				c->last_bit_idx = index;
		}
		if (!c->timeout_cnt) {
			tfa1_flush(o, c, d2i(10 * log10((double)c->rssi_i)));
			c->mark_lvl = 0;
			c->rssi_i = 0;
			c->last_bit_idx = 0;
		}
	}
	c->last_i = iq[0];
	c->last_q = iq[1];
	return triggered;
}

/* ---- TFA_2 / TFA_3 / TX22 ------------------------------------------------------------------ */
static void tfa2_decoder_reset(chan *c)
{
	c->sr_cnt = -1;
	c->sr = 0;
	c->byte_cnt = 0;
}
/* tfa2_decoder::flush_tfa, tfa2.cpp:219-279 */
static void tfa2_flush_tfa(orc_t *o, chan *c, int rssi, int offset)
{
	if (c->byte_cnt >= 7) {
		const uint8_t *r = c->rdata;
		int id = (c->type << 28) | (r[2] << 8) | (r[3] & 0xc0);
		double temp = ((r[3] & 0xf) * 100 + (r[4] >> 4) * 10 + (r[4] & 0xf));
		temp = DIV10(temp) - 40;
		int hum = r[5];
		uint8_t crc_val = r[6], crc_calc = orc_crc8(&r[2], 4);
		if (hum == 0x7d)
			id |= 1;
		if (crc_val == crc_calc) {
			if (hum > 100)
				hum = 0;
			orc_frame *f = emit_frame(o, c, ORC_FRAME_OK, rssi, offset, c->byte_cnt);
			snprintf(f->line, sizeof(f->line), "TFA%i ID %06x %+.1lf %i%% RSSI %i Offset %.0lfkHz",
				 c->type + 1, id, temp, hum, rssi, -1536.0 * offset / 131072);
			emit_record(o, c, f, (uint64_t)(int64_t)id, temp, hum, 0, 0, rssi);
		} else {
			c->bad++;
			emit_frame(o, c, ORC_FRAME_BAD_CRC, rssi, offset, c->byte_cnt);
		}
	}
	tfa2_decoder_reset(c);
}
/* tfa2_decoder::flush_tx22, tfa2.cpp:72-217 */
static void tfa2_flush_tx22(orc_t *o, chan *c, int rssi, int offset)
{
	if (c->byte_cnt >= 7 && c->byte_cnt < 64) {
		const uint8_t *r = c->rdata;
		uint8_t crc_val = 0, crc_calc = 0;
		int ok = 0;
		int id = 0, error = 0, lowbat = 0, num = 0;
		if ((r[2] >> 4) == 0xa) {
			id = ((r[2] & 0xf) << 2) | (r[3] >> 6);
			error = !((r[3] >> 4) & 1);
			lowbat = (r[3] >> 3) & 1;
			num = r[3] & 7;
			crc_val = r[2 * num + 4];
			crc_calc = orc_crc8(&r[2], 2 + 2 * num);
			ok = (crc_val == crc_calc);
		}
		if (ok) {
			int have_temp = 0, have_hum = 0, have_rain = 0, have_wind = 0, have_gust = 0;
			double temp = 0, hum = 0, rain = 0, wdir = 0, wspeed = 0, wgust = 0;
			for (int n = 0; n < num; n++) {
				const uint8_t *w = &r[4 + n * 2];
				switch (w[0] >> 4) {
				case 0: {
					double v = (w[0] & 0xf) * 100 + (w[1] >> 4) * 10 + (w[1] & 0xf);
					temp = DIV10(v) - 40;
					have_temp = 1;
					break;
				}
				case 1:
					hum = (w[0] & 0xf) * 100 + (w[1] >> 4) * 10 + (w[1] & 0xf);
					have_hum = 1;
					break;
				case 2:
					rain = ((w[0] & 0xf) << 8) + w[1];
					have_rain = 1;
					break;
				case 3:
					wdir = (w[0] & 0xf) * 22.5;
					wspeed = DIV10((double)w[1]);
					have_wind = 1;
					break;
				case 4:
					wgust = DIV10((double)(((w[0] & 0xf) << 8) + w[1]));
					have_gust = 1;
					break;
				default:
					break;
				}
			}
			int alarm = error | lowbat;
			int new_id = (c->type << 28) | (id << 4);
			orc_frame *f = emit_frame(o, c, ORC_FRAME_OK, rssi, offset, c->byte_cnt);
			char *p = f->line;
			size_t cap = sizeof(f->line);
			int k = snprintf(p, cap, "TX22 ID %x, ", new_id);
			if (have_temp) k += snprintf(p + k, cap - k, "temp %g, ", temp);
			if (have_hum) k += snprintf(p + k, cap - k, "hum %g, ", hum);
			if (have_rain) k += snprintf(p + k, cap - k, "rain %g, ", rain);
			if (have_wind) k += snprintf(p + k, cap - k, "speed %g, dir %g, ", wspeed, wdir);
			if (have_gust) k += snprintf(p + k, cap - k, "gust %g, ", wgust);
			snprintf(p + k, cap - k, "RSSI %i, offset %.0lfkHz", rssi, -1536.0 * offset / 131072);
			if (have_temp) emit_record(o, c, f, (uint64_t)(int64_t)new_id, temp, hum, 0, alarm, rssi);
			if (have_rain) emit_record(o, c, f, (uint64_t)(int64_t)(new_id | 2), rain, 0, 0, alarm, rssi);
			if (have_wind) emit_record(o, c, f, (uint64_t)(int64_t)(new_id | 3), wspeed, wdir, 0, alarm, rssi);
			if (have_gust) emit_record(o, c, f, (uint64_t)(int64_t)(new_id | 4), wgust, 0, 0, alarm, rssi);
		} else {
			/* the reference only counts bad++ when dbg is set (tfa2.cpp:205-206); not modelled */
			emit_frame(o, c, crc_val != crc_calc ? ORC_FRAME_BAD_CRC : ORC_FRAME_BAD_SANITY, rssi, offset,
				   c->byte_cnt);
		}
	}
	tfa2_decoder_reset(c);
}
static void tfa2_flush(orc_t *o, chan *c, int rssi, int offset)
{
	if (c->type == ORC_TX22)
		tfa2_flush_tx22(o, c, rssi, offset);
	else
		tfa2_flush_tfa(o, c, rssi, offset);
}
/* tfa2_decoder::store_bit, tfa2.cpp:281-314 */
static void tfa2_bit(orc_t *o, chan *c, int bit)
{
	c->sr = (c->sr << 1) | (uint32_t)bit;
	if ((c->sr & 0xffff) == 0x2dd4) {
		c->sr_cnt = 0;
		c->rdata[0] = (c->sr >> 8) & 0xff;
		c->byte_cnt = 1;
		c->invert = 0;
	}
	if (((~c->sr) & 0xffff) == 0x2dd4) {
		if (o)
			o->inverted++;
		c->sr_cnt = 0;
		c->rdata[0] = (uint8_t)~((c->sr >> 8) & 0xff);
		c->byte_cnt = 1;
		c->invert = 1;
	}
	if (c->sr_cnt == 0) {
		if (c->byte_cnt < 256)
			c->rdata[c->byte_cnt] = c->invert ? (uint8_t)~(c->sr & 0xff) : (uint8_t)(c->sr & 0xff);
		c->byte_cnt++;
	}
	if (c->sr_cnt >= 0)
		c->sr_cnt = (c->sr_cnt + 1) & 7;
}
/* tfa2_demod::reset, tfa2.cpp:325-334 (leaves the biquad and last_bit_idx alone) */
static void tfa2_reset(chan *c)
{
	c->offset = 0;
	c->bitcnt = 0;
	c->dmin = 32767;
	c->dmax = -32767;
	c->last_bit = 0;
	c->rssi_i = 0;
	c->est_spb = c->spb;
}
/* tfa2_demod::demod, tfa2.cpp:346-442 */
static int tfa2_sample(orc_t *o, chan *c, int thresh, int pwr, int index, const int16_t *iq)
{
	int triggered = 0;
	if (pwr > thresh) {
		if (!c->timeout_cnt)
			tfa2_reset(c);
		c->timeout_cnt = (int)(16 * c->spb);
	}
	if (c->timeout_cnt) {
		triggered = 1;
		int dev = orc_fm_dev(iq[0], iq[1], c->last_i, c->last_q);
		if (o->tap_mask & 1)
			tap_put(o, c, 0, &dev);
		double y = biquad_step(&c->lp, dev);
		if (o->tap_mask & 4)
			tap_put(o, c, 2, &y);
		int ld = (int)y;
		if (c->bitcnt < 10) {
			if (ld > c->dmax)
				c->dmax = (7 * c->dmax + ld) / 8;
			if (ld < c->dmin)
				c->dmin = (7 * c->dmin + ld) / 8;
			c->offset = (c->dmax + c->dmin) / 2;
			if (c->bitcnt > 4)
				c->rssi_i = (int)((uint32_t)c->rssi_i +
						  (uint32_t)((int)((uint32_t)c->rssi_i + (uint32_t)(iq[0] * iq[0]) +
								   (uint32_t)(iq[1] * iq[1])) / 100));
		}
		c->timeout_cnt--;
		dev = ld;
		int noffset = (int)(0.9 * c->offset);
		int bit = 0;
		const int margin = 32;
		if (dev > noffset + (c->dmax / margin))
			bit = 1;
		if ((dev > noffset + c->dmax / margin || dev < noffset + c->dmin / margin) && bit != c->last_bit) {
			if (index > (c->last_bit_idx + 8)) {
				c->bitcnt++;
				int tdiff = index - c->last_bit_idx;
				if (tdiff > c->spb / 4 && tdiff < 32 * c->spb) {
					int bit_diff = (index - c->last_bit_idx) / 2;
					int numbits = (int)((bit_diff + (c->est_spb / 2)) / c->est_spb);
					if (numbits < 32)
						for (int n = 1; n < numbits; n++)
							tfa2_bit(o, c, c->last_bit);
					tfa2_bit(o, c, bit);
					c->last_bit = bit;
				}
			}
			if (index - c->last_bit_idx > 2)
				c->last_bit_idx = index;
		}
		if (!c->timeout_cnt) {
			for (int n = 0; n < 16; n++)
				tfa2_bit(o, c, c->last_bit);
			tfa2_flush(o, c, d2i(10 * log10((double)c->rssi_i)), c->offset);
			tfa2_reset(c);
		}
	}
	c->last_i = iq[0];
	c->last_q = iq[1];
	return triggered;
}

/* ---- WeatherHub ---------------------------------------------------------------------------- */
static uint32_t whb_crc_init(uint32_t stype, int *known)
{
	/* crc_initvals, whb.cpp:50-62 */
	static const uint32_t tab[][2] = { { 0x02, 0x97d97a26 }, { 0x03, 0xf59c5a1e }, { 0x04, 0x98e1d11f },
		{ 0x06, 0xa7a41254 }, { 0x07, 0x3303fb1d }, { 0x08, 0x29f0f49b }, { 0x09, 0xa7a41254 },
		{ 0x0b, 0xe7720ae4 }, { 0x10, 0x62d0afc1 }, { 0x11, 0x8cba0708 }, { 0x12, 0x5a9e30ae } };
	for (size_t k = 0; k < sizeof(tab) / sizeof(tab[0]); k++)
		if (tab[k][0] == stype) {
			*known = 1;
			return tab[k][1];
		}
	*known = 0;
	return 0;
}
static const uint32_t WHB_TIMEUNIT[4] = { 24 * 60 * 60, 60 * 60, 60, 1 }; /* whb.cpp:65-70 */
#define BE16(x) (((x)[0] << 8) | (x)[1])
#define BE24(x) (((x)[0] << 16) | ((x)[1] << 8) | (x)[2])
#define BE32(x) (((uint32_t)(x)[0] << 24) | ((x)[1] << 16) | ((x)[2] << 8) | (x)[3])

/* whb_decoder::cvt_temp, whb.cpp:109-123 */
static double whb_temp(uint16_t raw, int extended)
{
	if (extended == 1)
		return (raw & 0x800) ? DIV10((double)(-((raw ^ 0xfff) + 1))) : DIV10((double)raw);
	return (raw & 0x400) ? DIV10((double)(-((raw ^ 0x7ff) + 1))) : DIV10((double)raw);
}
/* the eleven payload parsers, whb.cpp:126-475; prints restated for dbg == 0 */
static void whb_payload(orc_t *o, chan *c, orc_frame *f, uint32_t stype, const uint8_t *m, uint64_t id, int rssi)
{
	char *L = f->line;
	size_t cap = sizeof(f->line);
	unsigned long long pid = (unsigned long long)id;
	uint64_t base = id << 4;
	int seq = BE16(m) & 0x3fff;
	switch (stype) {
	case 0x02: {
		uint16_t t = BE16(m + 2) & 0x7ff, tp = BE16(m + 4) & 0x7ff;
		snprintf(L, cap, "WHB02 ID %llx TEMP %g, PTEMP %g", pid, whb_temp(t, 0), whb_temp(tp, 0));
		emit_record(o, c, f, base, whb_temp(t, 0), 0, seq, 0, rssi);
		break;
	}
	case 0x03: {
		uint16_t t = BE16(m + 2) & 0x7ff, h = BE16(m + 4) & 0xff;
		uint16_t tp = BE16(m + 6) & 0x7ff, hp = BE16(m + 8) & 0xff;
		snprintf(L, cap, "WHB03 ID %llx TEMP %g HUM %i, PTEMP %g PHUM %i", pid, whb_temp(t, 0), h,
			 whb_temp(tp, 0), hp);
		emit_record(o, c, f, base, whb_temp(t, 0), h, seq, 0, rssi);
		break;
	}
	case 0x04: {
		uint16_t t = BE16(m + 2) & 0x7ff, h = BE16(m + 4) & 0xff;
		uint8_t wet = m[6];
		uint16_t tp = BE16(m + 7) & 0x7ff, hp = BE16(m + 9) & 0xff, wp = m[11];
		snprintf(L, cap, "WHB04 ID %llx TEMP %g HUM %i WET %i, PTEMP %g PHUM %i PWET %i", pid, whb_temp(t, 0), h,
			 (wet & 1) ^ 1, whb_temp(tp, 0), hp, (wp & 1) ^ 1);
		emit_record(o, c, f, base, whb_temp(t, 0), h, seq, 0, rssi);
		emit_record(o, c, f, base | 5, (wet & 1) ^ 1, 0, seq, 0, rssi);
		break;
	}
	case 0x06:
	case 0x09: {
		int ext = (stype == 0x09);
		uint16_t t = BE16(m + 2) & 0x7ff;
		uint16_t t2 = BE16(m + 4) & (ext ? 0xfff : 0x7ff), t2p = BE16(m + 10) & (ext ? 0xfff : 0x7ff);
		uint16_t h = BE16(m + 6) & 0xff, tp = BE16(m + 8) & 0x7ff, hp = BE16(m + 12) & 0xff;
		snprintf(L, cap, "WHB0%i ID %llxTEMP %g HUM %i TEMP2 %g, PTEMP %g PHUM %i PTEMP2 %g", ext ? 9 : 6, pid,
			 whb_temp(t, 0), h, whb_temp(t2, ext), whb_temp(tp, 0), hp, whb_temp(t2p, ext));
		emit_record(o, c, f, base, whb_temp(t, 0), h, seq, 0, rssi);
		emit_record(o, c, f, base | 1, whb_temp(t2, ext), 0, seq, 0, rssi);
		break;
	}
	case 0x07: {
		uint16_t t[4], h[4];
		for (int n = 0; n < 4; n++) {
			t[n] = BE16(m + 2 + 4 * n) & 0x07ff;
			h[n] = BE16(m + 4 + 4 * n) & 0x0ff;
		}
		snprintf(L, cap, "WHB07 ID %llx TEMP_IN %g HUM_IN %i TEMP_OUT %g HUM_OUT %i", pid, whb_temp(t[0], 0),
			 h[0], whb_temp(t[1], 0), h[1]);
		emit_record(o, c, f, base, whb_temp(t[0], 0), h[0], seq, 0, rssi);
		emit_record(o, c, f, base | 0xc, whb_temp(t[1], 0), h[1], seq, 0, rssi);
		break;
	}
	case 0x08: {
		uint16_t t = BE16(m + 2) & 0x07ff, cnt = BE16(m + 4);
		uint32_t times[10];
		for (int i = 0; i < 10; i++) {
			uint16_t x = BE16(m + 6 + 2 * i);
			times[i] = WHB_TIMEUNIT[(x >> 14) & 3] * (x & 0x3fff);
		}
		snprintf(L, cap, "WHB08 ID %llx cnt %i", pid, cnt);
		emit_record(o, c, f, base | 2, cnt, times[1], seq, 0, rssi);
		emit_record(o, c, f, base, whb_temp(t, 0), 0, seq, 0, rssi);
		break;
	}
	case 0x0b: {
		uint32_t v = BE32(m + 3);
		float dir = 22.5 * (v >> 28);
		float speed = DIV10((double)(((v >> 16) & 0xff) + 256 * ((v >> 25) & 1)));
		float gust = DIV10((double)(((v >> 8) & 0xff) + 256 * ((v >> 24) & 1)));
		uint32_t tm = (v & 0xff) * 2;
		snprintf(L, cap, "WHB0b ID %llx #%i DIR %f SPEED %f GUST %f time %i", pid, 0, dir, speed, gust, tm);
		emit_record(o, c, f, base | 3, speed, dir, BE24(m), 0, rssi);
		emit_record(o, c, f, base | 4, gust, 0, BE24(m), 0, rssi);
		break;
	}
	case 0x10: {
		uint8_t state[4];
		uint32_t times[4];
		for (int i = 0; i < 4; i++) {
			uint16_t x = BE16(m + 2 + 2 * i);
			state[i] = x >> 15;
			times[i] = WHB_TIMEUNIT[(x >> 13) & 3] * (x & 0x1fff);
		}
		snprintf(L, cap, "WHB10 ID %llx #%i %i %i", pid, 0, state[0], times[0]);
		emit_record(o, c, f, base | 5, state[0], times[1], seq, 0, rssi);
		break;
	}
	case 0x11: {
		uint16_t t[8], h[8];
		for (int n = 0; n < 8; n++) {
			t[n] = BE16(m + 2 + 4 * n) & 0x07ff;
			h[n] = BE16(m + 4 + 4 * n) & 0xff;
		}
		snprintf(L, cap, "WHB11 %llx TEMP1 %g HUM1 %i TEMP2 %g HUM2 %i TEMP3 %g HUM3 %i TEMP_IN %g HUM_IN %i", pid,
			 whb_temp(t[0], 0), h[0], whb_temp(t[1], 0), h[1], whb_temp(t[2], 0), h[2], whb_temp(t[3], 0), h[3]);
		emit_record(o, c, f, base, whb_temp(t[3], 0), h[3], seq, 0, rssi);
		for (int n = 0; n < 3; n++)
			emit_record(o, c, f, base | (uint64_t)(0xc + n), whb_temp(t[n], 0), h[n], seq, 0, rssi);
		break;
	}
	case 0x12: {
		uint16_t h[5] = { (uint16_t)(m[8] & 0x7f), (uint16_t)(m[2] & 0x7f), (uint16_t)(m[3] & 0x7f),
				  (uint16_t)(m[4] & 0x7f), (uint16_t)(m[5] & 0x7f) };
		uint16_t t = BE16(m + 6) & 0x7ff;
		snprintf(L, cap, "WHB12 %llx TEMP %g HUM %i HUM3h %i HUM24h %i HUM7d %i HUM30d %i", pid, whb_temp(t, 0),
			 h[0], h[1], h[2], h[3], h[4]);
		emit_record(o, c, f, base, whb_temp(t, 0), h[0], seq, 0, rssi);
		emit_record(o, c, f, base + 1, 0, h[1], seq, 0, rssi);
		for (int n = 0; n < 3; n++)
			emit_record(o, c, f, base + 0xc + n, 0, h[2 + n], seq, 0, rssi);
		break;
	}
	default:
		break;
	}
}
/* whb_decoder::flush, whb.cpp:477-564 */
static void whb_flush(orc_t *o, chan *c, int rssi, int offset)
{
	if (!(c->byte_cnt < 11 || c->byte_cnt > 60)) {
		const uint8_t *r = c->rdata;
		int plen = r[4];
		uint32_t stype = r[5];
		int known = 0;
		uint32_t init = whb_crc_init(stype, &known);
		uint32_t crc_calc = 0, crc_val = 0;
		int good = 0;
		if (plen <= 60 && known) {
			crc_calc = orc_crc32(&r[4], plen - 4 > 0 ? plen - 4 : 0, init);
			crc_val = BE32(r + plen);
			good = (crc_calc == crc_val);
		}
		if (good) {
			uint64_t id = 0;
			for (int k = 0; k < 6; k++)
				id = (id << 8) | r[5 + k];
			orc_frame *f = emit_frame(o, c, ORC_FRAME_OK, rssi, offset, c->byte_cnt);
			whb_payload(o, c, f, stype, &r[11], id, rssi);
		} else {
			c->bad++;
			orc_frame *f = emit_frame(o, c, crc_val != crc_calc ? ORC_FRAME_BAD_CRC : ORC_FRAME_BAD_SANITY,
						  rssi, offset, c->byte_cnt);
			if (plen <= 60 && !known)
				snprintf(f->line, sizeof(f->line),
					 "WHB: Probably unsupported sensor type %02x! Please report", stype);
		}
	}
	c->sr_cnt = -1;
	c->sr = 0;
	c->byte_cnt = 0;
	c->synced = 0;
}
/* whb_decoder::store_bit, whb.cpp:566-603 */
static void whb_bit(chan *c, int bit)
{
	if (bit == c->w_last_bit)
		c->w_psk = 1 - c->w_psk;
	if (c->w_psk == c->w_last_psk)
		c->w_nrzs = 1 - c->w_nrzs;
	c->w_last_bit = bit;
	c->w_last_psk = c->w_psk;
	int d = c->w_nrzs ^ ((c->w_lfsr >> 16) & 1) ^ ((c->w_lfsr >> 11) & 1);
	c->w_lfsr = (c->w_lfsr << 1) | (uint32_t)c->w_nrzs;
	c->sr = (c->sr >> 1) | ((uint32_t)d << 31);
	if (c->sr == 0x2bd42d4bu) {
		c->synced = 1;
		c->sr_cnt = 0;
		c->rdata[0] = c->sr & 0xff;
		c->rdata[1] = (c->sr >> 8) & 0xff;
		c->rdata[2] = (c->sr >> 16) & 0xff;
		c->byte_cnt = 3;
	}
	if (c->sr_cnt == 0) {
		if (c->byte_cnt < 256)
			c->rdata[c->byte_cnt] = (c->sr >> 24) & 0xff;
		c->byte_cnt++;
	}
	if (c->sr_cnt >= 0)
		c->sr_cnt = (c->sr_cnt + 1) & 7;
}
/* whb_demod::reset, whb.cpp:616-623 */
static void whb_reset(chan *c)
{
	c->offset = 0;
	c->bitcnt = 0;
	c->rssi_d = 0;
	c->step = c->last_peak = 0;
}
/* whb_demod::demod, whb.cpp:632-707 */
static int whb_sample(orc_t *o, chan *c, int thresh, int pwr, const int16_t *iq)
{
	int triggered = 0;
	if (pwr > thresh) {
		if (!c->timeout_cnt)
			whb_reset(c);
		c->timeout_cnt = (int)(8 * c->spb);
	}
	if (c->timeout_cnt) {
		triggered = 1;
		int dev = orc_fm_dev_nrzs(iq[0], iq[1], c->last_i, c->last_q);
		if (o->tap_mask & 2)
			tap_put(o, c, 1, &dev);
		double y = biquad_step(&c->lp, dev);
		if (o->tap_mask & 4)
			tap_put(o, c, 2, &y);
		dev = (int)y;
		if (!c->synced) {
			double a = biquad_step(&c->lp_avg, 0.5 * dev);
			if (o->tap_mask & 4)
				tap_put(o, c, 2, &a);
			c->avg_of = (int)a;
		}
		c->timeout_cnt--;
		int tdiff = (int)(c->step - c->last_peak);
		if (dev < c->avg_of && dev > c->last_dev && (tdiff > 3 * c->spb / 4)) {
			whb_bit(c, 0);
			c->bitcnt++;
			int bit0 = (int)((tdiff + c->spb / 2) / c->spb);
			for (int n = 1; n < bit0; n++) {
				whb_bit(c, 1);
				c->bitcnt++;
			}
			c->last_peak = c->step;
		}
		c->last_dev = dev;
		if (c->synced)
			c->rssi_d += (iq[0] * iq[0] + iq[1] * iq[1]);
		if (!c->timeout_cnt) {
			if (c->synced) {
				for (int n = 0; n < 16; n++)
					whb_bit(c, 0);
				whb_flush(o, c, d2i(10 * log10(1 + DIV4000(c->rssi_d))), c->offset);
			}
			whb_reset(c);
			c->rssi_d = 0;
		}
	}
	c->last_i = iq[0];
	c->last_q = iq[1];
	c->step++;
	return triggered;
}

/* ---- wiring --------------------------------------------------------------------------------- */
static void chan_init(chan *c, int kind, int type, double spb)
{
	memset(c, 0, sizeof(*c));
	c->kind = kind;
	c->type = type;
	c->sr_cnt = -1;
	c->seen.esz = sizeof(seen_t);
	c->spb = spb;
	if (type == ORC_TFA_WHB) {
		whb_reset(c);
		biquad_init(&c->lp, 2.0 / spb);      /* whb.cpp:610 */
		biquad_init(&c->lp_avg, 0.0025 / spb); /* whb.cpp:611 */
	} else if (type != ORC_TFA_1) {
		tfa2_reset(c);
		biquad_init(&c->lp, 0.5 / spb);      /* tfa2.cpp:321; iir_fac is 0.5 for every registration */
	}
}

/* registration order and spb constants: main.cpp:171-218 */
orc_t *orc_create(int types, int filter, int thresh)
{
	orc_t *o = (orc_t *)calloc(1, sizeof(*o));
	if (!o)
		return NULL;
	if (types & (1 << ORC_TFA_1)) chan_init(&o->ch[o->n_ch], o->n_ch, ORC_TFA_1, 10.0), o->n_ch++;
	if (types & (1 << ORC_TFA_2)) chan_init(&o->ch[o->n_ch], o->n_ch, ORC_TFA_2, (1536000 / 4.0) / 17240), o->n_ch++;
	if (types & (1 << ORC_TFA_3)) chan_init(&o->ch[o->n_ch], o->n_ch, ORC_TFA_3, (1536000 / 4.0) / 9600), o->n_ch++;
	if (types & (1 << ORC_TX22)) chan_init(&o->ch[o->n_ch], o->n_ch, ORC_TX22, (1536000 / 4.0) / 8842), o->n_ch++;
	if (types & (1 << ORC_TFA_WHB)) chan_init(&o->ch[o->n_ch], o->n_ch, ORC_TFA_WHB, (1536000 / 4.0) / 6000), o->n_ch++;
	decim_init(&o->dec, filter);
	/* fsk_demod::fsk_demod, fm_demod.cpp:19-32 */
	o->thresh = thresh;
	o->thresh_mode = 0;
	if (thresh == 0) {
		o->thresh = 500;
		o->thresh_mode = 1;
	}
	o->frames.esz = sizeof(orc_frame);
	o->records.esz = sizeof(orc_record);
	o->blocks.esz = sizeof(orc_block_trace);
	o->tap[0].esz = o->tap[1].esz = sizeof(int32_t);
	o->tap[2].esz = sizeof(double);
	o->tap_ch[0].esz = o->tap_ch[1].esz = o->tap_ch[2].esz = 1;
	return o;
}
void orc_destroy(orc_t *o)
{
	if (!o)
		return;
	for (int k = 0; k < o->n_ch; k++)
		free(o->ch[k].seen.p);
	free(o->frames.p);
	free(o->records.p);
	free(o->blocks.p);
	for (int k = 0; k < 3; k++) {
		free(o->tap[k].p);
		free(o->tap_ch[k].p);
	}
	free(o);
}
void orc_set_taps(orc_t *o, int mask) { o->tap_mask = mask; }

/* fsk_demod::process, fm_demod.cpp:34-74, with demodulator::start (decoder.cpp:118-122) inlined */
static void process_block(orc_t *o, const int16_t *d, int len)
{
	int triggered = 0;
	orc_block_trace tr;
	o->runs++;
	tr.thresh = o->thresh;
	for (int k = 0; k < o->n_ch; k++)
		if (o->ch[k].last_bit_idx)
			o->ch[k].last_bit_idx -= len;
	for (int i = 0; i < len; i += 2) {
		int pwr = abs(d[i]) + abs(d[i + 1]);
		int t = 0;
		o->cur_pos = o->pos_base + i / 2;
		for (int k = 0; k < o->n_ch; k++) {
			chan *c = &o->ch[k];
			if (c->type == ORC_TFA_1)
				t += tfa1_sample(o, c, o->thresh, pwr, i, d + i);
			else if (c->type == ORC_TFA_WHB)
				t += whb_sample(o, c, o->thresh, pwr, d + i);
			else
				t += tfa2_sample(o, c, o->thresh, pwr, i, d + i);
		}
		if (t)
			triggered++;
	}
	o->triggered_avg = (31 * o->triggered_avg + triggered) / 32;
	if (o->thresh_mode == 1 && (o->runs & 3) == 0) {
		if (o->triggered_avg >= len / 32)
			o->thresh += 2;
		else if (o->triggered_avg <= len / 64 && o->thresh > 50)
			o->thresh -= 2;
	}
	tr.triggered = triggered;
	tr.triggered_avg = o->triggered_avg;
	vec_push(&o->blocks, &tr);
	o->pos_base += len / 2;
}

/* engine::run replay branch, engine.cpp:63-93: whole 65536-byte blocks only */
long orc_process(orc_t *o, const uint8_t *iq, size_t nbytes)
{
	long blocks = 0;
	for (size_t off = 0; off + ORC_BLOCK_BYTES <= nbytes; off += ORC_BLOCK_BYTES) {
		size_t ld = decim_block(&o->dec, iq + off, ORC_BLOCK_BYTES, o->dbuf);
		process_block(o, o->dbuf, (int)ld);
		blocks++;
	}
	return blocks;
}

size_t orc_n_frames(const orc_t *o) { return o->frames.n; }
const orc_frame *orc_frames(const orc_t *o) { return (const orc_frame *)o->frames.p; }
size_t orc_n_records(const orc_t *o) { return o->records.n; }
const orc_record *orc_records(const orc_t *o) { return (const orc_record *)o->records.p; }
size_t orc_n_blocks(const orc_t *o) { return o->blocks.n; }
const orc_block_trace *orc_blocks(const orc_t *o) { return (const orc_block_trace *)o->blocks.p; }
size_t orc_n_tap(const orc_t *o, int kind) { return (kind >= 0 && kind < 3) ? o->tap[kind].n : 0; }
const void *orc_tap(const orc_t *o, int kind) { return (kind >= 0 && kind < 3) ? o->tap[kind].p : NULL; }
const uint8_t *orc_tap_chan(const orc_t *o, int kind) { return (kind >= 0 && kind < 3) ? (const uint8_t *)o->tap_ch[kind].p : NULL; }
int orc_thresh(const orc_t *o) { return o->thresh; }
long orc_inverted_syncs(const orc_t *o) { return o->inverted; }
void orc_clear_results(orc_t *o)
{
	o->frames.n = o->records.n = o->blocks.n = 0;
	o->tap[0].n = o->tap[1].n = o->tap[2].n = 0;
	o->tap_ch[0].n = o->tap_ch[1].n = o->tap_ch[2].n = 0;
}

/* main.cpp:45-50 (-X): decoder::store_bytes (decoder.cpp:35-40) then flush(0) */
int orc_parse(int type, const uint8_t *bytes, int len, orc_frame *frame, orc_record *recs, int max_recs)
{
	orc_t *o = orc_create(1 << type, 0, 0);
	if (!o || o->n_ch != 1) {
		orc_destroy(o);
		return -1;
	}
	chan *c = &o->ch[0];
	o->cur_pos = -1;
	if (len > 256)
		len = 256;
	memcpy(c->rdata, bytes, (size_t)len);
	c->byte_cnt = len;
	c->synced = 1;
	if (type == ORC_TFA_1)
		tfa1_flush(o, c, 0);
	else if (type == ORC_TFA_WHB)
		whb_flush(o, c, 0, 0);
	else
		tfa2_flush(o, c, 0, 0);
	int n = -1;
	if (o->frames.n) {
		if (frame)
			*frame = ((orc_frame *)o->frames.p)[0];
		n = (int)o->records.n;
		for (int k = 0; k < n && k < max_recs; k++)
			recs[k] = ((orc_record *)o->records.p)[k];
	}
	orc_destroy(o);
	return n;
}

/* decoder::execute_handler, decoder.cpp:67-96: "<handler> id temp hum seq alarm rssi flags ts" */
int orc_format_exec(const orc_record *r, char *buf, size_t cap)
{
	if (r->type != ORC_TFA_WHB) {
		uint64_t nid = r->id | (uint64_t)(int64_t)(r->type << 24);
		return snprintf(buf, cap, "%04llx %+.1f %g %i %i %i %i", (unsigned long long)nid, r->temp, r->humidity,
				r->sequence, r->alarm, r->rssi, r->flags & 0xff);
	}
	return snprintf(buf, cap, "%013llx %+.1f %g %i %i %i %i", (unsigned long long)r->id, r->temp, r->humidity,
			r->sequence, r->alarm, r->rssi, r->flags & 0xff);
}
