// demod_dev.cuh - device functions shared by the back-end kernels: region enumeration over the sparse
// decimated buffer, the discriminators and biquad, and the bit framers (store_bit: tfa1.cpp:120-134,
// tfa2.cpp:281-314, whb.cpp:566-603).  The demodulators themselves live in backend2.cu (TFA_1, TFA_2 family)
// and backend.cu (WeatherHub).  See backend.cu for the floating-point notes.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "tfr_dev.h"

namespace tfr {

// ------------------------------------------------------------------------------------------------
// region enumeration: the samples of a block the front-end kept = [0, carry_in) U segments
// ------------------------------------------------------------------------------------------------
struct Regions {
	int n;
	int start[kMaxSeg + 1];
	int end[kMaxSeg + 1];
};
__device__ __forceinline__ void build_regions(const TileDesc &td, int carry_in, Regions &r)
{
	r.n = 0;
	int cur_s = 0, cur_e = carry_in;   // may be empty
	const int ns = td.n_seg;
	for (int k = 0; k < ns; k++) {
		const int s = td.seg_start[k], e = s + td.seg_len[k];
		if (cur_e > cur_s && s <= cur_e) {
			cur_e = max(cur_e, e);
		} else {
			if (cur_e > cur_s) { r.start[r.n] = cur_s; r.end[r.n] = cur_e; r.n++; }
			cur_s = s;
			cur_e = e;
		}
	}
	if (cur_e > cur_s) { r.start[r.n] = cur_s; r.end[r.n] = cur_e; r.n++; }
}

__device__ __forceinline__ int pwr_of(uint32_t w)
{
	const int i = (int)(int16_t)(w & 0xffff), q = (int)(int16_t)(w >> 16);
	return abs(i) + abs(q);
}

// ------------------------------------------------------------------------------------------------
// discriminators and biquad
// ------------------------------------------------------------------------------------------------
// fm_dev_nrzs, dsp_stuff.cpp:269-279
__device__ __forceinline__ int fm_dev_nrzs(int ar, int aj, int br, int bj)
{
	int cr = ar * br + aj * bj;
	if (cr > 1000000000) cr = 1000000000;
	if (cr < -1000000000) cr = -1000000000;
	return cr;
}

// fm_dev, dsp_stuff.cpp:284-292.  The products are exact in double; what matters is the sign of a zero
// result (decides the +-pi branch) and the value atan2 returns on the axes and diagonals, where
// angle*16384/pi lands exactly on an integer before truncation.  Those cases return glibc's values.
__device__ __forceinline__ int fm_dev(int ar, int aj, int br, int bj)
{
	const double cr = __dadd_rn(__dmul_rn((double)aj, (double)bj), __dmul_rn((double)ar, (double)br));
	const double cj = __dsub_rn(__dmul_rn((double)br, (double)aj), __dmul_rn((double)ar, (double)bj));
	// what glibc's atan2 returns for (0,-1), (1,0), (1,1), (1,-1)
	const double PI = 0x1.921fb54442d18p+1, PI_2 = 0x1.921fb54442d18p+0, PI_4 = 0x1.921fb54442d18p-1,
		     PI3_4 = 0x1.2d97c7f3321d2p+1;
	double ang;
	if (cj == 0.0) {
		ang = (cr > 0.0 || (cr == 0.0 && !signbit(cr))) ? cj : copysign(PI, cj);
	} else if (cr == 0.0) {
		ang = copysign(PI_2, cj);
	} else if (fabs(cj) == fabs(cr)) {
		ang = copysign(cr > 0.0 ? PI_4 : PI3_4, cj);
	} else {
		ang = atan2(cj, cr);
	}
	return __double2int_rz(__dmul_rn(ang, 5215.189175235227));   // fl(16384/pi), see header
}

// fm_dev for the bulk of the samples: the same integer as fm_dev() above, ~8x fewer instructions.
// The result is trunc(angle * fl(16384/pi)); an approximation of the angle with absolute error e gives the same
// integer unless angle*K lies within e*K of a truncation boundary.  The angle comes from a degree-12 polynomial
// in q^2 (q = min/max of |cr|,|cj|, octant folding; max error 6e-12 rad = 3.1e-8 output units against glibc's
// atan2, tools/fit_atan.py) and a Newton reciprocal; every sample whose scaled angle is within 1e-6 of an
// integer - and the axis / diagonal / zero cases, whose value depends on signed zeros and on glibc's exact
// return values - takes fm_dev().  Products and sums are exact in int32 (|I|,|Q| <= 12.2k).
__device__ __forceinline__ int fm_dev_fast(int ar, int aj, int br, int bj)
{
	const int cri = aj * bj + ar * br, cji = br * aj - ar * bj;
	const int axi = abs(cri), ayi = abs(cji);
	const bool special = (cji == 0 || cri == 0 || axi == ayi);   // (the arithmetic below is then discarded)
	const int mxi = max(axi, ayi), mni = min(axi, ayi);
	// non-negative int -> double without I2F.F64: the low mantissa word of 2^52
	const double mx = __hiloint2double(0x43300000, mxi) - 4503599627370496.0;
	const double mn = __hiloint2double(0x43300000, mni) - 4503599627370496.0;
	float rf;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"((float)mxi));
	double r = (double)rf;                       // 2^-22 relative; two Newton steps -> double precision
	double e = fma(-mx, r, 1.0);
	r = fma(r, e, r);
	e = fma(-mx, r, 1.0);
	r = fma(r, e, r);
	const double q = mn * r, z = q * q;
	double p = 0x1.b79940e5536f0p-12;
	p = fma(p, z, -0x1.a2d5e8268d02fp-9);
	p = fma(p, z, 0x1.7676be8c4ef52p-7);
	p = fma(p, z, -0x1.aa15aa0c3e761p-6);
	p = fma(p, z, 0x1.658dd34696c8fp-5);
	p = fma(p, z, -0x1.ed56cc14c7895p-5);
	p = fma(p, z, 0x1.33018fb5c9b3fp-4);
	p = fma(p, z, -0x1.72a787e7c5385p-4);
	p = fma(p, z, 0x1.c6df953d05217p-4);
	p = fma(p, z, -0x1.248fb94625489p-3);
	p = fma(p, z, 0x1.99997c8640331p-3);
	p = fma(p, z, -0x1.55555513c3688p-2);
	p = fma(p, z, 0x1.ffffffffe73b8p-1);
	double a = q * p;                            // atan(min/max) in (0, pi/4)
	if (ayi > axi) a = 0x1.921fb54442d18p+0 - a;
	if (cri < 0) a = 0x1.921fb54442d18p+1 - a;
	const double v = a * 5215.189175235227;      // in (0, 16384)
	const double t = __dadd_rz(v, 4503599627370496.0);
	const double fr = v - (t - 4503599627370496.0);   // v - floor(v), exact
	if (special || fr < 1e-6 || fr > 1.0 - 1e-6) return fm_dev(ar, aj, br, bj);
	const int n = __double2loint(t);
	return cji < 0 ? -n : n;
}

// iir2::step, dsp_stuff.cpp:47-55, association as built
__device__ __forceinline__ double biquad_step(Biquad &f, const BiquadCoef &k, double dn)
{
	const double t1 = __dadd_rn(__dmul_rn(k.b2, f.d2), __dmul_rn(k.a1, f.y0));
	const double t2 = __dadd_rn(__dmul_rn(k.b0, dn), __dmul_rn(k.b1, f.d1));
	const double y = __dadd_rn(__dadd_rn(t1, t2), __dmul_rn(k.a2, f.y1));
	f.y1 = f.y0;
	f.y0 = y;
	f.d2 = f.d1;
	f.d1 = dn;
	return y;
}

// (int)y, truncation toward zero, for |y| < 2^31 without the variable-latency F2I.F64: |y| + 2^52 rounded toward
// zero leaves floor(|y|) in the low mantissa word; the sign comes back from y's high word
__device__ __forceinline__ int trunc_to_int(double y)
{
	const int lo = __double2loint(__dadd_rz(fabs(y), 4503599627370496.0));
	return (__double2hiint(y) < 0) ? -lo : lo;
}
// exact int -> double without the (slow, variable latency) I2F.F64: 2^52 + 2^31 + v, minus the offset
__device__ __forceinline__ double int_to_double(int v)
{
	return __dsub_rn(__hiloint2double(0x43300000, (int)((uint32_t)v + 0x80000000u)), 4503601774854144.0);
}

// ------------------------------------------------------------------------------------------------
// walker context
// ------------------------------------------------------------------------------------------------
struct Walk {
	const BackParams *p;
	int stream, demod;
	int64_t pos;          // 384 kS/s position of the current sample
	DemodState s;
	uint32_t tap_n[3];
};

__device__ __forceinline__ void tap_i32(Walk &w, int kind, int v)
{
	if (!w.p->tap_cap) return;
	const uint32_t n = w.tap_n[kind]++;
	if (n < w.p->tap_cap)
		w.p->tap_i32[kind][((size_t)w.stream * kMaxDemods + w.demod) * w.p->tap_cap + n] = v;
}
__device__ __forceinline__ void tap_f64(Walk &w, double v)
{
	if (!w.p->tap_cap) return;
	const uint32_t n = w.tap_n[2]++;
	if (n < w.p->tap_cap) w.p->tap_f64[((size_t)w.stream * kMaxDemods + w.demod) * w.p->tap_cap + n] = v;
}

// a decoder::flush that passed its length gate becomes a candidate frame for parse_kernel
static __device__ void emit_frame(Walk &w, double rssi_raw, int offset)
{
	const uint32_t k = atomicAdd(&w.p->counters->n_frames, 1u);
	if (k >= w.p->max_frames) {
		w.p->counters->overflow = 1;
		return;
	}
	DevFrame &f = w.p->frames[k];
	f.status = -2;   // not a frame until everything below is in place
	f.stream = w.stream;
	f.demod = w.demod;
	f.type = w.p->cfg->d[w.demod].type;
	f.byte_cnt = w.s.byte_cnt;
	f.offset = offset;
	f.n_records = 0;
	f.first_record = w.p->slot_tag;   // until parsed: whose parse_kernel this frame is for
	f.pos = w.pos;
	f.rssi_raw = rssi_raw;
	for (int n = 0; n < kMaxRdata; n++) f.rdata[n] = w.s.rdata[n];
	__threadfence();
	f.status = -1;
}

// ---- TFA_1 ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tfa1_bit(DemodState &s, int bit)
{
	s.sr = (s.sr >> 1) | ((uint32_t)bit << 31);
	if ((s.sr & 0xffff) == 0xd42d) {
		s.sr_cnt = 0;
		s.byte_cnt = 0;
	}
	if (s.sr_cnt == 0) {
		if (s.byte_cnt < kRdataBytes) s.rdata[s.byte_cnt] = (uint8_t)(s.sr & 0xff);
		s.byte_cnt++;
	}
	if (s.sr_cnt >= 0) s.sr_cnt = (s.sr_cnt + 1) & 7;
}
// ---- TFA_2 / TFA_3 / TX22 ---------------------------------------------------------------------------
__device__ __forceinline__ void tfa2_bit(DemodState &s, int bit)
{
	s.sr = (s.sr << 1) | (uint32_t)bit;
	if ((s.sr & 0xffff) == 0x2dd4) {
		s.sr_cnt = 0;
		s.rdata[0] = (uint8_t)((s.sr >> 8) & 0xff);
		s.byte_cnt = 1;
		s.invert = 0;
	}
	if (((~s.sr) & 0xffff) == 0x2dd4) {
		s.sr_cnt = 0;
		s.rdata[0] = (uint8_t)~((s.sr >> 8) & 0xff);
		s.byte_cnt = 1;
		s.invert = 1;
		s.inv_cnt++;   // the reference prints "Inverted SYNC" here
	}
	if (s.sr_cnt == 0) {
		if (s.byte_cnt < kRdataBytes)
			s.rdata[s.byte_cnt] = s.invert ? (uint8_t)~(s.sr & 0xff) : (uint8_t)(s.sr & 0xff);
		s.byte_cnt++;
	}
	if (s.sr_cnt >= 0) s.sr_cnt = (s.sr_cnt + 1) & 7;
}
__device__ __forceinline__ void tfa2_reset(DemodState &s)
{
	s.offset = 0;
	s.bitcnt = 0;
	s.dmin = 32767;
	s.dmax = -32767;
	s.last_bit = 0;
	s.rssi_i = 0;
	s.inv_cnt = 0;
}
// ---- WeatherHub ---------------------------------------------------------------------------------
__device__ __forceinline__ void whb_bit(DemodState &s, int bit)
{
	if (bit == s.w_last_bit) s.w_psk = 1 - s.w_psk;
	if (s.w_psk == s.w_last_psk) s.w_nrzs = 1 - s.w_nrzs;
	s.w_last_bit = bit;
	s.w_last_psk = s.w_psk;
	const int d = s.w_nrzs ^ ((s.w_lfsr >> 16) & 1) ^ ((s.w_lfsr >> 11) & 1);
	s.w_lfsr = (s.w_lfsr << 1) | (uint32_t)s.w_nrzs;
	s.sr = (s.sr >> 1) | ((uint32_t)d << 31);
	if (s.sr == 0x2bd42d4bu) {
		s.synced = 1;
		s.sr_cnt = 0;
		s.rdata[0] = (uint8_t)(s.sr & 0xff);
		s.rdata[1] = (uint8_t)((s.sr >> 8) & 0xff);
		s.rdata[2] = (uint8_t)((s.sr >> 16) & 0xff);
		s.byte_cnt = 3;
	}
	if (s.sr_cnt == 0) {
		if (s.byte_cnt < kRdataBytes) s.rdata[s.byte_cnt] = (uint8_t)((s.sr >> 24) & 0xff);
		s.byte_cnt++;
	}
	if (s.sr_cnt >= 0) s.sr_cnt = (s.sr_cnt + 1) & 7;
}
static __device__ void whb_flush(Walk &w)
{
	DemodState &s = w.s;
	if (!(s.byte_cnt < 11 || s.byte_cnt > 60)) emit_frame(w, s.rssi_d, s.offset);
	s.sr_cnt = -1;
	s.sr = 0;
	s.byte_cnt = 0;
	s.synced = 0;
}
}  // namespace tfr
