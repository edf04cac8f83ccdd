// fir_exact.cuh - the exact-floor arithmetic of the reference's two decimator stages (dsp_stuff.cpp:172-230), shared by
// every kernel that computes decimated samples: the dense front-end (frontend.cu), the screening front-end's candidate
// check and the window kernel of frontend_screen.cu.
#pragma once
#include "frontend_common.cuh"

namespace tfr {

// ------------------------------------------------------------------------------------------------
// exact-floor accumulator bookkeeping
// ------------------------------------------------------------------------------------------------
// byte b -> float z = 33664 + b, built by integer ops in the [32768,65536) binade (ulp 1/256): bits =
// 0x47038000 + (b << 8).  z = 33*1024 + (b-128), so with c1 = t2/1024 every stage-1 FMA adds
// floor((b-128)*t2/1024) + 33*t2 : the wanted floor((x*t2)>>16 for x=(b-128)<<6) plus an integer.
constexpr uint32_t kCvtBase = 0x47038000u;
constexpr int kCvtMul = 33;
// stage-1 accumulator: starts at kA1, ends at kM1 + y1 with kM1 a multiple of 65536 so that the
// offset it injects into stage 2, kM1*t1/65536 = 163*t1, is again an integer.
constexpr int kM1Mul = 163;
constexpr int kM1 = kM1Mul * 65536;                       // 10,682,368
constexpr int kA1 = kM1 - kCvtMul * t2_sum();             //  8,433,616  (>= 2^23 + 8520)
static_assert(kA1 - 8520 >= (1 << 23) && kM1 + 8520 < (1 << 24), "stage-1 accumulator leaves the integer binade");
// stage 2 runs as two 10-tap chains (taps 0..9 and 10..19) so that each chain's offset fits the binade
constexpr int kA2a = (1 << 23) + 1300000;
constexpr int kA2b = (1 << 23) + 100000;

__host__ __device__ constexpr bool chain_in_range(bool wide, int lo, int hi, int start)
{
	int s = start;
	for (int n = lo; n < hi; n++) {
		s += kM1Mul * t1_tap(wide, n);
		if (s - 13000 < (1 << 23) || s + 13000 >= (1 << 24)) return false;
	}
	return true;
}
static_assert(chain_in_range(false, 0, 10, kA2a) && chain_in_range(false, 10, 20, kA2b), "narrow stage-2 chain range");
static_assert(chain_in_range(true, 0, 10, kA2a) && chain_in_range(true, 10, 20, kA2b), "wide stage-2 chain range");

// y2 = (bitsA - 0x4B000000 + 2^23 - kA2a - offA) + (same for B) with off = 163*sum(taps of the chain)
__host__ __device__ constexpr uint32_t y2_bias(bool wide)
{
	return 2u * (0x4B000000u - (1u << 23)) + (uint32_t)kA2a + (uint32_t)kA2b +
	       (uint32_t)(kM1Mul * t1_sum(wide, 0, 10)) + (uint32_t)(kM1Mul * t1_sum(wide, 10, 20));
}

// two bytes (I,Q) of a raw sample -> packed floats 33664+b
__device__ __forceinline__ f2 cvt_iq(uint32_t w, int half)
{
	uint32_t bi = __byte_perm(w, 0, half ? 0x4424 : 0x4404) + kCvtBase;
	uint32_t bq = __byte_perm(w, 0, half ? 0x4434 : 0x4414) + kCvtBase;
	return pack2(__uint_as_float(bi), __uint_as_float(bq));
}

template <bool WIDE>
__device__ __forceinline__ f2 c2pair(int n)
{
	float c = (float)t1_tap(WIDE, n) * (1.0f / 65536.0f);
	return pack2(c, c);
}
__device__ __forceinline__ f2 c1pair(int n)
{
	float c = (float)t2_tap(n) * (1.0f / 1024.0f);
	return pack2(c, c);
}

// stage 1: one 8-tap output from 8 consecutive converted raw samples
__device__ __forceinline__ f2 stage1(const f2 *x)
{
	f2 acc = pack2((float)kA1, (float)kA1);
#pragma unroll
	for (int n = 0; n < 8; n++) acc = fma2_rm(x[n], c1pair(n), acc);
	return acc;
}

// The per-thread cascade of frontend.cu as a function: NIT x 16 consecutive outputs from 96 + NIT x 128 consecutive raw
// bytes.  halo(q), q = 0..5: the q-th 16 bytes of the 96 in front of the thread's first raw byte (raw samples -48..-1);
// row(q): the q-th 16 bytes of its own; out(m, yi, yq) receives output m = 0 .. 16 NIT - 1 in order.
template <bool WIDE, int NIT, class Halo, class Row, class Out>
__device__ __forceinline__ void fir_cascade(const Halo &halo, const Row &row, const Out &out)
{
	f2 ring[32];   // stage-1 outputs (kM1 + y1), slot = index mod 32
	f2 xh[6];      // last 6 converted raw samples
	{
		f2 hx[48];
#pragma unroll
		for (int q = 0; q < 6; q++) {
			const uint4 v = halo(q);
			hx[8 * q + 0] = cvt_iq(v.x, 0); hx[8 * q + 1] = cvt_iq(v.x, 1);
			hx[8 * q + 2] = cvt_iq(v.y, 0); hx[8 * q + 3] = cvt_iq(v.y, 1);
			hx[8 * q + 4] = cvt_iq(v.z, 0); hx[8 * q + 5] = cvt_iq(v.z, 1);
			hx[8 * q + 6] = cvt_iq(v.w, 0); hx[8 * q + 7] = cvt_iq(v.w, 1);
		}
		// y1[j] (j = -18..-1) needs x[2j-6 .. 2j+1]; hx[k] = x[k-48]
#pragma unroll
		for (int j = -18; j < 0; j++) ring[(j + 32) & 31] = stage1(&hx[2 * j + 42]);
#pragma unroll
		for (int k = 0; k < 6; k++) xh[k] = hx[42 + k];
	}
#pragma unroll 1
	for (int it = 0; it < NIT; it++) {
#pragma unroll
		for (int s = 0; s < 8; s++) {
			const uint4 v = row(it * 8 + s);
			f2 x[14];
#pragma unroll
			for (int k = 0; k < 6; k++) x[k] = xh[k];
			x[6] = cvt_iq(v.x, 0); x[7] = cvt_iq(v.x, 1);
			x[8] = cvt_iq(v.y, 0); x[9] = cvt_iq(v.y, 1);
			x[10] = cvt_iq(v.z, 0); x[11] = cvt_iq(v.z, 1);
			x[12] = cvt_iq(v.w, 0); x[13] = cvt_iq(v.w, 1);
#pragma unroll
			for (int jj = 0; jj < 4; jj++) ring[(4 * s + jj) & 31] = stage1(&x[2 * jj]);
#pragma unroll
			for (int k = 0; k < 6; k++) xh[k] = x[8 + k];
#pragma unroll
			for (int mm = 0; mm < 2; mm++) {
				const int m = 2 * s + mm;
				f2 a = pack2((float)kA2a, (float)kA2a), b = pack2((float)kA2b, (float)kA2b);
#pragma unroll
				for (int n = 0; n < 10; n++) {
					a = fma2_rm(ring[(2 * m - 18 + n + 32) & 31], c2pair<WIDE>(n), a);
					b = fma2_rm(ring[(2 * m - 8 + n + 32) & 31], c2pair<WIDE>(n + 10), b);
				}
				uint32_t ai, aq, bi, bq;
				unpack2(a, ai, aq);
				unpack2(b, bi, bq);
				out(it * 16 + m, (int)(ai + bi - y2_bias(WIDE)), (int)(aq + bq - y2_bias(WIDE)));
			}
		}
	}
}

}  // namespace tfr
