// decim.cu - downconvert(passes)::process_iq (dsp_stuff.cpp:232-264) for ANY number of passes: the decimation
// sweep of BASELINE configs[4] (/2 .. /32).  The decode path itself always runs passes = 2 through the fused
// front-end (frontend.cu); this file is the stand-alone cascade: passes-1 first-stage filters (process2x1,
// 8 taps, dsp_stuff.cpp:204-230), then the 20-tap filter (process2x, narrow or wide, :172-202), one launch per
// stage, int16 I,Q between stages exactly as the reference stores them in place.
//
// One stage: out[k] = sum_n floor(x[2k-(T-2)+n] * t[n] / 2^16) per channel, x[<0] = 0 (zero hist0).  As in the
// front-end every tap product is ONE round-toward-minus-infinity FMA on an accumulator kept inside
// [2^23, 2^24), where the fp32 grid is the integers: fma.rm(x, t/2^16, acc) == acc + floor(x*t/2^16) exactly
// (x and t/2^16 are exact floats, the FMA rounds once), I and Q ride in one fma.rm.f32x2.  int16 inputs are
// converted with I2F (no byte structure to exploit): the accumulator starts at 2^23+2^22 and moves by less than
// 20k (sum|t|/2^16 <= 1.43, |x| <= 13.7k).  Raw bytes (the first stage) use the front-end's integer-only
// byte->float trick, whose known integer offset is removed at the end (see kCvtBase below).
//
// A thread produces 8 consecutive outputs from 2*8+T-2 consecutive inputs held in registers; a warp reads one
// contiguous span (32- or 64-bit loads, every sector fully used, the T-2 sample overlap between neighbours
// hits L1) and writes 32 B per thread.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "fir_taps.h"
#include "tfr_dev.h"

namespace tfr {

typedef unsigned long long f2;   // packed f32x2: lo = I, hi = Q

__device__ __forceinline__ f2 dpack2(float lo, float hi)
{
	f2 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ f2 dfma2_rm(f2 a, f2 b, f2 c)
{
	f2 d;
	asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}

constexpr int kStageOut = 8;          // outputs per thread
constexpr int kStageThreads = 256;
constexpr float kAcc0 = 12582912.0f;  // 2^23 + 2^22
constexpr uint32_t kAcc0Bits = 0x4B400000u;

template <int TAPS, bool WIDE>
__host__ __device__ constexpr int stage_tap_int(int n)
{
	return TAPS == 8 ? t2_tap(n) : t1_tap(WIDE, n);
}
// Raw bytes are turned into floats with integer instructions only, as in the front-end: bits 0x47038000 + (b << 8)
// are the float 33664 + b = 33*1024 + (b-128) (binade [2^15, 2^16), ulp 1/256).  With the tap scaled by 1/1024
// (x = (b-128)<<6, so x*t/2^16 = (b-128)*t/1024) every FMA then adds floor((b-128)*t/1024) + 33*t: the wanted
// term plus an integer that is removed when the accumulator bits are read back.
constexpr uint32_t kCvtBase = 0x47038000u;
constexpr int kCvtMul = 33;
template <int TAPS, bool WIDE>
__host__ __device__ constexpr int u8_offset()
{
	int s = 0;
	for (int n = 0; n < TAPS; n++) s += kCvtMul * stage_tap_int<TAPS, WIDE>(n);
	return s;
}
template <int TAPS, bool WIDE>
__host__ __device__ constexpr bool u8_chain_in_range()
{
	int s = (1 << 23) + (1 << 22);
	for (int n = 0; n < TAPS; n++) {
		s += kCvtMul * stage_tap_int<TAPS, WIDE>(n);
		if (s - 20000 < (1 << 23) || s + 20000 >= (1 << 24)) return false;
	}
	return true;
}
static_assert(u8_chain_in_range<8, false>() && u8_chain_in_range<20, false>() && u8_chain_in_range<20, true>(),
	      "u8 stage accumulator leaves the integer binade");

// U8IN: `in` is raw rtl-sdr offset-binary IQ (2 B per pair), converted with (b-128)<<6 (engine.cpp:77-78);
// otherwise int16 I,Q (4 B per pair) as left by the previous stage.
template <int TAPS, bool WIDE, bool U8IN>
__global__ void __launch_bounds__(kStageThreads) decim_stage_kernel(const void *__restrict__ in, int16_t *__restrict__ out, long long n_in,
								   long long n_out)
{
	constexpr int W = 2 * kStageOut + TAPS - 2;   // input pairs a thread needs
	const long long t = (long long)blockIdx.x * kStageThreads + threadIdx.x;
	const long long k0 = t * kStageOut;
	if (k0 >= n_out) return;
	const long long j0 = 2 * k0 - (TAPS - 2);     // first input pair (even; negative at the start of the buffer)
	f2 x[W];
	if (U8IN) {
		// one 32-bit word = two pairs; j0 is even, so the span is word aligned
		const uint32_t *src = reinterpret_cast<const uint32_t *>(in);
#pragma unroll
		for (int w = 0; w < W / 2; w++) {
			const long long j = j0 + 2 * w;
			uint32_t v = 0x80808080u;                                  // zero signal
			if (j >= 0 && j + 1 < n_in) v = __ldg(src + (j >> 1));
			else if (j >= 0 && j < n_in) v = (v & 0xffff0000u) | reinterpret_cast<const uint16_t *>(in)[j];
			x[2 * w] = dpack2(__uint_as_float(__byte_perm(v, 0, 0x4404) + kCvtBase), __uint_as_float(__byte_perm(v, 0, 0x4414) + kCvtBase));
			x[2 * w + 1] = dpack2(__uint_as_float(__byte_perm(v, 0, 0x4424) + kCvtBase), __uint_as_float(__byte_perm(v, 0, 0x4434) + kCvtBase));
		}
	} else {
		// one 64-bit word = two pairs
		const uint2 *src = reinterpret_cast<const uint2 *>(in);
#pragma unroll
		for (int w = 0; w < W / 2; w++) {
			const long long j = j0 + 2 * w;
			uint2 v = make_uint2(0u, 0u);
			if (j >= 0 && j + 1 < n_in) v = __ldg(src + (j >> 1));
			else if (j >= 0 && j < n_in) v.x = reinterpret_cast<const uint32_t *>(in)[j];
			x[2 * w] = dpack2((float)(int)(int16_t)(v.x & 0xffff), (float)(int)(int16_t)(v.x >> 16));
			x[2 * w + 1] = dpack2((float)(int)(int16_t)(v.y & 0xffff), (float)(int)(int16_t)(v.y >> 16));
		}
	}
	uint32_t o[kStageOut];
#pragma unroll
	for (int k = 0; k < kStageOut; k++) {
		f2 acc = dpack2(kAcc0, kAcc0);
#pragma unroll
		for (int n = 0; n < TAPS; n++) {
			const float c = (float)stage_tap_int<TAPS, WIDE>(n) * (U8IN ? 1.0f / 1024.0f : 1.0f / 65536.0f);
			acc = dfma2_rm(x[2 * k + n], dpack2(c, c), acc);
		}
		constexpr uint32_t bias = kAcc0Bits + (U8IN ? (uint32_t)u8_offset<TAPS, WIDE>() : 0u);
		const uint32_t yi = (uint32_t)(acc & 0xffffffffull) - bias, yq = (uint32_t)(acc >> 32) - bias;
		o[k] = (yi & 0xffffu) | (yq << 16);
	}
	uint32_t *dst = reinterpret_cast<uint32_t *>(out) + k0;
	if (k0 + kStageOut <= n_out) {
		reinterpret_cast<uint4 *>(dst)[0] = make_uint4(o[0], o[1], o[2], o[3]);
		reinterpret_cast<uint4 *>(dst)[1] = make_uint4(o[4], o[5], o[6], o[7]);
	} else {
#pragma unroll
		for (int k = 0; k < kStageOut; k++)
			if (k0 + k < n_out) dst[k] = o[k];
	}
}

template <int TAPS, bool WIDE, bool U8IN>
static cudaError_t launch_stage(const void *in, int16_t *out, long long n_in, long long n_out, cudaStream_t s)
{
	if (n_out <= 0) return cudaSuccess;
	const long long threads = (n_out + kStageOut - 1) / kStageOut;
	const long long blocks = (threads + kStageThreads - 1) / kStageThreads;
	if (blocks > 0x7fffffffll) return cudaErrorInvalidValue;
	decim_stage_kernel<TAPS, WIDE, U8IN><<<(unsigned)blocks, kStageThreads, 0, s>>>(in, out, n_in, n_out);
	return cudaGetLastError();
}

// iq: device pointer to n_pairs raw IQ pairs; tmp[0], tmp[1]: device int16 buffers of >= n_pairs/2 and >= n_pairs/4
// pairs; the result (n_pairs >> passes pairs) is in *result (one of the two).  Stages run back to back on `s`.
// in_i16: iq holds int16 I,Q pairs instead of raw bytes (the data fsk_demod::process(int16_t*,int) / process_iq get)
cudaError_t launch_downconvert(const void *iq, long long n_pairs, int passes, int wide, int16_t *tmp0, int16_t *tmp1, int16_t **result,
			       cudaStream_t s, bool in_i16)
{
	const void *cur = iq;
	long long n = n_pairs;
	int16_t *bufs[2] = { tmp0, tmp1 };
	int16_t *dst = nullptr;
	for (int p = 0; p < passes; p++) {
		dst = bufs[p & 1];
		const long long no = n / 2;
		const bool last = (p == passes - 1), first = (p == 0) && !in_i16;
		cudaError_t e;
		if (!last) e = first ? launch_stage<8, false, true>(cur, dst, n, no, s) : launch_stage<8, false, false>(cur, dst, n, no, s);
		else if (wide) e = first ? launch_stage<20, true, true>(cur, dst, n, no, s) : launch_stage<20, true, false>(cur, dst, n, no, s);
		else e = first ? launch_stage<20, false, true>(cur, dst, n, no, s) : launch_stage<20, false, false>(cur, dst, n, no, s);
		if (e != cudaSuccess) return e;
		cur = dst;
		n = no;
	}
	*result = dst;
	return cudaSuccess;
}

}  // namespace tfr
