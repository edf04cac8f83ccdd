// decim_fused.cu - downconvert(passes)::process_iq (dsp_stuff.cpp:232-264) as ONE kernel for passes = 1..5 (/2 .. /32),
// with the history carried from call to call: the streaming `downconvert` class of dsp_stuff.h:46-56.
//
// decim.cu runs the cascade as one launch per stage with int16 I,Q between the stages in HBM; the /8 cascade then
// moves 2 + 2*2 + 2*1 + 0.5 = 8.5 bytes per raw sample instead of the 2 it reads and 0.5 it writes.  Here a CTA takes a
// tile of 4096 (/2, /4) or 2048 (deeper cascades) raw samples plus the halo the stages need in front of it (18 / 42 / 90 / 186 / 378 raw samples for
// passes 1..5), stages the bytes in shared memory and runs every stage out of shared memory into shared memory: stage
// outputs stay on chip as exact float pairs (the next stage's FMA operand, no int16 round trip, no conversion), only
// the last stage writes int16 I,Q.  The arithmetic is the front-end's: every tap product is one round-toward-minus-
// infinity FMA on an accumulator kept in [2^23, 2^24), fma.rm(x, tap/2^16, acc) == acc + floor(x*tap/2^16) exactly
// (dsp_stuff.cpp:194-195, 222-223), I and Q in one fma.rm.f32x2.
//
// History: every stage's hist0 (the last taps-2 inputs of that stage, dsp_stuff.cpp:180-182,200-201) is a function of
// the raw samples before the buffer, and a stream that starts with all-zero histories is indistinguishable from one
// that was fed zero signal (byte 128) before: so the carried state is simply the last 384 raw samples of the previous
// call (or 128s), and the first tile's halo is read from it.
#include <cuda_runtime.h>
#include <stdint.h>

#include "fir_taps.h"
#include "tfr_dev.h"

namespace tfr {

typedef unsigned long long f2;   // packed f32x2: lo = I, hi = Q
__device__ __forceinline__ f2 fpack2(float lo, float hi)
{
	f2 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ void funpack2(f2 v, uint32_t &lo, uint32_t &hi)
{
	asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ f2 ffma2_rm(f2 a, f2 b, f2 c)
{
	f2 d;
	asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}
__device__ __forceinline__ f2 fadd2_rn(f2 a, f2 b)
{
	f2 d;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}

#ifndef TFR_DC_TILE
#define TFR_DC_TILE 0
#endif
#ifndef TFR_DC_THREADS
#define TFR_DC_THREADS 128
#endif
// raw samples per CTA tile: the deeper cascades want more resident CTAs rather than less halo (measured on B200 for
// 2^30 samples, tile 4096 -> 2048: /8 1.42 -> 1.58, /16 1.21 -> 1.38, /32 0.96 -> 1.20 TB/s; /2 and /4 lose a little)
__host__ __device__ constexpr int dc_tile(int passes) { return TFR_DC_TILE ? TFR_DC_TILE : (passes >= 3 ? 2048 : 4096); }
constexpr int kDcThreads = TFR_DC_THREADS;
constexpr int kDcR = 8;              // outputs per thread and pass over a stage
constexpr int kDcHist = 384;         // raw samples of history kept between calls (needs 378 for passes = 5)
constexpr float kDcAcc0 = 12582912.0f;        // 2^23 + 2^22: accumulator start for float inputs (|sum| < 20k)
constexpr uint32_t kDcAcc0Bits = 0x4B400000u;
// raw bytes -> float with integer instructions only (as the front-end): bits 0x47038000 + (b << 8) = 33664 + b =
// 33*1024 + (b-128); with the tap scaled by 1/1024 every FMA adds floor((b-128)*t/1024) + 33*t
constexpr uint32_t kDcCvtBase = 0x47038000u;
constexpr int kDcCvtMul = 33;

__host__ __device__ constexpr int dc_tap(int taps, bool wide, int n) { return taps == 8 ? t2_tap(n) : t1_tap(wide, n); }
__host__ __device__ constexpr int dc_u8_offset(int taps, bool wide)
{
	int s = 0;
	for (int n = 0; n < taps; n++) s += kDcCvtMul * dc_tap(taps, wide, n);
	return s;
}
__host__ __device__ constexpr int dc_round8(int v) { return (v + 7) & ~7; }
// samples of level k (0 = raw) a tile needs, level P = the tile's own outputs; every level rounded up to whole
// groups of eight (a thread computes eight outputs at a time) and large enough for the (rounded) level above it
template <int P>
struct DcGeom {
	int len[P + 1];
	constexpr DcGeom() : len()
	{
		len[P] = dc_round8(dc_tile(P) >> P);
		for (int k = P; k >= 1; k--) len[k - 1] = dc_round8(2 * len[k] + (k == P ? 18 : 6));
	}
};
// shared-memory placement of a level: level 0 as raw bytes (2 per sample), the others as float pairs with two pairs
// of padding after every sixteen (a thread's eight outputs read from sixteen consecutive inputs: a stride of 18 pairs
// = 9 x 16 bytes keeps the 128-bit loads of a quarter warp on different banks)
__host__ __device__ constexpr int dc_f2_slots(int n) { return n + 2 * ((n + 15) / 16) + 2; }
template <int P>
struct DcSmem {
	int off[P + 1];
	int total;
	constexpr DcSmem() : off(), total(0)
	{
		constexpr DcGeom<P> g{};
		int o = 0;
		off[0] = 0;
		o += (2 * g.len[0] + 15) & ~15;
		for (int k = 1; k < P; k++) {
			off[k] = o;
			o += 8 * dc_f2_slots(g.len[k]);
		}
		off[P] = o;
		total = o;
	}
};
__device__ __forceinline__ int dc_slot(int i) { return i + 2 * (i >> 4); }

// One stage of the tile: n_out outputs (a multiple of eight) from level IN to level OUT in shared memory, or - FINAL -
// to global memory as int16 I,Q.  A thread takes eight consecutive outputs at a time.
template <int TAPS, bool WIDE, bool U8IN, bool FINAL>
__device__ __forceinline__ void dc_stage(const uint8_t *in, uint8_t *out, int n_out, uint32_t *gout, long long g0, long long g_n)
{
	constexpr int W = 2 * kDcR + TAPS - 2;   // inputs a thread needs
	for (int o0 = threadIdx.x * kDcR; o0 < n_out; o0 += kDcThreads * kDcR) {
		f2 x[W];
		if (U8IN) {
			// input 2*o0 + k, two bytes each; 2*o0 is a multiple of 16 samples = 32 bytes
			const uint4 *src = reinterpret_cast<const uint4 *>(in + 4 * o0);
#pragma unroll
			for (int q = 0; q < (W + 7) / 8; q++) {
				const uint4 v = src[q];
				const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
				for (int j = 0; j < 4; j++) {
					if (8 * q + 2 * j < W)
						x[8 * q + 2 * j] = fpack2(__uint_as_float(__byte_perm(w[j], 0, 0x4404) + kDcCvtBase),
									  __uint_as_float(__byte_perm(w[j], 0, 0x4414) + kDcCvtBase));
					if (8 * q + 2 * j + 1 < W)
						x[8 * q + 2 * j + 1] = fpack2(__uint_as_float(__byte_perm(w[j], 0, 0x4424) + kDcCvtBase),
									      __uint_as_float(__byte_perm(w[j], 0, 0x4434) + kDcCvtBase));
				}
			}
		} else {
			// input 2*o0 + k as float pairs; 2*o0 is a multiple of 16: slot 18*(o0/8) + k + 2*(k/16), 128-bit loads of two
			const f2 *src = reinterpret_cast<const f2 *>(in) + 18 * (o0 / kDcR);
#pragma unroll
			for (int k = 0; k < W; k += 2) {
				const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(src + k + 2 * (k >> 4));
				x[k] = v.x;
				if (k + 1 < W) x[k + 1] = v.y;
			}
		}
		uint32_t ow[kDcR];
		f2 of[kDcR];
#pragma unroll
		for (int k = 0; k < kDcR; k++) {
			f2 acc = fpack2(kDcAcc0, kDcAcc0);
#pragma unroll
			for (int n = 0; n < TAPS; n++) {
				const float c = (float)dc_tap(TAPS, WIDE, n) * (U8IN ? 1.0f / 1024.0f : 1.0f / 65536.0f);
				acc = ffma2_rm(x[2 * k + n], fpack2(c, c), acc);
			}
			constexpr int off = U8IN ? dc_u8_offset(TAPS, WIDE) : 0;
			if (FINAL) {
				uint32_t ai, aq;
				funpack2(acc, ai, aq);
				const uint32_t yi = ai - (kDcAcc0Bits + (uint32_t)off), yq = aq - (kDcAcc0Bits + (uint32_t)off);
				ow[k] = (yi & 0xffffu) | (yq << 16);
			} else {
				// the result as an exact float pair: the accumulator minus its (integer, < 2^24) starting offset
				constexpr float neg = -(kDcAcc0 + (float)off);
				of[k] = fadd2_rn(acc, fpack2(neg, neg));
			}
		}
		if (FINAL) {
			const long long g = g0 + o0;   // global index of the thread's first output
			if (g + kDcR <= g_n) {
				reinterpret_cast<uint4 *>(gout + g)[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
				reinterpret_cast<uint4 *>(gout + g)[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
			} else {
#pragma unroll
				for (int k = 0; k < kDcR; k++)
					if (g + k < g_n) gout[g + k] = ow[k];
			}
		} else {
			// outputs o0 .. o0+7: slots dc_slot(o0) .. +7 (inside one group of sixteen), four 128-bit stores
			f2 *dst = reinterpret_cast<f2 *>(out) + dc_slot(o0);
#pragma unroll
			for (int k = 0; k < kDcR; k += 2) {
				ulonglong2 v;
				v.x = of[k];
				v.y = of[k + 1];
				*reinterpret_cast<ulonglong2 *>(dst + k) = v;
			}
		}
	}
}

// iq: n_pairs raw IQ pairs (u8); hist: the kDcHist raw samples before iq[0] (null: zero signal); out: n_pairs >> P
// int16 I,Q pairs
template <int P, bool WIDE>
__global__ void __launch_bounds__(kDcThreads) dc_fused_kernel(const uint8_t *__restrict__ iq, const uint8_t *__restrict__ hist, long long n_pairs,
							      uint32_t *__restrict__ out, long long n_out)
{
	extern __shared__ __align__(16) uint8_t smem[];
	constexpr DcGeom<P> G{};
	constexpr DcSmem<P> S{};
	const long long tile = blockIdx.x;
	// raw sample index of smem level-0 element 0: the tile's first raw sample minus the halo the stages consume:
	// level k element 0 sits 18 (last stage) or 6 inputs before input 2*0 of the level above: accumulate down to raw
	int back = 0;
#pragma unroll
	for (int k = P; k >= 1; k--) back = 2 * back + (k == P ? 18 : 6);
	const long long r0 = tile * (long long)dc_tile(P) - back;

	// ---- stage the raw bytes: 16 bytes (8 samples) of shared memory per thread and turn; the tile's first byte is
	// 4-byte aligned in iq (the halo is an even number of samples) but not 16-byte aligned: four 32-bit loads
	{
		const int n_bytes = 2 * G.len[0];
		const long long b0 = 2 * r0;   // byte offset of smem byte 0 inside iq (negative: history)
		for (int i = threadIdx.x * 16; i < n_bytes; i += kDcThreads * 16) {
			const long long gb = b0 + i;
			uint4 v;
			if (gb >= 0 && gb + 16 <= 2 * n_pairs) {
				const uint32_t *src = reinterpret_cast<const uint32_t *>(iq + gb);
				v = make_uint4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
			} else {
				uint32_t w[4] = { 0u, 0u, 0u, 0u };
#pragma unroll
				for (int j = 0; j < 16; j++) {
					const long long g = gb + j;
					uint32_t b = 128u;   // zero signal before the stream and after the buffer
					if (g >= 0) {
						if (g < 2 * n_pairs) b = iq[g];
					} else if (hist && g >= -2 * kDcHist) {
						b = hist[2 * kDcHist + g];
					}
					w[j >> 2] |= b << (8 * (j & 3));
				}
				v = make_uint4(w[0], w[1], w[2], w[3]);
			}
			*reinterpret_cast<uint4 *>(smem + S.off[0] + i) = v;
		}
	}
	__syncthreads();
	// ---- the stages
	const long long g0 = tile * (long long)(dc_tile(P) >> P);
	if constexpr (P == 1) {
		dc_stage<20, WIDE, true, true>(smem + S.off[0], nullptr, G.len[1], out, g0, n_out);
	} else {
		dc_stage<8, false, true, false>(smem + S.off[0], smem + S.off[1], G.len[1], nullptr, 0, 0);
		__syncthreads();
#pragma unroll
		for (int k = 2; k < P; k++) {
			dc_stage<8, false, false, false>(smem + S.off[k - 1], smem + S.off[k], G.len[k], nullptr, 0, 0);
			__syncthreads();
		}
		dc_stage<20, WIDE, false, true>(smem + S.off[P - 1], nullptr, G.len[P], out, g0, n_out);
	}
}

template <int P>
static cudaError_t dc_launch_p(const uint8_t *iq, const uint8_t *hist, long long n_pairs, int wide, uint32_t *out, cudaStream_t s)
{
	constexpr DcSmem<P> S{};
	const long long n_out = n_pairs >> P;
	if (n_out <= 0) return cudaSuccess;
	const long long tiles = (n_out + (dc_tile(P) >> P) - 1) / (dc_tile(P) >> P);
	if (tiles > 0x7fffffffll) return cudaErrorInvalidValue;
	cudaError_t e;
	if (wide) {
		e = cudaFuncSetAttribute(dc_fused_kernel<P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S.total);
		if (e != cudaSuccess) return e;
		dc_fused_kernel<P, true><<<(unsigned)tiles, kDcThreads, S.total, s>>>(iq, hist, n_pairs, out, n_out);
	} else {
		e = cudaFuncSetAttribute(dc_fused_kernel<P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S.total);
		if (e != cudaSuccess) return e;
		dc_fused_kernel<P, false><<<(unsigned)tiles, kDcThreads, S.total, s>>>(iq, hist, n_pairs, out, n_out);
	}
	return cudaGetLastError();
}

// passes 1..5; hist may be null (zero history)
cudaError_t launch_downconvert_fused(const uint8_t *iq, const uint8_t *hist, long long n_pairs, int passes, int wide, int16_t *out, cudaStream_t s)
{
	uint32_t *o = reinterpret_cast<uint32_t *>(out);
	switch (passes) {
	case 1: return dc_launch_p<1>(iq, hist, n_pairs, wide, o, s);
	case 2: return dc_launch_p<2>(iq, hist, n_pairs, wide, o, s);
	case 3: return dc_launch_p<3>(iq, hist, n_pairs, wide, o, s);
	case 4: return dc_launch_p<4>(iq, hist, n_pairs, wide, o, s);
	case 5: return dc_launch_p<5>(iq, hist, n_pairs, wide, o, s);
	}
	return cudaErrorInvalidValue;
}

// the history for the next call: the last kDcHist raw samples of (old history, iq[0 .. n_pairs))
__global__ void dc_hist_kernel(const uint8_t *iq, long long n_pairs, const uint8_t *old_hist, uint8_t *new_hist)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;   // byte of the new history
	if (i >= 2 * kDcHist) return;
	const long long g = 2 * n_pairs - 2 * kDcHist + i;   // byte offset in iq (negative: still history)
	uint8_t b = 128;
	if (g >= 0) b = iq[g];
	else if (old_hist && g >= -2 * kDcHist) b = old_hist[2 * kDcHist + g];
	new_hist[i] = b;
}
cudaError_t launch_dc_hist(const uint8_t *iq, long long n_pairs, const uint8_t *old_hist, uint8_t *new_hist, cudaStream_t s)
{
	dc_hist_kernel<<<(2 * kDcHist + 255) / 256, 256, 0, s>>>(iq, n_pairs, old_hist, new_hist);
	return cudaGetLastError();
}

}  // namespace tfr
