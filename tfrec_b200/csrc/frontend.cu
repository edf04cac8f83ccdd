// frontend.cu - raw u8 IQ -> exact two-stage FIR decimate /4 -> power trigger -> sparse window store.
//
// Replaces, bit-exactly, the per-block work of engine.cpp:77-78 ((u8-128)<<6), downconvert::process_iq
// (dsp_stuff.cpp:243-264: process2x1 8 taps, then process2x 20 taps, each tap floored separately with
// (x*tap)>>16, dsp_stuff.cpp:194-195,222-223) and the pwr=|I|+|Q| trigger test of fm_demod.cpp:45.
//
// One CTA = one reference block (65536 B -> 8192 decimated samples).  128 threads, each owns 512
// consecutive raw bytes (64 outputs).  The block's bytes arrive by TMA bulk copies (one 512-byte row
// per thread, rows padded to 528 B so that the 128-bit shared loads of a warp are conflict free)
// signalled on one mbarrier.  Arithmetic: every tap product is ONE round-toward-minus-infinity FMA on
// an accumulator kept inside [2^23, 2^24), where the fp32 grid is exactly the integers, so that
//     fma.rm(x, tap/2^16, acc) == acc + floor(x*tap / 2^16)           (exact, no separate floor)
// and I/Q share taps, so both channels ride in one fma.rm.f32x2 (SASS FFMA2.RM).  The offsets that
// keep the accumulators inside the binade are integer multiples of the tap denominators and are
// removed once, in integer arithmetic, when the accumulator bits are read back (see constants below).
//
// Output: decimated int16 I,Q are written ONLY where a demodulator can be active: [trigger,
// trigger+t_max) for every sample with pwr > thresh_lo, plus the first t_max samples of every block
// (they may be covered by a trigger in the previous block) and the block's last sample (lead-in for
// the next block).  A TileDesc per block lists the segments.  The sparse buffer is direct mapped
// (sample m of block b lives at dec[b*8192+m]) so the back-end addresses by position.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "fir_exact.cuh"


namespace tfr {

// ------------------------------------------------------------------------------------------------
// shared memory layout
// ------------------------------------------------------------------------------------------------
constexpr int kRowBytes = 512;               // raw bytes per thread
constexpr int kRowStride = 528;              // +16: 33 x 16 B, odd -> conflict-free LDS.128 across a quarter warp
constexpr int kRow0 = 128;                   // byte offset of row 0; the 96 halo bytes sit at [16,112)
constexpr int kHaloOff = 16;
constexpr int kSmemBytes = kRow0 + kThreads * kRowStride;  // 67,712 B -> 3 CTAs / SM

// one block of one stream: everything between the TMA loads and the descriptor.  `phase` is the parity of the mbarrier
// at smem[0] for this use (a CTA that takes several blocks in turn flips it)
template <bool WIDE>
__device__ __forceinline__ void fe_block(const FrontParams &p, const StreamJob &job, StreamState *st, int tile, uint8_t *smem, EpiShared &es,
					 uint32_t phase)
{
	const int tid = threadIdx.x;
	const uint32_t bar = smem_u32(smem);
	const uint8_t *src = job.iq + (size_t)tile * kBlockBytes;
	if (tid == 0) {
		mbar_expect_tx(bar, kBlockBytes + kHistBytes);
		const uint8_t *hsrc = (tile == 0) ? st->hist[st->hist_parity & 1] : src - kHistBytes;
		tma_bulk_g2s(smem_u32(smem + kHaloOff), hsrc, kHistBytes, bar);
	}
	tma_bulk_g2s(smem_u32(smem + kRow0 + tid * kRowStride), src + tid * kRowBytes, kRowBytes, bar);

	int thresh_lo = st->thresh;
	if (st->thresh_mode) thresh_lo = p.margin ? thresh_lo - p.margin : st->spec_lo;

	mbar_wait(bar, phase);

	// ------------------------------------------------------------------ per-thread FIR cascade
	uint8_t *row = smem + kRow0 + tid * kRowStride;
	const uint4 *halo = reinterpret_cast<const uint4 *>(row - 16 - kHistBytes);  // 48 raw samples before the row

	f2 ring[32];   // stage-1 outputs (kM1 + y1), slot = index mod 32
	f2 xh[6];      // last 6 converted raw samples

	{
		// prologue: y1[-18..-1] from raw x[-42..-1]; halo holds x[-48..-1]
		f2 hx[48];
#pragma unroll
		for (int q = 0; q < 6; q++) {
			uint4 v = halo[q];
			hx[8 * q + 0] = cvt_iq(v.x, 0); hx[8 * q + 1] = cvt_iq(v.x, 1);
			hx[8 * q + 2] = cvt_iq(v.y, 0); hx[8 * q + 3] = cvt_iq(v.y, 1);
			hx[8 * q + 4] = cvt_iq(v.z, 0); hx[8 * q + 5] = cvt_iq(v.z, 1);
			hx[8 * q + 6] = cvt_iq(v.w, 0); hx[8 * q + 7] = cvt_iq(v.w, 1);
		}
		// y1[j] (j=-18..-1) needs x[2j-6 .. 2j+1]; hx[k] = x[k-48]  ->  hx[2j+42 .. 2j+49]
#pragma unroll
		for (int j = -18; j < 0; j++) ring[(j + 32) & 31] = stage1(&hx[2 * j + 42]);
#pragma unroll
		for (int k = 0; k < 6; k++) xh[k] = hx[42 + k];
	}

	unsigned long long trig64 = 0ull;   // bit m = pwr(output m) > thresh_lo
	const uint4 *rowv = reinterpret_cast<const uint4 *>(row);
	uint32_t *outw = reinterpret_cast<uint32_t *>(row);   // outputs overwrite the front of the own row

#pragma unroll 1
	for (int it = 0; it < 4; it++) {
		uint32_t t16 = 0u;   // trigger bits of this iteration's 16 outputs (static bit positions)
#pragma unroll
		for (int s = 0; s < 8; s++) {
			const uint4 v = rowv[it * 8 + s];
			f2 x[14];
#pragma unroll
			for (int k = 0; k < 6; k++) x[k] = xh[k];
			x[6] = cvt_iq(v.x, 0); x[7] = cvt_iq(v.x, 1);
			x[8] = cvt_iq(v.y, 0); x[9] = cvt_iq(v.y, 1);
			x[10] = cvt_iq(v.z, 0); x[11] = cvt_iq(v.z, 1);
			x[12] = cvt_iq(v.w, 0); x[13] = cvt_iq(v.w, 1);
			// four stage-1 outputs j = 4s..4s+3 (relative to the iteration), y1[j] needs x[2(j-4s) .. +7] of this window
#pragma unroll
			for (int jj = 0; jj < 4; jj++) ring[(4 * s + jj) & 31] = stage1(&x[2 * jj]);
#pragma unroll
			for (int k = 0; k < 6; k++) xh[k] = x[8 + k];
			// two stage-2 outputs m = 2s, 2s+1: y1 indices 2m-18 .. 2m+1
			uint32_t o[2];
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
				const int yi = (int)(ai + bi - y2_bias(WIDE));
				const int yq = (int)(aq + bq - y2_bias(WIDE));
				const int pwr = abs(yi) + abs(yq);
				if (pwr > thresh_lo) t16 |= 1u << m;
				o[mm] = __byte_perm((uint32_t)yi, (uint32_t)yq, 0x5410);
			}
			*reinterpret_cast<uint2 *>(outw + it * 16 + 2 * s) = make_uint2(o[0], o[1]);
		}
		trig64 |= (unsigned long long)t16 << (16 * it);
	}
	block_epilogue(p, job, tile, [&](int m) -> uint32_t {
		return *reinterpret_cast<const uint32_t *>(smem + kRow0 + (m >> 6) * kRowStride + (m & 63) * 4);
	}, es, trig64);
}

template <bool WIDE>
__global__ void __launch_bounds__(kThreads, 3) frontend_kernel(const FrontParams p)
{
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ EpiShared es;
	const int stream = blockIdx.y;
	const StreamJob job = p.jobs[stream];
	StreamState *st = p.st + stream;
	const int tile = (p.use_progress ? (int)st->t2_done : p.tile0) + blockIdx.x;   // block index inside the submit
	if (tile >= (int)job.n_blocks) return;
	if (threadIdx.x == 0) mbar_init(smem_u32(smem), 1);
	__syncthreads();
	fe_block<WIDE>(p, job, st, tile, smem, es, 0u);
}

// The same block bookkeeping for input that is ALREADY decimated: int16 I,Q at 384 kS/s, what the reference hands to
// fsk_demod::process(int16_t *data_iq, int len) (fm_demod.cpp:34-74, len = 16384 per block).  No filter: a thread
// copies its 64 samples into the row layout above, tests them against the trigger bound and joins the epilogue.
__global__ void __launch_bounds__(kThreads, 3) frontend_i16_kernel(const FrontParams p)
{
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ EpiShared es;
	const int tid = threadIdx.x;
	const int stream = blockIdx.y;
	const StreamJob job = p.jobs[stream];
	StreamState *st = p.st + stream;
	const int tile = (p.use_progress ? (int)st->t2_done : p.tile0) + blockIdx.x;
	if (tile >= (int)job.n_blocks) return;
	int thresh_lo = st->thresh;
	if (st->thresh_mode) thresh_lo = p.margin ? thresh_lo - p.margin : st->spec_lo;
	const uint4 *src = reinterpret_cast<const uint4 *>(job.iq + (size_t)tile * (kBlockDec * 4)) + tid * 16;
	uint4 *row = reinterpret_cast<uint4 *>(smem + kRow0 + tid * kRowStride);
	unsigned long long trig64 = 0ull;
#pragma unroll 4
	for (int q = 0; q < 16; q++) {
		const uint4 v = src[q];
		row[q] = v;
		const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int yi = (int)(int16_t)(w[j] & 0xffff), yq = (int)(int16_t)(w[j] >> 16);
			if (abs(yi) + abs(yq) > thresh_lo) trig64 |= 1ull << (4 * q + j);
		}
	}
	block_epilogue(p, job, tile, [&](int m) -> uint32_t {
		return *reinterpret_cast<const uint32_t *>(smem + kRow0 + (m >> 6) * kRowStride + (m & 63) * 4);
	}, es, trig64);
}

// after the last front-end launch and threshold walk of a call: remember the FIR history for the next call
// (other parity) and rewind the walk cursor
__global__ void save_history_kernel(const StreamJob *jobs, StreamState *st, int n_streams, int decimated)
{
	const int stream = blockIdx.x;
	if (stream >= n_streams) return;
	const StreamJob job = jobs[stream];
	if (job.n_blocks == 0) return;
	StreamState *s = st + stream;
	const int par = (s->hist_parity & 1) ^ 1;
	const uint8_t *src = job.iq + (size_t)job.n_blocks * kBlockBytes - kHistBytes;
	// (a call that was fed decimated samples has no raw tail: the filter history restarts from zero signal)
	if (threadIdx.x < kHistBytes) s->hist[par][threadIdx.x] = decimated ? (uint8_t)128 : src[threadIdx.x];
	__syncthreads();
	if (threadIdx.x == 0) {
		s->hist_parity = par;
		s->t2_done = 0;   // the threshold walk of this call is complete; the next call starts at block 0
		s->spec_lo = s->thresh - spec_margin(s->thresh);   // the next call's speculative bound
	}
}

cudaError_t launch_frontend(const FrontParams &p, int n_streams, int wide, cudaStream_t stream)
{
	// the opt-in shared-memory size is a per-device function attribute: set it on every device we meet
	static bool attr_done[64] = { false };
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	if (dev < 64 && !attr_done[dev]) {
		e = cudaFuncSetAttribute(frontend_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
		if (e != cudaSuccess) return e;
		e = cudaFuncSetAttribute(frontend_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
		if (e != cudaSuccess) return e;
		attr_done[dev] = true;
	}
	if (p.n_tiles <= 0 || n_streams <= 0) return cudaSuccess;
	if (getenv("TFR_DEBUG"))
		fprintf(stderr, "[tfr] frontend grid=(%d,%d) tile0=%d t_max=%d keep_all=%d wide=%d smem=%d\n", p.n_tiles, n_streams,
			p.tile0, p.t_max, p.keep_all, wide, kSmemBytes);
	dim3 grid(p.n_tiles, n_streams);
	if (wide)
		frontend_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(p);
	else
		frontend_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(p);
	return cudaGetLastError();
}

cudaError_t launch_frontend_i16(const FrontParams &p, int n_streams, cudaStream_t stream)
{
	static bool attr_done[64] = { false };
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	if (dev < 64 && !attr_done[dev]) {
		e = cudaFuncSetAttribute(frontend_i16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
		if (e != cudaSuccess) return e;
		attr_done[dev] = true;
	}
	if (p.n_tiles <= 0 || n_streams <= 0) return cudaSuccess;
	frontend_i16_kernel<<<dim3(p.n_tiles, n_streams), kThreads, kSmemBytes, stream>>>(p);
	return cudaGetLastError();
}

cudaError_t launch_save_history(const StreamJob *jobs, StreamState *st, int n_streams, int decimated, cudaStream_t stream)
{
	save_history_kernel<<<n_streams, 128, 0, stream>>>(jobs, st, n_streams, decimated);
	return cudaGetLastError();
}

}  // namespace tfr
