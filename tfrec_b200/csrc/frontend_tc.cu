// frontend_tc.cu - the front-end with the byte->float conversion done by the tensor core.
//
// Same contract, same arithmetic and same output as frontend.cu (engine.cpp:77-78, dsp_stuff.cpp:172-264,
// fm_demod.cpp:45, bit for bit); what changes is how a raw byte becomes the float the exact-floor FMA chain eats.
// frontend.cu spends two ALU instructions per byte on it (PRMT + IADD) plus the shared-memory loads - a quarter of
// the instructions of a kernel that is bound by instruction issue.  Here the block's bytes are never touched by a
// thread: TMA tensor copies drop them into shared memory in the 128-byte-swizzled K-major operand layout (row r =
// the 512 bytes thread r owns, four boxes of 128 rows x 128 bytes), and one thread issues tcgen05.mma (kind::i8,
// u8 x u8 -> s32, M=128 N=32 K=32) against a 128 x identity matrix, on top of a constant:
//     D[r][j] = 114688 + 128 * byte[r][j]
// read as a float that word is the subnormal (896 + b) * 2^-142 - linear in b with no binade to leave - and
// 896 = 1024 - 128, so one round-down FMA with c = tap * 2^100 adds tap + floor((b-128)*tap/1024) to an accumulator
// kept in [2^23, 2^24) * 2^-32: the reference's floor((x*tap)>>16) for x = (b-128)<<6 plus an integer that is removed
// at the end.  The accumulators sit in tensor memory (two buffers of 64 columns, so the MMAs of the next 64 bytes run
// under the FMAs of the current ones); a thread pulls its sixteen next floats with one tcgen05.ld.
//
// A thread's first stage-1 outputs need the six raw samples before its row: a fifth, 32-byte wide TMA box fetches the
// 32 bytes in front of every row (tensor map shifted by -32 bytes: row r of it is the tail of row r-1) and goes through
// the same MMA.  Its first nine stage-2 outputs also need the left neighbour's last 18 stage-1 outputs: the thread
// runs its chains over zeros there, the neighbour sums the missing taps at the end of its own row, and the two exact
// integer partial sums meet through a shuffle (shared memory across warps; the block's first thread gets its share
// from the 96 history bytes).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "frontend_common.cuh"

namespace tfr {

// ------------------------------------------------------------------------------------------------
// exact-floor bookkeeping for this kernel
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kZMul = 128u;                  // TMEM word = kZBase + kZMul * byte
constexpr uint32_t kZBase = 896u * kZMul;         // 114,688 = 4 * 224 * 128 (the constant MMA)
constexpr float kC1Scale = 0x1p100f;              // z * c1 = (896+b) * t2/1024 * 2^-32
constexpr float kS1 = 0x1p-32f;                   // stage-1 accumulators live scaled by 2^-32 (c1 stays finite)
constexpr float kC2Scale = 0x1p16f;               // (kM1+y1)*2^-32 * t1*2^16 = (kM1+y1) * t1/65536
constexpr int kTcM1Mul = 130;
constexpr int kTcM1 = kTcM1Mul * 65536;           // 8,519,680
constexpr int kTcA1 = kTcM1 - t2_sum();           // 8,451,536: every tap adds 1*t2 on top of its floor
static_assert(kTcA1 - 8520 >= (1 << 23) && kTcM1 + 8520 < (1 << 24), "stage-1 accumulator leaves the integer binade");
constexpr int kSpill = 9;                         // outputs of a thread that also depend on its left neighbour

// ------------------------------------------------------------------------------------------------
// shared memory map (dynamic, 1024-byte aligned: the 128-byte swizzle repeats every 1024 bytes)
// ------------------------------------------------------------------------------------------------
constexpr int kBoxBytes = 128 * 128;              // one TMA box: 128 rows x 128 bytes
constexpr int kOffHalo = 4 * kBoxBytes;           // 128 rows x 32 bytes, 32-byte swizzle
constexpr int kOffBid = kOffHalo + 4096;          // 32 x 32 u8, 128 on the diagonal
constexpr int kOffBc = kOffBid + 1024;            // 32 x 32 u8, 128 in columns k < 4
constexpr int kOffAc = kOffBc + 1024;             // constant A operand, every byte 224 (two 128-byte core matrices)
constexpr int kOffHist = kOffAc + 256;            // 96 raw bytes before the block
constexpr int kOffBar = kOffHist + 128;           // 4 box barriers, halo, full[2], empty[2]
constexpr int kOffTmem = kOffBar + 9 * 8;
constexpr int kTcSmemBytes = kOffTmem + 8;        // 72,144 B -> 3 CTAs / SM
constexpr uint32_t kTmemCols = 128;               // two buffers of 64 columns
// u8 x u8 -> s32, M = 128, N = 32, both operands K-major (cute::UMMA::InstrDescriptor: c_format [4,6) = 2 (S32),
// a/b_format = 0 (U8), n_dim [17,23) = N>>3, m_dim [24,29) = M>>4)
constexpr uint32_t kIdesc = (2u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

// ------------------------------------------------------------------------------------------------
// PTX: tensor-map TMA, tcgen05
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_tensor_2d(uint32_t dst, const void *tmap, int c0, int c1, uint32_t bar)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
		     "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
		     : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start [0,14), leading byte offset [16,30), stride byte
// offset [32,46) - all in 16-byte units -, version 1 at [46,48), swizzle mode at [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout)
{
	return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
	       (1ull << 46) | ((uint64_t)layout << 61);
}
constexpr uint32_t kLayoutNone = 0, kLayoutSw128 = 2, kLayoutSw32 = 6;
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t accumulate)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"setp.ne.b32 p, %4, 0;\n"
		"tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
		"}\n" ::"r"(d_tmem),
		"l"(da), "l"(db), "r"(kIdesc), "r"(accumulate)
		: "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// sixteen consecutive columns of the thread's own TMEM lane (32x32b: lane = the warp's base lane + lane id)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
		     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
		       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
		     : "r"(taddr)
		     : "memory");
}
// the registers of an outstanding tcgen05.ld may only be read after this; naming them as operands keeps the
// compiler from moving their first use in front of it
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&v)[16])
{
	asm volatile("tcgen05.wait::ld.sync.aligned;"
		     : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
		       "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
		     :
		     : "memory");
}

// ------------------------------------------------------------------------------------------------
// FIR pieces
// ------------------------------------------------------------------------------------------------
template <bool WIDE>
__device__ __forceinline__ f2 tc_c2pair(int n)
{
	const float c = (float)t1_tap(WIDE, n) * kC2Scale;
	return pack2(c, c);
}
__device__ __forceinline__ f2 tc_c1pair(int n)
{
	const float c = (float)t2_tap(n) * kC1Scale;
	return pack2(c, c);
}
__device__ __forceinline__ f2 tc_stage1(const f2 *x)
{
	f2 acc = pack2((float)kTcA1 * kS1, (float)kTcA1 * kS1);
#pragma unroll
	for (int n = 0; n < 8; n++) acc = fma2_rm(x[n], tc_c1pair(n), acc);
	return acc;
}
// taps [LO, HI) of the output whose tap 0 sits in ring slot R0 (mod 32): adds the accumulator bits to si / sq
template <bool WIDE, int LO, int HI, int R0>
__device__ __forceinline__ void tc_chain(const f2 (&ring)[32], uint32_t &si, uint32_t &sq)
{
	if constexpr (LO < HI) {
		constexpr float a0 = (float)ChainC<kTcM1Mul, WIDE, LO, HI>::k.start;
		f2 acc = pack2(a0, a0);
#pragma unroll
		for (int n = LO; n < HI; n++) acc = fma2_rm(ring[(R0 + n + 64) & 31], tc_c2pair<WIDE>(n), acc);
		uint32_t ai, aq;
		unpack2(acc, ai, aq);
		si += ai;
		sq += aq;
	}
}
// exact sum of the per-tap floors of taps [N0, N1) (+ ADD on both channels), as two chains split at tap 10 (one chain's
// injected offsets alone would not fit the binade)
template <bool WIDE, int N0, int N1, int R0, int ADD>
__device__ __forceinline__ void tc_taps(const f2 (&ring)[32], int &yi, int &yq)
{
	constexpr int A0 = N0 < 10 ? N0 : 10, A1 = N1 < 10 ? N1 : 10;
	constexpr int B0 = N0 > 10 ? N0 : 10, B1 = N1 > 10 ? N1 : 10;
	uint32_t si = 0, sq = 0;
	tc_chain<WIDE, A0, A1, R0>(ring, si, sq);
	tc_chain<WIDE, B0, B1, R0>(ring, si, sq);
	constexpr uint32_t bias = ChainC<kTcM1Mul, WIDE, A0, A1>::bias + ChainC<kTcM1Mul, WIDE, B0, B1>::bias - (uint32_t)ADD;
	yi = (int)(si - bias);
	yq = (int)(sq - bias);
}
// what a thread's last 18 stage-1 outputs add to its right neighbour's first nine outputs: taps [0, 18-2M) of the
// neighbour's output M, biased by 16384 per channel and packed.  The ring holds the thread's last 32 stage-1 outputs,
// slot = index mod 32; the neighbour's output M starts at this thread's stage-1 output 128+2M-18, slot 14+2M.
template <bool WIDE, int M>
__device__ __forceinline__ uint32_t tc_spill_out(const f2 (&ring)[32])
{
	int yi, yq;
	tc_taps<WIDE, 0, 18 - 2 * M, 14 + 2 * M, 16384>(ring, yi, yq);
	return pack_iq(yi, yq);
}
// a raw byte from shared memory (history path of the block's first thread) in the tensor core's format
__device__ __forceinline__ f2 tc_cvt_iq(uint32_t w, int half)
{
	const uint32_t bi = ((w >> (16 * half)) & 0xffu) * kZMul + kZBase;
	const uint32_t bq = ((w >> (16 * half + 8)) & 0xffu) * kZMul + kZBase;
	return pack2(__uint_as_float(bi), __uint_as_float(bq));
}

// where output j (0..63) of thread r lives in shared memory.  The outputs of a 64-byte phase go where the bytes are
// dead by then: phase 0 into the halo tile, an odd phase into the first half of the box row the phase itself read (its
// MMA is complete before the thread sees the data), an even phase into the second half of the previous box's row;
// 16-byte chunks of a box row are XOR-swizzled by r so that a warp's stores spread over the banks.
__device__ __forceinline__ uint32_t tc_out_off(int r, int j)
{
	const int p = j >> 3, k = j & 7;
	if (p == 0) return (uint32_t)(kOffHalo + r * 32 + k * 4);
	const int box = (p - 1) >> 1, wi = ((p & 1) ? 0 : 8) + k;
	return (uint32_t)(box * kBoxBytes + r * 128 + ((((wi >> 2) ^ (r & 7))) << 4) + (wi & 3) * 4);
}

// One step: sixteen floats (eight raw samples) from tensor memory -> four stage-1 outputs into ring slots 4S..4S+3 ->
// the two stage-2 outputs 2S, 2S+1 of the current group of sixteen.
template <bool WIDE, int S>
__device__ __forceinline__ void tc_step(const uint32_t (&v)[16], f2 (&ring)[32], f2 (&xh)[6], uint8_t *orow, int xm, int wofs, uint32_t &t16,
					int thresh_lo)
{
	f2 x[14];
#pragma unroll
	for (int k = 0; k < 6; k++) x[k] = xh[k];
#pragma unroll
	for (int k = 0; k < 8; k++) x[6 + k] = pack2(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1]));
#pragma unroll
	for (int jj = 0; jj < 4; jj++) ring[(4 * S + jj) & 31] = tc_stage1(&x[2 * jj]);
#pragma unroll
	for (int k = 0; k < 6; k++) xh[k] = x[8 + k];
	uint32_t o[2];
#pragma unroll
	for (int mm = 0; mm < 2; mm++) {
		const int M = 2 * S + mm;
		int yi, yq;
		if (mm == 0) tc_taps<WIDE, 0, 20, 4 * S - 18, 0>(ring, yi, yq);
		else tc_taps<WIDE, 0, 20, 4 * S - 16, 0>(ring, yi, yq);
		if (abs(yi) + abs(yq) > thresh_lo) t16 |= 1u << M;
		o[mm] = pack_iq(yi, yq);
	}
	// outputs 2(S&3), 2(S&3)+1 of the phase: one 8-byte store
	const int wi = wofs + 2 * (S & 3);
	*reinterpret_cast<uint2 *>(orow + ((((wi >> 2) ^ xm)) << 4) + (wi & 3) * 4) = make_uint2(o[0], o[1]);
}

template <bool WIDE>
__global__ void __launch_bounds__(kThreads, 3) frontend_tc_kernel(const FrontParams p)
{
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ EpiShared es;
	__shared__ uint32_t s_spill[kWarps][kSpill + 1];   // [w]: what the thread left of warp w's lane 0 spills into it
	__shared__ f2 s_y1h[18];                           // stage-1 outputs -18..-1 of the block, from the 96-byte history

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int stream = blockIdx.y;
	const StreamJob job = p.jobs[stream];
	StreamState *st = p.st + stream;
	const int tile = (p.use_progress ? (int)st->t2_done : p.tile0) + blockIdx.x;   // block index inside the submit
	if (tile >= (int)job.n_blocks) return;

	const uint32_t sbase = smem_u32(smem);
	const uint32_t bar_box = sbase + kOffBar, bar_halo = bar_box + 32, bar_full = bar_box + 40, bar_empty = bar_box + 56;
	if (tid == 0) {
		for (int q = 0; q < 7; q++) mbar_init(bar_box + 8 * q, 1);        // 4 boxes, halo, full[0], full[1]
		for (int u = 0; u < 2; u++) mbar_init(bar_empty + 8 * u, kWarps);
	}
	// the constant operands: 16-byte unit u of an N x K = 32 x 32 K-major tile without swizzle holds row n = u & 31,
	// bytes k = 16 (u >> 5) ... + 15
	if (tid < 64) {
		const int n = tid & 31, k0 = 16 * (tid >> 5);
		const int d = n - k0;   // position of the diagonal element inside this unit, if 0 <= d < 16
		const uint32_t one = (d >= 0 && d < 16) ? 0x80u << (8 * (d & 3)) : 0u;
		reinterpret_cast<uint4 *>(smem + kOffBid)[tid] = make_uint4((d >> 2) == 0 ? one : 0u, (d >> 2) == 1 ? one : 0u, (d >> 2) == 2 ? one : 0u, (d >> 2) == 3 ? one : 0u);
		reinterpret_cast<uint4 *>(smem + kOffBc)[tid] = make_uint4(k0 == 0 ? 0x80808080u : 0u, 0u, 0u, 0u);
	} else if (tid < 80) {
		reinterpret_cast<uint4 *>(smem + kOffAc)[tid - 64] = make_uint4(0xe0e0e0e0u, 0xe0e0e0e0u, 0xe0e0e0e0u, 0xe0e0e0e0u);
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kOffTmem), "n"(kTmemCols) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = *reinterpret_cast<const volatile uint32_t *>(smem + kOffTmem);
	const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);

	const uint64_t d_bid = umma_desc(sbase + kOffBid, 512, 128, kLayoutNone);
	const uint64_t d_bc = umma_desc(sbase + kOffBc, 512, 128, kLayoutNone);
	const uint64_t d_ac = umma_desc(sbase + kOffAc, 128, 0, kLayoutNone);
	// the MMAs of 64-byte phase ph (two 32-column halves) into buffer ph & 1
	auto issue_phase = [&](int ph) {
		const uint32_t col = tmem + (uint32_t)((ph & 1) * 64);
#pragma unroll
		for (int j = 0; j < 2; j++) {
			const uint64_t da = umma_desc(sbase + (ph >> 1) * kBoxBytes + ((ph & 1) * 2 + j) * 32, 16, 1024, kLayoutSw128);
			umma_i8(col + 32 * j, d_ac, d_bc, 0u);
			umma_i8(col + 32 * j, da, d_bid, 1u);
		}
		umma_commit(bar_full + 8 * (ph & 1));
	};

	if (tid == 0) {
		const uint8_t *tm = reinterpret_cast<const uint8_t *>(p.tmaps) + (size_t)stream * 256;
		asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm) : "memory");
		asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm + 128) : "memory");
		const uint8_t *src = job.iq + (size_t)tile * kBlockBytes;
		mbar_expect_tx(bar_halo, 4096 + kHistBytes);
		tma_tensor_2d(sbase + kOffHalo, tm + 128, 0, tile * 128 - 1, bar_halo);
		const uint8_t *hsrc = (tile == 0) ? st->hist[st->hist_parity & 1] : src - kHistBytes;
		tma_bulk_g2s(sbase + kOffHist, hsrc, kHistBytes, bar_halo);
#pragma unroll
		for (int q = 0; q < 4; q++) {
			mbar_expect_tx(bar_box + 8 * q, kBoxBytes);
			tma_tensor_2d(sbase + q * kBoxBytes, tm, 128 * q, tile * 128, bar_box + 8 * q);
		}
		// the 32 bytes in front of every row -> buffer 1, columns 64..95; then phase 0 -> buffer 0
		mbar_wait(bar_halo, 0);
		tc_fence_after();
		umma_i8(tmem + 64, d_ac, d_bc, 0u);
		umma_i8(tmem + 64, umma_desc(sbase + kOffHalo, 16, 256, kLayoutSw32), d_bid, 1u);
		umma_commit(bar_full + 8);
		mbar_wait(bar_box, 0);
		tc_fence_after();
		issue_phase(0);
	}
	__syncwarp();

	int thresh_lo = st->thresh;
	if (st->thresh_mode) thresh_lo = p.margin ? thresh_lo - p.margin : st->spec_lo;

	// ------------------------------------------------------------------ the six raw samples before the row
	f2 xh[6];
	{
		uint32_t v[16];
		mbar_wait(bar_full + 8, 0);
		tc_fence_after();
		tmem_ld16(tlane + 64 + 16, v);   // bytes 16..31 of the halo row = raw samples -8..-1
		tmem_wait_ld(v);
#pragma unroll
		for (int k = 0; k < 6; k++) xh[k] = pack2(__uint_as_float(v[4 + 2 * k]), __uint_as_float(v[5 + 2 * k]));
		tc_fence_before();
		__syncwarp();
		if (lane == 0) mbar_arrive(bar_empty + 8);
		if (tid == 0) {
			if (tile == 0) {   // nothing in front of the submit: the carried history (also landed on bar_halo, seen above)
				const uint32_t *hw = reinterpret_cast<const uint32_t *>(smem + kOffHist + 84);
#pragma unroll
				for (int k = 0; k < 3; k++) {
					xh[2 * k] = tc_cvt_iq(hw[k], 0);
					xh[2 * k + 1] = tc_cvt_iq(hw[k], 1);
				}
			}
			mbar_wait(bar_empty + 8, 0);
			tc_fence_after();
			issue_phase(1);
		}
		__syncwarp();
	}

	// ------------------------------------------------------------------ per-thread FIR cascade
	f2 ring[32];   // stage-1 outputs (kM1 + y1) * 2^-32, slot = index mod 32
	{
		const f2 zero = pack2((float)kTcM1 * kS1, (float)kTcM1 * kS1);   // left neighbour's outputs: summed by the neighbour
#pragma unroll
		for (int k = 0; k < 32; k++) ring[k] = zero;
	}
	unsigned long long trig64 = 0ull;
#pragma unroll 1
	for (int q = 0; q < 4; q++) {
		uint32_t t16 = 0u;
		uint32_t va[16], vb[16];   // the next step's columns are in flight while the current step computes
		// ---- phase 2q: buffer 0 (its use number q); outputs into the halo tile (q = 0) or the previous box
		{
			uint8_t *orow = (q == 0) ? smem + kOffHalo + tid * 32 : smem + (q - 1) * kBoxBytes + tid * 128;
			const int xm = (q == 0) ? 0 : (tid & 7), wofs = (q == 0) ? 0 : 8;
			mbar_wait(bar_full, (uint32_t)(q & 1));
			tc_fence_after();
			tmem_ld16(tlane + 0, va);
			tmem_wait_ld(va);
			tmem_ld16(tlane + 16, vb);
			tc_step<WIDE, 0>(va, ring, xh, orow, xm, wofs, t16, thresh_lo);
			if (tid == 0 && q > 0) {   // phase 2q+1 into buffer 1, free once every warp has pulled phase 2q-1
				mbar_wait(bar_empty + 8, (uint32_t)(q & 1));
				tc_fence_after();
				issue_phase(2 * q + 1);
			}
			__syncwarp();
			tmem_wait_ld(vb);
			tmem_ld16(tlane + 32, va);
			tc_step<WIDE, 1>(vb, ring, xh, orow, xm, wofs, t16, thresh_lo);
			tmem_wait_ld(va);
			tmem_ld16(tlane + 48, vb);
			tc_step<WIDE, 2>(va, ring, xh, orow, xm, wofs, t16, thresh_lo);
			tmem_wait_ld(vb);
			tc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_empty);
			tc_step<WIDE, 3>(vb, ring, xh, orow, xm, wofs, t16, thresh_lo);
		}
		// ---- phase 2q+1: buffer 1 (its use number q+1); outputs into the first half of box q's rows
		{
			uint8_t *orow = smem + q * kBoxBytes + tid * 128;
			const int xm = tid & 7, wofs = 0;
			mbar_wait(bar_full + 8, (uint32_t)((q + 1) & 1));
			tc_fence_after();
			tmem_ld16(tlane + 64, va);
			tmem_wait_ld(va);
			tmem_ld16(tlane + 80, vb);
			tc_step<WIDE, 4>(va, ring, xh, orow, xm, wofs, t16, thresh_lo);
			if (tid == 0 && q < 3) {   // phase 2q+2 into buffer 0, free once every warp has pulled phase 2q
				mbar_wait(bar_box + 8 * (q + 1), 0);
				mbar_wait(bar_empty, (uint32_t)(q & 1));
				tc_fence_after();
				issue_phase(2 * q + 2);
			}
			__syncwarp();
			tmem_wait_ld(vb);
			tmem_ld16(tlane + 96, va);
			tc_step<WIDE, 5>(vb, ring, xh, orow, xm, wofs, t16, thresh_lo);
			tmem_wait_ld(va);
			tmem_ld16(tlane + 112, vb);
			tc_step<WIDE, 6>(va, ring, xh, orow, xm, wofs, t16, thresh_lo);
			tmem_wait_ld(vb);
			tc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_empty + 8);
			tc_step<WIDE, 7>(vb, ring, xh, orow, xm, wofs, t16, thresh_lo);
		}
		trig64 |= (unsigned long long)t16 << (16 * q);
	}

	// ------------------------------------------------------------------ the left neighbour's share of outputs 0..8
	uint32_t sp[kSpill];
	sp[0] = tc_spill_out<WIDE, 0>(ring); sp[1] = tc_spill_out<WIDE, 1>(ring); sp[2] = tc_spill_out<WIDE, 2>(ring);
	sp[3] = tc_spill_out<WIDE, 3>(ring); sp[4] = tc_spill_out<WIDE, 4>(ring); sp[5] = tc_spill_out<WIDE, 5>(ring);
	sp[6] = tc_spill_out<WIDE, 6>(ring); sp[7] = tc_spill_out<WIDE, 7>(ring); sp[8] = tc_spill_out<WIDE, 8>(ring);
	if (lane == 31 && warp + 1 < kWarps) {
#pragma unroll
		for (int m = 0; m < kSpill; m++) s_spill[warp + 1][m] = sp[m];
	}
	if (warp == 0) {
		// the block's first thread has no neighbour in this CTA: its share comes from the 96 history bytes (raw samples
		// -48..-1), once per block, the 18 stage-1 outputs and the nine sums spread over lanes
		mbar_wait(bar_halo, 0);
		if (lane < 18) {
			// stage-1 output -18+lane needs raw samples -42+2*lane .. -35+2*lane
			const uint32_t *hw = reinterpret_cast<const uint32_t *>(smem + kOffHist + 12 + 4 * lane);
			f2 x[8];
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const uint32_t w = hw[k];
				x[2 * k] = tc_cvt_iq(w, 0);
				x[2 * k + 1] = tc_cvt_iq(w, 1);
			}
			s_y1h[lane] = tc_stage1(x);
		}
		__syncwarp();
		if (lane < kSpill) {
			// output `lane` of thread 0: taps 0 .. 17-2*lane on stage-1 outputs -18+2*lane ..; the remaining taps up to 17
			// run on a zero, so that every lane carries the same constant offsets
			const f2 zero = pack2((float)kTcM1 * kS1, (float)kTcM1 * kS1);
			constexpr float a0 = (float)ChainC<kTcM1Mul, WIDE, 0, 10>::k.start, b0 = (float)ChainC<kTcM1Mul, WIDE, 10, 18>::k.start;
			f2 a = pack2(a0, a0), b = pack2(b0, b0);
#pragma unroll
			for (int n = 0; n < 18; n++) {
				const f2 y = (n <= 17 - 2 * lane) ? s_y1h[2 * lane + n] : zero;
				if (n < 10) a = fma2_rm(y, tc_c2pair<WIDE>(n), a);
				else b = fma2_rm(y, tc_c2pair<WIDE>(n), b);
			}
			uint32_t ai, aq, bi, bq;
			unpack2(a, ai, aq);
			unpack2(b, bi, bq);
			constexpr uint32_t bias = ChainC<kTcM1Mul, WIDE, 0, 10>::bias + ChainC<kTcM1Mul, WIDE, 10, 18>::bias - 16384u;
			s_spill[0][lane] = pack_iq((int)(ai + bi - bias), (int)(aq + bq - bias));
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 0) {   // every thread has pulled its last tensor-memory columns
		tc_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
	}
	trig64 &= ~0x1ffull;
#pragma unroll
	for (int m = 0; m < kSpill; m++) {
		uint32_t in = __shfl_up_sync(0xffffffffu, sp[m], 1);
		if (lane == 0) in = s_spill[warp][m];
		// outputs 0..7 are phase 0 (halo tile), output 8 is word 0 of box 0's row
		uint32_t *slot = reinterpret_cast<uint32_t *>(m < 8 ? smem + kOffHalo + tid * 32 + m * 4 : smem + tid * 128 + ((tid & 7) << 4));
		const uint32_t own = *slot;
		const int yi = (int)(int16_t)(own & 0xffff) + (int)(in & 0xffff) - 16384;
		const int yq = ((int)own >> 16) + (int)(in >> 16) - 16384;
		*slot = pack_iq(yi, yq);
		if (abs(yi) + abs(yq) > thresh_lo) trig64 |= 1ull << m;
	}

	block_epilogue(p, job, tile, [&](int m) -> uint32_t {
		return *reinterpret_cast<const uint32_t *>(smem + tc_out_off(m >> 6, m & 63));
	}, es, trig64);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
				  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
				  CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn()
{
	static EncodeTiledFn fn = nullptr;
	static bool tried = false;
	if (!tried) {
		tried = true;
		void *ptr = nullptr;
		cudaDriverEntryPointQueryResult qres;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
			fn = reinterpret_cast<EncodeTiledFn>(ptr);
		else
			cudaGetLastError();
	}
	return fn;
}

bool frontend_tc_available() { return encode_fn() != nullptr; }

// the two tensor maps of one stream's submit: [0] rows of 512 bytes, boxes of 128 rows x 128 bytes, 128-byte swizzle;
// [1] the 32 bytes in front of every row (the same memory seen from 32 bytes before the end of row 0, row r of this map
// is the tail of row r-1), boxes of 128 rows x 32 bytes, 32-byte swizzle.  out = 256 bytes, 64-byte aligned.
int frontend_tc_encode(const void *iq, uint32_t n_blocks, void *out)
{
	EncodeTiledFn fn = encode_fn();
	if (!fn || !iq || !n_blocks) return -1;
	CUtensorMap *tm = reinterpret_cast<CUtensorMap *>(out);
	const cuuint64_t rows = (cuuint64_t)n_blocks * 128;
	const cuuint64_t stride[1] = { 512 };
	const cuuint32_t estr[2] = { 1, 1 };
	{
		const cuuint64_t dim[2] = { 512, rows };
		const cuuint32_t box[2] = { 128, 128 };
		if (fn(&tm[0], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(iq), dim, stride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
		       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
			return -2;
	}
	{
		const cuuint64_t dim[2] = { 32, rows };
		const cuuint32_t box[2] = { 32, 128 };
		void *base = const_cast<uint8_t *>(reinterpret_cast<const uint8_t *>(iq) + 480);
		if (fn(&tm[1], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dim, stride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
		       CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
			return -3;
	}
	return 0;
}

cudaError_t launch_frontend_tc(const FrontParams &p, int n_streams, int wide, cudaStream_t stream)
{
	static bool attr_done[64] = { false };
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	if (dev < 64 && !attr_done[dev]) {
		e = cudaFuncSetAttribute(frontend_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
		if (e != cudaSuccess) return e;
		e = cudaFuncSetAttribute(frontend_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
		if (e != cudaSuccess) return e;
		if (getenv("TFR_DEBUG")) {
			int nb = 0;
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, frontend_tc_kernel<false>, kThreads, kTcSmemBytes);
			fprintf(stderr, "[tfr] frontend_tc: %d CTAs per SM, %d B dynamic shared memory\n", nb, kTcSmemBytes);
		}
		attr_done[dev] = true;
	}
	if (p.n_tiles <= 0 || n_streams <= 0) return cudaSuccess;
	dim3 grid(p.n_tiles, n_streams);
	if (wide)
		frontend_tc_kernel<true><<<grid, kThreads, kTcSmemBytes, stream>>>(p);
	else
		frontend_tc_kernel<false><<<grid, kThreads, kTcSmemBytes, stream>>>(p);
	return cudaGetLastError();
}

}  // namespace tfr
