// frontend_screen.cu - the screening front-end: the tensor core proves which decimated samples can NOT trigger, the
// exact FIR runs only where a demodulator will look.
//
// Same contract as frontend.cu (engine.cpp:77-78, dsp_stuff.cpp:172-264, fm_demod.cpp:45), bit for bit in everything
// a consumer reads.  What it exploits: downstream of the decimator the reference only ever (a) compares
// pwr = |I|+|Q| of every decimated sample with the threshold and (b) reads the samples inside demodulator windows
// (3-4 % of a noise stream).  The per-tap floors of dsp_stuff.cpp:194-195,222-223 make the exact value expensive
// (18 dependent-rounding products per raw sample, frontend.cu is bound by FP32 issue at 35 % of the HBM roofline),
// but they move the result by a BOUNDED amount away from the plain linear filter:
//     y  = sum_k floor(s1[.]*t1[k] / 2^16),   s1 = sum_n floor(x[.]*t2[n] / 2^16),   x = (b-128) << 6
//     y  = L / 2^26 - E,    L = sum_i (b_i - 128) * T[i],    T = t1 (*) upsampled t2   (46 taps, exact integers)
//     E in (-8*N1, 8*P1 + 20),  P1 / N1 = sum of the positive / negative t1 taps / 2^16  (each floor loses [0,1))
// L is a plain integer dot product of raw BYTES with constants - tensor-core work (tcgen05.mma kind::i8, u8 x s8 ->
// s32, exact).  The 29-bit taps are rounded to 16 bits (two s8 planes, T16 = 256*hi + lo; the rounding error is bounded
// by 128 * sum|T - 2^k*T16| and is part of the slack), so with c = L16 - centre and slack = the half width of E plus the
// rounding bound, per channel:       |y| <= |c| / 2^q + slack/2.
// A sample with |cI| + |cQ| <= (thresh_lo - slack) << q therefore has pwr <= thresh_lo and is provably no trigger.  The
// rest (a handful per block on noise) are CANDIDATES: their exact values are computed with the reference arithmetic
// (fir_exact.cuh, one warp per candidate, lane = stage-2 tap) and tested for real.  Blocks with many candidates
// (telegram bursts, start-up) run the whole exact cascade in place, like frontend.cu.  The decimated samples
// the demodulators read are produced after the threshold walk, when the windows are known, by decwin_kernel (below):
// exact FIR over [window start - 1, window end] only.
//
// The block as a GEMM: A = the block's bytes as 128 rows x 512 bytes (row r = the 64 outputs thread r owns in
// frontend.cu), TMA tensor boxes in the 128-byte-swizzled K-major layout.  Output j of a row needs raw samples
// 4j-42 .. 4j+3 of that row, so a 32-byte K slice s (16 samples) only reaches outputs 4s .. 4s+14: the Toeplitz band is
// the SAME 32 x 64 matrix for every slice (N = 64: 16 outputs x {I,Q} x {lo,hi} plane), accumulated into D columns
// 16s.. of tensor memory.  The last 84 bytes of the previous row come from a fifth box (the same tensor map one row
// up, column 384) through the same band matrix shifted by 16 / 32 / 48 columns.  The first MMA of a block multiplies a
// constant A operand with a constant B: it initialises all 256 columns and subtracts the centre.
//
// One CTA = four screening warps + one producer warp (work items, TMA, MMAs), persistent over blocks, 256 columns of
// tensor memory, 2 CTAs per SM.  The producer runs one block ahead: while the screeners evaluate block i's candidates,
// the MMAs of block i+1 run and the boxes of block i+2 are in flight.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fir_exact.cuh"

namespace tfr {

// ------------------------------------------------------------------------------------------------
// constant operands (built on the host by screen_build_consts, one copy per handle in device memory)
// ------------------------------------------------------------------------------------------------
constexpr int kScBand = 0;                        // N x K = 64 x 32 s8, K-major core matrices: byte(n, k) at 16 n + 1024 (k >> 4) + (k & 15)
constexpr int kScConst = 2048;                    // N x K = 256 x 32 s8: byte(n, k) at 16 n + 4096 (k >> 4) + (k & 15)
constexpr int kScAc = kScConst + 8192;            // constant A operand: 8 rows x 16 bytes of 128, then 8 rows x 16 bytes of 1
constexpr int kScBytes = kScAc + 256;             // 10,496
constexpr int kCandMax = 48;                      // more candidates than this in a block: hand it to the dense kernel

// ------------------------------------------------------------------------------------------------
// shared memory map (dynamic, 1024-byte aligned)
// ------------------------------------------------------------------------------------------------
constexpr int kBox = 128 * 128;                   // one TMA box: 128 rows x 128 bytes
constexpr int kOffHaloBox = 4 * kBox;             // box 4: the last 128 bytes of the row before every row
constexpr int kOffConsts = 5 * kBox;
constexpr int kOffBarS = kOffConsts + kScBytes;   // mbarriers: 5 TMA boxes, MMAs done, item published, tensor memory read out
constexpr int kOffTmemS = kOffBarS + 8 * 8;
constexpr int kScreenSmem = kOffTmemS + 16;       // 92,480 B -> 2 CTAs / SM
constexpr uint32_t kScreenCols = 256;

// u8 x s8 -> s32, M = 128, K-major both (cute::UMMA::InstrDescriptor: c_format [4,6) = 2 (S32), a_format [7,10) = 0 (U8),
// b_format [10,13) = 1 (S8), n_dim [17,23) = N >> 3, m_dim [24,29) = M >> 4)
__host__ __device__ constexpr uint32_t screen_idesc(uint32_t n) { return (2u << 4) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24); }

__device__ __forceinline__ void sc_tma_2d(uint32_t dst, const void *tmap, int c0, int c1, uint32_t bar)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
		     "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
		     : "memory");
}
__device__ __forceinline__ uint64_t sc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout)
{
	return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
	       ((uint64_t)layout << 61);
}
constexpr uint32_t kScNone = 0, kScSw128 = 2;
__device__ __forceinline__ void sc_mma(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"setp.ne.b32 p, %4, 0;\n"
		"tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
		"}\n" ::"r"(d_tmem),
		"l"(da), "l"(db), "r"(idesc), "r"(accumulate)
		: "memory");
}
__device__ __forceinline__ void sc_commit(uint32_t bar)
{
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void sc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void sc_ld16(uint32_t taddr, uint32_t (&v)[16])
{
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
		     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
		       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
		     : "r"(taddr)
		     : "memory");
}
// 32 consecutive columns (8 outputs) per instruction: a round trip to tensor memory costs ~150 cycles whatever it carries
__device__ __forceinline__ void sc_ld32(uint32_t taddr, uint32_t (&v)[32])
{
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
		     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
		     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
		       "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
		       "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
		       "=r"(v[31])
		     : "r"(taddr)
		     : "memory");
}
// the registers of outstanding tcgen05.ld may only be read after this (the asm volatile statements keep their order)
__device__ __forceinline__ void sc_wait_ld_all() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void sc_wait_ld(uint32_t (&v)[16])
{
	asm volatile("tcgen05.wait::ld.sync.aligned;"
		     : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
		       "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
		     :
		     : "memory");
}

// stage-2 taps by lane (lanes 20..31: 0)
__device__ const int kT1Narrow[32] = TFR_T1N;
__device__ const int kT1Wide[32] = TFR_T1W;
constexpr int kA2One = (1 << 23) + 4000000;   // one stage-2 tap per lane: 163 * t1 + floor part stays inside the integer binade
static_assert(kA2One - 163 * 3198 - 3000 >= (1 << 23) && kA2One + 163 * 17421 + 3000 < (1 << 24), "single-tap accumulator range");

#ifdef TFR_SCREEN_PROFILE
__device__ unsigned long long g_scprof[8];   // thread 0: cycles waiting for loads, in the MMAs, readback, candidate list, evaluation, epilogue; blocks
#define SCPROF(k) do { if (tid == 0) { const long long t_ = clock64(); prof[k] += t_ - pt; pt = t_; } } while (0)
#else
#define SCPROF(k) do { } while (0)
#endif

struct ScreenShared {
	uint16_t cand[kCandMax + 2];
	uint32_t res_w[kCandMax + 2];
	int res_p[kCandMax + 2];
	int warp_cnt[kWarps];
	int item[2];   // the work item of round r (slot r & 1), published by the producer; n_items ends the CTA
};

constexpr int kScreenThreads = kThreads + 32;   // four screening warps (one per tensor-memory lane quadrant) + the producer warp
// the screening warps' own barrier (the producer warp never joins)
__device__ __forceinline__ void sc_sync_screeners() { asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory"); }
struct ScreenersSync {
	__device__ __forceinline__ void operator()() const { sc_sync_screeners(); }
};

template <bool WIDE>
__global__ void __launch_bounds__(kScreenThreads, 2) frontend_screen_kernel(const FrontParams p)
{
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ ScreenShared ss;
	__shared__ EpiShared es;   // burst blocks done in place

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t sbase = smem_u32(smem);
	// mbarriers: 5 TMA boxes, MMAs done, item published, tensor memory read out
	const uint32_t bar_box = sbase + kOffBarS, bar_mma = bar_box + 40, bar_item = bar_box + 48, bar_tfree = bar_box + 56;

	// ---- once per CTA: constants, barriers, tensor memory
	{
		const uint4 *src = reinterpret_cast<const uint4 *>(p.screen_consts);
		uint4 *dst = reinterpret_cast<uint4 *>(smem + kOffConsts);
		for (int k = tid; k < kScBytes / 16; k += kScreenThreads) dst[k] = src[k];
	}
	if (tid == 0) {
		for (int q = 0; q < 7; q++) mbar_init(bar_box + 8 * q, 1);
		mbar_init(bar_tfree, kWarps);
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kOffTmemS), "n"(kScreenCols) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	sc_fence_before();
	__syncthreads();
	sc_fence_after();
	const uint32_t tmem = *reinterpret_cast<const volatile uint32_t *>(smem + kOffTmemS);
	const int n_items = p.n_tiles * p.n_streams;

	if (warp == kWarps) {
		// =====================================================================================================
		// producer (one thread): work items, TMA loads, MMAs.  Round r: publish the item, wait until the screeners have
		// read round r-1 out of tensor memory, run the MMAs, and as soon as they have consumed shared memory send the
		// next block's loads - which fly while the screeners evaluate round r's candidates.
		// =====================================================================================================
		if (lane == 0) {
			const uint64_t d_band = sc_desc(sbase + kOffConsts + kScBand, 1024, 128, kScNone);
			const uint64_t d_bconst = sc_desc(sbase + kOffConsts + kScConst, 4096, 128, kScNone);
			const uint64_t d_ac = sc_desc(sbase + kOffConsts + kScAc, 128, 0, kScNone);
			// Work items (stream, block) are handed out by a counter: a burst block done in place holds its CTA ten
			// times as long as a screened one.  Streams may be shorter than the launch's tile range.
			auto valid_item = [&](int it) {
				if (it >= n_items) return true;
				const int s_ = it / p.n_tiles;
				return p.tile0 + (it - s_ * p.n_tiles) < (int)p.jobs[s_].n_blocks;
			};
			int fenced_stream = -1;
			auto issue_loads = [&](int it) {
				const int s_ = it / p.n_tiles, t_ = p.tile0 + (it - s_ * p.n_tiles);
				const uint8_t *tm = reinterpret_cast<const uint8_t *>(p.tmaps) + (size_t)s_ * 256;
				if (s_ != fenced_stream) {
					asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm) : "memory");
					fenced_stream = s_;
				}
				mbar_expect_tx(bar_box + 32, kBox);
				sc_tma_2d(sbase + kOffHaloBox, tm, 384, t_ * 128 - 1, bar_box + 32);
#pragma unroll
				for (int q = 0; q < 4; q++) {
					mbar_expect_tx(bar_box + 8 * q, kBox);
					sc_tma_2d(sbase + q * kBox, tm, 128 * q, t_ * 128, bar_box + 8 * q);
				}
			};
			int item = (int)atomicAdd(p.work_ctr, 1u);
			while (!valid_item(item)) item = (int)atomicAdd(p.work_ctr, 1u);
			if (item < n_items) issue_loads(item);
			for (int round = 0;; round++) {
				const uint32_t ph = (uint32_t)(round & 1);
				// the counter's answer for the round after this one is on its way while the MMAs run
				int nx = (item < n_items) ? (int)atomicAdd(p.work_ctr, 1u) : n_items;
				if (round) {
					mbar_wait(bar_tfree, ph ^ 1u);   // round - 1 has been read out of tensor memory
					sc_fence_after();
				}
				// (published only now: the screeners are past round - 1's item, so the barrier never runs two phases ahead of
				// a waiter - a parity wait cannot tell phase r from phase r + 2)
				ss.item[round & 1] = item < n_items ? item : n_items;
				mbar_arrive(bar_item);
				if (item >= n_items) break;
				const int stream = item / p.n_tiles;
				const int tile = p.tile0 + (item - stream * p.n_tiles);
				// all 256 columns = -(128 * sum T16 + centre), in planes: needs nothing of the block, goes first
				sc_mma(tmem, d_ac, d_bconst, screen_idesc(256), 0u);
				mbar_wait(bar_box + 32, ph);
				if (tile == 0) {
					// nothing in front of the submit (TMA filled row -1 with zeros): the carried history, 96 bytes in front
					// of row 0.  The same bytes go to the slot's copy for decwin_kernel, which runs after save_history_kernel.
					StreamState *st = p.st + stream;
					const uint4 *hs = reinterpret_cast<const uint4 *>(st->hist[st->hist_parity & 1]);
					uint4 *hd = reinterpret_cast<uint4 *>(smem + kOffHaloBox + 32);
					uint4 *hc = reinterpret_cast<uint4 *>(p.hist_copy + (size_t)stream * kHistBytes);
#pragma unroll
					for (int k = 0; k < 6; k++) {
						const uint4 v = hs[k];
						hd[k] = v;
						hc[k] = v;
					}
					asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
				}
				mbar_wait(bar_box, ph);
				sc_fence_after();
				// the row before: byte slices -96.., -64.., -32.. reach outputs 0..2, 0..6, 0..10
#pragma unroll
				for (int s = 1; s <= 3; s++) {
					const uint64_t da = sc_desc(sbase + kOffHaloBox + 32 * s, 16, 1024, kScSw128);
					const uint64_t db = sc_desc(sbase + kOffConsts + kScBand + (64 - 16 * s) * 16, 1024, 128, kScNone);
					sc_mma(tmem, da, db, screen_idesc(16 * s), 1u);
				}
#pragma unroll
				for (int s = 0; s < 16; s++) {
					if (s && (s & 3) == 0) {
						mbar_wait(bar_box + 8 * (s >> 2), ph);
						sc_fence_after();
					}
					const uint64_t da = sc_desc(sbase + (s >> 2) * kBox + (s & 3) * 32, 16, 1024, kScSw128);
					const uint32_t n = (s <= 12) ? 64u : (uint32_t)(256 - 16 * s);
					sc_mma(tmem + 16 * s, da, d_band, screen_idesc(n), 1u);
				}
				sc_commit(bar_mma);
				while (!valid_item(nx)) nx = (int)atomicAdd(p.work_ctr, 1u);
				// shared memory is free once the MMAs are done: the next block's boxes
				mbar_wait(bar_mma, ph);
				if (nx < n_items) issue_loads(nx);
				item = nx;
			}
		}
	} else {
		// =====================================================================================================
		// screeners (four warps, thread t = row t of the block = tensor-memory lane t)
		// =====================================================================================================
		const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
		uint32_t st_sparse = 0, st_dense = 0, st_cand = 0, st_true = 0;   // thread 0's tallies
#ifdef TFR_SCREEN_PROFILE
		long long prof[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }, pt = clock64();
#endif
		for (int round = 0;; round++) {
			const uint32_t ph = (uint32_t)(round & 1);
			mbar_wait(bar_item, ph);
			const int item = ss.item[round & 1];
			if (item >= n_items) break;
			const int stream = item / p.n_tiles;
			const int tile = p.tile0 + (item - stream * p.n_tiles);
			const StreamJob job = p.jobs[stream];
			StreamState *st = p.st + stream;
			const size_t gtile = (size_t)job.dec_off + tile;
			const uint8_t *blk = job.iq + (size_t)tile * kBlockBytes;
			const uint8_t *hist = st->hist[st->hist_parity & 1];
			int thresh_lo = st->thresh;
			if (st->thresh_mode) thresh_lo = p.margin ? thresh_lo - p.margin : st->spec_lo;
			const int lim = thresh_lo - p.screen_slack;
			const int thr = (lim < 0) ? -1 : (lim << p.screen_shift);

			// ---- screen values out of tensor memory: 32 columns = 8 outputs x (I lo, I hi, Q lo, Q hi)
			mbar_wait(bar_mma, ph);
			sc_fence_after();
			SCPROF(0);
			unsigned long long cand64 = 0ull;
			int32_t *dbg = p.screen_dbg ? p.screen_dbg + (gtile * kBlockDec + (size_t)tid * kOutPerThread) * 2 : nullptr;
			{
				// two loads of 32 columns in flight (64 columns each fill the register file: 254 registers, and the back-end
				// kernels of the previous call no longer fit beside the two screening CTAs of an SM)
				uint32_t va[32], vb[32];
				auto screen8 = [&](const uint32_t (&v)[32], int j0) {
					unsigned m8 = 0u;   // (constant bit positions: a predicated OR per output instead of a 64-bit variable shift)
#pragma unroll
					for (int k = 0; k < 8; k++) {
						const int ci = (int)(v[4 * k + 1] << 8) + (int)v[4 * k], cq = (int)(v[4 * k + 3] << 8) + (int)v[4 * k + 2];
						if (abs(ci) + abs(cq) > thr) m8 |= 1u << k;
						if (dbg) { dbg[2 * (j0 + k)] = ci; dbg[2 * (j0 + k) + 1] = cq; }
					}
					cand64 |= (unsigned long long)m8 << j0;
				};
				sc_ld32(tlane, va);
				sc_ld32(tlane + 32, vb);
#pragma unroll 1
				for (int g = 0; g < 4; g++) {
					sc_wait_ld_all();   // both loads have landed
					screen8(va, 16 * g);
					if (g < 3) sc_ld32(tlane + 64 * (g + 1), va);
					screen8(vb, 16 * g + 8);
					if (g < 3) sc_ld32(tlane + 64 * (g + 1) + 32, vb);
				}
			}
			// this warp's quadrant of tensor memory is free for the next block's MMAs
			sc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_tfree);
			SCPROF(1);

			// ---- ordered candidate list
			const int nt = __popcll(cand64);
			int incl = nt;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const int o = __shfl_up_sync(0xffffffffu, incl, d);
				if (lane >= d) incl += o;
			}
			if (lane == 31) ss.warp_cnt[warp] = incl;
			sc_sync_screeners();
			int base = incl - nt;
			for (int w = 0; w < warp; w++) base += ss.warp_cnt[w];
			const int total = ss.warp_cnt[0] + ss.warp_cnt[1] + ss.warp_cnt[2] + ss.warp_cnt[3];
			const bool is_last = (tile == (int)job.n_blocks - 1);

			if (total > kCandMax) {
				// ---- a burst (telegram, start-up transient): the whole block with the exact cascade, here.  Shared
				// memory already belongs to the next block, so a thread reads its 512-byte row (and the 96 bytes in front
				// of it) from global memory - L2, the block has just been through - and writes its 64 samples straight to dec.
				const uint4 *row = reinterpret_cast<const uint4 *>(blk + (size_t)tid * 512);
				const uint4 *halo = (tid == 0 && tile == 0) ? reinterpret_cast<const uint4 *>(hist) : row - 6;
				uint32_t *dst = p.dec + gtile * kBlockDec + (size_t)tid * kOutPerThread;
				unsigned long long trig64 = 0ull;
				uint32_t prev = 0;
				fir_cascade<WIDE, 4>([&](int q) { return __ldg(halo + q); }, [&](int q) { return __ldg(row + q); },
						     [&](int m, int yi, int yq) {
							     if (abs(yi) + abs(yq) > thresh_lo) trig64 |= 1ull << m;
							     const uint32_t w = pack_iq(yi, yq);
							     if (m & 1) *reinterpret_cast<uint2 *>(dst + m - 1) = make_uint2(prev, w);
							     prev = w;
						     });
				FrontParams pk = p;
				pk.keep_all = 2;   // the samples are in dec already
				const uint32_t *blk_dec = p.dec + gtile * kBlockDec;
				block_epilogue(pk, job, tile, [&](int m) -> uint32_t { return blk_dec[m]; }, es, trig64, ScreenersSync());
				if (tid == 0) st_dense++;
				sc_sync_screeners();   // es is reused by the next burst block
			} else {
				if (nt) {
					unsigned long long msk = cand64;
					int k = base;
					while (msk) {
						const int b = __ffsll((long long)msk) - 1;
						msk &= msk - 1;
						ss.cand[k++] = (uint16_t)(tid * kOutPerThread + b);
					}
				}
				if (tid == 0 && is_last) ss.cand[total] = (uint16_t)(kBlockDec - 1);   // the lead-in sample of the next call
				sc_sync_screeners();
				SCPROF(2);
				// exact values with the reference arithmetic: lane k < 20 = stage-1 output 2m-18+k and stage-2 tap k
				const int n_eval = total + (is_last ? 1 : 0);
				const int t1 = WIDE ? kT1Wide[lane] : kT1Narrow[lane];
				const float c2 = (float)t1 * (1.0f / 65536.0f);
				const f2 c2p = pack2(c2, c2);
				for (int i0 = warp; i0 < n_eval; i0 += 4 * kWarps) {
					// (from global memory - L2 - because shared memory already belongs to the next block; bytes before a
					// submit's first block are the carried history.)  The loads of four candidates go out together.
					uint32_t w[4][4];
#pragma unroll
					for (int u = 0; u < 4; u++) {
						const int i = i0 + u * kWarps;
						if (i < n_eval && lane < 20) {
							const int o = 8 * (int)ss.cand[i] - 84 + 4 * lane;
#pragma unroll
							for (int q = 0; q < 4; q++) {
								const int oq = o + 4 * q;
								w[u][q] = (oq >= 0 || tile > 0) ? __ldg(reinterpret_cast<const uint32_t *>(blk + oq))
											: *reinterpret_cast<const uint32_t *>(hist + kHistBytes + oq);
							}
						}
					}
#pragma unroll
					for (int u = 0; u < 4; u++) {
						const int i = i0 + u * kWarps;
						if (i >= n_eval) break;   // warp uniform
						int fi = 0, fq = 0;
						if (lane < 20) {
							f2 x[8];
							x[0] = cvt_iq(w[u][0], 0); x[1] = cvt_iq(w[u][0], 1);
							x[2] = cvt_iq(w[u][1], 0); x[3] = cvt_iq(w[u][1], 1);
							x[4] = cvt_iq(w[u][2], 0); x[5] = cvt_iq(w[u][2], 1);
							x[6] = cvt_iq(w[u][3], 0); x[7] = cvt_iq(w[u][3], 1);
							const f2 y1 = stage1(x);
							const f2 acc = fma2_rm(y1, c2p, pack2((float)kA2One, (float)kA2One));
							uint32_t ai, aq;
							unpack2(acc, ai, aq);
							const int off = (kA2One - (1 << 23)) + kM1Mul * t1;
							fi = (int)(ai - 0x4B000000u) - off;
							fq = (int)(aq - 0x4B000000u) - off;
						}
						const int yi = __reduce_add_sync(0xffffffffu, fi), yq = __reduce_add_sync(0xffffffffu, fq);
						if (lane == 0) {
							ss.res_w[i] = pack_iq(yi, yq);
							ss.res_p[i] = abs(yi) + abs(yq);
						}
					}
				}
				sc_sync_screeners();
				SCPROF(3);
				// events in order, descriptor
				if (warp == 0) {
					uint32_t *ev = p.events + gtile * kMaxEvt;
					int nev = 0, last = -1;
					for (int b0 = 0; b0 < total; b0 += 32) {
						const int i = b0 + lane;
						const bool valid = i < total;
						const int pw = valid ? ss.res_p[i] : 0;
						const bool trig = valid && pw > thresh_lo;
						const unsigned mask = __ballot_sync(0xffffffffu, trig);
						if (trig) ev[nev + __popc(mask & ((1u << lane) - 1u))] = ((uint32_t)ss.cand[i] << 16) | (uint32_t)pw;
						if (mask) last = ss.cand[b0 + 31 - __clz(mask)];
						nev += __popc(mask);
					}
					if (lane == 0) {
						TileDesc *td = p.tiles + gtile;
						td->n_seg = 0;
						td->carry_out = (uint16_t)((last >= 0) ? max(last + p.t_max - kBlockDec, 0) : 0);
						td->n_trig = (uint32_t)nev;
						td->pad = 0;
						if (is_last) p.dec[gtile * kBlockDec + kBlockDec - 1] = ss.res_w[total];
						st_sparse++;
						st_cand += (uint32_t)total;
						st_true += (uint32_t)nev;
					}
				}
				SCPROF(4);
			}
#ifdef TFR_SCREEN_PROFILE
			prof[6]++;
#endif
		}
#ifdef TFR_SCREEN_PROFILE
		if (tid == 0)
			for (int k = 0; k < 7; k++) atomicAdd(&g_scprof[k], (unsigned long long)prof[k]);
#endif
		if (tid == 0 && p.screen_stat) {
			if (st_sparse) atomicAdd(p.screen_stat + 0, st_sparse);
			if (st_dense) atomicAdd(p.screen_stat + 1, st_dense);
			if (st_cand) atomicAdd(p.screen_stat + 2, st_cand);
			if (st_true) atomicAdd(p.screen_stat + 3, st_true);
		}
	}
	sc_fence_before();
	__syncthreads();
	if (warp == 0) {
		sc_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kScreenCols) : "memory");
	}
}

// ------------------------------------------------------------------------------------------------
// decwin_kernel: the decimated samples the demodulators read - exact FIR (fir_exact.cuh, the arithmetic of frontend.cu)
// over [start - 1, end] of every window of the demodulator with the longest timeout (its windows contain every other
// demodulator's: same triggers, longer hold).  A CTA takes kDwGroup consecutive windows at a time and spreads their
// aligned chunks of 16 outputs over its threads: per chunk 96 history bytes + 128 bytes from global memory, 18 + 32
// stage-1 outputs, 16 stage-2 outputs, one 64-byte store.  Chunks are computed whole: samples outside the window are
// exact as well.
// ------------------------------------------------------------------------------------------------
constexpr int kDwGroup = 6;
template <bool WIDE>
__global__ void __launch_bounds__(64) decwin_kernel(const BackParams p, const uint8_t *hist_copy, uint32_t *dec_out)
{
	const int stream = blockIdx.y;
	const StreamJob job = p.jobs[stream];
	if (job.n_blocks == 0) return;
	const int demod = p.demod;
	const uint32_t n_win = p.wincnt[stream].n[demod];
	const WinEntry *wl = p.wins + job.win_off + (size_t)demod * job.win_cap;
	const uint32_t call_len = job.n_blocks * (uint32_t)kBlockDec;
	uint32_t *out = dec_out + (size_t)job.dec_off * kBlockDec;
	const uint8_t *hist = hist_copy + (size_t)stream * kHistBytes;
	for (uint32_t w0 = blockIdx.x * kDwGroup; w0 < n_win; w0 += gridDim.x * kDwGroup) {
		// the group's chunk ranges [first, first + cnt), a chunk shared by two windows taken once
		uint32_t first[kDwGroup], cum[kDwGroup + 1];
		cum[0] = 0;
		uint32_t covered = 0;   // first chunk not yet taken
#pragma unroll
		for (int k = 0; k < kDwGroup; k++) {
			uint32_t f = 0, n = 0;
			if (w0 + k < n_win) {
				const WinEntry e = wl[w0 + k];
				if (e.start < call_len) {
					f = (e.start ? e.start - 1 : 0u) >> 4;
					const uint32_t l = min(e.end, call_len - 1) >> 4;
					if (k && f < covered) f = covered;
					if (l >= f) n = l - f + 1;
					covered = max(covered, l + 1);
				}
			}
			first[k] = f;
			cum[k + 1] = cum[k] + n;
		}
		for (uint32_t idx = threadIdx.x; idx < cum[kDwGroup]; idx += blockDim.x) {
			uint32_t c = 0;
#pragma unroll
			for (int k = 0; k < kDwGroup; k++)
				if (idx >= cum[k] && idx < cum[k + 1]) c = first[k] + (idx - cum[k]);
			// raw bytes 128c-96 .. 128c+127: x[-48..-1] and x[0..63] relative to the chunk
			const uint4 *row = reinterpret_cast<const uint4 *>(job.iq + (size_t)c * 128);
			const uint4 *halo = c ? row - 6 : reinterpret_cast<const uint4 *>(hist);
			uint32_t o[16];
			fir_cascade<WIDE, 1>([&](int q) { return __ldg(halo + q); }, [&](int q) { return __ldg(row + q); },
					     [&](int m, int yi, int yq) { o[m] = pack_iq(yi, yq); });
			uint4 *dst = reinterpret_cast<uint4 *>(out + (size_t)c * 16);
#pragma unroll
			for (int q = 0; q < 4; q++) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
		}
	}
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// The combined filter, its 16-bit rounding and the bound, for one filter type.  Returns the shift q (screen value =
// linear output * 2^q) and the slack in output units through *shift / *slack; fills the kScBytes blob.
void screen_build_consts(int wide, uint8_t *blob, int *shift, int *slack)
{
	static const int t2[8] = TFR_T2;
	static const int t1n[20] = TFR_T1N, t1w[20] = TFR_T1W;
	const int *t1 = wide ? t1w : t1n;
	long long T[46] = { 0 };
	for (int k = 0; k < 20; k++)
		for (int n = 0; n < 8; n++) T[2 * k + n] += (long long)t1[k] * t2[n];
	long long tmax = 0;
	for (int i = 0; i < 46; i++) tmax = std::max(tmax, T[i] < 0 ? -T[i] : T[i]);
	int k = 0;
	while (((tmax + (1ll << k) / 2) >> k) > 32767 - 128) k++;   // |T16| fits 256*hi + lo with hi, lo in [-128, 127]
	const int q = 26 - k;
	long long T16[46], sum16 = 0, err = 0;
	for (int i = 0; i < 46; i++) {
		T16[i] = (long long)floor((double)T[i] / (double)(1ll << k) + 0.5);
		sum16 += T16[i];
		const long long d = T[i] - T16[i] * (1ll << k);
		err += d < 0 ? -d : d;
	}
	// y = L/2^26 - E, E in (-8 N1, 8 P1 + 20)
	double P1 = 0, N1 = 0;
	for (int i = 0; i < 20; i++) (t1[i] > 0 ? P1 : N1) += fabs((double)t1[i]) / 65536.0;
	const double e_lo = -8.0 * N1, e_hi = 8.0 * P1 + 20.0;
	const double centre = 0.5 * (e_lo + e_hi), half = 0.5 * (e_hi - e_lo);
	const double round_err = 128.0 * (double)err / 67108864.0;   // |L - 2^k L16| / 2^26
	const long long centre_q = (long long)floor(centre * (double)(1ll << q) + 0.5);
	// per channel |y| <= |c|/2^q + half + round_err + 0.5/2^q; two channels, rounded up, +1 for good measure
	*slack = (int)ceil(2.0 * (half + round_err + 1.0 / (double)(1ll << q))) + 1;
	*shift = q;

	memset(blob, 0, kScBytes);
	auto lo_of = [](long long v) { long long l = ((v + 128) & 255) - 128; return l; };
	int8_t *band = reinterpret_cast<int8_t *>(blob + kScBand);
	for (int n = 0; n < 64; n++) {
		const int jp = n >> 2, c = (n >> 1) & 1, plane = n & 1;
		for (int kk = 0; kk < 32; kk++) {
			const int sg = kk >> 1, comp = kk & 1, i = sg - 4 * jp + 42;
			long long v = 0;
			if (comp == c && i >= 0 && i < 46) {
				const long long l = lo_of(T16[i]);
				v = plane ? (T16[i] - l) / 256 : l;
			}
			band[16 * n + 1024 * (kk >> 4) + (kk & 15)] = (int8_t)v;
		}
	}
	// constant: A row = 16 x 128, 16 x 1;  D = 128 * sum(B[0..15]) + sum(B[16..31]) = plane of -(128 * sum T16 + centre)
	const long long C = -(128 * sum16 + centre_q);
	const long long c_lo = lo_of(C), c_hi = (C - c_lo) / 256;
	int8_t *bc = reinterpret_cast<int8_t *>(blob + kScConst);
	for (int n = 0; n < 256; n++) {
		const int plane = n & 1;
		long long want = plane ? c_hi : c_lo;
		long long m = want / 128;   // toward zero; the remainder keeps the sign of want
		long long r = want - 128 * m;
		for (int kk = 0; kk < 16; kk++) {
			long long part = m / (16 - kk);   // spread m over the 16 entries, each within s8
			if (part > 127) part = 127;
			if (part < -127) part = -127;
			bc[16 * n + 4096 * 0 + kk] = (int8_t)part;
			m -= part;
		}
		bc[16 * n + 4096 * 1 + 0] = (int8_t)r;
	}
	uint8_t *ac = blob + kScAc;
	memset(ac, 128, 128);
	memset(ac + 128, 1, 128);
}

cudaError_t launch_frontend_screen(const FrontParams &p, int wide, int n_ctas, cudaStream_t stream)
{
	static bool attr_done[64] = { false };
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	if (dev < 64 && !attr_done[dev]) {
		e = cudaFuncSetAttribute(frontend_screen_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScreenSmem);
		if (e != cudaSuccess) return e;
		e = cudaFuncSetAttribute(frontend_screen_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScreenSmem);
		if (e != cudaSuccess) return e;
		if (getenv("TFR_DEBUG")) {
			int nb = 0;
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, frontend_screen_kernel<false>, kScreenThreads, kScreenSmem);
			fprintf(stderr, "[tfr] frontend_screen: %d CTAs per SM, %d B dynamic shared memory\n", nb, kScreenSmem);
		}
		attr_done[dev] = true;
	}
	const int items = p.n_tiles * p.n_streams;
	if (items <= 0) return cudaSuccess;
	const int grid = items < n_ctas ? items : n_ctas;
#ifdef TFR_SCREEN_PROFILE
	{
		static int calls = 0;
		if (++calls % 32 == 0) {
			unsigned long long r[8], z[8] = { 0 };
			cudaDeviceSynchronize();
			cudaMemcpyFromSymbol(r, g_scprof, sizeof(r));
			cudaMemcpyToSymbol(g_scprof, z, sizeof(z));
			const double n = (double)(r[6] ? r[6] : 1);
			fprintf(stderr, "[scprof] %llu blocks, cycles per block (screener thread 0): waiting for the MMAs %.0f, readback %.0f, candidate list %.0f, evaluation %.0f, events %.0f\n",
				r[6], r[0] / n, r[1] / n, r[2] / n, r[3] / n, r[4] / n);
		}
	}
#endif
	if (wide)
		frontend_screen_kernel<true><<<grid, kScreenThreads, kScreenSmem, stream>>>(p);
	else
		frontend_screen_kernel<false><<<grid, kScreenThreads, kScreenSmem, stream>>>(p);
	return cudaGetLastError();
}

cudaError_t launch_decwin(const BackParams &p, int wide, const uint8_t *hist_copy, uint32_t *dec_out, cudaStream_t s)
{
	if (p.max_blocks <= 0) return cudaSuccess;
	int gx = (p.max_blocks + kDwGroup - 1) / kDwGroup;   // about one window per block
	gx = gx < 1 ? 1 : (gx > 4096 ? 4096 : gx);
	if (wide)
		decwin_kernel<true><<<dim3(gx, p.n_streams), 64, 0, s>>>(p, hist_copy, dec_out);
	else
		decwin_kernel<false><<<dim3(gx, p.n_streams), 64, 0, s>>>(p, hist_copy, dec_out);
	return cudaGetLastError();
}

}  // namespace tfr
