// frontend_common.cuh - what the two front-end kernels (frontend.cu: shared-memory conversion; frontend_tc.cu:
// byte->float conversion by the tensor core) share: packed-FMA and mbarrier/TMA PTX helpers, the exact-floor chain
// bookkeeping of FIR stage 2, and the block epilogue (trigger events, segments, sparse store, descriptor).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tfr_dev.h"
#include "fir_taps.h"

namespace tfr {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kOutPerThread = 64;

// Stage 2: a run of taps [lo, hi) of one output is one FMA chain.  Stage-1 outputs arrive as kM1 + y1 with
// kM1 = M1MUL * 65536, so every tap injects the integer M1MUL * t1[n].  The chain's accumulator starts at a constant
// chosen so that every partial sum - the injected offsets plus the partial result, |.| < kYMargin - stays inside
// [2^23, 2^24); the constants are removed in integer arithmetic when the accumulator bits are read back.
constexpr int kYMargin = 13000;   // |sum of any subset of the per-tap floors| <= 1.423 * 8520 + 20
struct ChainK { int start; int off; bool ok; };
__host__ __device__ constexpr ChainK chain_k(int m1mul, bool wide, int lo, int hi)
{
	long long s = 0, mn = 0, mx = 0;
	for (int n = lo; n < hi; n++) {
		s += (long long)m1mul * t1_tap(wide, n);
		if (s < mn) mn = s;
		if (s > mx) mx = s;
	}
	const long long start = (1ll << 23) + kYMargin - mn;
	return ChainK{ (int)start, (int)s, start + mx + kYMargin < (1ll << 24) };
}
template <int M1MUL, bool WIDE, int LO, int HI>
struct ChainC {
	static constexpr ChainK k = chain_k(M1MUL, WIDE, LO, HI);
	static_assert(LO >= HI || k.ok, "stage-2 chain leaves the integer binade");
	// accumulator bits - bias = the chain's exact sum of floors
	static constexpr uint32_t bias = (LO < HI) ? (0x4B000000u - (1u << 23) + (uint32_t)k.start + (uint32_t)k.off) : 0u;
};

// ------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------
typedef unsigned long long f2;  // packed f32x2: lo = I, hi = Q

__device__ __forceinline__ f2 pack2(float lo, float hi)
{
	f2 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ void unpack2(f2 v, uint32_t &lo, uint32_t &hi)
{
	asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ f2 fma2_rm(f2 a, f2 b, f2 c)
{
	f2 d;
	asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra DONE_%=;\n"
		"bra WAIT_%=;\n"
		"DONE_%=:\n"
		"}\n" ::"r"(bar), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
		     "l"(src), "r"(bytes), "r"(bar)
		     : "memory");
}
__device__ __forceinline__ uint32_t pack_iq(int yi, int yq) { return __byte_perm((uint32_t)yi, (uint32_t)yq, 0x5410); }

// ------------------------------------------------------------------------------------------------
// block epilogue, shared by both front-end kernels: trigger bookkeeping (ordered event list, covered segments),
// sparse store of the decimated samples and the block descriptor.  On entry the block's 8192 outputs sit in shared
// memory wherever the kernel put them - `smp(m)` returns output m (I lo16, Q hi16), thread t owns outputs
// 64t .. 64t+63 - and trig64 holds the thread's 64 trigger bits.  The caller has NOT synchronised yet.
// ------------------------------------------------------------------------------------------------
struct EpiShared {
	int first[kThreads];   // position of the first trigger among the thread's 64 outputs, -1 if none
	int last[kThreads];
	int seg_start[kMaxSeg + 4], seg_end[kMaxSeg + 4];
	int nseg, ntrig;
	int warp_cnt[kWarps];
};
struct CtaSync {
	__device__ __forceinline__ void operator()() const { __syncthreads(); }
};
// `sync` is the barrier of the kThreads threads that call (the whole CTA, unless the kernel has other warps besides)
template <class Smp, class Sync = CtaSync>
__device__ __forceinline__ void block_epilogue(const FrontParams &p, const StreamJob &job, int tile, const Smp &smp, EpiShared &es,
					       unsigned long long trig64, const Sync &sync = Sync())
{
	const int tid = threadIdx.x;
	const uint32_t trig[2] = { (uint32_t)trig64, (uint32_t)(trig64 >> 32) };
	{
		int f = -1, l = -1;
		if (trig[0]) f = __ffs(trig[0]) - 1;
		else if (trig[1]) f = 32 + __ffs(trig[1]) - 1;
		if (trig[1]) l = 63 - __clz(trig[1]);
		else if (trig[0]) l = 31 - __clz(trig[0]);
		es.first[tid] = (f < 0) ? -1 : tid * kOutPerThread + f;
		es.last[tid] = (l < 0) ? -1 : tid * kOutPerThread + l;
		const int nt = __popc(trig[0]) + __popc(trig[1]);
		// ordered event list: exclusive scan of the per-thread trigger counts over the CTA
		int incl = nt;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const int o = __shfl_up_sync(0xffffffffu, incl, d);
			if ((tid & 31) >= d) incl += o;
		}
		if ((tid & 31) == 31) es.warp_cnt[tid >> 5] = incl;
		sync();
		int base = incl - nt;
		for (int w = 0; w < (tid >> 5); w++) base += es.warp_cnt[w];
		if (tid == kThreads - 1) es.ntrig = base + nt;
		if (nt && base < kMaxEvt) {
			uint32_t *ev = p.events + ((size_t)job.dec_off + tile) * kMaxEvt;
			unsigned long long msk = trig64;
			int k = base;
			while (msk && k < kMaxEvt) {
				const int b = __ffsll((long long)msk) - 1;
				msk &= msk - 1;
				const uint32_t w = smp(tid * kOutPerThread + b);   // this thread's own output b (I lo16, Q hi16)
				const int pi = (int)(int16_t)(w & 0xffff), pq = (int)(int16_t)(w >> 16);
				ev[k++] = ((uint32_t)(tid * kOutPerThread + b) << 16) | (uint32_t)(abs(pi) + abs(pq));
			}
		}
	}

	// warp 0: merge the per-thread trigger extents into covered segments [start, end)
	if (tid < 32) {
		// each lane scans 4 consecutive threads; a segment can only begin at a thread's first trigger
		// because t_max (>= 356) exceeds the 64 samples a thread owns
		int lastq = -1;          // last trigger seen before this lane's group (filled by the scan below)
		int grp_last = -1;
#pragma unroll
		for (int k = 0; k < 4; k++) grp_last = max(grp_last, es.last[tid * 4 + k]);
		// inclusive prefix max over lanes, then shift to exclusive
		int pm = grp_last;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			int o = __shfl_up_sync(0xffffffffu, pm, d);
			if (tid >= d) pm = max(pm, o);
		}
		lastq = __shfl_up_sync(0xffffffffu, pm, 1);
		if (tid == 0) lastq = -1;
		// walk the 4 threads of the group, emitting (start) markers and tracking chain ends
		int starts[4];
		int q = lastq;
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const int f = es.first[tid * 4 + k], l = es.last[tid * 4 + k];
			starts[k] = -1;
			if (f >= 0) {
				if (q < 0 || f - q > p.t_max) starts[k] = f;
				q = l;
			}
		}
		// number the starts across the warp
		int cnt = 0;
#pragma unroll
		for (int k = 0; k < 4; k++) cnt += (starts[k] >= 0);
		int incl = cnt;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			int o = __shfl_up_sync(0xffffffffu, incl, d);
			if (tid >= d) incl += o;
		}
		int base = incl - cnt;
		const int total = __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
		for (int k = 0; k < 4; k++)
			if (starts[k] >= 0) {
				if (base < kMaxSeg + 4) es.seg_start[base] = starts[k];
				base++;
			}
		if (tid == 0) es.nseg = min(total, kMaxSeg);
		__syncwarp();
		// segment i ends t_max after the last trigger that precedes segment i+1's start
		const int nseg = min(total, kMaxSeg);
		const int lastall = __shfl_sync(0xffffffffu, pm, 31);
		if (tid < nseg) {
			int endq;
			if (tid == nseg - 1) {
				endq = lastall;
			} else {
				// last trigger strictly before the next start: scan thread slots backwards from the owner of that start
				const int nxt = es.seg_start[tid + 1];
				int t = nxt / kOutPerThread - 1;
				while (t >= 0 && es.last[t] < 0) t--;
				endq = (t >= 0) ? es.last[t] : -1;
			}
			es.seg_end[tid] = endq + p.t_max;   // exclusive; may exceed the block -> carry_out
		}
	}
	sync();

	// ------------------------------------------------------------------ sparse store + descriptor
	const size_t gtile = (size_t)job.dec_off + tile;
	uint32_t *dst = p.dec + gtile * kBlockDec;
	const int nseg = es.nseg;
	auto sample = [&](int m) -> uint32_t { return smp(m); };
	if (p.keep_all == 2) {
		// the caller has already written every sample of the block to dec (frontend_screen.cu, a burst block done in place)
	} else if (p.keep_all) {
		for (int m = tid; m < kBlockDec; m += kThreads) dst[m] = sample(m);
	} else {
		// head [0, t_max) and the last sample are always kept (needed when the previous block's trigger
		// reaches into this one, and as the next block's lead-in sample)
		int covered = min(p.t_max, kBlockDec);
		for (int m = tid; m < covered; m += kThreads) dst[m] = sample(m);
		if (tid == 0) dst[kBlockDec - 1] = sample(kBlockDec - 1);
		for (int sgi = 0; sgi < nseg; sgi++) {
			const int s0 = max(max(es.seg_start[sgi] - 1, covered), 0);   // one lead-in sample
			const int e0 = min(es.seg_end[sgi], kBlockDec);
			for (int m = s0 + tid; m < e0; m += kThreads) dst[m] = sample(m);
			covered = max(covered, e0);
		}
	}
	if (tid < kMaxSeg) {
		TileDesc *td = p.tiles + gtile;
		const bool on = tid < nseg;
		const int s0 = on ? es.seg_start[tid] : 0;
		const int e0 = on ? min(es.seg_end[tid], kBlockDec) : 0;
		td->seg_start[tid] = (uint16_t)s0;
		td->seg_len[tid] = (uint16_t)(e0 - s0);
		if (tid == 0) {
			td->n_seg = (uint16_t)nseg;
			const int over = nseg ? es.seg_end[nseg - 1] - kBlockDec : 0;
			td->carry_out = (uint16_t)max(over, 0);
			td->n_trig = (uint32_t)es.ntrig;
			td->pad = 0;
		}
	}
}

}  // namespace tfr
