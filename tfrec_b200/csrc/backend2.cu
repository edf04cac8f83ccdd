// backend2.cu - the parallel back-end.
//
// The reference runs every demodulator as one serial state machine over the stream.  What is actually
// serial is small: a demodulator only runs inside *windows* (a trigger at sample t keeps it active for
// t .. t+T_d-1, retriggerable: tfa1.cpp:147-148, tfa2.cpp:351-356), window extents are a pure function
// of the trigger positions, and almost all demodulator state is reset when a window closes.  What does
// survive a window boundary is
//     TFA_1            the decoder's 32-bit shift register            (tfa1.cpp:115-117)
//     TFA_2/3, TX22    the biquad state and last_bit_idx              (tfa2.cpp:325-334)
// So:  thresh2_kernel  walks the per-block trigger EVENT lists the front-end wrote (cheap), reproduces
//                      fsk_demod::process' per-block bookkeeping (fm_demod.cpp:51-73) and lists every
//                      demodulator's windows (start, end) in stream order;
//      devfm_kernel    computes fm_dev (dsp_stuff.cpp:284-292, FP64 atan2) for every stored sample, in parallel;
//      win_kernel      ONE THREAD PER WINDOW runs slicer + framer with a *speculated* carry-in: the shift
//                      register is assumed empty, last_bit_idx is assumed far in the past, and the biquad is
//                      warmed up over the preceding windows' samples; each run records what is needed to
//                      check those assumptions;
//      verify_kernel   one thread per (stream, demod) walks the window records in order with the TRUE carried
//                      state, proves for each window that its assumed carry-in was equivalent to the true
//                      one (bitwise for the biquad outputs) and re-runs the window inline when it was not.
// The result is exact by construction; speculation only decides how often the slow path is taken
// (Counters::n_reruns).  WeatherHub keeps the serial walker of backend.cu (its 0.0025/spb averaging
// biquad, whb.cpp:611, remembers ~10^5 samples, so windows are not independent in any useful sense).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "demod_dev.cuh"

namespace tfr {
#ifdef TFR_WIN_PROFILE
__device__ unsigned long long g_slprof[8];   // of one 2784-sample window: fast groups, general groups, cycles in each
#endif


#ifndef TFR_CHAIN_MAX
#define TFR_CHAIN_MAX 8
#endif
#ifndef TFR_WARM_X4
#define TFR_WARM_X4 12   // biquad warm-up length of a speculated window in quarter timeouts
#endif
constexpr int kChainMax = TFR_CHAIN_MAX;
#ifndef TFR_CHAIN_GROUP
#define TFR_CHAIN_GROUP 1
#endif
constexpr uint32_t kChainGroup = TFR_CHAIN_GROUP;
#ifdef TFR_WIN_PROFILE
__device__ unsigned long long g_winprof[16];
#endif
__device__ __forceinline__ bool win_near(const WinEntry *wl, uint32_t w, int timeout)
{
	return wl[w].start - wl[w - 1].end <= (uint32_t)timeout + 1u;
}
// pure function of the window list, so that every thread agrees on where chains begin
__device__ __forceinline__ bool chain_head(const WinEntry *wl, uint32_t w, int timeout)
{
	uint32_t k = 0;
	while (w - k > 0 && k < 8u * kChainMax && win_near(wl, w - k, timeout)) k++;
	if (k) return (k % kChainMax) == 0;
	// A window that is not near could be speculated on its own, but every speculated window pays a biquad warm-up of
	// several timeouts (and some of them a repair): only every kChainGroup-th one is a head, the thread carries the
	// true filter state and last_bit_idx through the windows in between.  (Measured slower for every G > 1.)
	return (w % kChainGroup) == 0;
}


// ------------------------------------------------------------------------------------------------
// walk_table_kernel: one warp per block of the launch, ahead of thresh2_kernel on the same stream.
//
// The walk's chain (threshold -> triggered count -> average -> threshold) is serial over a stream's blocks, but what a
// block contributes is a function of the block and ITS threshold alone, and the threshold moves in steps of two around
// where the launch found it.  So every block is evaluated here, in parallel, under the kWalkNT thresholds
// base + 2k (base = threshold at the start of the launch - kWalkNT): first and last trigger, samples its own triggers
// cover, and the trigger gaps no shorter than the shortest demodulator timeout - the only triggers besides the first
// that can open a window (tfa1.cpp:147-148, tfa2.cpp:351-356).  Burst blocks (no complete event list) are scanned
// here too, by all warps at once instead of by the stream's one.  The walk then costs a table look-up per block and
// never repeats a block because the threshold stepped.
//   entry: x = first | last << 16      y = covered | min(gaps, 3) << 16 | has << 31      z, w = the first two gaps,
//   (trigger before) << 16 | (trigger after)
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kWalkSet = 0x80000000u;   // s_trig item: only moves last_trig
// thresholds above the table's last column at which walk_table_kernel also evaluates a block (lanes 17..31)
__device__ const int kWalkLadder[15] = { 4, 8, 16, 24, 32, 48, 64, 96, 128, 192, 256, 384, 512, 768, 1024 };
__global__ void __launch_bounds__(128) walk_table_kernel(const BackParams p)
{
	const int stream = blockIdx.y;
	const int lane = threadIdx.x & 31;
	const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
	const StreamJob job = p.jobs[stream];
	if (job.n_blocks == 0) return;
	const StreamState *st = p.st + stream;
	const int b0 = (int)st->t2_done;
	const int b = b0 + i;
	if (b >= min(b0 + p.n_tiles, (int)job.n_blocks)) return;
	const DevConfig &cfg = *p.cfg;
	const int base = st->thresh - kWalkNT;
	if (i == 0 && lane == 0) p.walk_base[stream] = base;
	int t_min = 0x7fffffff;
	for (int d = 0; d < cfg.n_demods; d++) t_min = min(t_min, cfg.d[d].timeout);
	const int t_max = cfg.t_max;
	// Lane 16 evaluates the block under the lowest threshold its lists are complete for, lanes 17..31 under a ladder of
	// thresholds above the table.  The triggers above a threshold shrink as it rises, so an entry that is the same at
	// both ends of a range of thresholds holds for the whole range (first/last trigger and coverage are monotone in the
	// set; a gap of the middle set spans no trigger of the larger one and is one of the smaller one's): bit 30 of column
	// 0 says so for the range below the table, bits 26..29 of column 15 hold the number of ladder steps its entry stays
	// the same for, and the walk follows the threshold out of the table that far - which it does after every telegram.
	const int mode = st->thresh_mode;
	const int theta_lo = (mode == 1) ? (p.margin ? st->thresh - p.margin : st->spec_lo) : st->thresh;
	const int theta_top = base + 2 * (kWalkNT - 1);
	const int theta_hi = theta_top + kWalkLadder[14];
	const int kk = lane;
	const int theta = (lane < kWalkNT) ? base + 2 * lane : (lane == kWalkNT ? theta_lo : theta_top + kWalkLadder[lane - kWalkNT - 1]);
	const int theta_min = min(base, theta_lo);
	const size_t g = (size_t)job.dec_off + b;
	const uint32_t n = p.tiles[g].n_trig;
	int f = -1, l = 0, cov = 0, cend = 0, ng = 0;
	uint32_t g0 = 0, g1 = 0, g2 = 0, g3 = 0;
	// triggers pf .. pl (less than t_max apart, in order after everything seen so far)
	auto upd = [&](int pf, int pl) {
		if (f < 0) {
			f = pf;
		} else if (pf - l >= t_min) {
			const uint32_t gp = ((uint32_t)l << 16) | (uint32_t)pf;
			if (ng == 0) g0 = gp;
			else if (ng == 1) g1 = gp;
			else if (ng == 2) g2 = gp;
			else if (ng == 3) g3 = gp;
			ng = min(ng + 1, 5);
		}
		const int hi = min(pl + t_max, kBlockDec);
		cov += max(hi - max(pf, cend), 0);
		cend = hi;
		l = pl;
	};
	if (n <= (uint32_t)kMaxEvt) {
		const uint32_t *ev = p.events + g * kMaxEvt;
		for (uint32_t e0 = 0; e0 < n; e0 += 32) {
			const uint32_t mine = (e0 + (uint32_t)lane < n) ? ev[e0 + lane] : 0u;
			const int cnt = (int)min(32u, n - e0);
			for (int q = 0; q < cnt; q++) {
				const uint32_t e = __shfl_sync(0xffffffffu, mine, q);
				if ((int)(e & 0xffff) > theta) upd((int)(e >> 16), (int)(e >> 16));
			}
		}
	} else {
		// a burst block: every sample that can be a trigger lies in a stored segment
		const TileDesc &td = p.tiles[g];
		const uint32_t *d = p.dec + g * kBlockDec;
		const int ns = td.n_seg;
		for (int sgi = 0; sgi < ns; sgi++) {
			const int a = td.seg_start[sgi], e_ = a + td.seg_len[sgi];
			for (int mb = a; mb < e_; mb += 256) {
				int pw[8];
#pragma unroll
				for (int q = 0; q < 8; q++) {
					const int m = mb + 32 * q + lane;
					pw[q] = (m < e_) ? pwr_of(d[m]) : (int)0x80000000;
				}
#pragma unroll
				for (int q = 0; q < 8; q++) {
					const int m0 = mb + 32 * q;
					const unsigned lo = __ballot_sync(0xffffffffu, pw[q] > theta_min);
					if (!lo) continue;
					unsigned mm = lo;
					if (lo != __ballot_sync(0xffffffffu, pw[q] > theta_hi) || p.walk_dbg == 4) {
						for (int k = 0; k < 32; k++) {
							const unsigned mk = __ballot_sync(0xffffffffu, pw[q] > __shfl_sync(0xffffffffu, theta, k));
							if (kk == k) mm = mk;
						}
					}
					if (mm) upd(m0 + __ffs(mm) - 1, m0 + 31 - __clz(mm));
				}
			}
		}
	}
	const bool has = f >= 0;
	const uint32_t fl = has ? ((uint32_t)f | ((uint32_t)l << 16)) : 0u;
	uint32_t cv = (uint32_t)cov | ((uint32_t)ng << 16) | (has ? 0x80000000u : 0u);
	if (p.walk_gap) {
		// lane 16 against column 0, lanes 17.. against column 15
		const int ref = (lane <= kWalkNT) ? 0 : kWalkNT - 1;
		// (every shuffle by every lane: no short-circuit between them)
		const uint32_t r_fl = __shfl_sync(0xffffffffu, fl, ref), r_cv = __shfl_sync(0xffffffffu, cv, ref);
		const uint32_t r_g0 = __shfl_sync(0xffffffffu, g0, ref), r_g1 = __shfl_sync(0xffffffffu, g1, ref);
		const uint32_t r_g2 = __shfl_sync(0xffffffffu, g2, ref), r_g3 = __shfl_sync(0xffffffffu, g3, ref);
		const bool same = (r_fl == fl) & (r_cv == cv) & (r_g0 == g0) & (r_g1 == g1) & (r_g2 == g2) & (r_g3 == g3);
		const unsigned eq = __ballot_sync(0xffffffffu, same);
		if (lane == 0 && ((eq >> kWalkNT) & 1u) && p.walk_dbg < 2) cv |= 0x40000000u;
		if (lane == kWalkNT - 1 && (p.walk_dbg < 1 || p.walk_dbg == 4 || p.walk_dbg == 32 || (p.walk_dbg == 8 && n <= (uint32_t)kMaxEvt) || (p.walk_dbg == 16 && n > (uint32_t)kMaxEvt))) cv |= (uint32_t)(__ffs(~(eq >> (kWalkNT + 1))) - 1) << 26;   // (15 lanes: at most 15)
	}
	if (lane < kWalkNT) {
		if (p.walk_gap) {   // for walk_cta_kernel: the chain's half apart (it is staged in shared memory), four gaps, 5 = more
			reinterpret_cast<uint2 *>(p.walk_tab)[g * kWalkNT + lane] = make_uint2(fl, cv);
			reinterpret_cast<uint4 *>(p.walk_gap)[g * kWalkNT + lane] = make_uint4(g0, g1, g2, g3);
		} else {            // for thresh2_kernel: two gaps, 3 = more
			reinterpret_cast<uint4 *>(p.walk_tab)[g * kWalkNT + lane] =
				make_uint4(fl, (uint32_t)cov | ((uint32_t)min(ng, 3) << 16) | (has ? 0x80000000u : 0u), g0, g1);
		}
	}
}

// ------------------------------------------------------------------------------------------------
// thresh2_kernel: one warp per stream; reproduces fsk_demod::process' per-block bookkeeping
// (fm_demod.cpp:51-73) from the front-end's event lists and lists every demodulator's windows.
//
// Block b's threshold depends on the triggered counts of the blocks before it, so the walk over blocks is a
// serial chain.  What is on the chain is kept minimal.  Blocks are taken in aligned chunks of 32, lane L owning
// block chunk+L:
//   1. every lane walks ITS block's events against the current threshold (assumed constant over the chunk):
//      first/last trigger, samples covered by its own triggers, count, and the first four trigger positions
//   2. the coverage carried in from the previous block is one shuffle away (a trigger covers t_max < 8192
//      samples, so it never reaches past the next block)
//   3. the chain itself: triggered_avg = (31*avg + triggered)/32 and the +-2 threshold step every 4th block,
//      evaluated by all lanes from shuffled counts - a handful of integer instructions per block.  The blocks
//      up to and including the first threshold change are accepted; the rest of the chunk is redone with the new
//      threshold (in steady state the threshold sits in its dead band and whole chunks are accepted)
//   4. the accepted blocks' true triggers (about one per block) are fed, in order, to the per-demod window
//      bookkeeping (lane d = demod d).
// Blocks with more than kMaxEvt events (a burst) have no complete list: the warp scans their stored samples.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32, 1) thresh2_kernel(const BackParams p)
{
	const int stream = blockIdx.x;
	const int lane = threadIdx.x;
	__shared__ uint32_t s_trig[192];   // the triggers of a chunk's accepted blocks, in order (at most four / six items per block)
	__shared__ uint4 s_tab[32][kWalkNT + 1];   // the chunk's rows of the walk table (padded: conflict-free 16-byte stores)
	__shared__ int s_t[32], s_avg[32];         // a pass' triggered counts and running averages
	if (stream >= p.n_streams) return;
	const StreamJob job = p.jobs[stream];
	const DevConfig &cfg = *p.cfg;
	WinCount *part = (p.part_idx >= 0) ? p.partcnt + (size_t)p.part_idx * p.n_streams + stream : nullptr;
	if (job.n_blocks == 0) {
		if (part && lane < cfg.n_demods) part->n[lane] = 0;
		return;
	}
	StreamState *st = p.st + stream;
	const int b0 = (int)st->t2_done;
	if (b0 >= (int)job.n_blocks) {   // this stream finished in an earlier launch of the call: every window is final
		if (part && lane < cfg.n_demods) part->n[lane] = p.wincnt[stream].n[lane];
		return;
	}
	const int t_max = cfg.t_max;
	const int nd = cfg.n_demods;
	const uint32_t call_len = job.n_blocks * (uint32_t)kBlockDec;

	int thresh = st->thresh, avg = st->triggered_avg, runs = st->runs;
	int c = st->any_timeout;              // samples at the head of the next block still covered by a trigger
	const int mode = st->thresh_mode;
	// the bound the front-end launches of these blocks kept samples and events for
	const int thresh_lo = (mode == 1) ? (p.margin ? thresh - p.margin : st->spec_lo) : thresh;
	long long last_trig;                  // position of the latest trigger (negative: in an earlier call)
	// per-demod window bookkeeping lives in lane d
	const int T_d = (lane < nd) ? cfg.d[lane].timeout : 0x7fffffff;
	uint32_t n_win = 0, cum = 0, open = 0, open_start = 0;
	WinEntry *wl = (lane < nd) ? p.wins + job.win_off + (size_t)lane * job.win_cap : nullptr;
	if (b0 == 0) {
		last_trig = -(long long)st->trig_age;
		if (lane < nd && st->win_cont[lane]) {   // a window was open when the previous call ended
			WinEntry e = { 0u, 0xffffffffu, 0u, kWinCont };
			wl[0] = e;
			n_win = 1;
			open = 1;
		}
	} else {
		last_trig = st->call_last_trig;
		if (lane < nd) {
			n_win = st->win_n[lane];
			cum = st->win_cum[lane];
			open = st->win_open[lane];
			if (open && n_win) open_start = wl[n_win - 1].start;
		}
	}
	unsigned long long act_lane = 0;
	const int t_end = min(b0 + p.n_tiles, (int)job.n_blocks);
	int t_stop = t_end;                   // first block NOT walked (a violated bound stops the walk early)
	const bool tab = p.walk_tab != nullptr;   // walk_table_kernel ran ahead of this launch
#ifdef TFR_WALK_PROFILE
	long long wp_t0 = clock64(), wp_chain = 0, wp_items = 0, wp_feed = 0, wp_stage = 0, wp_old = 0, wp_a;
	int wp_fast = 0, wp_slow = 0, wp_nitems = 0, wp_ncomplex = 0;
#define WP(x) x
#else
#define WP(x)
#endif
	const int tab_base = tab ? p.walk_base[stream] : 0;

	// one trigger at position t (inside the call), in order: the window lists
	auto win_trigger = [&](uint32_t t) {
		if (lane < nd) {
			if (!open || (long long)t - last_trig >= T_d) {
				if (open) {   // close the previous window: it flushed T_d-1 samples after its last trigger
					const uint32_t end = (uint32_t)(last_trig + T_d - 1);
					wl[n_win - 1].end = end;
					cum += end - open_start + 1;
				}
				if (n_win < job.win_cap) {
					WinEntry e = { t, 0xffffffffu, cum, 0u };
					wl[n_win] = e;
					n_win++;
					open_start = t;
				} else {
					p.counters->overflow = 1;
				}
				open = 1;
			}
		}
		last_trig = t;
	};

	// lane L's block of a chunk: event count and the first sixteen events (a block holds 5-10 events above the
	// speculative bound; a list that does not fit the registers costs one dependent L2 round trip per four events
	// on the stream's serial chain)
	auto fetch = [&](int chunk, uint32_t &n, uint4 &ea, uint4 &eb, uint4 &ec, uint4 &ed) {
		n = 0;
		ea = eb = ec = ed = make_uint4(0, 0, 0, 0);
		const int b = chunk + lane;
		if (b >= b0 && b < t_end) {
			const size_t g = (size_t)job.dec_off + b;
			n = p.tiles[g].n_trig;
			const uint4 *ev = reinterpret_cast<const uint4 *>(p.events + g * kMaxEvt);
			ea = ev[0];
			eb = ev[1];
			ec = ev[2];
			ed = ev[3];
		}
	};

	// lane L's row of the walk table
	auto fetch_row = [&](int chunk, uint4 (&row)[kWalkNT]) {
		const int b = chunk + lane;
		if (b >= b0 && b < t_end) {
			const uint4 *src = reinterpret_cast<const uint4 *>(p.walk_tab) + ((size_t)job.dec_off + b) * kWalkNT;
#pragma unroll
			for (int k = 0; k < kWalkNT; k++) row[k] = src[k];
		} else {
#pragma unroll
			for (int k = 0; k < kWalkNT; k++) row[k] = make_uint4(0, 0, 0, 0);
		}
	};
	auto stage_row = [&](const uint4 (&row)[kWalkNT]) {
		__syncwarp();
#pragma unroll
		for (int k = 0; k < kWalkNT; k++) s_tab[lane][k] = row[k];
		__syncwarp();
	};
	// every trigger of block chunk+j above theta_j to the window lists, in order (all lanes call this together)
	auto feed_block = [&](int chunk, int j, int theta_j) {
		const uint32_t base = (uint32_t)(chunk + j) * kBlockDec;
		const size_t gj = (size_t)job.dec_off + chunk + j;
		const uint32_t nj = p.tiles[gj].n_trig;
		if (nj <= (uint32_t)kMaxEvt) {
			const uint32_t *ev = p.events + gj * kMaxEvt;
			for (uint32_t e0 = 0; e0 < nj; e0 += 32) {
				const bool have = e0 + (uint32_t)lane < nj;
				const uint32_t e = have ? ev[e0 + lane] : 0u;
				unsigned mask = __ballot_sync(0xffffffffu, have && (int)(e & 0xffff) > theta_j);
				while (mask) {
					const int k = __ffs(mask) - 1;
					mask &= mask - 1;
					win_trigger(base + (__shfl_sync(0xffffffffu, e, k) >> 16));
				}
			}
		} else {
			const TileDesc &td = p.tiles[gj];
			const uint32_t *d = p.dec + gj * kBlockDec;
			const int ns = td.n_seg;
			for (int sgi = 0; sgi < ns; sgi++) {
				const int a = td.seg_start[sgi], b = a + td.seg_len[sgi];
				for (int mb = a; mb < b; mb += 256) {
					uint32_t v[8];
#pragma unroll
					for (int q = 0; q < 8; q++) {
						const int m = mb + 32 * q + lane;
						v[q] = (m < b) ? d[m] : 0u;
					}
#pragma unroll
					for (int q = 0; q < 8; q++) {
						const int m0 = mb + 32 * q, m = m0 + lane;
						const unsigned mask = __ballot_sync(0xffffffffu, (m < b) && (pwr_of(v[q]) > theta_j));
						if (mask) {
							const int pf = __ffs(mask) - 1, pl = 31 - __clz(mask);
							win_trigger(base + m0 + pf);
							if (pl != pf) last_trig = base + m0 + pl;   // < 32 apart: they only move the tail
						}
					}
				}
			}
		}
	};

	bool stop = false;
	int chunk = b0 & ~31;
	uint32_t n = 0, n_nx = 0;
	uint4 ea, eb, ec, ed, ea_nx, eb_nx, ec_nx, ed_nx;
	ea = eb = ec = ed = ea_nx = eb_nx = ec_nx = ed_nx = make_uint4(0, 0, 0, 0);
	uint4 row_nx[kWalkNT];
	if (tab) {
		fetch_row(chunk, row_nx);
		stage_row(row_nx);
	} else {
		fetch(chunk, n, ea, eb, ec, ed);
	}
	int pos = b0;
	while (pos < t_end && !stop) {
		// the next chunk's inputs are in flight while this chunk is walked
		if (tab) fetch_row(chunk + 32, row_nx);
		else fetch(chunk + 32, n_nx, ea_nx, eb_nx, ec_nx, ed_nx);
		bool have_ev = !tab;
		const int jend = min(32, t_end - chunk);
		const size_t g = (size_t)job.dec_off + chunk + lane;
		while (pos < chunk + jend && !stop) {
			const int j0 = pos - chunk;
			if (tab && (unsigned)((thresh - tab_base) >> 1) < (unsigned)kWalkNT && !(mode == 1 && thresh < thresh_lo)) {
				// ---- from the table: every lane looks its block up under the threshold in force (a step of the threshold
				// ends the pass, the next one costs a look-up again); on the serial chain are two instructions per block
				WP(wp_a = clock64(); wp_fast++;)
				const int k = (thresh - tab_base) >> 1;
				const bool active = lane >= j0 && lane < jend;
				const uint2 e = active ? *reinterpret_cast<const uint2 *>(&s_tab[lane][k]) : make_uint2(0u, 0u);
				const bool has = (e.y >> 31) != 0;
				const int f = (int)(e.x & 0xffff), l = (int)(e.x >> 16);
				const int out = has ? max(l + t_max - kBlockDec, 0) : 0;
				int c_in = __shfl_up_sync(0xffffffffu, out, 1);
				if (lane == j0) c_in = c;
				const int trig = (int)(e.y & 0xffff) + (has ? min(c_in, f) : c_in);
				s_t[lane] = trig;
				__syncwarp();
				const int used = thresh;
				int acc_end = jend;
				for (int j = j0; j < jend;) {
					// up to the next threshold decision (every 4th block, fm_demod.cpp:58-73)
					const int nstep = min((mode == 1) ? 4 - (runs & 3) : 4, jend - j);
					const int t0 = s_t[j], t1 = s_t[min(j + 1, 31)], t2 = s_t[min(j + 2, 31)], t3 = s_t[min(j + 3, 31)];
					avg = (int)((uint32_t)(31 * avg + t0) >> 5);   // (both are >= 0: the same as the reference's signed /32)
					s_avg[j] = avg;
					if (nstep > 1) { avg = (int)((uint32_t)(31 * avg + t1) >> 5); s_avg[j + 1] = avg; }
					if (nstep > 2) { avg = (int)((uint32_t)(31 * avg + t2) >> 5); s_avg[j + 2] = avg; }
					if (nstep > 3) { avg = (int)((uint32_t)(31 * avg + t3) >> 5); s_avg[j + 3] = avg; }
					runs += nstep;
					j += nstep;
					if (mode == 1 && (runs & 3) == 0) {
						if (avg >= kIdxPerBlock / 32) thresh += 2;
						else if (avg <= kIdxPerBlock / 64 && thresh > 50) thresh -= 2;
						if (thresh != used) { acc_end = j; break; }
					}
				}
				__syncwarp();
				const bool accepted = lane >= j0 && lane < acc_end;
				if (accepted) {
					BlockTrace bt = { used, trig, s_avg[lane] };
					p.trace[g] = bt;
					act_lane += (unsigned long long)trig;
				}
				c = __shfl_sync(0xffffffffu, out, acc_end - 1);
				WP(wp_chain += clock64() - wp_a; wp_a = clock64();)
				// ---- window lists.  Per block: its first trigger, the gaps that can open a window, its last trigger, as
				// items in shared memory (T: a trigger, S: only moves last_trig); then lanes = items, demodulator by
				// demodulator: which items open a window is a ballot, the windows' running sample count a scan
				const bool acc_has = accepted && has;
				unsigned wm = __ballot_sync(0xffffffffu, acc_has);
				const int ngl = acc_has ? (int)((e.y >> 16) & 3u) : 0;
				if (wm) {
					// a block with more gaps than the table holds feeds all its triggers one by one, in its place
					unsigned cm = __ballot_sync(0xffffffffu, ngl == 3);
					const int mine = (acc_has && ngl < 3) ? 1 + 2 * ngl + (l != f ? 1 : 0) : 0;
					int incl = mine;
#pragma unroll
					for (int d = 1; d < 32; d <<= 1) {
						const int o = __shfl_up_sync(0xffffffffu, incl, d);
						if (lane >= d) incl += o;
					}
					const int total = __shfl_sync(0xffffffffu, incl, 31);
					const uint32_t base = (uint32_t)(chunk + lane) * kBlockDec;
					const int at0 = incl - mine;
					int at = at0;
					if (mine) {
						s_trig[at++] = base + (uint32_t)f;
						if (ngl > 0) {
							const uint32_t gz = s_tab[lane][k].z;
							s_trig[at++] = kWalkSet | (base + (gz >> 16));
							s_trig[at++] = base + (gz & 0xffff);
						}
						if (ngl > 1) {
							const uint32_t gw = s_tab[lane][k].w;
							s_trig[at++] = kWalkSet | (base + (gw >> 16));
							s_trig[at++] = base + (gw & 0xffff);
						}
						if (l != f) s_trig[at++] = kWalkSet | (base + (uint32_t)l);
					}
					__syncwarp();
					int cur = 0;
					for (;;) {
						const int jc = cm ? __ffs(cm) - 1 : 0;
						const int seg_end = cm ? __shfl_sync(0xffffffffu, at0, jc) : total;
						for (; cur < seg_end; cur += 32) {
							const int cntb = min(32, seg_end - cur);
							const bool valid = lane < cntb;
							const uint32_t v = valid ? s_trig[cur + lane] : 0u;
							const bool is_t = valid && !(v & kWalkSet);
							const uint32_t t = v & ~kWalkSet;
							const uint32_t pv = __shfl_up_sync(0xffffffffu, t, 1);
							const long long prev = (lane == 0) ? last_trig : (long long)pv;
							const unsigned tm = __ballot_sync(0xffffffffu, is_t);
							for (int d = 0; d < nd; d++) {
								const int T_dd = __shfl_sync(0xffffffffu, T_d, d);
								const uint32_t open_d = __shfl_sync(0xffffffffu, open, d);
								unsigned m = __ballot_sync(0xffffffffu, is_t && ((long long)t - prev >= (long long)T_dd));
								if (!open_d) m |= tm & (0u - tm);
								if (!m) continue;
								const uint32_t nwin_d = __shfl_sync(0xffffffffu, n_win, d);
								const uint32_t cum_d = __shfl_sync(0xffffffffu, cum, d);
								const uint32_t ostart_d = __shfl_sync(0xffffffffu, open_start, d);
								const unsigned pm = m & ((1u << lane) - 1u);
								const bool opens = (m >> lane) & 1u;
								const uint32_t t_before = __shfl_sync(0xffffffffu, t, pm ? 31 - __clz(pm) : 0);
								const bool had_open = pm ? true : (open_d != 0);
								const uint32_t os_prev = pm ? t_before : ostart_d;
								const uint32_t end = (uint32_t)(prev + (long long)T_dd - 1);   // of the window this item's opening closes
								uint32_t len = (opens && had_open) ? end - os_prev + 1u : 0u;
#pragma unroll
								for (int q = 1; q < 32; q <<= 1) {
									const uint32_t o = __shfl_up_sync(0xffffffffu, len, q);
									if (lane >= q) len += o;
								}
								// a window opened here is closed by the next opening of the batch, if there is one (every lane writes
								// its own entry whole; only the batch's first opening touches an entry that was there before)
								const unsigned nm = m & ~((2u << lane) - 1u);
								const uint32_t end_next = __shfl_sync(0xffffffffu, end, nm ? __ffs(nm) - 1 : 0);
								WinEntry *wd = p.wins + job.win_off + (size_t)d * job.win_cap;
								if (opens) {
									const uint32_t idx = nwin_d + (uint32_t)__popc(pm);
									if (!pm && had_open && idx - 1u < job.win_cap) wd[idx - 1u].end = end;
									if (idx < job.win_cap) {
										WinEntry we = { t, nm ? end_next : 0xffffffffu, cum_d + len, 0u };
										wd[idx] = we;
									} else {
										p.counters->overflow = 1;
									}
								}
								const int last_o = 31 - __clz(m);
								const uint32_t cum_new = cum_d + __shfl_sync(0xffffffffu, len, last_o);
								const uint32_t os_new = __shfl_sync(0xffffffffu, t, last_o);
								if (lane == d) {
									const uint32_t n_new = nwin_d + (uint32_t)__popc(m);
									cum = cum_new;
									open = 1;
									if (n_new <= job.win_cap) {
										n_win = n_new;
										open_start = os_new;
									} else {
										// (the list is full: the call reports TFR_E_OVERFLOW, what follows is truncated)
										n_win = job.win_cap;
									}
								}
							}
							last_trig = (long long)__shfl_sync(0xffffffffu, t, cntb - 1);
						}
						cur = seg_end;
						if (!cm) break;
						cm &= cm - 1;
						WP(long long wp_b = clock64(); wp_ncomplex++;)
						feed_block(chunk, jc, used);
						WP(wp_feed += clock64() - wp_b;)
					}
					__syncwarp();
					WP(wp_nitems += total;)
				}
				WP(wp_items += clock64() - wp_a;)
				pos = chunk + acc_end;
				continue;
			}
			if (!have_ev) {
				fetch(chunk, n, ea, eb, ec, ed);
				have_ev = true;
			}
			WP(wp_slow++; wp_a = clock64();)
			const int theta = thresh;
			const bool active = lane >= j0 && lane < jend;
			// ---- 1. own block against theta
			int f = -1, l = -1, cov = 0, cnt = 0, cend = 0;
			uint32_t tp0 = 0, tp1 = 0, tp2 = 0, tp3 = 0;
			auto own = [&](uint32_t e) {
				if ((int)(e & 0xffff) > theta) {
					const int t = (int)(e >> 16);
					if (f < 0) f = t;
					const int hi = min(t + t_max, kBlockDec);
					cov += max(hi - max(t, cend), 0);
					cend = hi;
					l = t;
					if (cnt == 0) tp0 = t; else if (cnt == 1) tp1 = t; else if (cnt == 2) tp2 = t; else if (cnt == 3) tp3 = t;
					cnt++;
				}
			};
			if (active && n <= (uint32_t)kMaxEvt) {
				if (n > 0) own(ea.x);
				if (n > 1) own(ea.y);
				if (n > 2) own(ea.z);
				if (n > 3) own(ea.w);
				if (n > 4) own(eb.x);
				if (n > 5) own(eb.y);
				if (n > 6) own(eb.z);
				if (n > 7) own(eb.w);
				if (n > 8) {
					own(ec.x);
					if (n > 9) own(ec.y);
					if (n > 10) own(ec.z);
					if (n > 11) own(ec.w);
					if (n > 12) own(ed.x);
					if (n > 13) own(ed.y);
					if (n > 14) own(ed.z);
					if (n > 15) own(ed.w);
					const uint4 *ev4 = reinterpret_cast<const uint4 *>(p.events + g * kMaxEvt);
					for (uint32_t j = 16; j < n; j += 4) {
						const uint4 q = ev4[j >> 2];
						own(q.x);
						if (j + 1 < n) own(q.y);
						if (j + 2 < n) own(q.z);
						if (j + 3 < n) own(q.w);
					}
				}
			}
			unsigned dense = __ballot_sync(0xffffffffu, active && n > (uint32_t)kMaxEvt);
			while (dense) {
				// a burst block: every sample that can be a trigger is stored (it lies in a segment); scan the segments
				const int j = __ffs(dense) - 1;
				dense &= dense - 1;
				const size_t gj = (size_t)job.dec_off + chunk + j;
				const TileDesc &td = p.tiles[gj];
				const uint32_t *d = p.dec + gj * kBlockDec;
				const int ns = td.n_seg;
				int df = -1, dl = -1, dcov = 0, dcnt = 0, dend = 0;
				for (int sgi = 0; sgi < ns; sgi++) {
					const int a = td.seg_start[sgi], b = a + td.seg_len[sgi];
					for (int mb = a; mb < b; mb += 256) {
						// eight loads per lane in flight: this scan sits on the stream's serial chain, one L2 round
						// trip per 32 samples would cost ~0.1 ms per burst block
						uint32_t v[8];
#pragma unroll
						for (int q = 0; q < 8; q++) {
							const int m = mb + 32 * q + lane;
							v[q] = (m < b) ? d[m] : 0u;
						}
#pragma unroll
						for (int q = 0; q < 8; q++) {
							const int m0 = mb + 32 * q, m = m0 + lane;
							const unsigned mask = __ballot_sync(0xffffffffu, (m < b) && (pwr_of(v[q]) > theta));
							if (mask) {
								const int pf = m0 + __ffs(mask) - 1, pl = m0 + 31 - __clz(mask);
								if (df < 0) df = pf;
								const int hi = min(pl + t_max, kBlockDec);   // triggers of one chunk are < 32 < t_max apart
								dcov += max(hi - max(pf, dend), 0);
								dend = hi;
								dl = pl;
								dcnt += __popc(mask);
							}
						}
					}
				}
				if (lane == j) { f = df; l = dl; cov = dcov; cnt = dcnt ? dcnt + 1000 : 0; }   // +1000: positions not in tp0..3
			}
			// ---- 2. coverage carried in
			const bool has = cnt > 0;
			const int out = has ? max(l + t_max - kBlockDec, 0) : 0;
			int c_in = __shfl_up_sync(0xffffffffu, out, 1);
			if (lane == j0) c_in = c;
			const int trig = cov + (has ? min(c_in, f) : c_in);
			// ---- 3. the chain
			int acc_end = jend, my_used = 0, my_avg = 0;
			for (int j = j0; j < jend; j++) {
				if (mode == 1 && thresh < thresh_lo) {   // the front-end's bound does not cover this block: hand back
					stop = true;
					acc_end = j;
					break;
				}
				const int tj = __shfl_sync(0xffffffffu, trig, j);
				runs++;
				const int used = thresh;
				avg = (31 * avg + tj) / 32;
				if (lane == j) { my_used = used; my_avg = avg; }
				if (mode == 1 && (runs & 3) == 0) {
					if (avg >= kIdxPerBlock / 32) thresh += 2;
					else if (avg <= kIdxPerBlock / 64 && thresh > 50) thresh -= 2;
					if (thresh != used) { acc_end = j + 1; break; }
				}
			}
			const bool accepted = lane >= j0 && lane < acc_end;
			if (accepted) {
				BlockTrace bt = { my_used, trig, my_avg };
				p.trace[g] = bt;
				act_lane += (unsigned long long)trig;
			}
			if (acc_end > j0) c = __shfl_sync(0xffffffffu, out, acc_end - 1);
			// ---- 4. window lists: the accepted blocks' triggers in order
			unsigned wm = __ballot_sync(0xffffffffu, accepted && has);
			// The usual case - no accepted block holds more than four triggers - without a warp-wide broadcast per block:
			// every lane drops its triggers into an ordered list in shared memory (one scan), then the demodulator lanes run
			// the window bookkeeping over the list on their own.  (Per block the loop below costs five shuffles and a
			// divergent call; on the stream's serial chain every instruction waits for the one before it.)
			if (wm && !__any_sync(0xffffffffu, accepted && cnt > 4)) {
				const int mine = (accepted && has) ? cnt : 0;
				int incl = mine;
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					const int o = __shfl_up_sync(0xffffffffu, incl, d);
					if (lane >= d) incl += o;
				}
				const int total = __shfl_sync(0xffffffffu, incl, 31);
				const uint32_t base = (uint32_t)(chunk + lane) * kBlockDec;
				int at = incl - mine;
				if (mine > 0) s_trig[at] = base + tp0;
				if (mine > 1) s_trig[at + 1] = base + tp1;
				if (mine > 2) s_trig[at + 2] = base + tp2;
				if (mine > 3) s_trig[at + 3] = base + tp3;
				__syncwarp();
				if (lane < nd)
					for (int i = 0; i < total; i++) win_trigger(s_trig[i]);
				last_trig = s_trig[total - 1];
				__syncwarp();
				wm = 0;
			}
			while (wm) {
				const int j = __ffs(wm) - 1;
				wm &= wm - 1;
				const int cj = __shfl_sync(0xffffffffu, cnt, j);
				const uint32_t q0 = __shfl_sync(0xffffffffu, tp0, j), q1 = __shfl_sync(0xffffffffu, tp1, j);
				const uint32_t q2 = __shfl_sync(0xffffffffu, tp2, j), q3 = __shfl_sync(0xffffffffu, tp3, j);
				const uint32_t base = (uint32_t)(chunk + j) * kBlockDec;
				if (cj <= 4) {
					win_trigger(base + q0);
					if (cj > 1) win_trigger(base + q1);
					if (cj > 2) win_trigger(base + q2);
					if (cj > 3) win_trigger(base + q3);
				} else {
					feed_block(chunk, j, theta);
				}
			}
			WP(wp_old += clock64() - wp_a;)
			pos = chunk + acc_end;
		}
		if (stop) {
			t_stop = pos;
			break;
		}
		chunk += 32;
		if (tab) {
			WP(wp_a = clock64();)
			stage_row(row_nx);
			WP(wp_stage += clock64() - wp_a;)
		} else {
			n = n_nx;
			ea = ea_nx;
			eb = eb_nx;
			ec = ec_nx;
			ed = ed_nx;
		}
	}

#ifdef TFR_WALK_PROFILE
	if (stream == 0 && lane == 0)
		printf("walk: blocks %d total %lld chain %lld items %lld (feed %lld) stage %lld old %lld | passes fast %d slow %d items %d complex %d\n",
		       t_end - b0, clock64() - wp_t0, wp_chain, wp_items, wp_feed, wp_stage, wp_old, wp_fast, wp_slow, wp_nitems, wp_ncomplex);
#endif
	const bool finished = (t_stop == (int)job.n_blocks);
	if (finished && lane < nd) {
		uint32_t cont = 0;
		if (open) {
			const long long end = last_trig + T_d - 1;   // >= 0 because an open window means last_trig > -T_d
			wl[n_win - 1].end = (uint32_t)end;
			cum += (uint32_t)end - open_start + 1;
			cont = (end >= (long long)call_len) ? 1u : 0u;   // still running when the data ends: the next call resumes it
		}
		st->win_cont[lane] = cont;
		p.wincnt[stream].n[lane] = n_win;
		p.wincnt[stream].cum[lane] = cum;
		// the leading windows that need the previous call's final state (kLateRow): window 0, and for the filtered
		// demodulators every window whose warm-up history (TFR_WARM_X4 quarter timeouts of earlier windows' samples) would
		// reach window 0, up to the next chain head
		uint32_t W = 0;
		if (n_win) {
			W = 1;
			if (cfg.d[lane].kind != K_TFA1) {
				const uint32_t want = ((uint32_t)TFR_WARM_X4 * (uint32_t)T_d) / 4u;
				const uint32_t c1 = (n_win > 1) ? wl[1].cum : 0u;   // samples of window 0
				while (W < n_win && !(wl[W].cum - c1 >= want && chain_head(wl, W, T_d))) W++;
			}
		}
		p.partcnt[(size_t)kLateRow * p.n_streams + stream].n[lane] = W;
	}
	if (lane < nd) {
		st->win_n[lane] = n_win;
		st->win_cum[lane] = cum;
		st->win_open[lane] = open;
		if (part) {
			// the windows a back-end part may take now: the leading ones that are closed and end before the blocks walked
			// so far, cut at a chain head.  A window that is still to come starts at or after `bound`; it can only be near
			// (inside the chain of) the last listed window if that one ended within T_d+1 samples of bound.
			uint32_t W = n_win;
			if (!finished) {
				const uint32_t bound = (uint32_t)t_stop * (uint32_t)kBlockDec;
				while (W > 0 && wl[W - 1].end >= bound) W--;
				if (cfg.d[lane].kind != K_TFA1) {
					if (W == n_win && W > 0 && bound - wl[W - 1].end <= (uint32_t)T_d + 1u) W--;
					while (W > 0 && W < n_win && !chain_head(wl, W, T_d)) W--;
				}
			}
			part->n[lane] = W;
		}
	}
	for (int o = 16; o; o >>= 1) act_lane += __shfl_xor_sync(0xffffffffu, act_lane, o);
	if (lane == 0) {
		st->thresh = thresh;
		st->triggered_avg = avg;
		st->runs = runs;
		st->any_timeout = c;
		st->call_last_trig = (int32_t)max(last_trig, (long long)INT32_MIN / 2);
		if (finished) {
			const long long age = (long long)call_len - last_trig;
			st->trig_age = (int32_t)min(age, (long long)INT32_MAX / 2);
		}
		st->t2_done = (uint32_t)t_stop;   // == n_blocks when finished; submit_epilogue_kernel clears it for the next call
		if (p.progress) p.progress[stream] = (uint32_t)t_stop;
		if (act_lane) atomicAdd(&p.counters->active_samples, act_lane);
	}
}

// ------------------------------------------------------------------------------------------------
// walk_cta_kernel: the walk from the table (walk_table_kernel) with one CTA of eight warps per stream.
//
// thresh2_kernel is one warp, and a lone warp pays a pipeline latency for every instruction: even from the table it
// spends ~250 cycles per block.  Here a super-chunk of 256 blocks is taken at once, thread = block:
//   A  all threads look their block up under the threshold in force and derive its triggered count (the coverage
//      carried in from the block before is one shared-memory read away)
//   B  warp 0 runs the chain  avg = (31 avg + triggered) / 32, threshold step every 4th block  (fm_demod.cpp:58-73)
//      over the counts until the threshold moves; the blocks up to there are accepted, the rest is looked up again
//   C  every accepted block turns its table entry into items - T: a trigger that may open a window, S: only moves
//      last_trig - packed in order by a block-wide scan (a block with more gaps than the table holds walks its own
//      event list)
//   D  warp d keeps demodulator d's window list: lanes = items, the openings are a ballot, the windows' running
//      sample counts a scan.
// Where the walk cannot go on from the table (the front-end's bound fails, or the threshold left the table's range)
// it stops like thresh2_kernel does; the next launch, or the host's epochs, carry on from there.
// ------------------------------------------------------------------------------------------------
// CT = threads = blocks of a super-chunk.  256 (a warp per demodulator in D) is what runs; 64 (two warps, <= 176 registers)
// would fit beside the two front-end CTAs an SM holds, which leave 11.7 k registers and 40 KB of shared memory, but measured
// slower at 8, 16 and 32 streams (TFR_WALK_CT=64 selects it; profiles/r2_walk_table.txt).
struct WalkDem { uint32_t n_win, cum, open, open_start; };

template <int CT>
__global__ void __launch_bounds__(CT, 1) walk_cta_kernel(const BackParams p)
{
	constexpr int kWalkCta = CT, kWalkItems = CT * 8;
	const int stream = blockIdx.x;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	__shared__ uint2 s_tab[kWalkCta][kWalkNT + 1];   // (padded: conflict-free 8-byte accesses along a column)
	__shared__ int s_t[kWalkCta], s_out[kWalkCta], s_avg[kWalkCta];
	__shared__ uint32_t s_items[kWalkItems];
	__shared__ int s_scan[kWalkCta / 32];
	__shared__ int s_ctl[8];
	__shared__ WalkDem s_dem[kMaxDemods];
	if (stream >= p.n_streams) return;
	const StreamJob job = p.jobs[stream];
	const DevConfig &cfg = *p.cfg;
	const int nd = cfg.n_demods;
	WinCount *part = (p.part_idx >= 0) ? p.partcnt + (size_t)p.part_idx * p.n_streams + stream : nullptr;
	if (job.n_blocks == 0) {
		if (part && tid < nd) part->n[tid] = 0;
		return;
	}
	StreamState *st = p.st + stream;
	const int b0 = (int)st->t2_done;
	if (b0 >= (int)job.n_blocks) {   // this stream finished in an earlier launch of the call: every window is final
		if (part && tid < nd) part->n[tid] = p.wincnt[stream].n[tid];
		return;
	}
	const int t_max = cfg.t_max;
	const uint32_t call_len = job.n_blocks * (uint32_t)kBlockDec;
	int t_min = 0x7fffffff;
	for (int d = 0; d < nd; d++) t_min = min(t_min, cfg.d[d].timeout);

	int thresh = st->thresh, avg = st->triggered_avg, runs = st->runs;
	int c = st->any_timeout;
	const int mode = st->thresh_mode;
	const int thresh_lo = (mode == 1) ? (p.margin ? thresh - p.margin : st->spec_lo) : thresh;
	long long last_trig;
	// the carried window state, by thread d as in thresh2_kernel; warp d works on it below
	const int T_d = (tid < nd) ? cfg.d[tid].timeout : 0x7fffffff;
	uint32_t n_win = 0, cum = 0, open = 0, open_start = 0;
	WinEntry *wl = (tid < nd) ? p.wins + job.win_off + (size_t)tid * job.win_cap : nullptr;
	if (b0 == 0) {
		last_trig = -(long long)st->trig_age;
		if (tid < nd && st->win_cont[tid]) {   // a window was open when the previous call ended
			WinEntry e = { 0u, 0xffffffffu, 0u, kWinCont };
			wl[0] = e;
			n_win = 1;
			open = 1;
		}
	} else {
		last_trig = st->call_last_trig;
		if (tid < nd) {
			n_win = st->win_n[tid];
			cum = st->win_cum[tid];
			open = st->win_open[tid];
			if (open && n_win) open_start = wl[n_win - 1].start;
		}
	}
	if (tid < nd) {
		WalkDem wdm = { n_win, cum, open, open_start };
		s_dem[tid] = wdm;
	}
#ifdef TFR_WALK_PROFILE
	if (tid == 0) s_ctl[6] = s_ctl[7] = 0;
#endif
	__syncthreads();
	unsigned long long act = 0;
	const int t_end = min(b0 + p.n_tiles, (int)job.n_blocks);
	int t_stop = t_end;
	const int tab_base = p.walk_base[stream];
	const uint2 *chain_tab = reinterpret_cast<const uint2 *>(p.walk_tab);
	const uint4 *gap_tab = reinterpret_cast<const uint4 *>(p.walk_gap);
#ifdef TFR_WALK_PROFILE
	long long wp_t0 = clock64(), wp_ab = 0, wp_c = 0, wp_d = 0, wp_a;
	int wp_pass = 0, wp_nitems = 0, wp_ncomplex = 0;
#endif

	auto fetch_row = [&](int sc, uint4 (&row)[kWalkNT / 2]) {
		const int b = sc + tid;
		if (b >= b0 && b < t_end) {
			const uint4 *src = reinterpret_cast<const uint4 *>(chain_tab + ((size_t)job.dec_off + b) * kWalkNT);
#pragma unroll
			for (int k = 0; k < kWalkNT / 2; k++) row[k] = src[k];
		} else {
#pragma unroll
			for (int k = 0; k < kWalkNT / 2; k++) row[k] = make_uint4(0, 0, 0, 0);
		}
	};
	auto stage_row = [&](const uint4 (&row)[kWalkNT / 2]) {
		__syncthreads();
#pragma unroll
		for (int k = 0; k < kWalkNT / 2; k++) {
			s_tab[tid][2 * k] = make_uint2(row[k].x, row[k].y);
			s_tab[tid][2 * k + 1] = make_uint2(row[k].z, row[k].w);
		}
		__syncthreads();
	};
	// a block's table entry under a threshold the table does not hold (the threshold drifted more than eight steps
	// inside one launch: start-up transients), from the block's own event list - what walk_table_kernel does, by one thread
	auto eval_block = [&](size_t gb, int theta_b, uint2 &e, uint4 &gp) {
		int f = -1, l = 0, cov = 0, cend = 0, ng = 0;
		gp = make_uint4(0, 0, 0, 0);
		auto upd = [&](int t) {
			if (f < 0) {
				f = t;
			} else if (t - l >= t_min) {
				const uint32_t q = ((uint32_t)l << 16) | (uint32_t)t;
				if (ng == 0) gp.x = q;
				else if (ng == 1) gp.y = q;
				else if (ng == 2) gp.z = q;
				else if (ng == 3) gp.w = q;
				ng = min(ng + 1, 5);
			}
			const int hi = min(t + t_max, kBlockDec);
			cov += max(hi - max(t, cend), 0);
			cend = hi;
			l = t;
		};
		const uint32_t nj = p.tiles[gb].n_trig;
		if (nj <= (uint32_t)kMaxEvt) {
			const uint32_t *ev = p.events + gb * kMaxEvt;
			for (uint32_t q = 0; q < nj; q++) {
				const uint32_t v = ev[q];
				if ((int)(v & 0xffff) > theta_b) upd((int)(v >> 16));
			}
		} else {
			const TileDesc &td = p.tiles[gb];
			const uint32_t *d = p.dec + gb * kBlockDec;
			for (int sgi = 0; sgi < (int)td.n_seg; sgi++) {
				const int a = td.seg_start[sgi], b = a + td.seg_len[sgi];
				for (int m = a; m < b; m++)
					if (pwr_of(d[m]) > theta_b) upd(m);
			}
		}
		const bool has = f >= 0;
		e.x = has ? ((uint32_t)f | ((uint32_t)l << 16)) : 0u;
		e.y = (uint32_t)cov | ((uint32_t)ng << 16) | (has ? 0x80000000u : 0u);
	};
	// the same for a burst block (no complete list: its stored samples are scanned), by all lanes of a warp together
	auto eval_burst_warp = [&](size_t gb, int theta_b, uint2 &e, uint4 &gp) {
		int f = -1, l = 0, cov = 0, cend = 0, ng = 0;
		gp = make_uint4(0, 0, 0, 0);
		const TileDesc &td = p.tiles[gb];
		const uint32_t *d = p.dec + gb * kBlockDec;
		const int ns = td.n_seg;
		for (int sgi = 0; sgi < ns; sgi++) {
			const int a = td.seg_start[sgi], b = a + td.seg_len[sgi];
			for (int mb = a; mb < b; mb += 256) {
				uint32_t v[8];
#pragma unroll
				for (int q = 0; q < 8; q++) {
					const int m = mb + 32 * q + lane;
					v[q] = (m < b) ? d[m] : 0u;
				}
#pragma unroll
				for (int q = 0; q < 8; q++) {
					const int m0 = mb + 32 * q, m = m0 + lane;
					const unsigned mask = __ballot_sync(0xffffffffu, (m < b) && (pwr_of(v[q]) > theta_b));
					if (mask) {
						const int pf = m0 + __ffs(mask) - 1, pl = m0 + 31 - __clz(mask);   // (< 32 apart: no gap in between)
						if (f < 0) {
							f = pf;
						} else if (pf - l >= t_min) {
							const uint32_t qg = ((uint32_t)l << 16) | (uint32_t)pf;
							if (ng == 0) gp.x = qg;
							else if (ng == 1) gp.y = qg;
							else if (ng == 2) gp.z = qg;
							else if (ng == 3) gp.w = qg;
							ng = min(ng + 1, 5);
						}
						const int hi = min(pl + t_max, kBlockDec);
						cov += max(hi - max(pf, cend), 0);
						cend = hi;
						l = pl;
					}
				}
			}
		}
		const bool has = f >= 0;
		e.x = has ? ((uint32_t)f | ((uint32_t)l << 16)) : 0u;
		e.y = (uint32_t)cov | ((uint32_t)ng << 16) | (has ? 0x80000000u : 0u);
	};
	// the items of one block from its own triggers (a block with more gaps than its table entry holds): count, or write
	auto gen_items = [&](size_t gb, int theta_b, uint32_t base, uint32_t *dst) -> int {
		int cnt = 0, first = -1, prev = -1;
		auto trig = [&](int t) {
			if (first < 0) {
				first = t;
				if (dst) dst[cnt] = base + (uint32_t)t;
				cnt++;
			} else if (t - prev >= t_min) {
				if (dst) {
					dst[cnt] = kWalkSet | (base + (uint32_t)prev);
					dst[cnt + 1] = base + (uint32_t)t;
				}
				cnt += 2;
			}
			prev = t;
		};
		const uint32_t nj = p.tiles[gb].n_trig;
		if (nj <= (uint32_t)kMaxEvt) {
			const uint32_t *ev = p.events + gb * kMaxEvt;
			for (uint32_t q = 0; q < nj; q++) {
				const uint32_t e = ev[q];
				if ((int)(e & 0xffff) > theta_b) trig((int)(e >> 16));
			}
		} else {
			const TileDesc &td = p.tiles[gb];
			const uint32_t *d = p.dec + gb * kBlockDec;
			for (int sgi = 0; sgi < (int)td.n_seg; sgi++) {
				const int a = td.seg_start[sgi], b = a + td.seg_len[sgi];
				for (int m = a; m < b; m++)
					if (pwr_of(d[m]) > theta_b) trig(m);
			}
		}
		if (prev != first) {
			if (dst) dst[cnt] = kWalkSet | (base + (uint32_t)prev);
			cnt++;
		}
		return cnt;
	};

	bool stop = false;
	int sc = b0;
	uint4 row[kWalkNT / 2];
	fetch_row(sc, row);
	stage_row(row);
	int pos = b0;
	while (pos < t_end && !stop) {
		fetch_row(sc + kWalkCta, row);   // in flight while this super-chunk is walked
		const int jend = min(kWalkCta, t_end - sc);
		const size_t g = (size_t)job.dec_off + sc + tid;
		uint32_t my_fl = 0, my_cv = 0;
		int my_k = 0, my_used = 0;
		bool mine_acc = false, my_own = false;   // my_own: the entry came from eval_block, its gaps are in my_gaps
		uint4 my_gaps = make_uint4(0, 0, 0, 0);
#ifdef TFR_WALK_PROFILE
		wp_a = clock64();
#endif
		// ---- A, B
		while (pos < sc + jend) {
			const int k = (thresh - tab_base) >> 1;
			if (mode == 1 && thresh < thresh_lo) {   // the front-end's bound does not cover this block: hand back
				stop = true;
				break;
			}
			const bool in_tab = (unsigned)k < (unsigned)kWalkNT;
			const int j0 = pos - sc;
			const bool active = tid >= j0 && tid < jend;
			uint2 e = make_uint2(0u, 0u);
			uint4 e_gaps = make_uint4(0, 0, 0, 0);
			const int kq = min(max(k, 0), kWalkNT - 1);
			bool own = false;
			if (active) {
				e = s_tab[tid][kq];
				// off the table: the nearest column holds if its entry is proven constant out to here (walk_table_kernel)
				if (k < 0) {
					own = !((e.y >> 30) & 1u);
				} else if (!in_tab) {
					const int run = (int)((e.y >> 26) & 15u);
					own = !(run > 0 && thresh <= tab_base + 2 * (kWalkNT - 1) + kWalkLadder[max(run - 1, 0)]);
				}
			}
			if (!in_tab && p.walk_dbg == 32 && active && !own && p.tiles[g].n_trig <= (uint32_t)kMaxEvt) {
				uint2 e2;
				uint4 g2;
				eval_block(g, thresh, e2, g2);
				if (e2.x != e.x || ((e2.y ^ e.y) & 0x8007ffffu))
					printf("MISMATCH stream %d block %d thresh %d base %d k %d n %u tab %08x %08x eval %08x %08x\n", stream, sc + tid, thresh,
					       tab_base, k, p.tiles[g].n_trig, e.x, e.y, e2.x, e2.y);
			}
			if (!in_tab) {
				// ... else the block is evaluated here: a list by its thread, a burst block by the warp
				const bool burst = own && p.tiles[g].n_trig > (uint32_t)kMaxEvt;
				if (own && !burst) eval_block(g, thresh, e, e_gaps);
				unsigned bm = __ballot_sync(0xffffffffu, burst);
				while (bm) {
					const int jb = __ffs(bm) - 1;
					bm &= bm - 1;
					uint2 eb;
					uint4 gb4;
					eval_burst_warp((size_t)job.dec_off + sc + 32 * warp + jb, thresh, eb, gb4);
					if (lane == jb) {
						e = eb;
						e_gaps = gb4;
					}
				}
#ifdef TFR_WALK_PROFILE
				if (own) atomicAdd(&s_ctl[6], 1);
				if (burst) atomicAdd(&s_ctl[7], 1);
#endif
			}
			const bool has = (e.y >> 31) != 0;
			const int f = (int)(e.x & 0xffff), l = (int)(e.x >> 16);
			const int out = has ? max(l + t_max - kBlockDec, 0) : 0;
			s_out[tid] = out;
			__syncthreads();
			const int c_in = (tid == j0) ? c : s_out[max(tid - 1, 0)];
			const int trig = (int)(e.y & 0xffff) + (has ? min(c_in, f) : c_in);
			s_t[tid] = trig;
			__syncthreads();
			const int used = thresh;
			if (warp == 0) {
				// the averages as if the threshold stayed (two instructions per block on the chain, nothing else) ...
				// (eight counts are loaded, eight steps taken, eight averages stored: a load behind the store of the step
				// before would sit on the chain)
				// ... in stretches of 32 blocks; the first decision (every 4th block, fm_demod.cpp:58-73) that moves the
				// threshold ends the pass
				int a = avg;
				int first = kWalkCta;
				for (int seg = j0; seg < jend && first == kWalkCta; seg += 32) {
					const int send = min(seg + 32, jend);
					int tn[8];
#pragma unroll
					for (int q = 0; q < 8; q++) tn[q] = s_t[min(seg + q, kWalkCta - 1)];
					for (int j = seg; j < send; j += 8) {
						int tc[8], r[8];
#pragma unroll
						for (int q = 0; q < 8; q++) tc[q] = tn[q];
#pragma unroll
						for (int q = 0; q < 8; q++) tn[q] = s_t[min(j + 8 + q, kWalkCta - 1)];   // the next eight, off the chain
#pragma unroll
						for (int q = 0; q < 8; q++) {
							if (j + q < send) a = (int)((uint32_t)(31 * a + tc[q]) >> 5);   // (both >= 0: the reference's signed /32)
							r[q] = a;
						}
#pragma unroll
						for (int q = 0; q < 8; q++)
							if (j + q < send) s_avg[j + q] = r[q];
					}
					__syncwarp();
					if (mode == 1) {
						const int j = seg + lane;
						bool moves = false;
						if (j < send && ((runs + (j - j0 + 1)) & 3) == 0) {
							const int aj = s_avg[j];
							moves = aj >= kIdxPerBlock / 32 || (aj <= kIdxPerBlock / 64 && thresh > 50);
						}
						const unsigned mv = __ballot_sync(0xffffffffu, moves);
						if (mv) first = seg + __ffs(mv) - 1;
					}
				}
				int acc_end = jend;
				if (first < kWalkCta) {
					acc_end = first + 1;
					thresh += (s_avg[first] >= kIdxPerBlock / 32) ? 2 : -2;
				}
				if (lane == 0) {
					s_ctl[0] = thresh;
					s_ctl[1] = s_avg[acc_end - 1];
					s_ctl[2] = runs + (acc_end - j0);
					s_ctl[3] = acc_end;
				}
			}
			__syncthreads();
			thresh = s_ctl[0];
			avg = s_ctl[1];
			runs = s_ctl[2];
			const int acc_end = s_ctl[3];
			if (tid >= j0 && tid < acc_end) {
				BlockTrace bt = { used, trig, s_avg[tid] };
				p.trace[g] = bt;
				act += (unsigned long long)trig;
				mine_acc = true;
				my_fl = e.x;
				my_cv = e.y;
				my_k = kq;
				my_used = used;
				my_own = own;
				my_gaps = e_gaps;
				const int ng_e = (int)((e.y >> 16) & 7u);
				if (has && !own && ng_e > 0 && ng_e <= 4) my_gaps = gap_tab[g * kWalkNT + kq];   // (in flight until the items are made)
			}
			c = s_out[acc_end - 1];
			pos = sc + acc_end;
			__syncthreads();   // (the next pass writes s_out, s_t and s_ctl again)
#ifdef TFR_WALK_PROFILE
			wp_pass++;
#endif
		}
#ifdef TFR_WALK_PROFILE
		wp_ab += clock64() - wp_a; wp_a = clock64();
#endif
		// ---- C: the accepted blocks' items, in order
		const bool acc_has = mine_acc && (my_cv >> 31);
		const int ngl = acc_has ? (int)((my_cv >> 16) & 7u) : 0;
		const uint32_t f_l = my_fl & 0xffff, l_l = my_fl >> 16;
		const uint32_t base = (uint32_t)(sc + tid) * kBlockDec;
		int cnt = 0;
		if (acc_has) cnt = (ngl <= 4) ? 1 + 2 * ngl + (l_l != f_l ? 1 : 0) : gen_items(g, my_used, base, nullptr);
		int incl = cnt;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const int o = __shfl_up_sync(0xffffffffu, incl, d);
			if (lane >= d) incl += o;
		}
		if (lane == 31) s_scan[warp] = incl;
		__syncthreads();
		int total = 0;
#pragma unroll
		for (int w = 0; w < kWalkCta / 32; w++) {
			const int sw = s_scan[w];
			if (w < warp) incl += sw;
			total += sw;
		}
		const uint4 gaps = my_gaps;
		int done = 0;
		while (done < total) {   // (more than once only if the items of 256 blocks do not fit s_items)
			if (tid == 0) s_ctl[4] = done;
			__syncthreads();
			const bool in = cnt > 0 && incl - cnt >= done && incl <= done + kWalkItems;
			if (in) {
				atomicMax(&s_ctl[4], incl);
				uint32_t *dst = s_items + (incl - cnt - done);
				if (ngl <= 4) {
					int at = 0;
					dst[at++] = base + f_l;
					if (ngl > 0) { dst[at++] = kWalkSet | (base + (gaps.x >> 16)); dst[at++] = base + (gaps.x & 0xffff); }
					if (ngl > 1) { dst[at++] = kWalkSet | (base + (gaps.y >> 16)); dst[at++] = base + (gaps.y & 0xffff); }
					if (ngl > 2) { dst[at++] = kWalkSet | (base + (gaps.z >> 16)); dst[at++] = base + (gaps.z & 0xffff); }
					if (ngl > 3) { dst[at++] = kWalkSet | (base + (gaps.w >> 16)); dst[at++] = base + (gaps.w & 0xffff); }
					if (l_l != f_l) dst[at++] = kWalkSet | (base + l_l);
				} else {
					gen_items(g, my_used, base, dst);
#ifdef TFR_WALK_PROFILE
					wp_ncomplex++;
#endif
				}
			}
			__syncthreads();
			const int n = s_ctl[4] - done;
			if (n == 0) break;   // (cannot happen: a block holds at most 48 items)
#ifdef TFR_WALK_PROFILE
			wp_c += clock64() - wp_a; wp_a = clock64();
			wp_nitems += n;
#endif
			// ---- D: the demodulators over the warps
			for (int dd = warp; dd < nd; dd += kWalkCta / 32) {
				const int T_w = cfg.d[dd].timeout;
				WinEntry *w_wl = p.wins + job.win_off + (size_t)dd * job.win_cap;
				const WalkDem wdm = s_dem[dd];
				uint32_t w_n = wdm.n_win, w_cum = wdm.cum, w_open = wdm.open, w_os = wdm.open_start;
				long long lt = last_trig;
				for (int cur = 0; cur < n; cur += 32) {
					const int cntb = min(32, n - cur);
					const bool valid = lane < cntb;
					const uint32_t v = valid ? s_items[cur + lane] : 0u;
					const bool is_t = valid && !(v & kWalkSet);
					const uint32_t t = v & ~kWalkSet;
					const uint32_t pv = __shfl_up_sync(0xffffffffu, t, 1);
					const long long prev = (lane == 0) ? lt : (long long)pv;
					const unsigned tm = __ballot_sync(0xffffffffu, is_t);
					unsigned m = __ballot_sync(0xffffffffu, is_t && ((long long)t - prev >= (long long)T_w));
					if (!w_open) m |= tm & (0u - tm);
					if (m) {
						const unsigned pm = m & ((1u << lane) - 1u);
						const bool opens = (m >> lane) & 1u;
						const uint32_t t_before = __shfl_sync(0xffffffffu, t, pm ? 31 - __clz(pm) : 0);
						const bool had_open = pm ? true : (w_open != 0);
						const uint32_t os_prev = pm ? t_before : w_os;
						const uint32_t end = (uint32_t)(prev + (long long)T_w - 1);   // of the window this item's opening closes
						uint32_t len = (opens && had_open) ? end - os_prev + 1u : 0u;
#pragma unroll
						for (int q = 1; q < 32; q <<= 1) {
							const uint32_t o = __shfl_up_sync(0xffffffffu, len, q);
							if (lane >= q) len += o;
						}
						// a window opened here is closed by the next opening of the batch, if there is one (every lane writes
						// its own entry whole; only the batch's first opening touches an entry that was there before)
						const unsigned nm = m & ~((2u << lane) - 1u);
						const uint32_t end_next = __shfl_sync(0xffffffffu, end, nm ? __ffs(nm) - 1 : 0);
						if (opens) {
							const uint32_t idx = w_n + (uint32_t)__popc(pm);
							if (!pm && had_open && idx - 1u < job.win_cap) w_wl[idx - 1u].end = end;
							if (idx < job.win_cap) {
								WinEntry we = { t, nm ? end_next : 0xffffffffu, w_cum + len, 0u };
								w_wl[idx] = we;
							} else {
								p.counters->overflow = 1;
							}
						}
						const int last_o = 31 - __clz(m);
						w_cum += __shfl_sync(0xffffffffu, len, last_o);
						const uint32_t os_new = __shfl_sync(0xffffffffu, t, last_o);
						const uint32_t n_new = w_n + (uint32_t)__popc(m);
						w_open = 1;
						if (n_new <= job.win_cap) {
							w_n = n_new;
							w_os = os_new;
						} else {
							w_n = job.win_cap;   // (the list is full: the call reports TFR_E_OVERFLOW, what follows is truncated)
						}
					}
					lt = (long long)__shfl_sync(0xffffffffu, t, cntb - 1);
				}
				if (lane == 0) {
					WalkDem o = { w_n, w_cum, w_open, w_os };
					s_dem[dd] = o;
				}
			}
			if (n) last_trig = (long long)(s_items[n - 1] & ~kWalkSet);
			done += n;
			__syncthreads();   // (the next round writes s_items and s_ctl again)
#ifdef TFR_WALK_PROFILE
			wp_d += clock64() - wp_a; wp_a = clock64();
#endif
		}
#ifdef TFR_WALK_PROFILE
		wp_c += clock64() - wp_a;
#endif
		if (stop) {
			t_stop = pos;
			break;
		}
		sc += kWalkCta;
		stage_row(row);
	}
#ifdef TFR_WALK_PROFILE
	if (tid == 0 && (stream == 0 || clock64() - wp_t0 > 150000))
		printf("walk_cta[%d]: blocks %d total %lld A+B %lld C %lld D %lld | passes %d items %d complex(t0) %d evals %d (burst blocks %d)\n", stream,
		       t_end - b0, clock64() - wp_t0, wp_ab, wp_c, wp_d, wp_pass, wp_nitems, wp_ncomplex, s_ctl[6], s_ctl[7]);
#endif

	// ---- the state goes back to thread d; from here on as in thresh2_kernel
	__syncthreads();
	if (tid < nd) {
		const WalkDem wdm = s_dem[tid];
		n_win = wdm.n_win;
		cum = wdm.cum;
		open = wdm.open;
		open_start = wdm.open_start;
	}
	const bool finished = (t_stop == (int)job.n_blocks);
	if (finished && tid < nd) {
		uint32_t cont = 0;
		if (open) {
			const long long end = last_trig + T_d - 1;   // >= 0 because an open window means last_trig > -T_d
			wl[n_win - 1].end = (uint32_t)end;
			cum += (uint32_t)end - open_start + 1;
			cont = (end >= (long long)call_len) ? 1u : 0u;   // still running when the data ends: the next call resumes it
		}
		st->win_cont[tid] = cont;
		p.wincnt[stream].n[tid] = n_win;
		p.wincnt[stream].cum[tid] = cum;
		uint32_t W = 0;   // the leading windows that need the previous call's final state (see thresh2_kernel)
		if (n_win) {
			W = 1;
			if (cfg.d[tid].kind != K_TFA1) {
				const uint32_t want = ((uint32_t)TFR_WARM_X4 * (uint32_t)T_d) / 4u;
				const uint32_t c1 = (n_win > 1) ? wl[1].cum : 0u;
				while (W < n_win && !(wl[W].cum - c1 >= want && chain_head(wl, W, T_d))) W++;
			}
		}
		p.partcnt[(size_t)kLateRow * p.n_streams + stream].n[tid] = W;
	}
	if (tid < nd) {
		st->win_n[tid] = n_win;
		st->win_cum[tid] = cum;
		st->win_open[tid] = open;
		if (part) {
			uint32_t W = n_win;
			if (!finished) {
				const uint32_t bound = (uint32_t)t_stop * (uint32_t)kBlockDec;
				while (W > 0 && wl[W - 1].end >= bound) W--;
				if (cfg.d[tid].kind != K_TFA1) {
					if (W == n_win && W > 0 && bound - wl[W - 1].end <= (uint32_t)T_d + 1u) W--;
					while (W > 0 && W < n_win && !chain_head(wl, W, T_d)) W--;
				}
			}
			part->n[tid] = W;
		}
	}
	for (int o = 16; o; o >>= 1) act += __shfl_xor_sync(0xffffffffu, act, o);
	if (lane == 0 && act) atomicAdd(&p.counters->active_samples, act);
	if (tid == 0) {
		st->thresh = thresh;
		st->triggered_avg = avg;
		st->runs = runs;
		st->any_timeout = c;
		st->call_last_trig = (int32_t)max(last_trig, (long long)INT32_MIN / 2);
		if (finished) {
			const long long age = (long long)call_len - last_trig;
			st->trig_age = (int32_t)min(age, (long long)INT32_MAX / 2);
		}
		st->t2_done = (uint32_t)t_stop;
		if (p.progress) p.progress[stream] = (uint32_t)t_stop;
	}
}

// ------------------------------------------------------------------------------------------------
// devfm_kernel: fm_dev for every stored sample of a block (one CTA per block)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) devfm_kernel(const BackParams p)
{
	// One CTA per block; the kernel is a chain of dependent loads (job -> descriptor -> samples), so what matters is
	// how many of them are in flight: every thread reads the (L1-resident after the first touch) segment table itself
	// instead of waiting at a barrier for a serially built region list, and strides the kept samples by the CTA width.
	const int stream = blockIdx.y;
	const StreamJob job = p.jobs[stream];
	const int tile = p.tile0 + blockIdx.x;   // a back-end part covers the blocks [tile0, tile0 + n_tiles)
	if (tile >= (int)job.n_blocks) return;
	const size_t gtile = (size_t)job.dec_off + tile;
	const TileDesc &td = p.tiles[gtile];
	const int n_seg = td.n_seg;
	const int carry_in = (tile == 0) ? p.st[stream].carry_in : p.tiles[gtile - 1].carry_out;
	if (n_seg == 0 && carry_in == 0) return;
	const uint32_t *d = p.dec + gtile * kBlockDec;
	int32_t *o = p.devfm + gtile * kBlockDec;
	uint32_t prev_last;
	if (tile == 0) prev_last = ((uint32_t)(uint16_t)p.st[stream].last_i) | ((uint32_t)(uint16_t)p.st[stream].last_q << 16);
	else prev_last = d[-1];   // the previous block's last sample is always stored
	// the kept samples are [0, carry_in) U segments (build_regions): walk them in order, skipping what is covered
	int covered = 0;
	for (int k = -1; k < n_seg; k++) {
		int a = 0, b = carry_in;
		if (k >= 0) {
			const int s0 = td.seg_start[k];
			a = max(s0, covered);
			b = s0 + td.seg_len[k];
		}
		for (int m = a + (int)threadIdx.x; m < b; m += (int)blockDim.x) {
			const uint32_t cw = d[m], lw = (m == 0) ? prev_last : d[m - 1];
			o[m] = fm_dev_fast((int)(int16_t)(cw & 0xffff), (int)(int16_t)(cw >> 16), (int)(int16_t)(lw & 0xffff),
					   (int)(int16_t)(lw >> 16));
		}
		covered = max(covered, b);
	}
}

// ------------------------------------------------------------------------------------------------
// devfm_win_kernel: fm_dev for the samples the demodulators will actually read.
//
// devfm_kernel above works from the block descriptors, i.e. on every sample the front-end kept: with the auto
// threshold that is everything above the call's speculative LOWER bound, about a quarter of all samples, of which the
// demodulators - which run on the true threshold - read a seventh.  Once the threshold walk has listed the windows the
// needed samples are known: a demodulator with timeout T is active on the union of [t, t+T-1] over the true triggers
// t, so the windows of the fm_dev-using demodulator with the LONGEST timeout (BackParams::demod) contain every sample
// any of them reads, warm-up histories included (those are earlier windows).  One CTA per window.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) devfm_win_kernel(const BackParams p)
{
	const int stream = blockIdx.y;
	const StreamJob job = p.jobs[stream];
	if (job.n_blocks == 0) return;
	const int demod = p.demod;
	const uint32_t n_win = p.wincnt[stream].n[demod];
	const WinEntry *wl = p.wins + job.win_off + (size_t)demod * job.win_cap;
	const uint32_t call_len = job.n_blocks * (uint32_t)kBlockDec;
	const uint32_t *d = p.dec + (size_t)job.dec_off * kBlockDec;
	int32_t *o = p.devfm + (size_t)job.dec_off * kBlockDec;
	const StreamState *st = p.st + stream;
	const uint32_t prev_last = ((uint32_t)(uint16_t)st->last_i) | ((uint32_t)(uint16_t)st->last_q << 16);
	for (uint32_t w = blockIdx.x; w < n_win; w += gridDim.x) {
		const WinEntry e = wl[w];
		if (e.start >= call_len) break;
		const uint32_t last = min(e.end, call_len - 1);
		for (uint32_t m = e.start + threadIdx.x; m <= last; m += blockDim.x) {
			const uint32_t cw = d[m], lw = (m == 0) ? prev_last : d[m - 1];   // the sample before a window's trigger is always stored
			o[m] = fm_dev_fast((int)(int16_t)(cw & 0xffff), (int)(int16_t)(cw >> 16), (int)(int16_t)(lw & 0xffff),
					   (int)(int16_t)(lw >> 16));
		}
	}
}

// fm_dev of sample 0 once the previous call's last sample is known (devfm_win_kernel of a call may run beside the previous
// call's verifier, before submit_epilogue_kernel has rolled StreamState::last_i/q forward)
__global__ void devfm_first_kernel(const BackParams p)
{
	const int stream = blockIdx.x * blockDim.x + threadIdx.x;
	if (stream >= p.n_streams) return;
	const StreamJob job = p.jobs[stream];
	if (job.n_blocks == 0) return;
	const StreamState *st = p.st + stream;
	const uint32_t cw = p.dec[(size_t)job.dec_off * kBlockDec];
	p.devfm[(size_t)job.dec_off * kBlockDec] = fm_dev_fast((int)(int16_t)(cw & 0xffff), (int)(int16_t)(cw >> 16), (int)st->last_i, (int)st->last_q);
}

// ------------------------------------------------------------------------------------------------
// window context shared by the speculative kernels and the verifier's slow path
// ------------------------------------------------------------------------------------------------
struct WinCtx {
	const BackParams *p;
	int stream, demod;
	const uint32_t *dec;     // this stream's sparse decimated samples, indexed by position in the call
	const int32_t *devfm;
	int32_t *ld;             // filter chains: this demodulator's pre-filtered slicer inputs (int)y, indexed like devfm; else null
	uint32_t call_len;
	uint32_t prev_last;      // sample before position 0
	int64_t base_pos;        // blocks_done * 8192
};

static __device__ int put_frame(const WinCtx &c, int reuse, const DemodState &s, uint32_t pos, double rssi_raw, int offset)
{
	uint32_t k = (uint32_t)reuse;
	if (reuse < 0) {
		k = atomicAdd(&c.p->counters->n_frames, 1u);
		if (k >= c.p->max_frames) {
			c.p->counters->overflow = 1;
			return -1;
		}
	}
	DevFrame &f = c.p->frames[k];
	f.status = -2;   // not a frame until everything below is in place (parse_kernel of the previous call may be looking)
	f.stream = c.stream;
	f.demod = c.demod;
	f.type = c.p->cfg->d[c.demod].type;
	f.byte_cnt = s.byte_cnt;
	f.offset = offset;
	f.n_records = 0;
	f.first_record = c.p->slot_tag;   // until parsed: whose parse_kernel this frame is for
	f.pos = c.base_pos + pos;
	f.rssi_raw = rssi_raw;
	for (int n = 0; n < kMaxRdata; n++) f.rdata[n] = s.rdata[n];
	__threadfence();
	f.status = -1;
	return (int)k;
}
// a re-run that no longer yields a frame retires the slot its first run claimed
static __device__ void drop_frame(const WinCtx &c, int idx)
{
	if (idx >= 0) c.p->frames[idx].status = -2;
}

// tfa2_decoder::store_bit prints "Inverted SYNC" whenever the inverted sync word passes the shift register
// (tfa2.cpp:294-300), noise windows included.  A window that saw some leaves ONE notice in the frame list: status 3,
// byte_cnt = how many, pos = the window's end (where its frame, if any, follows).  Slot handling as for frames.
static __device__ int put_notice(const WinCtx &c, int reuse, int count, uint32_t pos)
{
	uint32_t k = (uint32_t)reuse;
	if (reuse < 0) {
		k = atomicAdd(&c.p->counters->n_frames, 1u);
		if (k >= c.p->max_frames) {
			c.p->counters->overflow = 1;
			return -1;
		}
	}
	DevFrame &f = c.p->frames[k];
	f.stream = c.stream;
	f.demod = c.demod;
	f.type = c.p->cfg->d[c.demod].type;
	f.status = 3;
	f.byte_cnt = count;
	f.offset = 0;
	f.n_records = 0;
	f.first_record = 0;
	f.pos = c.base_pos + pos;
	f.rssi_raw = 1.0;
	for (int n = 0; n < kMaxRdata; n++) f.rdata[n] = 0;
	return (int)k;
}
static __device__ int window_notice(const WinCtx &c, int pad3, int count, uint32_t end)
{
	int ni = pad3 - 1;
	if (count) {
		ni = put_notice(c, ni, count, end);
	} else {
		drop_frame(c, ni);
		ni = -1;
	}
	return ni + 1;   // the record's new pad3
}

// last_bit_idx bookkeeping: the reference keeps it relative to the current block and subtracts len at every
// block start unless it is 0 (demodulator::start, decoder.cpp:118-122) - so 0 stays 0 forever
__device__ __forceinline__ int lbi_at_block(int v, int vblock, int block)
{
	if (v == 0) return 0;
	const long long r = (long long)v - (long long)kIdxPerBlock * (block - vblock);
	return (int)max(r, (long long)INT32_MIN / 2);
}

// ---- TFA_1 window ------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// The serial walks read the sparse buffers 16 samples per iteration with 128-bit loads (all four in flight
// together), so one L2 round trip is paid per 16 samples instead of per sample.  Chunks are 16-aligned and
// never cross the end of the call (a call is a whole number of 8192-sample blocks).
struct Chunk16 { int v[16]; };
__device__ __forceinline__ Chunk16 load16(const int32_t *p)
{
	Chunk16 c;
	const int4 a = reinterpret_cast<const int4 *>(p)[0], b = reinterpret_cast<const int4 *>(p)[1];
	const int4 d = reinterpret_cast<const int4 *>(p)[2], e = reinterpret_cast<const int4 *>(p)[3];
	c.v[0] = a.x; c.v[1] = a.y; c.v[2] = a.z; c.v[3] = a.w; c.v[4] = b.x; c.v[5] = b.y; c.v[6] = b.z; c.v[7] = b.w;
	c.v[8] = d.x; c.v[9] = d.y; c.v[10] = d.z; c.v[11] = d.w; c.v[12] = e.x; c.v[13] = e.y; c.v[14] = e.z; c.v[15] = e.w;
	return c;
}
// Order-sensitive hash of the slicer input sequence: two 32-bit weighted sums, sum ld[i]*(2i+1) and
// sum ld[i]*(i*i+i+1), packed into 64 bits.  Additive (no dependency chain; the walks are instruction
// bound) and cheap (one 32-bit IMAD each).  The sequences compared differ, if at all, by +-1 in isolated
// places: one change moves the first sum by an odd number, two changes cannot cancel in both sums.
struct LdHash {
	uint32_t a, b, w1, w2, dw2;   // running sums and the current position weights
	__device__ __forceinline__ void init() { a = b = 0; w1 = 1; w2 = 1; dw2 = 2; }
	__device__ __forceinline__ void add(int ld)
	{
		a += (uint32_t)ld * w1;
		b += (uint32_t)ld * w2;
		w1 += 2;
		w2 += dw2;   // i*i+i+1 -> (i+1)^2+(i+1)+1 adds 2i+2
		dw2 += 2;
	}
	__device__ __forceinline__ unsigned long long value() const { return ((unsigned long long)b << 32) | a; }
};
// Bits are not pushed into the framer from inside the per-sample walk: a warp walks 32 windows in lockstep, and
// a bit burst on ONE lane (up to 31 framer steps for one edge, tfa2.cpp:400-407) would be executed, masked, by
// all of them at nearly every sample.  The walk only appends run-length entries ((count << 1) | bit) to a small
// per-thread list; the list is drained in a loop in which every active lane performs exactly one framer step per
// iteration, so the lanes stay converged.  The demodulators never read framer state (no has_sync feedback in
// tfa1_demod / tfa2_demod), so deferring the framer changes nothing.
constexpr int kBitRuns = 64;
struct BitRuns {
	uint16_t run[kBitRuns];
	int n;
};

// tfa1_demod::demod, tfa1.cpp:143-190 over the samples [start, min(end, call_len-1)].
// exact: s holds the true carried state (continuation or first window of the call); otherwise s.sr == 0
// is a speculation that is checked with head31/nbits.
static __device__ void run_tfa1_window(const WinCtx &c, const WinEntry &e, DemodState &s, WinRec &rec, bool resume)
{
	if (!resume) {
		s.mark_lvl = 0;
		s.rssi_i = 0;
		s.last_bit_idx = 0;
		s.sr_cnt = -1;
		s.byte_cnt = 0;
		s.rdata[10] = 0;
	}
	uint32_t head = 0;
	int nbits = 0;
	int mark = s.mark_lvl, rssi = s.rssi_i, lbi = s.last_bit_idx;   // hot state in registers
	BitRuns br;
	br.n = 0;
	auto drain = [&]() {
		int i = 0, rem = 0, b = 0;
		for (;;) {
			if (rem == 0) {
				if (i == br.n) break;
				b = br.run[i] & 1;
				rem = br.run[i] >> 1;
				i++;
				if (rem == 0) continue;
			}
			if (nbits < 31) head |= (uint32_t)b << nbits;
			nbits++;
			tfa1_bit(s, b);
			rem--;
		}
		br.n = 0;
	};
	const uint32_t last = min(e.end, c.call_len - 1);
	const bool taps = c.p->tap_cap != 0;
	int32_t *tap = taps ? c.p->tap_i32[1] + ((size_t)c.stream * kMaxDemods + c.demod) * c.p->tap_cap : nullptr;
	uint32_t lw = (e.start == 0) ? c.prev_last : c.dec[e.start - 1];
	// four samples per 128-bit load, the next load issued before the current four are walked; the walk itself is
	// ONE copy of the per-sample code (rotating the vector) - unrolled copies thrash the instruction cache
	const int4 *src4 = reinterpret_cast<const int4 *>(c.dec);
	if ((e.start & ~31u) + 32 <= last) prefetch_l1(c.dec + (e.start & ~31u) + 32);
	int4 nxt = src4[(e.start & ~3u) >> 2];
	for (uint32_t cb = e.start & ~3u; cb <= last; cb += 4) {
	int4 v4 = nxt;
	if (cb + 4 <= last) nxt = src4[(cb + 4) >> 2];
	if ((cb & 31u) == 0 && cb + 64 <= last) prefetch_l1(c.dec + cb + 64);   // the line after next: an L2 round trip is ~1 us
#pragma unroll 1
	for (int kk = 0; kk < 4; kk++) {
		const uint32_t m = cb + kk;
		const uint32_t cw = (uint32_t)v4.x;
		v4.x = v4.y; v4.y = v4.z; v4.z = v4.w;
		if (m < e.start || m > last) continue;
		const int index = 2 * (int)(m & (kBlockDec - 1));
		if (index == 0 && m != e.start && lbi) lbi -= kIdxPerBlock;
		const int dev = fm_dev_nrzs((int)(int16_t)(cw & 0xffff), (int)(int16_t)(cw >> 16), (int)(int16_t)(lw & 0xffff),
					    (int)(int16_t)(lw >> 16));
		lw = cw;
		if (taps) {
			const uint32_t ti = e.cum + (m - e.start);
			if (ti < c.p->tap_cap) tap[ti] = dev;
		}
		if (dev > mark) mark = dev;
		else mark = __double2int_rz(__dmul_rn((double)mark, 0.95));
		if (mark > rssi) rssi = mark;
		if (dev < mark / 2) {
			if (lbi) {
				const int gap = index - lbi;
				if (gap > 4) {
					// `for (n = 22; n <= gap; n += 20) bit(1); bit(0);`  (tfa1.cpp:168-173)
					// a run entry holds 15 bits of count: a gap beyond ~655k index units (a carrier held above the
					// threshold for 0.85 s without a dip) is split into several entries
					for (int ones = (gap >= 22) ? (gap - 22) / 20 + 1 : 0; ones > 0; ones -= 32767) {
						if (br.n + 2 > kBitRuns) drain();
						br.run[br.n++] = (uint16_t)((min(ones, 32767) << 1) | 1);
					}
					if (br.n + 1 > kBitRuns) drain();
					br.run[br.n++] = (uint16_t)(1 << 1);
				}
			}
			if (index - lbi > 2) lbi = index;
		}
	}
	}
	drain();
	s.mark_lvl = mark;
	s.rssi_i = rssi;
	s.last_bit_idx = lbi;
	rec.head31 = head;
	rec.nbits = nbits;
	rec.sr_final = s.sr;
	if (last == e.end) {
		// tfa1_decoder::flush gate (tfa1.cpp:48); the parser runs in parse_kernel
		int fi = -1;
		if (s.byte_cnt >= 10) fi = put_frame(c, rec.frame_idx, s, e.end, (double)s.rssi_i, 0);
		else drop_frame(c, rec.frame_idx);
		rec.frame_idx = fi;
		rec.flags = (rec.flags & kRecExact) | kRecRan;
	} else {
		s.timeout_cnt = (int)(e.end - last);
		rec.flags = (rec.flags & kRecExact) | kRecRan | kRecUnfinished;
	}
}

// ---- TFA_2 / TFA_3 / TX22 window ---------------------------------------------------------------------
// tfa2_demod::demod, tfa2.cpp:346-442 over [start, min(end, call_len-1)].  `far`: last_bit_idx is assumed
// to be so old that the first edge candidate only re-arms it (index > lbi+8 true, tdiff >= 32*spb).
static __device__ void run_tfa2_window(const WinCtx &c, const DemodCfg &cfg, const WinEntry &e, DemodState &s, WinRec &rec,
				       bool resume, bool far)
{
	if (!resume) {
		tfa2_reset(s);
		s.sr_cnt = -1;
		s.sr = 0;
		s.byte_cnt = 0;
	}
	const uint32_t last = min(e.end, c.call_len - 1);
	const bool taps = c.p->tap_cap != 0;
	const size_t tbase = ((size_t)c.stream * kMaxDemods + c.demod) * c.p->tap_cap;
	bool have_edge = false;
	rec.flags &= ~kRecEdge;
	// hot state in registers
	Biquad lp = s.lp;
	const BiquadCoef k = cfg.lp;
	int bitcnt = s.bitcnt, dmin = s.dmin, dmax = s.dmax, offset = s.offset, last_bit = s.last_bit, rssi = s.rssi_i,
	    lbi = s.last_bit_idx;
	const int td_lo = cfg.td_lo, td_hi = cfg.td_hi;   // tfa2.cpp:391 `tdiff>spb/4 && tdiff<32*spb` for integer tdiff
	LdHash hash;
	hash.init();
	BitRuns br;
	br.n = 0;
	auto drain = [&]() {
		int i = 0, rem = 0, b = 0;
		for (;;) {
			if (rem == 0) {
				if (i == br.n) break;
				b = br.run[i] & 1;
				rem = br.run[i] >> 1;
				i++;
				if (rem == 0) continue;
			}
			tfa2_bit(s, b);
			rem--;
		}
		br.n = 0;
	};
	// the slicer levels only move while bitcnt < 10: keep them out of the per-sample chain
	int noffset = __double2int_rz(__dmul_rn(0.9, (double)offset));
	int hi = noffset + dmax / 32, lo = noffset + dmin / 32;
	// filter chains (biq_kernel) ran ahead: the slicer inputs (int)y are read instead of being filtered here
	const bool pre = c.ld != nullptr;
	const int32_t *fsrc = pre ? c.ld : c.devfm;
	const int4 *src4 = reinterpret_cast<const int4 *>(fsrc);
	prefetch_l1(c.dec + (e.start & ~31u));        // the first bits' rssi reads I/Q (tfa2.cpp:373)
	if ((e.start & ~31u) + 32 <= last) {
		prefetch_l1(fsrc + (e.start & ~31u) + 32);
		prefetch_l1(c.dec + (e.start & ~31u) + 32);
	}
	// sample m is an edge candidate: (ld > hi || ld < lo) && bit != last_bit  (tfa2.cpp:386-412)
	auto edge = [&](uint32_t m, int ld) {
		const int index = 2 * (int)(m & (kBlockDec - 1));
		const int bit = ld > hi ? 1 : 0;
		if (far && !have_edge) {
			// speculated: index > last_bit_idx+8, tdiff >= 32*spb  ->  bitcnt++, no bits, re-arm
			rec.first_edge = index;
			rec.first_edge_block = (int)(m >> 13);
			rec.flags |= kRecEdge;
			bitcnt++;
			lbi = index;
		} else {
			if (index > lbi + 8) {
				bitcnt++;
				const int tdiff = index - lbi;
				if (tdiff >= td_lo && tdiff <= td_hi) {
					const int bit_diff = tdiff / 2;
					const int numbits = cfg.nbits[bit_diff];
					if (br.n + 2 > kBitRuns) drain();
					if (numbits < 32 && numbits > 1) br.run[br.n++] = (uint16_t)(((numbits - 1) << 1) | last_bit);
					br.run[br.n++] = (uint16_t)((1 << 1) | bit);
					last_bit = bit;
				}
			}
			if (index - lbi > 2) lbi = index;
		}
		have_edge = true;
	};
#ifdef TFR_WIN_PROFILE
	long long pf_fast = 0, pf_gen = 0, pf_cfast = 0, pf_cgen = 0, pf_t0 = clock64();
#endif
	int4 nxt = src4[(e.start & ~3u) >> 2];
	for (uint32_t cb = e.start & ~3u; cb <= last; cb += 4) {
	int4 v4 = nxt;
	if (cb + 4 <= last) nxt = src4[(cb + 4) >> 2];
	if ((cb & 31u) == 0 && cb + 64 <= last) {   // the line after next: an L2 round trip is ~1 us
		prefetch_l1(fsrc + cb + 64);
		if (bitcnt < 10) prefetch_l1(c.dec + cb + 64);
	}
#ifdef TFR_WIN_PROFILE
	const long long pg0 = clock64();
#endif
	if (cb >= e.start && cb + 3 <= last) {
		// A whole group of four.  The four filter steps, truncations and hash terms are ONE basic block shared by
		// every lane of the warp, whatever phase its window is in (the FP64 recurrence is the only chain); the
		// per-sample state machine follows: level tracking while bitcnt < 10 (tfa2.cpp:365-374), then the - rare -
		// edge candidates, in order.  (A window's first and last partial group take the per-sample loop below.)
		if ((cb & (kBlockDec - 1)) == 0 && cb != e.start && lbi) lbi -= kIdxPerBlock;   // demodulator::start of a new block
		int l0 = v4.x, l1 = v4.y, l2 = v4.z, l3 = v4.w;
		if (!pre) {
		const double y0 = biquad_step(lp, k, int_to_double(v4.x));
		const double y1 = biquad_step(lp, k, int_to_double(v4.y));
		const double y2 = biquad_step(lp, k, int_to_double(v4.z));
		const double y3 = biquad_step(lp, k, int_to_double(v4.w));
		l0 = trunc_to_int(y0); l1 = trunc_to_int(y1); l2 = trunc_to_int(y2); l3 = trunc_to_int(y3);
		hash.add(l0);
		hash.add(l1);
		hash.add(l2);
		hash.add(l3);
		if (taps) {
			const uint32_t ti = e.cum + (cb - e.start);
			if (ti + 3 < c.p->tap_cap) {
				c.p->tap_i32[0][tbase + ti] = v4.x; c.p->tap_f64[tbase + ti] = y0;
				c.p->tap_i32[0][tbase + ti + 1] = v4.y; c.p->tap_f64[tbase + ti + 1] = y1;
				c.p->tap_i32[0][tbase + ti + 2] = v4.z; c.p->tap_f64[tbase + ti + 2] = y2;
				c.p->tap_i32[0][tbase + ti + 3] = v4.w; c.p->tap_f64[tbase + ti + 3] = y3;
			} else {
				const int dv[4] = { v4.x, v4.y, v4.z, v4.w };
				const double yv[4] = { y0, y1, y2, y3 };
				for (int j = 0; j < 4; j++)
					if (ti + j < c.p->tap_cap) { c.p->tap_i32[0][tbase + ti + j] = dv[j]; c.p->tap_f64[tbase + ti + j] = yv[j]; }
			}
		}
		}
		auto sample = [&](uint32_t m, int ld) {
			if (bitcnt < 10) {
				if (ld > dmax || ld < dmin) {
					// the levels are pure functions of dmax/dmin (tfa2.cpp:366-368, 380-381): recompute them only when one moved
					if (ld > dmax) dmax = (7 * dmax + ld) / 8;
					if (ld < dmin) dmin = (7 * dmin + ld) / 8;
					offset = (dmax + dmin) / 2;
					noffset = __double2int_rz(__dmul_rn(0.9, (double)offset));
					hi = noffset + dmax / 32;
					lo = noffset + dmin / 32;
				}
				if (bitcnt > 4) {
					const uint32_t cw = c.dec[m];
					const int i = (int)(int16_t)(cw & 0xffff), q = (int)(int16_t)(cw >> 16);
					const uint32_t sum = (uint32_t)rssi + (uint32_t)(i * i) + (uint32_t)(q * q);
					rssi = (int)((uint32_t)rssi + (uint32_t)((int)sum / 100));
				}
			}
			if ((ld > hi || ld < lo) && (int)(ld > hi) != last_bit) edge(m, ld);
		};
		if (bitcnt < 10) {
			sample(cb, l0);
			sample(cb + 1, l1);
			sample(cb + 2, l2);
			sample(cb + 3, l3);
		} else {
			// The levels are frozen (tfa2.cpp:365 `if (bitcnt<10)`): which of the four samples is an edge candidate
			// depends on last_bit alone - (ld > hi || ld < lo) && (ld > hi) != last_bit is `ld > hi` for last_bit 0 and
			// `ld < lo && !(ld > hi)` for last_bit 1.  The candidates are taken in order, but in ONE loop: every lane of
			// the warp handles its first candidate in the same pass, whatever sample of the group it sits on (four
			// per-sample regions ran the edge code up to four times per group, each time for one or two lanes).
			const unsigned h = (unsigned)(l0 > hi) | ((unsigned)(l1 > hi) << 1) | ((unsigned)(l2 > hi) << 2) | ((unsigned)(l3 > hi) << 3);
			const unsigned l = (unsigned)(l0 < lo) | ((unsigned)(l1 < lo) << 1) | ((unsigned)(l2 < lo) << 2) | ((unsigned)(l3 < lo) << 3);
			const unsigned c0 = h, c1 = l & ~h;
			unsigned todo = 0xfu;
			for (;;) {
				const unsigned m = (last_bit ? c1 : c0) & todo;
				if (!m) break;
				const int kq = __ffs(m) - 1;
				edge(cb + kq, kq == 0 ? l0 : kq == 1 ? l1 : kq == 2 ? l2 : l3);
				todo = (0xfu << (kq + 1)) & 0xfu;
			}
		}
#ifdef TFR_WIN_PROFILE
		pf_fast++; pf_cfast += clock64() - pg0;
#endif
		continue;
	}
#pragma unroll 1
	for (int kk = 0; kk < 4; kk++) {
		const uint32_t m = cb + kk;
		const int dev0 = v4.x;
		v4.x = v4.y; v4.y = v4.z; v4.z = v4.w;
		if (m < e.start || m > last) continue;
		const int index = 2 * (int)(m & (kBlockDec - 1));
		if (index == 0 && m != e.start && lbi) lbi -= kIdxPerBlock;
		int ld = dev0;
		if (!pre) {
			const double y = biquad_step(lp, k, int_to_double(dev0));
			if (taps) {
				const uint32_t ti = e.cum + (m - e.start);
				if (ti < c.p->tap_cap) {
					c.p->tap_i32[0][tbase + ti] = dev0;
					c.p->tap_f64[tbase + ti] = y;
				}
			}
			ld = trunc_to_int(y);
			hash.add(ld);
		}
		if (bitcnt < 10) {
			if (ld > dmax || ld < dmin) {
				// the levels are pure functions of dmax/dmin (tfa2.cpp:366-368, 380-381): recompute them only when one moved
				if (ld > dmax) dmax = (7 * dmax + ld) / 8;
				if (ld < dmin) dmin = (7 * dmin + ld) / 8;
				offset = (dmax + dmin) / 2;
				noffset = __double2int_rz(__dmul_rn(0.9, (double)offset));
				hi = noffset + dmax / 32;
				lo = noffset + dmin / 32;
			}
			if (bitcnt > 4) {
				const uint32_t cw = c.dec[m];
				const int i = (int)(int16_t)(cw & 0xffff), q = (int)(int16_t)(cw >> 16);
				const uint32_t sum = (uint32_t)rssi + (uint32_t)(i * i) + (uint32_t)(q * q);
				rssi = (int)((uint32_t)rssi + (uint32_t)((int)sum / 100));
			}
		}
		if ((ld > hi || ld < lo) && (int)(ld > hi) != last_bit) edge(m, ld);
	}
#ifdef TFR_WIN_PROFILE
	pf_gen++; pf_cgen += clock64() - pg0;
#endif
	}
#ifdef TFR_WIN_PROFILE
	atomicAdd(&g_slprof[0], (unsigned long long)pf_fast); atomicAdd(&g_slprof[1], (unsigned long long)pf_gen); atomicAdd(&g_slprof[2], (unsigned long long)pf_cfast); atomicAdd(&g_slprof[3], (unsigned long long)pf_cgen); atomicAdd(&g_slprof[4], (unsigned long long)(clock64() - pf_t0)); atomicAdd(&g_slprof[5], 1ull);
#endif
	s.lp = lp;
	s.bitcnt = bitcnt; s.dmin = dmin; s.dmax = dmax; s.offset = offset; s.last_bit = last_bit; s.rssi_i = rssi;
	s.last_bit_idx = lbi;
	rec.e_y0 = lp.y0;
	rec.e_y1 = lp.y1;
	rec.ld_hash = hash.value();
	rec.lbi_end = lbi;
	rec.lbi_end_block = (int)(last >> 13);
	if (!far && have_edge) rec.flags |= kRecEdge;
	if (last == e.end) {
		if (br.n + 1 > kBitRuns) drain();
		br.run[br.n++] = (uint16_t)((16 << 1) | last_bit);   // 16 x store_bit(last_bit) before the flush (tfa2.cpp:430-433)
	}
	drain();
	if (last == e.end) {
		const bool gate = (cfg.kind == K_TX22) ? (s.byte_cnt >= 7 && s.byte_cnt < 64) : (s.byte_cnt >= 7);
		int fi = -1;
		if (gate) fi = put_frame(c, rec.frame_idx, s, e.end, (double)s.rssi_i, s.offset);
		else drop_frame(c, rec.frame_idx);
		rec.frame_idx = fi;
		rec.pad3 = window_notice(c, rec.pad3, s.inv_cnt, e.end);
		s.sr_cnt = -1;
		s.sr = 0;
		s.byte_cnt = 0;
		tfa2_reset(s);
		s.timeout_cnt = 0;
		rec.flags = (rec.flags & (kRecExact | kRecEdge | kRecLbiIn)) | kRecRan;
	} else {
		// the data ended inside the window: what it has seen so far is reported now (the reference printed it long ago)
		rec.pad3 = window_notice(c, rec.pad3, s.inv_cnt, last);
		s.inv_cnt = 0;   // reported: the continuation in the next call starts counting afresh
		s.timeout_cnt = (int)(e.end - last);
		rec.flags = (rec.flags & (kRecExact | kRecEdge | kRecLbiIn)) | kRecRan | kRecUnfinished;
	}
}

__device__ __forceinline__ WinCtx make_ctx(const BackParams &p, int stream, int demod, const StreamJob &job, const StreamState *st)
{
	WinCtx c;
	c.p = &p;
	c.stream = stream;
	c.demod = demod;
	c.dec = p.dec + (size_t)job.dec_off * kBlockDec;
	c.devfm = p.devfm ? p.devfm + (size_t)job.dec_off * kBlockDec : nullptr;
	c.ld = (p.ld && p.fm_slot[demod] >= 0) ? p.ld + (size_t)p.fm_slot[demod] * p.ld_stride + (size_t)job.dec_off * kBlockDec : nullptr;
	c.call_len = job.n_blocks * (uint32_t)kBlockDec;
	c.prev_last = ((uint32_t)(uint16_t)st->last_i) | ((uint32_t)(uint16_t)st->last_q << 16);
	c.base_pos = job.base_blocks * (int64_t)kBlockDec;
	return c;
}

// ------------------------------------------------------------------------------------------------
// Filter chains: biq_kernel + biq_verify_kernel (TFA_2 / TFA_3 / TX22).
//
// The biquad is the one part of a TFA_2-family demodulator that never resets (tfa2.cpp:325-334), so a window's first
// filter state depends on everything before it.  Speculating that state per WINDOW (win_kernel's warm-up over the
// preceding three timeouts) makes every sample pass the filter four times, inside the slow, divergent slicer threads.
// Here the filter runs ahead of the slicers and on its own: one thread per CHAIN of kBiqK consecutive windows, a
// warm-up of TFR_BIQ_WARM_X4 quarter timeouts once per chain, then nothing but the recurrence (four dependent FP64
// operations per sample), writing the slicer input (int)y of every window sample to `ld`.  biq_verify_kernel then
// proves the chains link up - a chain's assumed start outputs must be bitwise the outputs its predecessor left - and
// re-filters the ones that do not, in parallel rounds, exactly like verify_kernel does for whole windows.  After it
// `ld` holds what the reference's filter would have produced, bit for bit, and the window kernels only slice.
// ------------------------------------------------------------------------------------------------
#ifndef TFR_BIQ_WARM_X4
#define TFR_BIQ_WARM_X4 16   // warm-up length in quarter timeouts (amortised over kBiqK windows)
#endif
constexpr int kBiqThreads = 64;
#ifdef TFR_BIQ_PROFILE
__device__ unsigned long long g_biqprof[8];   // max chain cycles, its steps; sum of cycles, sum of steps, chains; max steps, its cycles
#endif

// The recurrence over the samples [a, b] of one window: (int)y into ld (STORE; taps from index tap0 on), filter state
// in lp_io.  Whole 128-byte lines run as batches of 32 samples with the eight loads of the NEXT batch in flight while
// the current one is filtered; the partial lines at both ends go sample by sample (a neighbouring window of another
// chain may share a 16-byte group with this span, so only the span's own elements are written).
// Measured on B200 (in-kernel clocks, bench workload, ~1 warp per scheduler: the kernel is latency bound): this form
// 140 cycles per sample; four samples per load 177; one masked loop over lines for all lanes of a warp 250; the same
// with the lines loaded and stored cooperatively through shared memory 420.  The recurrence itself needs 32.
template <bool STORE>
static __device__ __forceinline__ void biq_span(const WinCtx &c, const BiquadCoef &k, Biquad &lp_io, uint32_t a, uint32_t b, uint32_t tap0)
{
	Biquad lp = lp_io;   // keep the recurrence in registers
	const bool taps = STORE && c.p->tap_cap != 0;
	const size_t tbase = ((size_t)c.stream * kMaxDemods + c.demod) * c.p->tap_cap;
	auto step = [&](uint32_t m, int dv) -> int {
		const double y = biquad_step(lp, k, int_to_double(dv));
		if (taps) {
			const uint32_t ti = tap0 + (m - a);
			if (ti < c.p->tap_cap) {
				c.p->tap_i32[0][tbase + ti] = dv;
				c.p->tap_f64[tbase + ti] = y;
			}
		}
		return trunc_to_int(y);
	};
	uint32_t m = a;
	const uint32_t first_full = (a + 31u) & ~31u;
	if (first_full + 31 <= b) prefetch_l1(c.devfm + first_full);   // on its way while the head is walked
#pragma unroll 1
	for (; (m & 31u) && m <= b; m++) {
		const int l = step(m, c.devfm[m]);
		if (STORE) c.ld[m] = l;
	}
	if (m + 31 <= b) {
		const int4 *src4 = reinterpret_cast<const int4 *>(c.devfm);
		int4 *dst4 = reinterpret_cast<int4 *>(c.ld);
		int4 cur[8], nx[8];
#pragma unroll
		for (int q = 0; q < 8; q++) cur[q] = src4[(m >> 2) + q];
		for (; m + 31 <= b; m += 32) {
			const bool more = m + 63 <= b;
			if (!more && m + 32 <= b) prefetch_l1(c.devfm + m + 32);   // the tail's line
#pragma unroll
			for (int q = 0; q < 8; q++) nx[q] = more ? src4[((m + 32) >> 2) + q] : cur[q];
#pragma unroll
			for (int q = 0; q < 8; q++) {
				int4 o;
				o.x = step(m + 4 * q, cur[q].x);
				o.y = step(m + 4 * q + 1, cur[q].y);
				o.z = step(m + 4 * q + 2, cur[q].z);
				o.w = step(m + 4 * q + 3, cur[q].w);
				if (STORE) dst4[(m >> 2) + q] = o;
			}
#pragma unroll
			for (int q = 0; q < 8; q++) cur[q] = nx[q];
		}
	}
#pragma unroll 1
	for (; m <= b; m++) {
		const int l = step(m, c.devfm[m]);
		if (STORE) c.ld[m] = l;
	}
	lp_io = lp;
}
// the windows [v, w_end) of one demodulator: window v from sample `from` on, the others whole; windows before w0 are
// warm-up history, from w0 on (int)y goes to ld.  r: the filter outputs at the first stored sample and after the last.
static __device__ __forceinline__ void biq_walk(const WinCtx &c, const BiquadCoef &k, const WinEntry *wl, uint32_t v, uint32_t from, uint32_t w0,
						uint32_t w_end, Biquad &lp, BiqRec &r)
{
	for (uint32_t u = v; u < w0; u++) biq_span<false>(c, k, lp, (u == v) ? from : wl[u].start, wl[u].end, 0u);
	r.u_y0 = lp.y0;
	r.u_y1 = lp.y1;
	for (uint32_t u = w0; u < w_end; u++) {
		const WinEntry e = wl[u];
		biq_span<true>(c, k, lp, e.start, min(e.end, c.call_len - 1), e.cum);
	}
	r.e_y0 = lp.y0;
	r.e_y1 = lp.y1;
}
constexpr unsigned kFullMask = 0xffffffffu;
// the filter's input history at the end of window v (d1 = last, d2 = the one before), v >= 0
static __device__ __forceinline__ void biq_inputs_after(const WinCtx &c, const WinEntry *wl, int v, const DemodState &carry0, Biquad &lp)
{
	const WinEntry e = wl[v];
	const uint32_t last = min(e.end, c.call_len - 1);
	lp.d1 = (double)c.devfm[last];
	if (last >= e.start + 1) lp.d2 = (double)c.devfm[last - 1];
	else if (v > 0) lp.d2 = (double)c.devfm[min(wl[v - 1].end, c.call_len - 1)];   // one-sample window
	else lp.d2 = carry0.lp.d1;
}

__global__ void __launch_bounds__(kBiqThreads) biq_kernel(const BackParams p)
{
	const int stream = blockIdx.y;
	const int demod = (int)blockIdx.z;
	if (p.fm_slot[demod] < 0) return;
	const DemodCfg &cfg = p.cfg->d[demod];
	const StreamJob job = p.jobs[stream];
	if (job.n_blocks == 0) return;
	const StreamState *st = p.st + stream;
	uint32_t n_win = p.wincnt[stream].n[demod];
	const WinEntry *wl = p.wins + job.win_off + (size_t)demod * job.win_cap;
	BiqRec *bl = p.biq + job.win_off + (size_t)demod * job.win_cap;
	const WinCtx c = make_ctx(p, stream, demod, job, st);
	const BiquadCoef k = cfg.lp;
	while (n_win && wl[n_win - 1].start >= c.call_len) n_win--;   // windows beyond the data do not filter anything
	for (uint32_t ch = blockIdx.x * blockDim.x + threadIdx.x; ch * (uint32_t)kBiqK < n_win; ch += gridDim.x * blockDim.x) {
		const bool on = true;
		const uint32_t w0 = ch * (uint32_t)kBiqK, w_end = min(n_win, w0 + (uint32_t)kBiqK);
		Biquad lp = st->d[demod].lp;   // chain 0 starts from the state the previous call left: nothing to speculate
		BiqRec r;
		uint32_t v = w0, from = wl[w0].start;
		if (w0 != 0) {
			// warm-up over the samples of the preceding windows (the filter's actual history); reaching window 0 means
			// the true carried state can be used
			const uint32_t want = ((uint32_t)TFR_BIQ_WARM_X4 * (uint32_t)cfg.timeout) / 4u;
			uint32_t have = 0;
			while (v > 0 && have < want) {
				v--;
				const uint32_t len = wl[v].end - wl[v].start + 1;
				if (have + len >= want && v > 0) {
					from = wl[v].end + 1 - (want - have);
					have = want;
				} else {
					from = wl[v].start;
					have += len;
				}
			}
			if (!(v == 0 && from == wl[0].start)) lp.d1 = lp.d2 = lp.y0 = lp.y1 = 0.0;
		}
#ifdef TFR_BIQ_PROFILE
		const long long bp_t0 = clock64();
#endif
		biq_walk(c, k, wl, v, from, w0, w_end, lp, r);
#ifdef TFR_BIQ_PROFILE
		if (on) {
			const unsigned long long cyc = (unsigned long long)(clock64() - bp_t0);
			unsigned long long steps = wl[v].end - from + 1;
			for (uint32_t u = v + 1; u < w_end; u++) steps += min(wl[u].end, c.call_len - 1) - wl[u].start + 1;
			atomicMax(&g_biqprof[0], (cyc << 20) | min(steps, 0xfffffull));
			atomicAdd(&g_biqprof[2], cyc);
			atomicAdd(&g_biqprof[3], steps);
			atomicAdd(&g_biqprof[4], 1ull);
			atomicMax(&g_biqprof[5], (steps << 32) | min(cyc, 0xffffffffull));
		}
#endif
		if (on) bl[w0] = r;
	}
}

// one CTA per (stream, demod): proves that the chains link up, re-filters the ones that do not (rounds: a chain
// repaired from a predecessor that was itself still wrong is simply flagged again; the first bad chain of a round
// always starts from a proven state, so every round proves at least one more chain), then leaves the biquad state
// after the call's last window sample in StreamState::lp_next.
constexpr int kBiqVerThreads = 128;
__global__ void __launch_bounds__(kBiqVerThreads) biq_verify_kernel(const BackParams p)
{
	const int nd = p.cfg->n_demods;
	const int stream = blockIdx.x / nd, demod = blockIdx.x % nd;
	if (stream >= p.n_streams || p.fm_slot[demod] < 0) return;
	const DemodCfg &cfg = p.cfg->d[demod];
	const StreamJob job = p.jobs[stream];
	if (job.n_blocks == 0) return;
	StreamState *st = p.st + stream;
	uint32_t n_win = p.wincnt[stream].n[demod];
	const WinEntry *wl = p.wins + job.win_off + (size_t)demod * job.win_cap;
	BiqRec *bl = p.biq + job.win_off + (size_t)demod * job.win_cap;
	const WinCtx c = make_ctx(p, stream, demod, job, st);
	const BiquadCoef k = cfg.lp;
	const DemodState &carry0 = st->d[demod];
	while (n_win && wl[n_win - 1].start >= c.call_len) n_win--;   // windows beyond the data do not filter anything
	const uint32_t n_ch = (n_win + (uint32_t)kBiqK - 1) / (uint32_t)kBiqK;
	__shared__ uint32_t s_bad[kBiqVerThreads];
	__shared__ uint32_t s_n, s_first;
	uint32_t n_fix = 0;
	for (;;) {
		if (threadIdx.x == 0) {
			s_n = 0;
			s_first = 0xffffffffu;
		}
		__syncthreads();
		for (uint32_t ch = 1 + threadIdx.x; ch < n_ch; ch += kBiqVerThreads) {
			const BiqRec &a = bl[(ch - 1) * kBiqK], &b = bl[ch * kBiqK];
			if (__double_as_longlong(b.u_y0) != __double_as_longlong(a.e_y0) || __double_as_longlong(b.u_y1) != __double_as_longlong(a.e_y1)) {
				const uint32_t q = atomicAdd(&s_n, 1u);
				if (q < (uint32_t)kBiqVerThreads) s_bad[q] = ch;
				atomicMin(&s_first, ch);
			}
		}
		__syncthreads();
		const uint32_t nb = min(s_n, (uint32_t)kBiqVerThreads);
		if (nb == 0) break;
		if (threadIdx.x == 0 && s_n > (uint32_t)kBiqVerThreads) s_bad[0] = s_first;   // the certain one is in every round
		__syncthreads();
		if (threadIdx.x < nb) {
			const uint32_t ch = s_bad[threadIdx.x], w0 = ch * (uint32_t)kBiqK;
			Biquad lp;
			lp.y0 = bl[w0 - kBiqK].e_y0;
			lp.y1 = bl[w0 - kBiqK].e_y1;
			biq_inputs_after(c, wl, (int)w0 - 1, carry0, lp);
			BiqRec r;
			biq_walk(c, k, wl, w0, wl[w0].start, w0, min(n_win, w0 + (uint32_t)kBiqK), lp, r);
			bl[w0] = r;
			n_fix++;
		}
		__syncthreads();
	}
	if (n_fix) atomicAdd(&p.counters->rerun_biquad, n_fix);
	if (threadIdx.x == 0) {
		Biquad lp = carry0.lp;
		if (n_win) {
			lp.y0 = bl[(n_ch - 1) * kBiqK].e_y0;
			lp.y1 = bl[(n_ch - 1) * kBiqK].e_y1;
			biq_inputs_after(c, wl, (int)n_win - 1, carry0, lp);
		}
		st->lp_next[demod] = lp;
	}
}

// ------------------------------------------------------------------------------------------------
// win_kernel: one thread per window CHAIN; blockIdx.y = stream, blockIdx.z = registered demod (longest
// timeout first, so that the slowest windows start first).
//
// TFA_2 family: a window that begins within T_d+1 samples of its predecessor's end is `near`: the
// predecessor's last edge candidate can be less than 32 bit periods old (tfa2.cpp:395-397), so the "last
// edge is far in the past" speculation is unsafe for it.  Near windows are not speculated at all: the thread
// that ran the predecessor carries on into them with the biquad state and last_bit_idx it holds, exactly as
// the reference does.  A chain starts at a window that is not near (the speculation is then provably right
// unless last_bit_idx is the 0 sentinel, which verify_kernel catches) and is cut every kChainMax windows to
// bound the tail latency.  TFA_1 windows carry only the decoder shift register and stay one per thread.
// ------------------------------------------------------------------------------------------------
// the windows [lo, hi) of (stream, demod) a window-kernel launch works on (BackParams::part_lo / part_hi)
__device__ __forceinline__ void part_range(const BackParams &p, int stream, int demod, uint32_t &lo, uint32_t &hi)
{
	lo = (p.part_lo >= 0) ? p.partcnt[(size_t)p.part_lo * p.n_streams + stream].n[demod] : 0u;
	hi = (p.part_hi >= 0) ? p.partcnt[(size_t)p.part_hi * p.n_streams + stream].n[demod] : p.wincnt[stream].n[demod];
}

#ifndef TFR_LONG_X
#define TFR_LONG_X 4
#endif
// A chain is LONG when its windows add up to more than TFR_LONG_X timeouts (a telegram; noise that kept retriggering
// stays below that: a warp per chain only pays for the few really long ones).
// One thread walks a chain sample by sample at ~500 cycles per sample (every instruction waits for the one before);
// the few long chains of a call are therefore the window kernel's critical path.  They are taken out of the
// thread-per-chain kernel and run by winlong_kernel, one WARP per chain (below).
__device__ __forceinline__ bool chain_is_long(const WinEntry *wl, uint32_t w0, uint32_t n_win, uint32_t call_len, bool chains, int timeout)
{
	uint32_t tot = 0;
	for (uint32_t w = w0; w < n_win; w++) {
		if (w != w0 && (!chains || chain_head(wl, w, timeout))) break;
		if (wl[w].start >= call_len) break;
		tot += min(wl[w].end, call_len - 1) - wl[w].start + 1;
	}
	return tot > (uint32_t)TFR_LONG_X * (uint32_t)timeout;
}

// 64-thread CTAs, at least ten per SM (96 registers, no spills): the kernel is a crowd of divergent, latency-bound
// chains, so resident warps are what buys throughput (measured in one session: 128 registers / 12 warps per SM
// 2.17 ms back-end, 96 / 20 warps 2.02 ms, 80 / 24 warps with spills 2.05 ms, 64 / 32 warps 2.30 ms)
#ifndef TFR_WIN_MINBLOCKS
#define TFR_WIN_MINBLOCKS 10
#endif
#ifndef TFR_WIN_THREADS
#define TFR_WIN_THREADS 64
#endif
constexpr int kWinThreads = TFR_WIN_THREADS;
__global__ void __launch_bounds__(kWinThreads, TFR_WIN_MINBLOCKS) win_kernel(const BackParams p)
{
	const int stream = blockIdx.y;
	const int demod = (int)gridDim.z - 1 - (int)blockIdx.z;
	const DemodCfg &cfg = p.cfg->d[demod];
	if (cfg.kind == K_WHB) return;
	const StreamJob job = p.jobs[stream];
	if (job.n_blocks == 0) return;
	StreamState *st = p.st + stream;
	uint32_t w_lo, n_win;   // this launch's windows: [w_lo, n_win), both chain heads (or the ends of the list)
	part_range(p, stream, demod, w_lo, n_win);
	const WinEntry *wl = p.wins + job.win_off + (size_t)demod * job.win_cap;
	WinRec *rl = p.recs + job.win_off + (size_t)demod * job.win_cap;
	const WinCtx c = make_ctx(p, stream, demod, job, st);
	const bool chains = (cfg.kind != K_TFA1);

#ifdef TFR_WIN_PROFILE
	long long prof_t0 = clock64(), prof_warm = 0, prof_run = 0, prof_steps = 0;
#endif
	// thread t looks at the windows [t*G, (t+1)*G): normally the first one is the only head among them and its chain
	// covers the rest, so consecutive lanes all have work
	const uint32_t G = chains ? kChainGroup : 1u;
	for (uint32_t t0 = blockIdx.x * blockDim.x + threadIdx.x; w_lo + t0 * G < n_win; t0 += gridDim.x * blockDim.x)
	for (uint32_t w0 = w_lo + t0 * G; w0 < min(n_win, w_lo + (t0 + 1) * G); w0++) {
		if (chains && !chain_head(wl, w0, cfg.timeout)) continue;   // an earlier thread carries on into this window
		if (p.long_split && (p.long_all || chain_is_long(wl, w0, n_win, c.call_len, chains, cfg.timeout))) continue;   // winlong_kernel's
		DemodState s;
		bool have_lbi = false;   // s.last_bit_idx is the value the reference would hold (given the chain head's start state)
		for (uint32_t w = w0; w < n_win; w++) {
			if (w != w0 && (!chains || chain_head(wl, w, cfg.timeout))) break;
			const WinEntry e = wl[w];
			if (e.start >= c.call_len) break;
			WinRec rec;
			rec.frame_idx = -1;
			rec.flags = 0;
			rec.first_edge = rec.first_edge_block = 0;
			rec.pad = 0;
			rec.lbi_in = 0;
			rec.pad3 = 0;
			const bool cont = (e.flags & kWinCont) != 0;
			if (w == 0) {
				// the first window of a call starts from the true carried state: nothing to speculate
				s = st->d[demod];
				if (!cont && cfg.kind != K_TFA1) s.last_bit_idx = lbi_at_block(s.last_bit_idx, -1, (int)(e.start >> 13));
				if (cont && s.last_bit_idx) s.last_bit_idx -= kIdxPerBlock;   // demodulator::start for block 0
				rec.flags = kRecExact;
				have_lbi = true;
			} else if (w == w0) {
				memset(&s, 0, sizeof(s));
			}
			if (cfg.kind == K_TFA1) {
				run_tfa1_window(c, e, s, rec, cont);
			} else {
				bool far = true;
				if (w == 0) {
					far = false;
				} else if (w == w0 && c.ld) {
					// filter chains ran ahead (biq_kernel): the slicer reads their output, there is no filter state here
				} else if (w == w0) {
					// biquad warm-up over the samples of the preceding windows of this demod (they are the
					// filter's actual history); reaching window 0 means the true carried state can be used
					// (four timeouts.  The start-state error itself decays below one ulp within about one timeout - pole
					// radius 0.905 .. 0.946 - but the two trajectories then sit in a +-1 ulp dead band and only merge by
					// chance; measured on B200, windows that need the verifier's biquad-only repair / back-end time of a
					// 32 x 128 MiB call: 2 timeouts 21 % 2.04 ms, 2.5: 13 % 1.90, 3: 8 % 1.90, 3.5: 5 % 1.85, 4: 3.3 % 1.84)
					const uint32_t want = ((uint32_t)TFR_WARM_X4 * (uint32_t)cfg.timeout) / 4u;
					uint32_t have = 0;
					int v = (int)w;
					uint32_t from = e.start;
					while (v > 0 && have < want) {
						v--;
						const uint32_t len = wl[v].end - wl[v].start + 1;
						if (have + len >= want && v > 0) {
							from = wl[v].end + 1 - (want - have);
							have = want;
						} else {
							from = wl[v].start;
							have += len;
						}
					}
#ifdef TFR_WIN_PROFILE
					const long long pw0 = clock64();
#endif
					Biquad lp;
					lp.d1 = lp.d2 = lp.y0 = lp.y1 = 0.0;
					if (v == 0 && from == wl[0].start) lp = st->d[demod].lp;
					const BiquadCoef k = cfg.lp;
					// ONE loop for every lane of the warp: a step is an aligned group of four samples of the lane's current
					// history window, samples outside [m, b] masked.  (A loop nest per window - head, 16-sample chunks,
					// tail - ran the lanes of a warp through different loops at different times.)  One group is in flight
					// in registers, the line after next is prefetched into L1; a four-group register FIFO was measured
					// slower (156 registers: two CTAs fewer per SM).
					{
						const int4 *src4 = reinterpret_cast<const int4 *>(c.devfm);
						int u = v;
						uint32_t m = from, b = (u < (int)w) ? wl[u].end : 0u;
						int4 cur = make_int4(0, 0, 0, 0);
						if (u < (int)w) {
							cur = src4[m >> 2];
							prefetch_l1(c.devfm + (m & ~31u) + 32);
						}
						while (u < (int)w) {
							const uint32_t g = m & ~3u;
							const int4 v4 = cur;
							// where the next group is: the load goes out before the four filter steps
							int nu = u;
							uint32_t nm = g + 4, nb = b;
							if (nm > b) {
								nu = u + 1;
								if (nu < (int)w) {
									nm = wl[nu].start;
									nb = wl[nu].end;
									prefetch_l1(c.devfm + (nm & ~31u) + 32);
								}
							} else if ((g & 31u) == 0 && g + 64 <= b) {
								prefetch_l1(c.devfm + g + 64);   // the line after next
							}
							if (nu < (int)w) cur = src4[nm >> 2];
							if (g >= m) biquad_step(lp, k, int_to_double(v4.x));
							if (g + 1 >= m && g + 1 <= b) biquad_step(lp, k, int_to_double(v4.y));
							if (g + 2 >= m && g + 2 <= b) biquad_step(lp, k, int_to_double(v4.z));
							if (g + 3 <= b) biquad_step(lp, k, int_to_double(v4.w));
							u = nu;
							m = nm;
							b = nb;
						}
					}
					s.lp = lp;
#ifdef TFR_WIN_PROFILE
					prof_warm += clock64() - pw0;
#endif
				} else {
					// chain member: s is what the predecessor left (run_tfa2_window already did the end-of-window
					// reset); last_bit_idx moves to this window's first block (demodulator::start, decoder.cpp:118-122)
					far = !have_lbi;
					memset(s.rdata, 0, sizeof(s.rdata));   // every window starts from an empty frame buffer
					if (have_lbi) {
						const uint32_t plast = min(wl[w - 1].end, c.call_len - 1);
						s.last_bit_idx = lbi_at_block(s.last_bit_idx, (int)(plast >> 13), (int)(e.start >> 13));
						rec.flags |= kRecLbiIn;
						rec.lbi_in = s.last_bit_idx;
					}
				}
				rec.u_y0 = s.lp.y0;
				rec.u_y1 = s.lp.y1;
#ifdef TFR_WIN_PROFILE
				const long long pr0 = clock64();
#endif
				run_tfa2_window(c, cfg, e, s, rec, cont, far);
#ifdef TFR_WIN_PROFILE
				prof_run += clock64() - pr0;
				prof_steps += min(e.end, c.call_len - 1) - e.start + 1;
#endif
				if (rec.flags & kRecEdge) have_lbi = true;
			}
			if (rec.flags & kRecUnfinished) c.p->fin[(size_t)c.stream * kMaxDemods + demod] = s;
			rl[w] = rec;
			if (rec.flags & kRecUnfinished) break;
		}
	}
#ifdef TFR_WIN_PROFILE
	{
		const int kd = cfg.kind;
		atomicMax(&g_winprof[kd], (unsigned long long)prof_warm);
		atomicMax(&g_winprof[4 + kd], (unsigned long long)prof_run);
		atomicMax(&g_winprof[8 + kd], (unsigned long long)prof_steps);
		atomicMax(&g_winprof[12 + kd], (unsigned long long)(clock64() - prof_t0));
	}
#endif
}

// ------------------------------------------------------------------------------------------------
// winlong_kernel: one WARP per long window chain.
//
// Every lane holds the same copy of the demodulator state and runs the same control flow (warp-uniform); the lanes
// only differ inside a batch of 32 consecutive samples: coalesced loads, int->double, the filter's input-only
// terms, truncation, hash terms and the slicer's threshold tests are done by lane = sample; the filter recurrence -
// four dependent FP64 operations per sample, the only true chain - and the rare edge events (taken in order from a
// ballot mask) are executed by all lanes redundantly.  A sample then costs ~40 cycles instead of ~500.  The
// arithmetic, its order and every recorded value are those of run_tfa1_window / run_tfa2_window.
// ------------------------------------------------------------------------------------------------

// n <= 32 consecutive filter steps; dv = this lane's discriminator value devfm[base+lane] (0 for lane >= n), loaded by
// the caller one batch ahead (an L2/HBM round trip is about as long as the 32 serial steps).  Returns this lane's
// output y[base+lane]; lp moves to the state after the batch (uniform).
__device__ __forceinline__ double biquad_batch(Biquad &lp, const BiquadCoef &k, int dv, uint32_t n, int lane)
{
	const double xd = int_to_double(dv);
	double x1 = __shfl_up_sync(kFullMask, xd, 1), x2 = __shfl_up_sync(kFullMask, xd, 2);
	if (lane == 0) {
		x1 = lp.d1;
		x2 = lp.d2;
	} else if (lane == 1) {
		x2 = lp.d1;
	}
	// iir2::step as built (biquad_step): t2 = b0*x[n] + b1*x[n-1], t1 = b2*x[n-2] + a1*y[n-1], y = (t1 + t2) + a2*y[n-2]
	const double t2 = __dadd_rn(__dmul_rn(k.b0, xd), __dmul_rn(k.b1, x1));
	const double p2 = __dmul_rn(k.b2, x2);
	double y0 = lp.y0, y1 = lp.y1, mine = 0.0;
	if (n == 32) {
		// the common case, fully unrolled: the 64 broadcasts have constant source lanes and no dependence on the
		// recurrence, so they are all issued ahead of it and a step costs its four dependent FP64 operations
#pragma unroll
		for (int j = 0; j < 32; j++) {
			const double t2j = __shfl_sync(kFullMask, t2, j), p2j = __shfl_sync(kFullMask, p2, j);
			const double t1 = __dadd_rn(p2j, __dmul_rn(k.a1, y0));
			const double y = __dadd_rn(__dadd_rn(t1, t2j), __dmul_rn(k.a2, y1));
			y1 = y0;
			y0 = y;
			if (lane == j) mine = y;
		}
	} else {
		for (uint32_t j = 0; j < n; j++) {
			const double t2j = __shfl_sync(kFullMask, t2, (int)j), p2j = __shfl_sync(kFullMask, p2, (int)j);
			const double t1 = __dadd_rn(p2j, __dmul_rn(k.a1, y0));
			const double y = __dadd_rn(__dadd_rn(t1, t2j), __dmul_rn(k.a2, y1));
			y1 = y0;
			y0 = y;
			if (lane == (int)j) mine = y;
		}
	}
	const double xl = __shfl_sync(kFullMask, xd, (int)n - 1);
	const double xl2 = __shfl_sync(kFullMask, xd, n >= 2 ? (int)n - 2 : 0);
	lp.d2 = (n >= 2) ? xl2 : lp.d1;
	lp.d1 = xl;
	lp.y0 = y0;
	lp.y1 = y1;
	return mine;
}

static __device__ void run_tfa2_window_w(const WinCtx &c, const DemodCfg &cfg, const WinEntry &e, DemodState &s, WinRec &rec,
					 bool resume, bool far, int lane)
{
	if (!resume) {
		tfa2_reset(s);
		s.sr_cnt = -1;
		s.sr = 0;
		s.byte_cnt = 0;
	}
	const uint32_t last = min(e.end, c.call_len - 1);
	const bool taps = c.p->tap_cap != 0;
	const size_t tbase = ((size_t)c.stream * kMaxDemods + c.demod) * c.p->tap_cap;
	bool have_edge = false;
	rec.flags &= ~kRecEdge;
	Biquad lp = s.lp;
	const BiquadCoef k = cfg.lp;
	int bitcnt = s.bitcnt, dmin = s.dmin, dmax = s.dmax, offset = s.offset, last_bit = s.last_bit, rssi = s.rssi_i,
	    lbi = s.last_bit_idx;
	const int td_lo = cfg.td_lo, td_hi = cfg.td_hi;   // tfa2.cpp:391 `tdiff>spb/4 && tdiff<32*spb` for integer tdiff
	uint32_t ha = 0, hb = 0;   // this lane's share of the LdHash sums
	BitRuns br;
	br.n = 0;
	auto drain = [&]() {
		int i = 0, rem = 0, b = 0;
		for (;;) {
			if (rem == 0) {
				if (i == br.n) break;
				b = br.run[i] & 1;
				rem = br.run[i] >> 1;
				i++;
				if (rem == 0) continue;
			}
			tfa2_bit(s, b);
			rem--;
		}
		br.n = 0;
	};
	int noffset = __double2int_rz(__dmul_rn(0.9, (double)offset));
	int hi = noffset + dmax / 32, lo = noffset + dmin / 32;
	auto edge = [&](uint32_t m, int ld) {   // as in run_tfa2_window
		const int index = 2 * (int)(m & (kBlockDec - 1));
		const int bit = ld > hi ? 1 : 0;
		if (far && !have_edge) {
			rec.first_edge = index;
			rec.first_edge_block = (int)(m >> 13);
			rec.flags |= kRecEdge;
			bitcnt++;
			lbi = index;
		} else {
			if (index > lbi + 8) {
				bitcnt++;
				const int tdiff = index - lbi;
				if (tdiff >= td_lo && tdiff <= td_hi) {
					const int bit_diff = tdiff / 2;
					const int numbits = cfg.nbits[bit_diff];
					if (br.n + 2 > kBitRuns) drain();
					if (numbits < 32 && numbits > 1) br.run[br.n++] = (uint16_t)(((numbits - 1) << 1) | last_bit);
					br.run[br.n++] = (uint16_t)((1 << 1) | bit);
					last_bit = bit;
				}
			}
			if (index - lbi > 2) lbi = index;
		}
		have_edge = true;
	};
	// demodulator::start of every block that begins inside the window (not at its first sample), applied when the
	// walk reaches the block's first sample
	uint32_t nb = ((e.start >> 13) + 1u) << 13;
	auto cross = [&](uint32_t m) {
		while (nb <= m) {
			if (lbi) lbi -= kIdxPerBlock;
			nb += (uint32_t)kBlockDec;
		}
	};
	const bool pre = c.ld != nullptr;   // filter chains ran ahead: (int)y is read, not filtered here
	const int32_t *fsrc = pre ? c.ld : c.devfm;
	int dvn = (e.start + lane <= last) ? fsrc[e.start + lane] : 0;
	for (uint32_t base = e.start; base <= last; base += 32) {
		const uint32_t n = min(32u, last - base + 1);
		const bool on = (uint32_t)lane < n;
		const uint32_t m = base + lane;
		const int dv = dvn;
		dvn = (m + 32 <= last) ? fsrc[m + 32] : 0;   // the next batch, in flight during this one
		double y = 0.0;
		int ld = dv;
		if (!pre) {
			y = biquad_batch(lp, k, dv, n, lane);
			ld = trunc_to_int(y);
		}
		if (on && !pre) {
			const uint32_t i = m - e.start;
			ha += (uint32_t)ld * (2u * i + 1u);
			hb += (uint32_t)ld * (i * i + i + 1u);
			if (taps) {
				const uint32_t ti = e.cum + i;
				if (ti < c.p->tap_cap) {
					c.p->tap_i32[0][tbase + ti] = dv;
					c.p->tap_f64[tbase + ti] = y;
				}
			}
		}
		if (bitcnt < 10) {
			// the slicer levels still move (tfa2.cpp:365-374): this batch sample by sample, every lane the same steps
			const uint32_t cwl = on ? c.dec[m] : 0u;
			for (uint32_t j = 0; j < n; j++) {
				const int ldj = __shfl_sync(kFullMask, ld, (int)j);
				const uint32_t cw = __shfl_sync(kFullMask, cwl, (int)j);
				cross(base + j);
				if (bitcnt < 10) {
					if (ldj > dmax || ldj < dmin) {
						if (ldj > dmax) dmax = (7 * dmax + ldj) / 8;
						if (ldj < dmin) dmin = (7 * dmin + ldj) / 8;
						offset = (dmax + dmin) / 2;
						noffset = __double2int_rz(__dmul_rn(0.9, (double)offset));
						hi = noffset + dmax / 32;
						lo = noffset + dmin / 32;
					}
					if (bitcnt > 4) {
						const int i = (int)(int16_t)(cw & 0xffff), q = (int)(int16_t)(cw >> 16);
						const uint32_t sum = (uint32_t)rssi + (uint32_t)(i * i) + (uint32_t)(q * q);
						rssi = (int)((uint32_t)rssi + (uint32_t)((int)sum / 100));
					}
				}
				if ((ldj > hi || ldj < lo) && (int)(ldj > hi) != last_bit) edge(base + j, ldj);
			}
		} else {
			// levels frozen: the threshold tests of the 32 samples at once; an edge candidate is a sample beyond a level
			// whose bit differs from last_bit (which only an accepted edge changes), taken in order
			const bool valid = on && (ld > hi || ld < lo);
			const int bitl = ld > hi ? 1 : 0;
			unsigned done = 0;
			for (;;) {
				const unsigned mask = __ballot_sync(kFullMask, valid && bitl != last_bit) & ~done;
				if (!mask) break;
				const int j = __ffs((int)mask) - 1;
				done = (2u << j) - 1u;
				const int ldj = __shfl_sync(kFullMask, ld, j);
				cross(base + (uint32_t)j);
				edge(base + (uint32_t)j, ldj);
			}
		}
		cross(base + n - 1);
	}
	ha = __reduce_add_sync(kFullMask, ha);
	hb = __reduce_add_sync(kFullMask, hb);
	s.lp = lp;
	s.bitcnt = bitcnt; s.dmin = dmin; s.dmax = dmax; s.offset = offset; s.last_bit = last_bit; s.rssi_i = rssi;
	s.last_bit_idx = lbi;
	rec.e_y0 = lp.y0;
	rec.e_y1 = lp.y1;
	rec.ld_hash = ((unsigned long long)hb << 32) | ha;
	rec.lbi_end = lbi;
	rec.lbi_end_block = (int)(last >> 13);
	if (!far && have_edge) rec.flags |= kRecEdge;
	if (last == e.end) {
		if (br.n + 1 > kBitRuns) drain();
		br.run[br.n++] = (uint16_t)((16 << 1) | last_bit);   // 16 x store_bit(last_bit) before the flush (tfa2.cpp:430-433)
	}
	drain();
	if (last == e.end) {
		const bool gate = (cfg.kind == K_TX22) ? (s.byte_cnt >= 7 && s.byte_cnt < 64) : (s.byte_cnt >= 7);
		int fi = -1;
		if (lane == 0) {
			if (gate) fi = put_frame(c, rec.frame_idx, s, e.end, (double)s.rssi_i, s.offset);
			else drop_frame(c, rec.frame_idx);
		}
		fi = __shfl_sync(kFullMask, fi, 0);
		rec.frame_idx = fi;
		{   // every lane holds the same state; lane 0 owns the slot
			int np = rec.pad3;
			if (lane == 0) np = window_notice(c, rec.pad3, s.inv_cnt, e.end);
			rec.pad3 = __shfl_sync(kFullMask, np, 0);
		}
		s.sr_cnt = -1;
		s.sr = 0;
		s.byte_cnt = 0;
		tfa2_reset(s);
		s.timeout_cnt = 0;
		rec.flags = (rec.flags & (kRecExact | kRecEdge | kRecLbiIn)) | kRecRan;
	} else {
		int np = rec.pad3;
		if (lane == 0) np = window_notice(c, rec.pad3, s.inv_cnt, last);
		rec.pad3 = __shfl_sync(kFullMask, np, 0);
		s.inv_cnt = 0;
		s.timeout_cnt = (int)(e.end - last);
		rec.flags = (rec.flags & (kRecExact | kRecEdge | kRecLbiIn)) | kRecRan | kRecUnfinished;
	}
}

static __device__ void run_tfa1_window_w(const WinCtx &c, const WinEntry &e, DemodState &s, WinRec &rec, bool resume, int lane)
{
	if (!resume) {
		s.mark_lvl = 0;
		s.rssi_i = 0;
		s.last_bit_idx = 0;
		s.sr_cnt = -1;
		s.byte_cnt = 0;
		s.rdata[10] = 0;
	}
	uint32_t head = 0;
	int nbits = 0;
	int mark = s.mark_lvl, rssi = s.rssi_i, lbi = s.last_bit_idx;
	BitRuns br;
	br.n = 0;
	auto drain = [&]() {
		int i = 0, rem = 0, b = 0;
		for (;;) {
			if (rem == 0) {
				if (i == br.n) break;
				b = br.run[i] & 1;
				rem = br.run[i] >> 1;
				i++;
				if (rem == 0) continue;
			}
			if (nbits < 31) head |= (uint32_t)b << nbits;
			nbits++;
			tfa1_bit(s, b);
			rem--;
		}
		br.n = 0;
	};
	const uint32_t last = min(e.end, c.call_len - 1);
	const bool taps = c.p->tap_cap != 0;
	int32_t *tap = taps ? c.p->tap_i32[1] + ((size_t)c.stream * kMaxDemods + c.demod) * c.p->tap_cap : nullptr;
	uint32_t lw0 = (e.start == 0) ? c.prev_last : c.dec[e.start - 1];   // the sample before the batch
	uint32_t nb = ((e.start >> 13) + 1u) << 13;
	uint32_t cwn = (e.start + lane <= last) ? c.dec[e.start + lane] : 0u;
	for (uint32_t base = e.start; base <= last; base += 32) {
		const uint32_t n = min(32u, last - base + 1);
		const bool on = (uint32_t)lane < n;
		const uint32_t m = base + lane;
		const uint32_t cw = cwn;
		cwn = (m + 32 <= last) ? c.dec[m + 32] : 0u;   // the next batch, in flight during this one
		uint32_t lw = __shfl_up_sync(kFullMask, cw, 1);
		if (lane == 0) lw = lw0;
		lw0 = __shfl_sync(kFullMask, cw, (int)n - 1);
		const int dev = fm_dev_nrzs((int)(int16_t)(cw & 0xffff), (int)(int16_t)(cw >> 16), (int)(int16_t)(lw & 0xffff),
					    (int)(int16_t)(lw >> 16));
		if (taps && on) {
			const uint32_t ti = e.cum + (m - e.start);
			if (ti < c.p->tap_cap) tap[ti] = dev;
		}
		// the peak-hold / decay recurrence and the dip test chain every sample to the one before: all lanes, in order
		for (uint32_t j = 0; j < n; j++) {
			const int dj = __shfl_sync(kFullMask, dev, (int)j);
			const uint32_t mj = base + j;
			while (nb <= mj) {
				if (lbi) lbi -= kIdxPerBlock;
				nb += (uint32_t)kBlockDec;
			}
			const int index = 2 * (int)(mj & (kBlockDec - 1));
			if (dj > mark) mark = dj;
			else mark = __double2int_rz(__dmul_rn((double)mark, 0.95));
			if (mark > rssi) rssi = mark;
			if (dj < mark / 2) {
				if (lbi) {
					const int gap = index - lbi;
					if (gap > 4) {
						for (int ones = (gap >= 22) ? (gap - 22) / 20 + 1 : 0; ones > 0; ones -= 32767) {   // as in run_tfa1_window
							if (br.n + 2 > kBitRuns) drain();
							br.run[br.n++] = (uint16_t)((min(ones, 32767) << 1) | 1);
						}
						if (br.n + 1 > kBitRuns) drain();
						br.run[br.n++] = (uint16_t)(1 << 1);
					}
				}
				if (index - lbi > 2) lbi = index;
			}
		}
	}
	drain();
	s.mark_lvl = mark;
	s.rssi_i = rssi;
	s.last_bit_idx = lbi;
	rec.head31 = head;
	rec.nbits = nbits;
	rec.sr_final = s.sr;
	if (last == e.end) {
		int fi = -1;
		if (lane == 0) {
			if (s.byte_cnt >= 10) fi = put_frame(c, rec.frame_idx, s, e.end, (double)s.rssi_i, 0);
			else drop_frame(c, rec.frame_idx);
		}
		fi = __shfl_sync(kFullMask, fi, 0);
		rec.frame_idx = fi;
		rec.flags = (rec.flags & kRecExact) | kRecRan;
	} else {
		s.timeout_cnt = (int)(e.end - last);
		rec.flags = (rec.flags & kRecExact) | kRecRan | kRecUnfinished;
	}
}

constexpr int kLongWarps = 4;
__global__ void __launch_bounds__(32 * kLongWarps) winlong_kernel(const BackParams p)
{
	const int lane = threadIdx.x & 31;
	const int stream = blockIdx.y;
	const int demod = (int)gridDim.z - 1 - (int)blockIdx.z;
	const DemodCfg &cfg = p.cfg->d[demod];
	if (cfg.kind == K_WHB) return;
	const StreamJob job = p.jobs[stream];
	if (job.n_blocks == 0) return;
	StreamState *st = p.st + stream;
	uint32_t w_lo, n_win;   // this launch's windows: [w_lo, n_win)
	part_range(p, stream, demod, w_lo, n_win);
	const WinEntry *wl = p.wins + job.win_off + (size_t)demod * job.win_cap;
	WinRec *rl = p.recs + job.win_off + (size_t)demod * job.win_cap;
	const WinCtx c = make_ctx(p, stream, demod, job, st);
	const bool chains = (cfg.kind != K_TFA1);
	const uint32_t n_warps = gridDim.x * kLongWarps;
	for (uint32_t w0 = w_lo + blockIdx.x * kLongWarps + (threadIdx.x >> 5); w0 < n_win; w0 += n_warps) {   // warp-uniform
		if (chains && !chain_head(wl, w0, cfg.timeout)) continue;
		if (!p.long_all && !chain_is_long(wl, w0, n_win, c.call_len, chains, cfg.timeout)) continue;
		// the chain loop of win_kernel, every lane the same
		DemodState s;
		bool have_lbi = false;
		for (uint32_t w = w0; w < n_win; w++) {
			if (w != w0 && (!chains || chain_head(wl, w, cfg.timeout))) break;
			const WinEntry e = wl[w];
			if (e.start >= c.call_len) break;
			WinRec rec;
			rec.frame_idx = -1;
			rec.flags = 0;
			rec.first_edge = rec.first_edge_block = 0;
			rec.pad = 0;
			rec.lbi_in = 0;
			rec.pad3 = 0;
			rec.head31 = rec.sr_final = 0;
			rec.nbits = 0;
			rec.lbi_end = rec.lbi_end_block = 0;
			rec.u_y0 = rec.u_y1 = rec.e_y0 = rec.e_y1 = 0.0;
			rec.ld_hash = 0;
			rec.pad3 = 0;
			const bool cont = (e.flags & kWinCont) != 0;
			if (w == 0) {
				s = st->d[demod];
				if (!cont && cfg.kind != K_TFA1) s.last_bit_idx = lbi_at_block(s.last_bit_idx, -1, (int)(e.start >> 13));
				if (cont && s.last_bit_idx) s.last_bit_idx -= kIdxPerBlock;
				rec.flags = kRecExact;
				have_lbi = true;
			} else if (w == w0) {
				memset(&s, 0, sizeof(s));
			}
			if (cfg.kind == K_TFA1) {
				run_tfa1_window_w(c, e, s, rec, cont, lane);
			} else {
				bool far = true;
				if (w == 0) {
					far = false;
				} else if (w == w0 && c.ld) {
					// filter chains ran ahead: nothing to warm up
				} else if (w == w0) {
					// the same warm-up history as win_kernel, filtered in batches of 32
					const uint32_t want = ((uint32_t)TFR_WARM_X4 * (uint32_t)cfg.timeout) / 4u;
					uint32_t have = 0;
					int v = (int)w;
					uint32_t from = e.start;
					while (v > 0 && have < want) {
						v--;
						const uint32_t len = wl[v].end - wl[v].start + 1;
						if (have + len >= want && v > 0) {
							from = wl[v].end + 1 - (want - have);
							have = want;
						} else {
							from = wl[v].start;
							have += len;
						}
					}
					Biquad lp;
					lp.d1 = lp.d2 = lp.y0 = lp.y1 = 0.0;
					if (v == 0 && from == wl[0].start) lp = st->d[demod].lp;
					for (int u = v; u < (int)w; u++) {
						const uint32_t a = (u == v) ? from : wl[u].start, b = wl[u].end;
						int dvn = (a + lane <= b) ? c.devfm[a + lane] : 0;
						for (uint32_t base = a; base <= b; base += 32) {
							const int dv = dvn;
							dvn = (base + 32 + lane <= b) ? c.devfm[base + 32 + lane] : 0;
							biquad_batch(lp, cfg.lp, dv, min(32u, b - base + 1), lane);
						}
					}
					s.lp = lp;
				} else {
					far = !have_lbi;
					memset(s.rdata, 0, sizeof(s.rdata));
					if (have_lbi) {
						const uint32_t plast = min(wl[w - 1].end, c.call_len - 1);
						s.last_bit_idx = lbi_at_block(s.last_bit_idx, (int)(plast >> 13), (int)(e.start >> 13));
						rec.flags |= kRecLbiIn;
						rec.lbi_in = s.last_bit_idx;
					}
				}
				rec.u_y0 = s.lp.y0;
				rec.u_y1 = s.lp.y1;
				run_tfa2_window_w(c, cfg, e, s, rec, cont, far, lane);
				if (rec.flags & kRecEdge) have_lbi = true;
			}
			if (lane == 0) {
				if (rec.flags & kRecUnfinished) c.p->fin[(size_t)c.stream * kMaxDemods + demod] = s;
				rl[w] = rec;
			}
			if (rec.flags & kRecUnfinished) break;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// carried-state reconstruction from window records (valid when every earlier window has been proven)
// ------------------------------------------------------------------------------------------------
struct Lbi { int v, block; };
// last_bit_idx after window v: the newest window at or before v that either saw an edge candidate or is
// window 0 (which ran from the true state, so its recorded value is true even without an edge)
__device__ __forceinline__ Lbi lbi_after(const WinRec *rl, int v, const DemodState &carry0)
{
	for (; v >= 0; v--)
		if (v == 0 || (rl[v].flags & kRecEdge)) return Lbi{ rl[v].lbi_end, rl[v].lbi_end_block };
	return Lbi{ carry0.last_bit_idx, -1 };
}
// biquad state after window v (d1/d2 are the window's last discriminator values)
__device__ __forceinline__ Biquad biquad_after(const WinCtx &c, const WinEntry *wl, const WinRec *rl, int v, const DemodState &carry0)
{
	if (v < 0) return carry0.lp;
	Biquad lp;
	const WinEntry e = wl[v];
	const uint32_t last = min(e.end, c.call_len - 1);
	lp.y0 = rl[v].e_y0;
	lp.y1 = rl[v].e_y1;
	lp.d1 = (double)c.devfm[last];
	if (last >= e.start + 1) {
		lp.d2 = (double)c.devfm[last - 1];
	} else if (v > 0) {   // one-sample window: d2 is the previous window's last discriminator value
		lp.d2 = (double)c.devfm[min(wl[v - 1].end, c.call_len - 1)];
	} else {
		lp.d2 = carry0.lp.d1;
	}
	return lp;
}
// decoder shift register before TFA_1 window w: the newest 32 bits pushed by earlier windows
__device__ __forceinline__ uint32_t sr_before(const WinRec *rl, int w, const DemodState &carry0)
{
	uint32_t sr = 0;
	int nb = 0;   // bits of sr (from the top) already determined by newer windows
	for (int v = w - 1; v >= 0 && nb < 32; v--) {
		const WinRec &r = rl[v];
		sr |= (nb == 0) ? r.sr_final : (r.sr_final >> nb);
		nb += (v == 0) ? 32 : min(r.nbits, 32);   // window 0 ran from the true register: it is complete
	}
	if (nb < 32) sr |= (nb == 0) ? carry0.sr : (carry0.sr >> nb);
	return sr;
}
__device__ __forceinline__ bool tfa1_sync_same(const WinRec &rec, uint32_t sr_true)
{
	// would any sync decision among the first 31 shifts differ with the true shift register?
	uint32_t a = 0, b = sr_true;
	bool same = true;
	const int n = min(rec.nbits, 31);
	for (int k = 0; k < n; k++) {
		const uint32_t bitv = (rec.head31 >> k) & 1u;
		a = (a >> 1) | (bitv << 31);
		b = (b >> 1) | (bitv << 31);
		same &= ((a & 0xffff) == 0xd42d) == ((b & 0xffff) == 0xd42d);
	}
	return same;
}
__device__ __forceinline__ bool tfa2_edge_same(const WinRec &rec, const DemodCfg &cfg, Lbi l, uint32_t start)
{
	// a run that was handed an explicit last_bit_idx is right iff that value is the true one
	if (rec.flags & kRecLbiIn) return rec.lbi_in == lbi_at_block(l.v, l.block, (int)(start >> 13));
	if (!(rec.flags & kRecEdge)) return true;
	// the first edge candidate must behave the same with the true last_bit_idx
	const int v = lbi_at_block(l.v, l.block, rec.first_edge_block);
	const int tdiff = rec.first_edge - v;
	const bool c1 = rec.first_edge > v + 8;
	const bool c2 = (double)tdiff > __dmul_rn(cfg.spb, 0.25) && (double)tdiff < __dmul_rn(32.0, cfg.spb);
	return c1 && !c2 && (tdiff > 2);
}

// the biquad alone over a window from a given (true) start state: returns the hash of (int)y and leaves the
// end state in lp.  This is what the verifier runs when a window's assumed biquad state was not bitwise the
// true one: if the slicer inputs hash the same, everything the window produced stands.
static __device__ __forceinline__ unsigned long long biquad_only(const WinCtx &c, const DemodCfg &cfg, const WinEntry &e, Biquad &lp_io)
{
	Biquad lp = lp_io;   // keep the recurrence in registers (a by-reference state would live in local memory)
	const uint32_t last = min(e.end, c.call_len - 1);
	const BiquadCoef k = cfg.lp;
	const bool taps = c.p->tap_cap != 0;
	const size_t tbase = ((size_t)c.stream * kMaxDemods + c.demod) * c.p->tap_cap;
	LdHash hash;
	hash.init();
	auto step = [&](uint32_t m, int dv) {
		const double y = biquad_step(lp, k, int_to_double(dv));
		if (taps) {
			const uint32_t ti = e.cum + (m - e.start);
			if (ti < c.p->tap_cap) c.p->tap_f64[tbase + ti] = y;
		}
		hash.add(trunc_to_int(y));
	};
	uint32_t m = e.start;
	for (; (m & 15u) && m <= last; m++) step(m, c.devfm[m]);            // head up to the first 16-aligned sample
	if (m + 15 <= last) {                                                 // whole chunks: no per-sample predicate
		Chunk16 ck = load16(c.devfm + m);
		for (; m + 15 <= last; m += 16) {
			Chunk16 nx = ck;
			if (m + 31 <= last) nx = load16(c.devfm + m + 16);   // in flight while the current chunk is filtered
			if (m + 63 <= last) prefetch_l1(c.devfm + m + 48);   // a chunk is ~600 cycles of filter, an L2 round trip more
#pragma unroll
			for (int kk = 0; kk < 16; kk++) step(m + kk, ck.v[kk]);
			ck = nx;
		}
	}
	for (; m <= last; m++) step(m, c.devfm[m]);                          // tail
	lp_io = lp;
	return hash.value();
}

// ------------------------------------------------------------------------------------------------
// verify_kernel: one CTA per (stream, demod) - the exactness backstop, run in parallel rounds.
//
// A window's record is TRUE when it was produced from the start state the reference would have had.  Window
// 0 ran from the true carried state.  A later window is `good` when the carry-in its run assumed is
// equivalent to what its predecessor's record left behind:
//     TFA_1         the sync decisions of its first 31 shifts are the same with the predecessor chain's shift
//                   register as with the empty one it assumed (or it was re-run with exactly that register)
//     TFA_2 family  its assumed biquad outputs are bitwise the predecessor's end state, and its first edge
//                   candidate behaves the same with the predecessor chain's last_bit_idx
// If every window is good, every record is true by induction from window 0.  Each round recomputes the flags
// (all threads), collects the RUN STARTS (a bad window after a good one) and repairs every run in parallel,
// one thread per run, following the repaired state into the successors for as long as they disagree with it:
//     biquad state differs  -> re-run the biquad alone; if the slicer inputs (int)y hash the same, everything
//                              the window produced stands and it adopts the new start/end state; else full re-run
//     edge / shift register -> full re-run of the window from the predecessor chain's state
// Only the first run of a round is certain to start from a true state; a later run may be repaired from a
// stale predecessor and is then simply flagged again in the next round.  At least one window becomes true per
// round, so the loop ends, and it ends only when every link has been proven.  Finally thread 0 writes the
// state carried into the next call.
// ------------------------------------------------------------------------------------------------
#ifndef TFR_VERIFY_THREADS
#define TFR_VERIFY_THREADS 512
#endif
constexpr int kVerifyThreads = TFR_VERIFY_THREADS;
struct VerifyCounts { uint32_t cheap, full, sr, rounds; };
#ifdef TFR_VER_PROFILE
__device__ unsigned long long g_verprof[8];   // max over CTAs: cycles in flags / run starts / repair phases, rounds, whole CTA; max thread repair cycles, windows
#endif

__global__ void __launch_bounds__(kVerifyThreads) verify_kernel(const BackParams p)
{
	const int gid = blockIdx.x;
	const int tid = threadIdx.x;
	const int nd = p.cfg->n_demods;
	if (gid >= p.n_streams * nd) return;
	const int stream = gid / nd, demod = gid % nd;
	const DemodCfg &cfg = p.cfg->d[demod];
	if (cfg.kind == K_WHB) return;
	const StreamJob job = p.jobs[stream];
	if (job.n_blocks == 0) return;
	StreamState *st = p.st + stream;
	const uint32_t n_win = p.wincnt[stream].n[demod];
	const WinEntry *wl = p.wins + job.win_off + (size_t)demod * job.win_cap;
	WinRec *rl = p.recs + job.win_off + (size_t)demod * job.win_cap;
	const WinCtx c = make_ctx(p, stream, demod, job, st);
	const int last_block = (int)job.n_blocks - 1;
	const DemodState &carry0 = st->d[demod];   // read in place; overwritten only by the final store below
	const bool pre = c.ld != nullptr;
	__shared__ uint32_t s_runs[kVerifyThreads];
	__shared__ uint32_t s_nruns, s_first;
	uint32_t n_cheap = 0, n_full = 0, n_sr = 0, rounds = 0;
#ifdef TFR_VER_PROFILE
	long long vp_flags = 0, vp_starts = 0, vp_repair = 0, vp_t0 = clock64(), vp_mine = 0, vp_wins = 0;
#endif

	// is window w (> 0) consistent with the records before it?  edge_bad reports a wrong last_bit_idx assumption
	auto good = [&](uint32_t w, const WinRec &rec, bool &edge_bad) -> bool {
		edge_bad = false;
		if (cfg.kind == K_TFA1) {
			if (rec.flags & kRecExact) return true;
			const uint32_t sr = sr_before(rl, (int)w, carry0);
			if (rec.flags & kRecLbiIn) return (uint32_t)rec.lbi_in == sr;
			return tfa1_sync_same(rec, sr);
		}
		const WinRec &pr = rl[w - 1];
		// (with filter chains the slicer inputs were proven by biq_verify_kernel: only last_bit_idx is left to check)
		const bool bq = pre || ((__double_as_longlong(rec.u_y0) == __double_as_longlong(pr.e_y0)) &&
					(__double_as_longlong(rec.u_y1) == __double_as_longlong(pr.e_y1)));
		edge_bad = !tfa2_edge_same(rec, cfg, lbi_after(rl, (int)w - 1, carry0), wl[w].start);
		return bq && !edge_bad;
	};

	for (;;) {
#ifdef TFR_VER_PROFILE
		long long vp_a = clock64();
#endif
		if (tid == 0) {
			s_nruns = 0;
			s_first = 0xffffffffu;
		}
		// ---- flags
		for (uint32_t w = tid; w < n_win; w += kVerifyThreads) {
			int pad = kPadOk;
			if (w > 0) {
				bool eb;
				if (!good(w, rl[w], eb)) pad = 0;
			}
			rl[w].pad = pad;
		}
		__syncthreads();
#ifdef TFR_VER_PROFILE
		vp_flags += clock64() - vp_a; vp_a = clock64();
#endif
		// ---- run starts
		for (uint32_t w = 1 + tid; w < n_win; w += kVerifyThreads) {
			if (!(rl[w].pad & kPadOk) && (rl[w - 1].pad & kPadOk)) {
				const uint32_t k = atomicAdd(&s_nruns, 1u);
				if (k < (uint32_t)kVerifyThreads) s_runs[k] = w;   // the rest waits for the next round
				atomicMin(&s_first, w);
			}
		}
		__syncthreads();
		const uint32_t n_runs = min(s_nruns, (uint32_t)kVerifyThreads);
		if (n_runs == 0) break;
		// the first run is the one whose repair is certainly final: it must be in every round (a duplicate entry
		// only repeats the same deterministic work)
		if (tid == 0 && s_nruns > (uint32_t)kVerifyThreads) s_runs[0] = s_first;
		__syncthreads();
		rounds++;
#ifdef TFR_VER_PROFILE
		vp_starts += clock64() - vp_a; vp_a = clock64();
#endif
		// ---- repair, one thread per run
		if ((uint32_t)tid < n_runs) {
			const uint32_t w = s_runs[tid];
			if (cfg.kind == K_TFA1) {
				for (uint32_t v = w; v < n_win; v++) {
					WinRec rec = rl[v];
					if (v != w && !(rec.pad & kPadOk) && (rl[v - 1].pad & kPadOk)) break;   // another thread's run
					bool eb;
					if (good(v, rec, eb)) break;
					const uint32_t sr = sr_before(rl, (int)v, carry0);
					DemodState s;
					memset(&s, 0, sizeof(s));
					s.sr = sr;
					run_tfa1_window(c, wl[v], s, rec, false);
					rec.flags |= kRecLbiIn;
					rec.lbi_in = (int32_t)sr;
					if (rec.flags & kRecUnfinished) c.p->fin[(size_t)c.stream * kMaxDemods + demod] = s;
					rl[v] = rec;
					n_sr++;
				}
			} else {
				Biquad lp;
				lp.d1 = lp.d2 = lp.y0 = lp.y1 = 0.0;
				if (!pre) lp = biquad_after(c, wl, rl, (int)w - 1, carry0);
				for (uint32_t v = w; v < n_win; v++) {
					WinRec rec = rl[v];
					if (v != w && !(rec.pad & kPadOk) && (rl[v - 1].pad & kPadOk)) break;   // another thread's run
					const WinEntry e = wl[v];
					const Lbi l = lbi_after(rl, (int)v - 1, carry0);
					const bool bq_ok = pre || ((__double_as_longlong(rec.u_y0) == __double_as_longlong(lp.y0)) &&
								   (__double_as_longlong(rec.u_y1) == __double_as_longlong(lp.y1)));
					const bool edge_ok = tfa2_edge_same(rec, cfg, l, e.start);
					if (bq_ok && edge_ok) break;   // consistent from here on
					bool full = !edge_ok;
					if (!full) {
						Biquad t = lp;
						if (biquad_only(c, cfg, e, t) == rec.ld_hash) {
							rec.u_y0 = lp.y0;
							rec.u_y1 = lp.y1;
							rec.e_y0 = t.y0;
							rec.e_y1 = t.y1;
							if (rec.flags & kRecUnfinished) c.p->fin[(size_t)c.stream * kMaxDemods + demod].lp = t;
							lp = t;
							n_cheap++;
						} else {
							full = true;
						}
					}
					if (full) {
						DemodState s;
						memset(&s, 0, sizeof(s));
						s.lp = lp;
						s.last_bit_idx = lbi_at_block(l.v, l.block, (int)(e.start >> 13));
						rec.lbi_in = s.last_bit_idx;
						rec.u_y0 = lp.y0;
						rec.u_y1 = lp.y1;
						rec.flags &= ~kRecEdge;
						rec.flags |= kRecLbiIn;
						run_tfa2_window(c, cfg, e, s, rec, false, false);
						if (rec.flags & kRecUnfinished) c.p->fin[(size_t)c.stream * kMaxDemods + demod] = s;
						lp = s.lp;
						n_full++;
					}
					rl[v] = rec;
#ifdef TFR_VER_PROFILE
					vp_wins++;
#endif
					if (rec.flags & kRecUnfinished) break;
				}
			}
#ifdef TFR_VER_PROFILE
			vp_mine += clock64() - vp_a;
#endif
		}
		__syncthreads();
#ifdef TFR_VER_PROFILE
		vp_repair += clock64() - vp_a;
#endif
	}
#ifdef TFR_VER_PROFILE
	atomicMax(&g_verprof[5], (unsigned long long)vp_mine);
	atomicMax(&g_verprof[6], (unsigned long long)vp_wins);
	if (tid == 0 && gid < 5)
		printf("verify[%d]: windows %u rounds %u total %lld flags %lld starts %lld repair %lld | thread 0: cheap %u full %u sr %u\n", gid, n_win, rounds,
		       clock64() - vp_t0, vp_flags, vp_starts, vp_repair, n_cheap, n_full, n_sr);
	if (tid == 0) {
		atomicMax(&g_verprof[0], (unsigned long long)vp_flags); atomicMax(&g_verprof[1], (unsigned long long)vp_starts);
		atomicMax(&g_verprof[2], (unsigned long long)vp_repair); atomicMax(&g_verprof[3], (unsigned long long)rounds);
		atomicMax(&g_verprof[4], (unsigned long long)(clock64() - vp_t0));
	}
#endif

	if (n_cheap) atomicAdd(&p.counters->par_cheap, n_cheap);
	if (n_full) atomicAdd(&p.counters->ver_full, n_full);
	if (n_sr) atomicAdd(&p.counters->rerun_sr, n_sr);
	if (n_full + n_sr) atomicAdd(&p.counters->n_reruns, n_full + n_sr);
	if (tid != 0) return;
	if (rounds) atomicAdd(&p.counters->ver_checked, rounds);

	bool unfinished = false;
	if (n_win) unfinished = (rl[n_win - 1].flags & kRecUnfinished) != 0;
	if (cfg.kind == K_TFA1) {
		const uint32_t sr_end = sr_before(rl, (int)n_win, carry0);
		DemodState &out = st->d[demod];
		if (unfinished) {
			out = c.p->fin[(size_t)c.stream * kMaxDemods + demod];
		} else {
			// tfa1.cpp:179-184 + :115-117: everything but the shift register is reset by the flush
			out.mark_lvl = out.rssi_i = out.last_bit_idx = out.timeout_cnt = 0;
			out.sr_cnt = -1;
			out.byte_cnt = 0;
			out.rdata[10] = 0;
		}
		out.sr = sr_end;
	} else {
		const Lbi l = lbi_after(rl, (int)n_win - 1, carry0);
		const Biquad lp_end = pre ? st->lp_next[demod] : biquad_after(c, wl, rl, (int)n_win - 1, carry0);
		const int lbi_end = lbi_at_block(l.v, l.block, last_block);
		DemodState &out = st->d[demod];
		if (unfinished) {
			out = c.p->fin[(size_t)c.stream * kMaxDemods + demod];
			out.lp = lp_end;
		} else {
			tfa2_reset(out);
			out.timeout_cnt = 0;
			out.sr_cnt = -1;
			out.sr = 0;
			out.byte_cnt = 0;
			out.lp = lp_end;
		}
		out.last_bit_idx = lbi_end;
	}
	if (n_win) atomicAdd(&p.counters->n_windows, (unsigned long long)n_win);
	if (p.tap_cap) {
		uint32_t *tc = p.tap_cnt + ((size_t)stream * kMaxDemods + demod) * 3;
		const uint32_t n = p.wincnt[stream].cum[demod];
		// taps are written at (cum + offset): the number valid is the demod's active-sample count clipped to the call
		uint32_t valid = n;
		if (n_win) {
			const WinEntry e = wl[n_win - 1];
			if (e.end >= c.call_len) valid = n - (e.end - (c.call_len - 1));
		}
		if (cfg.kind == K_TFA1) {
			tc[1] = valid;
		} else {
			tc[0] = valid;
			tc[2] = valid;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
cudaError_t launch_thresh2(const BackParams &p, cudaStream_t s)
{
	if (p.walk_tab && p.n_tiles > 0) walk_table_kernel<<<dim3((p.n_tiles + 3) / 4, p.n_streams), 128, 0, s>>>(p);
	if (p.walk_tab && p.walk_gap) {
		const bool wide = p.walk_ct ? p.walk_ct >= 256 : true;   // (64 measured slower at 8, 16 and 32 streams: profiles/r2_walk_table.txt)
		if (wide) walk_cta_kernel<256><<<p.n_streams, 256, 0, s>>>(p);
		else walk_cta_kernel<64><<<p.n_streams, 64, 0, s>>>(p);
	}
	else thresh2_kernel<<<p.n_streams, 32, 0, s>>>(p);
	return cudaGetLastError();
}
cudaError_t launch_devfm(const BackParams &p, cudaStream_t s)
{
	if (p.n_tiles <= 0) return cudaSuccess;
	devfm_kernel<<<dim3(p.n_tiles, p.n_streams), 128, 0, s>>>(p);
	return cudaGetLastError();
}
cudaError_t launch_biq(const BackParams &p, int n_demods, cudaStream_t s)
{
	if (p.max_blocks <= 0) return cudaSuccess;
	int gx = (p.max_blocks * 2 / kBiqK + kBiqThreads - 1) / kBiqThreads;   // about one window per block and demodulator
	gx = gx < 1 ? 1 : (gx > 256 ? 256 : gx);
#ifdef TFR_BIQ_PROFILE
	{
		unsigned long long z[8] = { 0 }, r[8];
		cudaStreamSynchronize(s);
		cudaMemcpyToSymbol(g_biqprof, z, sizeof(z));
		biq_kernel<<<dim3(gx, p.n_streams, n_demods), kBiqThreads, 0, s>>>(p);
		cudaStreamSynchronize(s);
		cudaMemcpyFromSymbol(r, g_biqprof, sizeof(r));
		fprintf(stderr, "[biqprof] %llu chains, %.0f steps and %.0f cycles on average (%.1f cycles per step); slowest chain %llu cycles for %llu steps; longest chain %llu steps in %llu cycles\n",
			r[4], (double)r[3] / (r[4] + 1), (double)r[2] / (r[4] + 1), (double)r[2] / (r[3] + 1), r[0] >> 20, r[0] & 0xfffff, r[5] >> 32, r[5] & 0xffffffffull);
		biq_verify_kernel<<<p.n_streams * n_demods, kBiqVerThreads, 0, s>>>(p);
		return cudaGetLastError();
	}
#endif
	biq_kernel<<<dim3(gx, p.n_streams, n_demods), kBiqThreads, 0, s>>>(p);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return e;
	biq_verify_kernel<<<p.n_streams * n_demods, kBiqVerThreads, 0, s>>>(p);
	return cudaGetLastError();
}
cudaError_t launch_devfm_first(const BackParams &p, cudaStream_t s)
{
	devfm_first_kernel<<<(p.n_streams + 63) / 64, 64, 0, s>>>(p);
	return cudaGetLastError();
}
cudaError_t launch_devfm_win(const BackParams &p, cudaStream_t s)
{
	if (p.max_blocks <= 0) return cudaSuccess;
	int gx = p.max_blocks;   // about one window per block and demodulator
	gx = gx < 1 ? 1 : (gx > 2048 ? 2048 : gx);
	devfm_win_kernel<<<dim3(gx, p.n_streams), 128, 0, s>>>(p);
	return cudaGetLastError();
}
static dim3 win_grid(const BackParams &p, int n_demods, int threads)
{
	// windows per (stream, demod) are at most a few per block
	int gx = (p.n_tiles * 2 + threads - 1) / threads;   // n_tiles: the blocks this launch's windows come from
	gx = gx < 1 ? 1 : (gx > 512 ? 512 : gx);
	return dim3(gx, p.n_streams, n_demods);
}
cudaError_t launch_win(const BackParams &p, int n_demods, cudaStream_t s)
{
#ifdef TFR_WIN_PROFILE
	{
		unsigned long long z[16] = { 0 }, r[16];
		cudaStreamSynchronize(s);
		cudaMemcpyToSymbol(g_winprof, z, sizeof(z));
		cudaMemcpyToSymbol(g_slprof, z, 8 * sizeof(unsigned long long));
		win_kernel<<<win_grid(p, n_demods, kWinThreads), kWinThreads, 0, s>>>(p);
		cudaStreamSynchronize(s);
		cudaMemcpyFromSymbol(r, g_winprof, sizeof(r));
		unsigned long long q[8];
		cudaMemcpyFromSymbol(q, g_slprof, sizeof(q));
		fprintf(stderr, "[slprof] all tfa2 windows (%llu): %llu fast groups (%.0f cyc each), %llu general groups (%.0f cyc each), %.0f cyc per window\n", q[5], q[0], (double)q[2] / (q[0] + 1), q[1], (double)q[3] / (q[1] + 1), (double)q[4] / (q[5] + 1));
		for (int k = 0; k < 4; k++)
			fprintf(stderr, "[winprof] kind %d: max warm-up %llu cyc, max slicer %llu cyc over max %llu steps, max thread %llu cyc\n", k, r[k], r[4 + k], r[8 + k], r[12 + k]);
		return cudaGetLastError();
	}
#endif
	win_kernel<<<win_grid(p, n_demods, kWinThreads), kWinThreads, 0, s>>>(p);
	return cudaGetLastError();
}
cudaError_t launch_winlong(const BackParams &p, int n_demods, cudaStream_t s)
{
	// one warp per window slot; nearly all of them leave at once (not a chain head, or not long)
	int gx = (p.n_tiles * 2 + kLongWarps - 1) / kLongWarps;
	gx = gx < 1 ? 1 : (gx > 64 ? 64 : gx);   // the warps stride over the window list
	winlong_kernel<<<dim3(gx, p.n_streams, n_demods), 32 * kLongWarps, 0, s>>>(p);
	return cudaGetLastError();
}
cudaError_t launch_verify(const BackParams &p, int n_demods, cudaStream_t s)
{
#ifdef TFR_VER_PROFILE
	{
		unsigned long long z[8] = { 0 }, r[8];
		cudaStreamSynchronize(s);
		cudaMemcpyToSymbol(g_verprof, z, sizeof(z));
		verify_kernel<<<p.n_streams * n_demods, kVerifyThreads, 0, s>>>(p);
		cudaStreamSynchronize(s);
		cudaMemcpyFromSymbol(r, g_verprof, sizeof(r));
		fprintf(stderr, "[verprof] max over CTAs: flags %llu cyc, run starts %llu cyc, repair %llu cyc, rounds %llu, whole CTA %llu cyc | max thread: repair %llu cyc, %llu windows\n", r[0], r[1], r[2], r[3], r[4], r[5], r[6]);
		return cudaGetLastError();
	}
#endif
	verify_kernel<<<p.n_streams * n_demods, kVerifyThreads, 0, s>>>(p);
	return cudaGetLastError();
}

}  // namespace tfr
