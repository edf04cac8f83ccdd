// backend.cu - everything downstream of the decimator, on the device:
//
//   whb_kernel     WeatherHub demodulator + framer (whb_demod::demod whb.cpp:632-707, store_bit :566-603), one
//                  serial chain per stream over the demod's windows (TFA_1 and the TFA_2 family run
//                  window-parallel in backend2.cu)
//   parse_kernel   decoder::flush (tfa1.cpp:47-118, tfa2.cpp:64-279, whb.cpp:477-564) + CRC-8 / CRC-32
//                  (crc8.cpp, crc32.cpp), one warp per candidate frame
//
// Floating point follows the reference *as built by its own Makefile* (x86-64, -O3 -ffast-math, g++ 13):
// fm_dev scales with one multiply by fl(16384/pi), iir2::step sums as ((b2*dn2+a1*yn1)+(b0*dn+b1*dn1))+a2*yn2,
// x/10 is x*0.1.  All double arithmetic below is written with explicit round-to-nearest intrinsics so that
// nvcc cannot contract it into FMAs.  The one place a device libm call could disagree with glibc at a
// truncation knife edge - 10*log10(rssi) - is left to the host (rssi_raw travels in the frame).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "demod_dev.cuh"

namespace tfr {

// ------------------------------------------------------------------------------------------------
// whb_kernel: WeatherHub (whb.cpp:632-707), one CTA of two warps per stream.
//
// WeatherHub cannot be cut into independent windows: its averaging biquad (0.0025/spb, whb.cpp:611) has a
// time constant of ~5800 samples and is never reset, so avg_of at any sample depends on hundreds of earlier
// windows.  The demodulator therefore stays ONE serial chain per stream - but a tight one:
//   * the warp walks the demod's window list (threshold kernel), 32 samples per step: every lane loads one
//     stored sample and its predecessor (coalesced) and computes the discriminator Re(a*conj b) and I^2+Q^2;
//     the loads of the next 32 samples are issued before the current ones are consumed
//   * per 32-sample step only the two biquad RECURRENCES are serial (four dependent FP64 operations per sample
//     each, run redundantly by all lanes); everything around them - input-only filter terms, truncations,
//     comparisons - is done by the lanes in parallel between the two serial phases
//   * bits (one per >= 48 samples), the descrambler/framer and the frame buffer are the rare path (lane 0's
//     local memory is the reference's rdata[])
// ------------------------------------------------------------------------------------------------
template <bool TAPS>
__global__ void __launch_bounds__(64) whb_kernel(const BackParams p)
{
	const int stream = blockIdx.x;
	const int lane = threadIdx.x & 31;
	const int role = threadIdx.x >> 5;   // warp 0: pulse filter (stage A), warp 1: averaging filter, dips, framer (stage B)
	if (stream >= p.n_streams) return;
	int demod = -1;
	for (int k = 0; k < p.cfg->n_demods; k++)
		if (p.cfg->d[k].kind == K_WHB) demod = k;
	if (demod < 0) return;
	const DemodCfg cfg = p.cfg->d[demod];
	const StreamJob job = p.jobs[stream];
	if (job.n_blocks == 0) return;
	StreamState *st = p.st + stream;
	Walk w;
	w.p = &p;
	w.stream = stream;
	w.demod = demod;
	w.s = st->d[demod];
	for (int k = 0; k < 3; k++) w.tap_n[k] = p.tap_cap ? p.tap_cnt[((size_t)stream * kMaxDemods + demod) * 3 + k] : 0;
	constexpr bool taps = TAPS;   // compiled out of the production kernel: a branch in the loop body would keep the
	                              // scheduler from interleaving the two filter chains
	const uint32_t n_win = p.wincnt[stream].n[demod];
	const WinEntry *wl = p.wins + job.win_off + (size_t)demod * job.win_cap;
	const uint32_t *dec = p.dec + (size_t)job.dec_off * kBlockDec;
	const uint32_t call_len = job.n_blocks * (uint32_t)kBlockDec;
	const uint32_t prev_last = ((uint32_t)(uint16_t)st->last_i) | ((uint32_t)(uint16_t)st->last_q << 16);
	const int64_t base_pos = st->blocks_done * (int64_t)kBlockDec;

	// hot state in registers (identical in every lane)
	Biquad lp = w.s.lp, la = w.s.lp_avg;
	const BiquadCoef kp = cfg.lp, ka = cfg.lp_avg;
	int last_dev = w.s.last_dev, avg_of = w.s.avg_of, synced = w.s.synced;
	uint32_t step = w.s.step_lo, last_peak = w.s.last_peak;
	double rssi = w.s.rssi_d;
	const double spb12 = __dmul_rn(cfg.spb, 0.5);
	const int spb34 = (int)floor(__dmul_rn(cfg.spb, 0.75));   // (double)tdiff > 0.75*spb  <=>  tdiff > floor(0.75*spb)

	// per step: what the lanes prepare in parallel for the two serial recurrences, and what those produce
	__shared__ double s_t2[32], s_u[32], s_y[32], s_t2s[32], s_us[32], s_a[32];
	const size_t tbase = ((size_t)stream * kMaxDemods + demod) * (size_t)p.tap_cap;

	// ---- the stream's samples as a flat sequence of steps of <= 32 samples (a window is cut into equal steps, so
	// that a 513-sample window gives 17 steps of 30..31 samples and not 16 full ones plus one of a single sample)
	struct Step {
		uint32_t m0, last;    // first sample, last sample of the window inside this call
		int cnt;              // 0: no more steps
		uint32_t win_end;     // WinEntry::end
		bool first, final;    // first / last step of its window
		bool cont;            // the window was already open when the call began
	};
	uint32_t g_wi = 0, g_pos = 0, g_last = 0, g_size = 0, g_end = 0;
	bool g_open = false, g_cont = false;
	auto next_step = [&]() -> Step {
		Step q;
		q.cnt = 0;
		q.m0 = q.last = q.win_end = 0;
		q.first = q.final = q.cont = false;
		if (!g_open) {
			if (g_wi >= n_win) return q;
			const WinEntry e = wl[g_wi];
			if (e.start >= call_len) { g_wi = n_win; return q; }
			g_pos = e.start;
			g_end = e.end;
			g_last = min(e.end, call_len - 1);
			const uint32_t len = g_last - e.start + 1;
			g_size = (len + ((len + 31) / 32) - 1) / ((len + 31) / 32);
			g_cont = (e.flags & kWinCont) != 0;
			g_open = true;
			q.first = true;
		}
		q.m0 = g_pos;
		q.last = g_last;
		q.win_end = g_end;
		q.cont = g_cont;
		q.cnt = (int)min(g_size, g_last - g_pos + 1);
		g_pos += (uint32_t)q.cnt;
		if (g_pos > g_last) {
			q.final = true;
			g_open = false;
			g_wi++;
		}
		return q;
	};
	auto load = [&](const Step &q, uint32_t &cw, uint32_t &lw) {
		cw = lw = 0;
		const uint32_t m = q.m0 + (uint32_t)lane;
		if (lane < q.cnt) {
			cw = dec[m];
			lw = (m == 0) ? prev_last : dec[m - 1];
		}
	};
	auto phase_change = [&](int tdiff) {
		// one 0, then a 1 for every further bit period since the last one (whb.cpp:662-674)
		whb_bit(w.s, 0);
		w.s.bitcnt++;
		const int bit0 = __double2int_rz(__ddiv_rn(__dadd_rn((double)tdiff, spb12), cfg.spb));
		for (int n = 1; n < bit0; n++) {
			whb_bit(w.s, 1);
			w.s.bitcnt++;
		}
	};

	// Two warps, two pipeline stages.  In iteration i warp 0 runs stage A of step i+1 - loads, discriminator, the
	// pulse filter's recurrence, truncation (it never depends on the framer) - while warp 1 runs stage B of step
	// i: the averaging filter's recurrence, the dip decisions, framer and flush.  Stage A hands each sample's
	// (dev, 0.5*dev, I^2+Q^2, y) over in a double-buffered shared array; one __syncthreads per step.
	__shared__ int sb_dev[2][32], sb_pw[2][32];
	__shared__ double sb_x[2][32], sb_y[2][32];
	Step cur = next_step(), nxt = cur.cnt ? next_step() : cur;
	uint32_t cw = 0, lw = 0;
	auto stage_a = [&](const Step &q, uint32_t cwv, uint32_t lwv, int buf) {
		const int i = (int)(int16_t)(cwv & 0xffff), qq = (int)(int16_t)(cwv >> 16);
		const int cr = fm_dev_nrzs(i, qq, (int)(int16_t)(lwv & 0xffff), (int)(int16_t)(lwv >> 16));
		// the pulse filter's input-only terms, with the reference's roundings (iir2::step as built:
		// ((b2*dn2 + a1*yn1) + (b0*dn + b1*dn1)) + a2*yn2): u = b2*dn2 and t2 = b0*dn + b1*dn1
		const double d = (double)cr;
		double d1 = __shfl_up_sync(0xffffffffu, d, 1), d2 = __shfl_up_sync(0xffffffffu, d, 2);
		if (lane == 0) { d1 = lp.d1; d2 = lp.d2; }
		if (lane == 1) d2 = lp.d1;
		s_t2[lane] = __dadd_rn(__dmul_rn(kp.b0, d), __dmul_rn(kp.b1, d1));
		s_u[lane] = __dmul_rn(kp.b2, d2);
		// the filter's input history after this step
		const double dl = __shfl_sync(0xffffffffu, d, q.cnt - 1), dl2 = __shfl_sync(0xffffffffu, d, max(q.cnt - 2, 0));
		lp.d2 = (q.cnt >= 2) ? dl2 : lp.d1;
		lp.d1 = dl;
		if (taps) {
			const uint32_t ti = w.tap_n[1] + (uint32_t)lane;
			if (lane < q.cnt && ti < p.tap_cap) p.tap_i32[1][tbase + ti] = cr;
			w.tap_n[1] += (uint32_t)q.cnt;
		}
		__syncwarp();
		// the recurrence: four dependent FP64 operations per sample, operands fetched two samples ahead
		double y0 = lp.y0, y1 = lp.y1;
		double ua = s_u[0], ta = s_t2[0], ub = s_u[1], tb = s_t2[1];
#pragma unroll 4
		for (int k = 0; k < q.cnt; k++) {
			const int kn = min(k + 2, 31);
			const double un = s_u[kn], tn = s_t2[kn];
			const double y = __dadd_rn(__dadd_rn(__dadd_rn(ua, __dmul_rn(kp.a1, y0)), ta), __dmul_rn(kp.a2, y1));
			s_y[k] = y;
			y1 = y0;
			y0 = y;
			ua = ub; ta = tb; ub = un; tb = tn;
		}
		lp.y0 = y0;
		lp.y1 = y1;
		__syncwarp();
		const double y = s_y[lane];
		const int dev = trunc_to_int(y);
		sb_dev[buf][lane] = dev;
		sb_x[buf][lane] = __dmul_rn(0.5, int_to_double(dev));
		sb_pw[buf][lane] = i * i + qq * qq;
		if (taps) sb_y[buf][lane] = y;
	};
	if (role == 0 && cur.cnt) {
		load(cur, cw, lw);
		stage_a(cur, cw, lw, 0);
		if (nxt.cnt) load(nxt, cw, lw);
	}
	__syncthreads();
	for (int it = 0; cur.cnt; it++) {
		const Step nn = nxt.cnt ? next_step() : nxt;   // both warps step the (deterministic) generator alike
		if (role == 0) {
			if (nxt.cnt) {
				uint32_t cw2 = 0, lw2 = 0;
				if (nn.cnt) load(nn, cw2, lw2);   // in flight during this step's recurrence
				stage_a(nxt, cw, lw, (it + 1) & 1);
				cw = cw2;
				lw = lw2;
			}
		} else {
			const Step &q = cur;
			const int cnt = q.cnt, buf = it & 1;
			const int dev = sb_dev[buf][lane], pw_c = sb_pw[buf][lane];
			const double x_c = sb_x[buf][lane];
			const double y_c = taps ? sb_y[buf][lane] : 0.0;
			if (q.first && !q.cont) {   // whb.cpp:636-642: a trigger with the timeout expired starts a new window
				w.s.offset = 0;
				w.s.bitcnt = 0;
				rssi = 0.0;
				step = 0;
				last_peak = 0;
			}
			const bool run_b = !synced;
			double ya0 = la.y0, ya1 = la.y1;
			if (run_b) {
				// the averaging filter's input-only terms, then its recurrence
				double x1 = __shfl_up_sync(0xffffffffu, x_c, 1), x2 = __shfl_up_sync(0xffffffffu, x_c, 2);
				if (lane == 0) { x1 = la.d1; x2 = la.d2; }
				if (lane == 1) x2 = la.d1;
				s_t2s[lane] = __dadd_rn(__dmul_rn(ka.b0, x_c), __dmul_rn(ka.b1, x1));
				s_us[lane] = __dmul_rn(ka.b2, x2);
				__syncwarp();
				double va = s_us[0], wa = s_t2s[0], vb = s_us[1], wb = s_t2s[1];
#pragma unroll 4
				for (int k = 0; k < cnt; k++) {
					const int kn = min(k + 2, 31);
					const double vn = s_us[kn], wn = s_t2s[kn];
					const double a = __dadd_rn(__dadd_rn(__dadd_rn(va, __dmul_rn(ka.a1, ya0)), wa), __dmul_rn(ka.a2, ya1));
					s_a[k] = a;
					ya1 = ya0;
					ya0 = a;
					va = vb; wa = wb; vb = vn; wb = wn;
				}
				__syncwarp();
			}
			int dev_prev = __shfl_up_sync(0xffffffffu, dev, 1);
			if (lane == 0) dev_prev = last_dev;
			const bool rising = lane < cnt && dev > dev_prev;
			const uint32_t step0 = step;
			// dip decisions over a candidate mask, in sample order (only last_peak chains them); returns the sample
			// on which the sync word completed, or -1
			auto decide = [&](unsigned cm, bool stop_at_sync) -> int {
				while (cm) {
					const int k = __ffs(cm) - 1;
					cm &= cm - 1;
					const int tdiff = (int)(step0 + (uint32_t)k - last_peak);
					if (tdiff > spb34) {
						phase_change(tdiff);
						last_peak = step0 + (uint32_t)k;
						if (stop_at_sync && w.s.synced) return k;
					}
				}
				return -1;
			};
			// a frame is being received from sample k0 on (whb.cpp:653,677,693): avg_of frozen, rssi accumulating
			auto frame_part = [&](int k0) {
				decide(__ballot_sync(0xffffffffu, rising && lane >= k0 && dev < avg_of), false);
				long long pw = (lane >= k0 && lane < cnt) ? (long long)pw_c : 0ll;
				for (int o = 16; o; o >>= 1) pw += __shfl_xor_sync(0xffffffffu, pw, o);
				rssi = __dadd_rn(rssi, (double)pw);   // sums of I^2+Q^2 stay far below 2^53: every partial sum is exact
				if (taps) {
					const uint32_t ti = w.tap_n[2] + (uint32_t)(lane - k0);
					if (lane >= k0 && lane < cnt && ti < p.tap_cap) p.tap_f64[tbase + ti] = y_c;
					w.tap_n[2] += (uint32_t)(cnt - k0);
				}
			};
			if (run_b) {
				const double a_l = s_a[lane];
				const int avg = trunc_to_int(a_l);
				if (taps) {
					const uint32_t ti = w.tap_n[2] + 2u * (uint32_t)lane;
					if (lane < cnt && ti + 1 < p.tap_cap) {
						p.tap_f64[tbase + ti] = y_c;
						p.tap_f64[tbase + ti + 1] = a_l;
					}
				}
				const int ks = decide(__ballot_sync(0xffffffffu, rising && dev < avg), true);
				if (ks < 0) {
					const double xl = __shfl_sync(0xffffffffu, x_c, cnt - 1), xl2 = __shfl_sync(0xffffffffu, x_c, max(cnt - 2, 0));
					la.d2 = (cnt >= 2) ? xl2 : la.d1;
					la.d1 = xl;
					la.y0 = ya0;
					la.y1 = ya1;
					avg_of = __shfl_sync(0xffffffffu, avg, cnt - 1);
					if (taps) w.tap_n[2] += 2u * (uint32_t)cnt;
				} else {
					// the sync word completed on sample ks: the averaging filter stops there, the rest of the step is frame
					const double xs = __shfl_sync(0xffffffffu, x_c, ks), xs1 = __shfl_sync(0xffffffffu, x_c, max(ks - 1, 0));
					la.y1 = (ks >= 1) ? s_a[ks - 1] : la.y0;
					la.y0 = s_a[ks];
					la.d2 = (ks >= 1) ? xs1 : la.d1;
					la.d1 = xs;
					avg_of = __shfl_sync(0xffffffffu, avg, ks);
					synced = 1;
					rssi = __dadd_rn(rssi, (double)__shfl_sync(0xffffffffu, pw_c, ks));
					if (taps) w.tap_n[2] += 2u * (uint32_t)(ks + 1);
					frame_part(ks + 1);
				}
			} else {
				frame_part(0);
			}
			last_dev = __shfl_sync(0xffffffffu, dev, cnt - 1);
			step += (uint32_t)cnt;
			if (q.final) {
				if (q.last == q.win_end) {
					// the timeout ran out on this sample (whb.cpp:691-700)
					if (synced) {
						for (int n = 0; n < 16; n++) whb_bit(w.s, 0);
						w.s.rssi_d = rssi;
						w.pos = base_pos + q.win_end;
						if (lane == 0) whb_flush(w);
						else { w.s.sr_cnt = -1; w.s.sr = 0; w.s.byte_cnt = 0; w.s.synced = 0; }
						synced = 0;
					}
					w.s.offset = 0;
					w.s.bitcnt = 0;
					rssi = 0.0;
					step = 0;
					last_peak = 0;
					w.s.timeout_cnt = 0;
				} else {
					w.s.timeout_cnt = (int)(q.win_end - q.last);
				}
			}
		}
		__syncthreads();
		cur = nxt;
		nxt = nn;
	}
	// the carried state: warp 1 owns everything but the pulse filter
	if (role == 1 && lane == 0) {
		w.s.lp_avg = la;
		w.s.last_dev = last_dev;
		w.s.avg_of = avg_of;
		w.s.synced = synced;
		w.s.step_lo = step;
		w.s.last_peak = last_peak;
		w.s.rssi_d = rssi;
		st->d[demod] = w.s;
		if (p.tap_cap) p.tap_cnt[((size_t)stream * kMaxDemods + demod) * 3 + 2] = w.tap_n[2];
	}
	__syncthreads();
	if (role == 0 && lane == 0) {
		st->d[demod].lp = lp;
		if (p.tap_cap) p.tap_cnt[((size_t)stream * kMaxDemods + demod) * 3 + 1] = w.tap_n[1];
	}
}

// after the last epoch of a submit: roll positions, carry and last sample forward
__global__ void submit_epilogue_kernel(const BackParams p)
{
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= p.n_streams) return;
	const StreamJob job = p.jobs[s];
	if (job.n_blocks == 0) return;
	StreamState *st = p.st + s;
	const size_t glast = (size_t)job.dec_off + job.n_blocks - 1;
	const uint32_t lw = p.dec[glast * kBlockDec + kBlockDec - 1];
	st->last_i = (int16_t)(lw & 0xffff);
	st->last_q = (int16_t)(lw >> 16);
	st->carry_in = p.tiles[glast].carry_out;
	st->blocks_done += job.n_blocks;
}

// ------------------------------------------------------------------------------------------------
// parse_kernel: one warp per candidate frame
// ------------------------------------------------------------------------------------------------
// crc8.cpp: poly 0x31, init 0, MSB first
__device__ __forceinline__ uint8_t crc8_31(const uint8_t *d, int len)
{
	uint32_t c = 0;
	for (int n = 0; n < len; n++) {
		c ^= d[n];
		for (int m = 0; m < 8; m++) c = (c & 0x80) ? ((c << 1) ^ 0x31) & 0xff : (c << 1) & 0xff;
	}
	return (uint8_t)c;
}
// crc32.cpp: poly 0x04c11db7, caller init, MSB first.  The warp splits the message: lane k folds byte k
// through the (len-1-k) trailing bytes' worth of zero shifts, then the lanes XOR-reduce (CRC is linear);
// the init value rides with byte 0.
__device__ __forceinline__ uint32_t crc32_shift8(uint32_t c)
{
	for (int m = 0; m < 8; m++) c = (c & 0x80000000u) ? ((c << 1) ^ 0x04c11db7u) : (c << 1);
	return c;
}
__device__ uint32_t crc32_warp(const uint8_t *d, int len, uint32_t init, int lane)
{
	uint32_t acc = 0;
	for (int base = 0; base < len; base += 32) {
		// advance what has been accumulated so far by the bytes of this round
		const int n = min(32, len - base);
		uint32_t part = 0;
		if (lane < n) {
			part = (uint32_t)d[base + lane] << 24;
			if (base == 0 && lane == 0) part ^= init;
			// this byte enters the register and is then followed by (n-1-lane) more bytes of the round
			part = crc32_shift8(part);
			for (int k = 0; k < n - 1 - lane; k++) part = crc32_shift8(part);
		}
		// previous rounds' remainder is pushed through n more bytes
		if (lane == 0)
			for (int k = 0; k < n; k++) acc = crc32_shift8(acc);
		part ^= (lane == 0) ? acc : 0u;
		for (int o = 16; o; o >>= 1) part ^= __shfl_xor_sync(0xffffffffu, part, o);
		acc = part;
	}
	if (len <= 0) return init;
	return acc;
}

__device__ __forceinline__ uint32_t whb_crc_init(uint32_t stype, bool &known)
{
	known = true;
	switch (stype) {   // crc_initvals, whb.cpp:50-62
	case 0x02: return 0x97d97a26u;
	case 0x03: return 0xf59c5a1eu;
	case 0x04: return 0x98e1d11fu;
	case 0x06: return 0xa7a41254u;
	case 0x07: return 0x3303fb1du;
	case 0x08: return 0x29f0f49bu;
	case 0x09: return 0xa7a41254u;
	case 0x0b: return 0xe7720ae4u;
	case 0x10: return 0x62d0afc1u;
	case 0x11: return 0x8cba0708u;
	case 0x12: return 0x5a9e30aeu;
	}
	known = false;
	return 0;
}

struct RecOut {
	const BackParams *p;
	DevFrame *f;
	int frame_idx;
	int n;
	uint32_t first;
};
__device__ void put_record(RecOut &o, uint64_t id, double temp, double hum, int seq, int alarm)
{
	// records of one frame must be contiguous and ordered: reserve on first use (max 5 per frame)
	if (o.n == 0) o.first = atomicAdd(&o.p->counters->n_records, 5u);
	const uint32_t k = o.first + o.n;
	o.n++;
	if (k >= o.p->max_records) {
		o.p->counters->overflow = 1;
		return;
	}
	DevRecord &r = o.p->records[k];
	r.stream = o.f->stream;
	r.type = o.f->type;
	r.id = id;
	r.temp = temp;
	r.humidity = hum;
	r.alarm = alarm;
	r.flags = 0;
	r.sequence = seq;
	r.frame = o.frame_idx;
	r.pos = o.f->pos;
}

__device__ __forceinline__ double div10(double v) { return __dmul_rn(v, 0.1); }   // as built: x/10 -> x*0.1
__device__ __forceinline__ double bcd_temp(int v) { return __dsub_rn(div10((double)v), 40.0); }
__device__ __forceinline__ int be16(const uint8_t *x) { return (x[0] << 8) | x[1]; }
// whb_decoder::cvt_temp, whb.cpp:109-123
__device__ __forceinline__ double whb_temp(int raw, int ext)
{
	if (ext) return (raw & 0x800) ? div10((double)(-((raw ^ 0xfff) + 1))) : div10((double)raw);
	return (raw & 0x400) ? div10((double)(-((raw ^ 0x7ff) + 1))) : div10((double)raw);
}

// tfa1_decoder::flush, tfa1.cpp:56-113
__device__ void parse_tfa1(RecOut &o, const uint8_t *r)
{
	const int id = ((r[2] << 8) | r[3]) & 0x7fff;
	int batfail = (r[7] & 0x80) >> 7;
	double temp = bcd_temp((r[4] & 0xf) * 100 + (r[5] >> 4) * 10 + (r[5] & 0xf));
	int hum = r[6];
	const int seq = r[8] >> 4;
	const uint8_t crc_val = r[10], crc_calc = crc8_31(&r[2], 8);
	const bool sane = ((r[4] & 0xf0) == 0x80 || hum == 0x7f || hum == 0x6a) && hum <= 0x7f && (r[7] & 0x60) == 0x60 &&
			  (r[8] & 0xf) == 0 && r[9] == 0x56;
	if (crc_val == crc_calc && sane) {
		if (hum == 0x6a) hum = 0;
		if (r[5] == 0xff || r[5] == 0xaa || hum == 0x7f) {
			batfail = 2;
			hum = 0;
			temp = 0;
		}
		o.f->status = 0;
		put_record(o, (uint64_t)id, temp, (double)hum, seq, batfail);
	} else {
		o.f->status = (crc_val != crc_calc) ? 1 : 2;
	}
}
// tfa2_decoder::flush_tfa, tfa2.cpp:230-273
__device__ void parse_tfa2(RecOut &o, const uint8_t *r, int type)
{
	int id = (type << 28) | (r[2] << 8) | (r[3] & 0xc0);
	const double temp = bcd_temp((r[3] & 0xf) * 100 + (r[4] >> 4) * 10 + (r[4] & 0xf));
	int hum = r[5];
	const uint8_t crc_val = r[6], crc_calc = crc8_31(&r[2], 4);
	if (hum == 0x7d) id |= 1;
	if (crc_val == crc_calc) {
		if (hum > 100) hum = 0;
		o.f->status = 0;
		put_record(o, (uint64_t)(int64_t)id, temp, (double)hum, 0, 0);
	} else {
		o.f->status = 1;
	}
}
// tfa2_decoder::flush_tx22, tfa2.cpp:83-197
__device__ void parse_tx22(RecOut &o, const uint8_t *r, int type)
{
	if ((r[2] >> 4) != 0xa) {
		o.f->status = 2;
		return;
	}
	const int id = ((r[2] & 0xf) << 2) | (r[3] >> 6);
	const int error = !((r[3] >> 4) & 1), lowbat = (r[3] >> 3) & 1, num = r[3] & 7;
	const uint8_t crc_val = r[2 * num + 4], crc_calc = crc8_31(&r[2], 2 + 2 * num);
	if (crc_val != crc_calc) {
		o.f->status = 1;
		return;
	}
	bool have_temp = false, have_rain = false, have_wind = false, have_gust = false;
	double temp = 0, hum = 0, rain = 0, wdir = 0, wspeed = 0, wgust = 0;
	for (int n = 0; n < num; n++) {
		const uint8_t *w = &r[4 + n * 2];
		switch (w[0] >> 4) {
		case 0: temp = bcd_temp((w[0] & 0xf) * 100 + (w[1] >> 4) * 10 + (w[1] & 0xf)); have_temp = true; break;
		case 1: hum = (double)((w[0] & 0xf) * 100 + (w[1] >> 4) * 10 + (w[1] & 0xf)); break;
		case 2: rain = (double)(((w[0] & 0xf) << 8) + w[1]); have_rain = true; break;
		case 3: wdir = __dmul_rn((double)(w[0] & 0xf), 22.5); wspeed = div10((double)w[1]); have_wind = true; break;
		case 4: wgust = div10((double)(((w[0] & 0xf) << 8) + w[1])); have_gust = true; break;
		default: break;
		}
	}
	const int alarm = error | lowbat;
	const int new_id = (type << 28) | (id << 4);
	o.f->status = 0;
	if (have_temp) put_record(o, (uint64_t)(int64_t)new_id, temp, hum, 0, alarm);
	if (have_rain) put_record(o, (uint64_t)(int64_t)(new_id | 2), rain, 0, 0, alarm);
	if (have_wind) put_record(o, (uint64_t)(int64_t)(new_id | 3), wspeed, wdir, 0, alarm);
	if (have_gust) put_record(o, (uint64_t)(int64_t)(new_id | 4), wgust, 0, 0, alarm);
}
// the payload parsers decode_02 ... decode_12, whb.cpp:126-475
__device__ void parse_whb_payload(RecOut &o, uint32_t stype, const uint8_t *m, uint64_t id)
{
	static const uint32_t tu[4] = { 24 * 60 * 60, 60 * 60, 60, 1 };   // timeunit_tab, whb.cpp:65-70
	const uint64_t base = id << 4;
	const int seq = be16(m) & 0x3fff;
	switch (stype) {
	case 0x02: put_record(o, base, whb_temp(be16(m + 2) & 0x7ff, 0), 0, seq, 0); break;
	case 0x03: put_record(o, base, whb_temp(be16(m + 2) & 0x7ff, 0), (double)(be16(m + 4) & 0xff), seq, 0); break;
	case 0x04:
		put_record(o, base, whb_temp(be16(m + 2) & 0x7ff, 0), (double)(be16(m + 4) & 0xff), seq, 0);
		put_record(o, base | 5, (double)((m[6] & 1) ^ 1), 0, seq, 0);
		break;
	case 0x06:
	case 0x09: {
		const int ext = (stype == 0x09);
		const int t2 = be16(m + 4) & (ext ? 0xfff : 0x7ff);
		put_record(o, base, whb_temp(be16(m + 2) & 0x7ff, 0), (double)(be16(m + 6) & 0xff), seq, 0);
		put_record(o, base | 1, whb_temp(t2, ext), 0, seq, 0);
		break;
	}
	case 0x07:
		put_record(o, base, whb_temp(be16(m + 2) & 0x7ff, 0), (double)(be16(m + 4) & 0xff), seq, 0);
		put_record(o, base | 0xc, whb_temp(be16(m + 6) & 0x7ff, 0), (double)(be16(m + 8) & 0xff), seq, 0);
		break;
	case 0x08: {
		const int x1 = be16(m + 6 + 2);
		const uint32_t t1 = tu[(x1 >> 14) & 3] * (uint32_t)(x1 & 0x3fff);
		put_record(o, base | 2, (double)be16(m + 4), (double)t1, seq, 0);
		put_record(o, base, whb_temp(be16(m + 2) & 0x7ff, 0), 0, seq, 0);
		break;
	}
	case 0x0b: {
		const uint32_t v = ((uint32_t)m[3] << 24) | (m[4] << 16) | (m[5] << 8) | m[6];
		const float dir = __double2float_rn(__dmul_rn(22.5, (double)(v >> 28)));
		const float speed = __double2float_rn(div10((double)(((v >> 16) & 0xff) + 256 * ((v >> 25) & 1))));
		const float gust = __double2float_rn(div10((double)(((v >> 8) & 0xff) + 256 * ((v >> 24) & 1))));
		const int seq24 = (m[0] << 16) | (m[1] << 8) | m[2];
		put_record(o, base | 3, (double)speed, (double)dir, seq24, 0);
		put_record(o, base | 4, (double)gust, 0, seq24, 0);
		break;
	}
	case 0x10: {
		const int x0 = be16(m + 2), x1 = be16(m + 4);
		put_record(o, base | 5, (double)(x0 >> 15), (double)(tu[(x1 >> 13) & 3] * (uint32_t)(x1 & 0x1fff)), seq, 0);
		break;
	}
	case 0x11:
		put_record(o, base, whb_temp(be16(m + 2 + 12) & 0x7ff, 0), (double)(be16(m + 4 + 12) & 0xff), seq, 0);
		for (int n = 0; n < 3; n++)
			put_record(o, base | (uint64_t)(0xc + n), whb_temp(be16(m + 2 + 4 * n) & 0x7ff, 0),
				   (double)(be16(m + 4 + 4 * n) & 0xff), seq, 0);
		break;
	case 0x12:
		put_record(o, base, whb_temp(be16(m + 6) & 0x7ff, 0), (double)(m[8] & 0x7f), seq, 0);
		put_record(o, base + 1, 0, (double)(m[2] & 0x7f), seq, 0);
		for (int n = 0; n < 3; n++) put_record(o, base + 0xc + n, 0, (double)(m[3 + n] & 0x7f), seq, 0);
		break;
	default: break;
	}
}

__global__ void parse_kernel(const BackParams p)
{
	const int lane = threadIdx.x & 31;
	const uint32_t n_frames = min(p.counters->n_frames, p.max_frames);
	const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
	__shared__ uint8_t s_r[8][kMaxRdata + 8];
	uint8_t *r = s_r[threadIdx.x >> 5];
	for (uint32_t fi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; fi < n_frames; fi += n_warps) {
	if (p.frames[fi].status != -1) continue;   // >= 0: parsed by an earlier tfr_process; -2: retired by the verifier (dead)
	DevFrame *f = p.frames + fi;
	// a frame of the NEXT call (its early windows run beside this call's verifier and parsers) is left to that call's parse
	if (*reinterpret_cast<volatile int32_t *>(&f->first_record) != p.slot_tag) continue;
	__syncwarp();
	for (int k = lane; k < kMaxRdata; k += 32) r[k] = f->rdata[k];
	__syncwarp();
	RecOut o = { &p, f, (int)fi, 0, 0 };
	const int kind = p.cfg->d[f->demod].kind;
	if (kind == K_WHB) {
		// whb_decoder::flush, whb.cpp:493-547; the CRC-32 is the warp-cooperative part
		const int plen = r[4];
		const uint32_t stype = r[5];
		bool known;
		const uint32_t init = whb_crc_init(stype, known);
		bool good = false;
		uint32_t crc_calc = 0, crc_val = 0;
		if (plen <= 60 && known) {
			crc_calc = crc32_warp(&r[4], max(plen - 4, 0), init, lane);
			crc_val = ((uint32_t)r[plen] << 24) | (r[plen + 1] << 16) | (r[plen + 2] << 8) | r[plen + 3];
			good = (crc_calc == crc_val);
		}
		if (lane == 0) {
			if (good) {
				uint64_t id = 0;
				for (int k = 0; k < 6; k++) id = (id << 8) | r[5 + k];
				f->status = 0;
				parse_whb_payload(o, stype, &r[11], id);
			} else {
				f->status = (crc_val != crc_calc) ? 1 : 2;
			}
		}
	} else if (lane == 0) {
		if (kind == K_TFA1) parse_tfa1(o, r);
		else if (kind == K_TX22) parse_tx22(o, r, f->type);
		else parse_tfa2(o, r, f->type);
	}
	if (lane == 0) {
		f->n_records = o.n;
		f->first_record = (int)o.first;
	}
	}
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
cudaError_t launch_walk(const BackParams &p, int n_demods, cudaStream_t s)
{
	(void)n_demods;
	if (p.tap_cap) whb_kernel<true><<<p.n_streams, 64, 0, s>>>(p);
	else whb_kernel<false><<<p.n_streams, 64, 0, s>>>(p);
	return cudaGetLastError();
}
cudaError_t launch_submit_epilogue(const BackParams &p, cudaStream_t s)
{
	submit_epilogue_kernel<<<(p.n_streams + 127) / 128, 128, 0, s>>>(p);
	return cudaGetLastError();
}
cudaError_t launch_parse(const BackParams &p, cudaStream_t s)
{
	parse_kernel<<<32, 256, 0, s>>>(p);
	return cudaGetLastError();
}

}  // namespace tfr
