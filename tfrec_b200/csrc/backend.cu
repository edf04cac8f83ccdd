// backend.cu - everything downstream of the decimator, on the device:
//
//   walk_kernel    the per-sample demodulator state machines and bit framers
//                    TFA_1            tfa1_demod::demod      tfa1.cpp:143-190, store_bit :120-134
//                    TFA_2/TFA_3/TX22 tfa2_demod::demod      tfa2.cpp:346-442, store_bit :281-314
//                    WeatherHub       whb_demod::demod       whb.cpp:632-707,  store_bit :566-603
//                  run over the trigger windows the front-end kept, in stream order with carried state
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
// walk_kernel: one thread per (stream, demod), windows in stream order with carried state
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) walk_kernel(const BackParams p)
{
	const int gid = blockIdx.x * blockDim.x + threadIdx.x;
	const int nd = p.cfg->n_demods;
	if (gid >= p.n_streams * nd) return;
	Walk w;
	w.p = &p;
	w.stream = gid / nd;
	w.demod = gid % nd;
	const DemodCfg cfg = p.cfg->d[w.demod];
	if (cfg.kind != K_WHB) return;   // TFA_1 / TFA_2 family run window-parallel in backend2.cu
	const StreamJob job = p.jobs[w.stream];
	if (job.n_blocks == 0) return;
	StreamState *st = p.st + w.stream;
	w.s = st->d[w.demod];
	for (int k = 0; k < 3; k++) w.tap_n[k] = p.tap_cap ? p.tap_cnt[((size_t)w.stream * kMaxDemods + w.demod) * 3 + k] : 0;

	const int t_end = min(p.tile0 + p.n_tiles, (int)job.n_blocks);
	const int64_t base_block = st->blocks_done;
	int carry_in = entry_carry(p, job, st);
	uint32_t prev_last;   // the sample just before this epoch's first block (every demod's last_i/last_q)
	if (p.tile0 == 0)
		prev_last = ((uint32_t)(uint16_t)st->last_i) | ((uint32_t)(uint16_t)st->last_q << 16);
	else
		prev_last = p.dec[((size_t)job.dec_off + p.tile0 - 1) * kBlockDec + kBlockDec - 1];

	Regions reg;
	for (int tile = p.tile0; tile < t_end; tile++) {
		const size_t gtile = (size_t)job.dec_off + tile;
		const TileDesc &td = p.tiles[gtile];
		const uint32_t *d = p.dec + gtile * kBlockDec;
		const int thresh = p.trace[gtile].thresh;
		// demodulator::start, decoder.cpp:118-122
		if (w.s.last_bit_idx) w.s.last_bit_idx -= kIdxPerBlock;
		build_regions(td, carry_in, reg);
		int pos = 0;
		for (int r = 0; r < reg.n; r++) {
			const int a = reg.start[r], b = reg.end[r];
			if (cfg.kind == K_WHB) w.s.step_lo += (uint32_t)(a - pos);   // step++ runs on every sample (whb.cpp:705)
			uint32_t lw = (a == 0) ? prev_last : d[a - 1];
			for (int m = a; m < b; m++) {
				const uint32_t cw = d[m];
				const int i = (int)(int16_t)(cw & 0xffff), q = (int)(int16_t)(cw >> 16);
				const int li = (int)(int16_t)(lw & 0xffff), lq = (int)(int16_t)(lw >> 16);
				const int pwr = abs(i) + abs(q);
				w.pos = (base_block + tile) * (int64_t)kBlockDec + m;
				if (cfg.kind == K_TFA1) tfa1_sample(w, thresh, pwr, 2 * m, i, q, li, lq);
				else if (cfg.kind == K_WHB) whb_sample(w, cfg, thresh, pwr, i, q, li, lq);
				else tfa2_sample(w, cfg, thresh, pwr, 2 * m, i, q, li, lq);
				lw = cw;
			}
			pos = b;
		}
		if (cfg.kind == K_WHB) w.s.step_lo += (uint32_t)(kBlockDec - pos);
		carry_in = td.carry_out;
		prev_last = d[kBlockDec - 1];
	}
	st->d[w.demod] = w.s;
	if (p.tap_cap)
		for (int k = 0; k < 3; k++) p.tap_cnt[((size_t)w.stream * kMaxDemods + w.demod) * 3 + k] = w.tap_n[k];
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
	if (p.frames[fi].status >= 0) continue;   // parsed by an earlier tfr_process
	DevFrame *f = p.frames + fi;
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
	const int n = p.n_streams * n_demods;
	walk_kernel<<<(n + 31) / 32, 32, 0, s>>>(p);
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
