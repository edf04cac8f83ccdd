// tfr_api.cu - the C ABI declared in include/tfr.h: handle management, the submit/process/poll pump
// that stands in for the loop body of engine::run (engine.cpp:63-93), and the host-side finalisation
// of results (ordering, 10*log10 RSSI with the host libm, sensordata_t mirror).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/tfr.h"
#include "tfr_dev.h"

namespace tfr {
cudaError_t launch_frontend(const FrontParams &p, int n_streams, int wide, cudaStream_t stream);
cudaError_t launch_frontend_tc(const FrontParams &p, int n_streams, int wide, cudaStream_t stream);
bool frontend_tc_available();
int frontend_tc_encode(const void *iq, uint32_t n_blocks, void *out);
cudaError_t launch_save_history(const StreamJob *jobs, StreamState *st, int n_streams, int decimated, cudaStream_t stream);
cudaError_t launch_frontend_i16(const FrontParams &p, int n_streams, cudaStream_t stream);
cudaError_t launch_frontend_screen(const FrontParams &p, int wide, int n_ctas, cudaStream_t stream);
cudaError_t launch_decwin(const BackParams &p, int wide, const uint8_t *hist_copy, uint32_t *dec_out, cudaStream_t s);
void screen_build_consts(int wide, uint8_t *blob, int *shift, int *slack);
constexpr int kScreenConstBytes = 2048 + 8192 + 256;   // frontend_screen.cu: kScBytes
cudaError_t launch_thresh2(const BackParams &p, cudaStream_t s);
cudaError_t launch_devfm(const BackParams &p, cudaStream_t s);
cudaError_t launch_devfm_win(const BackParams &p, cudaStream_t s);
cudaError_t launch_devfm_first(const BackParams &p, cudaStream_t s);
cudaError_t launch_biq(const BackParams &p, int n_demods, cudaStream_t s);
cudaError_t launch_win(const BackParams &p, int n_demods, cudaStream_t s);
cudaError_t launch_winlong(const BackParams &p, int n_demods, cudaStream_t s);
cudaError_t launch_verify(const BackParams &p, int n_demods, cudaStream_t s);
cudaError_t launch_walk(const BackParams &p, int n_demods, cudaStream_t s);
cudaError_t launch_submit_epilogue(const BackParams &p, cudaStream_t s);
cudaError_t launch_parse(const BackParams &p, cudaStream_t s);
cudaError_t launch_downconvert_fused(const uint8_t *iq, const uint8_t *hist, long long n_pairs, int passes, int wide, int16_t *out, cudaStream_t s);
cudaError_t launch_dc_hist(const uint8_t *iq, long long n_pairs, const uint8_t *old_hist, uint8_t *new_hist, cudaStream_t s);
cudaError_t launch_downconvert(const void *iq, long long n_pairs, int passes, int wide, int16_t *tmp0, int16_t *tmp1, int16_t **result,
			       cudaStream_t s, bool in_i16);
}  // namespace tfr

using namespace tfr;

static_assert(sizeof(tfr_config) == 40 && sizeof(tfr_frame) == 112 && sizeof(tfr_record) == 72 && sizeof(tfr_stats) == 144,
	      "public struct layout changed: bump TFR_ABI_VERSION");

static constexpr int kFrontChunks = 8;   // front-end launches per call (each overlaps the previous chunk's threshold walk)
static thread_local std::string g_err;
static int fail(int code, const std::string &msg)
{
	g_err = msg;
	return code;
}
#define CU(call)                                                                                      \
	do {                                                                                          \
		cudaError_t e_ = (call);                                                              \
		if (e_ != cudaSuccess)                                                                \
			return fail(TFR_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));  \
	} while (0)

// iir2 coefficients exactly as the reference binary (x86-64, -ffast-math) computes them in iir2::set
// (dsp_stuff.cpp:36-45) for the cutoffs main.cpp registers; tests/test_abi.py checks them against the
// values read out of the reference object (tests/golden/biquad_coeffs.json).
static BiquadCoef coef_from_bits(const uint64_t b[5])
{
	BiquadCoef c;
	double v[5];
	memcpy(v, b, sizeof(v));
	c.b0 = v[0]; c.b1 = v[1]; c.b2 = v[2]; c.a1 = v[3]; c.a2 = v[4];
	return c;
}
static const uint64_t kCoefTfa2[5] = { 0x3f727f98b1037a14ull, 0x3f827f98b1037a14ull, 0x3f727f98b1037a14ull, 0x3ffcd1527f4a26e2ull, 0xbfea36a1c41c6995ull };
static const uint64_t kCoefTfa3[5] = { 0x3f57ed02b18a270dull, 0x3f67ed02b18a270dull, 0x3f57ed02b18a270dull, 0x3ffe397ac010fc89ull, 0xbfeca2cf85850d62ull };
static const uint64_t kCoefTx22[5] = { 0x3f5461fa1a309718ull, 0x3f6461fa1a309718ull, 0x3f5461fa1a309718ull, 0x3ffe5d4f47377e30ull, 0xbfece36282a35d90ull };
static const uint64_t kCoefWhbPulse[5] = { 0x3f814a67102a1ffdull, 0x3f914a67102a1ffdull, 0x3f814a67102a1ffdull, 0x3ffb949652fa3970ull, 0xbfe83dd316f714e0ull };
static const uint64_t kCoefWhbAvg[5] = { 0x3e502ae4cfc8910aull, 0x3e602ae4cfc8910aull, 0x3e502ae4cfc8910aull, 0x3ffffe9409fe171bull, 0xbfeffd283451f7d3ull };

struct PendingSubmit {
	const uint8_t *dev_ptr = nullptr;
	size_t nbytes = 0;
	bool pending = false;
	bool decimated = false;   // int16 I,Q at 384 kS/s (tfr_submit_decimated) instead of raw bytes
};

struct tfr_handle {
	tfr_config cfg;
	DevConfig dcfg;
	int device = 0;
	cudaStream_t stream = nullptr;
	// persistent device state
	DevConfig *d_cfg = nullptr;
	StreamState *d_state = nullptr;
	Counters *d_counters = nullptr;
	DevFrame *d_frames = nullptr;
	DevRecord *d_records = nullptr;
	uint32_t max_frames = 0, max_records = 0;
	// per-call work buffers (grown on demand), two slots: the back-end of call i (demodulators, verifier, parsers;
	// latency bound, a few SMs) runs on its own CUDA stream while the front-end and threshold walk of call i+1
	// (throughput bound) already use the other slot
	struct Slot {
		StreamJob *d_jobs = nullptr;
		uint8_t *d_tmaps = nullptr;              // [stream][2] CUtensorMap of the call's submits (frontend_tc_kernel)
		std::vector<uint8_t> h_tmaps;            // host copy (stays until the slot's next call)
		uint32_t *d_work_ctr = nullptr;          // [kFrontChunks] screening front-end: work counter of every chunk launch
		DemodState *d_fin = nullptr;             // [stream][kMaxDemods]: state left by a window the call's data ended in
		cudaEvent_t early_done = nullptr;        // decwin + fm_dev of the call's early back-end part are done
		uint8_t *d_hist_copy = nullptr;          // [stream][kHistBytes]: the FIR history the call started from
		cudaEvent_t raw_done = nullptr;          // decwin_kernel has read the call's raw bytes (the input arena may be rewritten)
		TileDesc *d_tiles = nullptr;
		uint32_t *d_dec = nullptr;
		BlockTrace *d_trace = nullptr;
		uint32_t *d_events = nullptr;
		uint32_t *d_walk_tab = nullptr;          // [cap_blocks][kWalkNT][4]: walk_table_kernel -> thresh2_kernel / walk_cta_kernel
		uint32_t *d_walk_gap = nullptr;          // [cap_blocks][kWalkNT][4]: ... -> walk_cta_kernel
		int32_t *d_walk_base = nullptr;          // [n_streams]
		int32_t *d_devfm = nullptr;
		int32_t *d_ld = nullptr;                 // filter chains: (int)y per fm demodulator, direct mapped like d_devfm
		BiqRec *d_biq = nullptr;                 // filter chain records, indexed like d_wins
		WinEntry *d_wins = nullptr;
		WinRec *d_recs = nullptr;
		WinCount *d_wincnt = nullptr;
		WinCount *d_partcnt = nullptr;           // [kMaxParts][stream]: window counts at the back-end part boundaries
		size_t cap_blocks = 0, cap_wins = 0;
		cudaEvent_t front_done = nullptr, back_done = nullptr;   // end of the slot's last front / back work
		cudaEvent_t fe0 = nullptr, fe1 = nullptr;                // around the speculative front-end launch
	};
	Slot slot[2];
	int cur = 0;                       // slot of the most recent tfr_process
	cudaStream_t stream_early2 = nullptr;  // ... of the calls in the other work-buffer slot: two calls' early parts may overlap
	cudaStream_t stream_early = nullptr;   // early back-end part of a call (decwin, fm_dev, the windows that do not need the previous
	                                       // call's final state): runs beside the previous call's verifier on the back stream
	int be_split = 1;                  // TFR_BE_SPLIT: 0 = the whole back-end of a call after the previous call's, 1 = split when the
	                                   // previous call is still in flight (default), 2 = always split (tests)
	std::vector<int64_t> stream_blocks;   // per stream: blocks decoded by the calls issued so far (StreamJob::base_blocks)
	cudaStream_t stream_be = nullptr;  // back-end stream
	cudaStream_t stream_long = nullptr; // winlong_kernel (the long window chains, one warp each) beside win_kernel; high priority
	cudaEvent_t long_ev[2] = { nullptr, nullptr };
	bool long_split = true;
	int be_parts = 1;                  // back-end parts per call (TFR_BE_PARTS): part j runs beside the front-end of the chunks after
	                                   // it.  Default 1: measured on B200 (profiles/r2_backend_parts.txt) the window kernels and the
	                                   // front-end slow each other down by more than the overlap buys (4.02 ms per call as one part,
	                                   // 4.62 / 4.69 / 4.91 ms as 2 / 4 / 8 parts)
	cudaEvent_t part_ev[kMaxParts] = { nullptr };
	// a part's window kernels run on streams of their own (a window chain is a long serial walk: parts queued one after
	// the other on one stream would add their latencies), joined before the verifier
	cudaStream_t part_stream[kMaxParts] = { nullptr }, part_long[kMaxParts] = { nullptr };
	cudaEvent_t part_fm[kMaxParts] = { nullptr }, part_done[kMaxParts] = { nullptr }, part_ldone[kMaxParts] = { nullptr };
	int t_min = 0x7fffffff;            // shortest demodulator timeout
	int walk_ct = 0, walk_dbg = 0;     // TFR_WALK_CT (CTA size of walk_cta_kernel, 0 = by the number of streams), TFR_WALK_DBG (experiments)
	bool walk_cta = true;              // ... by walk_cta_kernel, a CTA per stream (TFR_WALK=warp: by thresh2_kernel, a warp per stream)
	bool walk_table = true;            // threshold walk from the per-block table of walk_table_kernel (TFR_WALK_TAB=0: from the event lists)
	size_t min_chunk = 8192;           // blocks per front-end chunk launch at least (TFR_MIN_CHUNK: tests exercise chunks and parts on small inputs)
	bool use_screen = false;           // screening front-end (frontend_screen.cu): the tensor core proves which samples cannot trigger,
	                                   // exact FIR only for candidates and inside demodulator windows.  Default; TFR_FE=dense selects
	                                   // the dense exact kernel of frontend.cu for every block
	uint8_t *d_screen_consts = nullptr;
	int screen_shift = 0, screen_slack = 0;
	uint32_t *d_screen_stat = nullptr; // [4] sparse blocks, dense blocks, candidates, true bound-triggers (since create)
	// Bursty input (a fixed threshold inside the noise, a continuous strong signal): when most blocks come back from the screen
	// as bursts, screening only adds work - the in-place cascade runs at a third of the dense kernel's occupancy.  The counters
	// of every call are copied to the host behind its front-end; a call after one with more than a quarter of burst blocks takes
	// the dense kernel, every kDenseProbe-th call tries the screen again.
	uint32_t *h_screen_stat = nullptr;         // pinned [4]: snapshot behind the latest screened call's front-end
	cudaEvent_t ev_screen_stat = nullptr;
	uint32_t screen_seen[2] = { 0, 0 };        // sparse / burst blocks of the snapshot before
	bool screen_stat_pending = false;
	bool dense_mode = false;
	int dense_calls = 0;
	int32_t *d_screen_dbg = nullptr;   // TFR_FLAG_TAPS: screen values of the last call [block][8192][2]
	size_t screen_dbg_blocks = 0;
	int win_demod = -1;                // the demodulator with the longest timeout: its windows contain every sample any demodulator reads
	int n_sms = 148;
	bool use_tc = false;               // TFR_FE=tc: the front-end variant with tensor-core byte->float conversion (frontend_tc.cu,
	                                   // bit-identical, measured 28 % slower on B200: DESIGN.md 4.1b); default frontend.cu
	cudaStream_t stream_walk = nullptr;   // threshold walk of front-end chunk k, concurrent with the front-end of chunk k+1
	cudaStream_t stream_fe2 = nullptr;    // odd front-end chunks: consecutive chunk launches overlap their tails
	cudaEvent_t chunk_ev[8] = { nullptr };
	cudaEvent_t walk_ev = nullptr;
	cudaEvent_t dbg_ev[4] = { nullptr };   // TFR_DEBUG: walk end, back start, back end of the last call
	bool pipelined = true;             // false: the back-end of a call finishes before the next call starts (taps)
	cudaEvent_t span0 = nullptr, span1 = nullptr;   // first front-end start / last back-end end since the last tfr_sync
	bool span_open = false;
	double fe_ms_acc = 0;              // fallback front-end launches (timed synchronously, rare)
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> fe_pending;   // front-end event pairs not yet read
	bool has_fm = false, has_whb = false;
	int n_fm = 0;                      // fm_dev-using demodulators (each has a slot of Slot::d_ld)
	int fm_slot[kMaxDemods] = { -1, -1, -1, -1, -1 };
	bool biq_chains = false;           // TFR_BE=chains: filter chains ahead of the slicers (biq_kernel + biq_verify_kernel, exact, but on
	                                   // B200 the two latency-bound filter kernels cost 1.1 ms where they save 0.9: DESIGN.md 4.3b);
	                                   // default: the window kernels warm up and filter for themselves
	int fm_demod = -1;                 // the fm_dev-using demodulator with the longest timeout: its windows contain the others'
	// input arena for host submits: normally one chunk; more are added when a later submit does not
	// fit while earlier ones are still pending, and merged into one the next time the arena is idle
	struct Chunk { uint8_t *ptr; size_t cap, used; };
	std::vector<Chunk> arena;
	// taps
	uint32_t tap_cap = 0;
	int32_t *d_tap_i32[2] = { nullptr, nullptr };
	double *d_tap_f64 = nullptr;
	uint32_t *d_tap_cnt = nullptr;
	// bookkeeping
	std::vector<PendingSubmit> pend;
	std::vector<StreamJob> jobs;       // jobs of the last tfr_process
	std::vector<cudaEvent_t> ev;       // spare events for fallback front-end launches
	uint32_t *d_progress = nullptr;    // [stream] blocks walked by the threshold kernel
	uint32_t *h_progress = nullptr;    // pinned host copy
	cudaEvent_t ev_h2d0 = nullptr, ev_h2d1 = nullptr;
	bool h2d_timed = false;
	tfr_stats stats;
	// host copies of results
	bool results_valid = false;
	bool overflow_reported = false;    // TFR_E_OVERFLOW is returned by one poll, then the truncated results are served
	std::vector<tfr_frame> frames;
	std::vector<tfr_record> records;
};

static const char *kind_name(int k)
{
	static const char *n[] = { "TFA_1", "TFA_2", "TFA_3", "TX22", "WHB" };
	return n[k];
}

// demod registration in the reference's order with its spb constants, main.cpp:171-218
static void build_config(const tfr_config &c, DevConfig &d)
{
	memset(&d, 0, sizeof(d));
	d.filter = c.filter ? 1 : 0;
	d.thresh_cfg = c.thresh;
	d.flags = (int32_t)c.flags;
	d.n_streams = c.n_streams;
	auto add = [&](int kind, int type, double spb, int timeout, const uint64_t *lp, const uint64_t *lp_avg) {
		DemodCfg &q = d.d[d.n_demods++];
		q.kind = kind;
		q.type = type;
		q.spb = spb;
		q.timeout = timeout;
		if (lp) q.lp = coef_from_bits(lp);
		if (lp_avg) q.lp_avg = coef_from_bits(lp_avg);
		{
			// tfa2.cpp:391-398 in integers: (double)tdiff > spb/4  <=>  tdiff >= floor(spb/4) + 1;  (double)tdiff < 32*spb  <=>
			// tdiff <= ceil(32*spb) - 1;  numbits by table, computed here with the very operations the device used
			const volatile double spb_lo = spb * 0.25, spb_hi = 32.0 * spb, spb_half = spb * 0.5;
			q.td_lo = (int)floor(spb_lo) + 1;
			q.td_hi = (int)ceil(spb_hi) - 1;
			for (int bd = 0; bd < kNbitsTab; bd++) {
				const volatile double a = (double)bd + spb_half;
				const volatile double r = a / spb;
				const int nb = (int)r;
				q.nbits[bd] = (uint8_t)(nb > 255 ? 255 : nb);
			}
		}
		d.t_max = std::max(d.t_max, timeout);
	};
	if (c.types & (1 << TFR_TFA_1)) add(K_TFA1, TFR_TFA_1, 10.0, 400, nullptr, nullptr);   // 40*BITPERIOD, tfa1.cpp:34,148
	if (c.types & (1 << TFR_TFA_2)) { const double spb = (1536000 / 4.0) / 17240; add(K_TFA2, TFR_TFA_2, spb, (int)(16 * spb), kCoefTfa2, nullptr); }
	if (c.types & (1 << TFR_TFA_3)) { const double spb = (1536000 / 4.0) / 9600; add(K_TFA3, TFR_TFA_3, spb, (int)(16 * spb), kCoefTfa3, nullptr); }
	if (c.types & (1 << TFR_TX22)) { const double spb = (1536000 / 4.0) / 8842; add(K_TX22, TFR_TX22, spb, (int)(16 * spb), kCoefTx22, nullptr); }
	if (c.types & (1 << TFR_TFA_WHB)) { const double spb = (1536000 / 4.0) / 6000; add(K_WHB, TFR_TFA_WHB, spb, (int)(8 * spb), kCoefWhbPulse, kCoefWhbAvg); }
	if (d.n_demods == 0) d.t_max = 400;   // -T 0: decimator + trigger bookkeeping only
}

static void init_state(const DevConfig &d, StreamState &s)
{
	memset(&s, 0, sizeof(s));
	memset(s.hist, 128, sizeof(s.hist));   // zero FIR history == byte 128 (decimate::decimate, dsp_stuff.cpp:145-152)
	// fsk_demod::fsk_demod, fm_demod.cpp:19-32
	s.thresh = d.thresh_cfg;
	s.thresh_mode = 0;
	if (d.thresh_cfg == 0) {
		s.thresh = 500;
		s.thresh_mode = 1;
	}
	s.spec_lo = s.thresh - spec_margin(s.thresh);
	for (int k = 0; k < d.n_demods; k++) {
		DemodState &q = s.d[k];
		q.sr_cnt = -1;
		if (d.d[k].kind != K_TFA1) {   // tfa2_demod::reset (tfa2.cpp:325-334) / whb_demod::reset run in the constructors
			q.dmin = 32767;
			q.dmax = -32767;
		}
	}
}

extern "C" __attribute__((visibility("default"))) int tfr_abi_version(void) { return TFR_ABI_VERSION; }
extern "C" __attribute__((visibility("default"))) const char *tfr_last_error(void) { return g_err.c_str(); }

extern "C" __attribute__((visibility("default"))) void tfr_destroy(tfr_handle *h)
{
	if (!h) return;
	cudaSetDevice(h->device);
	if (h->stream) cudaStreamSynchronize(h->stream);
	if (h->stream_long) cudaStreamSynchronize(h->stream_long);
	for (auto st_ : h->part_stream) if (st_) cudaStreamSynchronize(st_);
	for (auto st_ : h->part_long) if (st_) cudaStreamSynchronize(st_);
	if (h->stream_early) cudaStreamSynchronize(h->stream_early);
	if (h->stream_early2) cudaStreamSynchronize(h->stream_early2);
	if (h->stream_be) cudaStreamSynchronize(h->stream_be);
	cudaFree(h->d_cfg); cudaFree(h->d_state); cudaFree(h->d_counters);
	cudaFree(h->d_frames); cudaFree(h->d_records);
	cudaFree(h->d_screen_consts); cudaFree(h->d_screen_stat); cudaFree(h->d_screen_dbg);
	if (h->h_screen_stat) cudaFreeHost(h->h_screen_stat);
	if (h->ev_screen_stat) cudaEventDestroy(h->ev_screen_stat);
	for (auto &sl : h->slot) {
		cudaFree(sl.d_work_ctr); cudaFree(sl.d_hist_copy); cudaFree(sl.d_fin); cudaFree(sl.d_walk_base);
		if (sl.early_done) cudaEventDestroy(sl.early_done);
		if (sl.raw_done) cudaEventDestroy(sl.raw_done);
		cudaFree(sl.d_jobs); cudaFree(sl.d_tmaps); cudaFree(sl.d_tiles); cudaFree(sl.d_dec); cudaFree(sl.d_trace); cudaFree(sl.d_events); cudaFree(sl.d_walk_tab); cudaFree(sl.d_walk_gap);
		cudaFree(sl.d_devfm); cudaFree(sl.d_wins); cudaFree(sl.d_recs); cudaFree(sl.d_wincnt); cudaFree(sl.d_partcnt);
		cudaFree(sl.d_ld); cudaFree(sl.d_biq);
		for (cudaEvent_t e : { sl.front_done, sl.back_done, sl.fe0, sl.fe1 })
			if (e) cudaEventDestroy(e);
	}
	for (auto &c : h->arena) cudaFree(c.ptr); cudaFree(h->d_tap_i32[0]); cudaFree(h->d_tap_i32[1]);
	cudaFree(h->d_tap_f64); cudaFree(h->d_tap_cnt); cudaFree(h->d_progress);
	if (h->h_progress) cudaFreeHost(h->h_progress);
	for (auto e : h->ev) cudaEventDestroy(e);
	if (h->ev_h2d0) cudaEventDestroy(h->ev_h2d0);
	if (h->ev_h2d1) cudaEventDestroy(h->ev_h2d1);
	if (h->span0) cudaEventDestroy(h->span0);
	if (h->span1) cudaEventDestroy(h->span1);
	if (h->stream_be) cudaStreamDestroy(h->stream_be);
	if (h->stream_early) cudaStreamDestroy(h->stream_early);
	if (h->stream_early2) cudaStreamDestroy(h->stream_early2);
	if (h->stream_long) cudaStreamDestroy(h->stream_long);
	for (auto &e : h->long_ev) if (e) cudaEventDestroy(e);
	for (auto st_ : h->part_stream) if (st_) cudaStreamDestroy(st_);
	for (auto st_ : h->part_long) if (st_) cudaStreamDestroy(st_);
	for (auto *arr : { h->part_ev, h->part_fm, h->part_done, h->part_ldone })
		for (int k = 0; k < kMaxParts; k++) if (arr[k]) cudaEventDestroy(arr[k]);
	if (h->stream_walk) cudaStreamDestroy(h->stream_walk);
	if (h->stream_fe2) cudaStreamDestroy(h->stream_fe2);
	for (auto e : h->chunk_ev) if (e) cudaEventDestroy(e);
	if (h->walk_ev) cudaEventDestroy(h->walk_ev);
	for (auto e : h->dbg_ev) if (e) cudaEventDestroy(e);
	if (h->stream) cudaStreamDestroy(h->stream);
	delete h;
}

extern "C" __attribute__((visibility("default"))) int tfr_create(const tfr_config *cfg, tfr_handle **out)
{
	if (!cfg || !out) return fail(TFR_E_INVAL, "tfr_create: null argument");
	if (cfg->struct_size != sizeof(tfr_config)) return fail(TFR_E_INVAL, "tfr_create: struct_size mismatch (ABI version?)");
	if (cfg->n_streams < 1 || cfg->n_streams > 65535) return fail(TFR_E_INVAL, "tfr_create: n_streams out of range");
	if (cfg->thresh < 0) return fail(TFR_E_INVAL, "tfr_create: negative threshold");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
		return fail(TFR_E_NODEVICE, "no CUDA device (this library has no CPU fallback)");
	if (cfg->device < 0 || cfg->device >= ndev) return fail(TFR_E_NODEVICE, "tfr_create: device ordinal out of range");
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, cfg->device));
	if (prop.major != 10)
		return fail(TFR_E_NODEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
						    ", kernels are built for sm_100a only");
	CU(cudaSetDevice(cfg->device));
	{   // the serial slow paths (window re-runs) keep a whole demodulator state on the stack
		size_t lim = 0;
		CU(cudaDeviceGetLimit(&lim, cudaLimitStackSize));
		if (lim < 4096) CU(cudaDeviceSetLimit(cudaLimitStackSize, 4096));
	}

	tfr_handle *h = new tfr_handle();
	h->cfg = *cfg;
	h->device = cfg->device;
	memset(&h->stats, 0, sizeof(h->stats));
	build_config(*cfg, h->dcfg);
	for (int k = 0; k < h->dcfg.n_demods; k++) {
		if (h->dcfg.d[k].kind == K_TFA2 || h->dcfg.d[k].kind == K_TFA3 || h->dcfg.d[k].kind == K_TX22) {
			h->has_fm = true;
			h->fm_slot[k] = h->n_fm++;
			if (h->fm_demod < 0 || h->dcfg.d[k].timeout > h->dcfg.d[h->fm_demod].timeout) h->fm_demod = k;
		}
		h->has_whb |= (h->dcfg.d[k].kind == K_WHB);
	}
	h->pend.resize(cfg->n_streams);
	h->max_frames = cfg->max_frames ? cfg->max_frames : 65536u;
	h->max_records = h->max_frames * 5u;
	auto bail = [&](int code) { tfr_destroy(h); return code; };
#define CUH(call)                                                                                          \
	do {                                                                                               \
		cudaError_t e_ = (call);                                                                   \
		if (e_ != cudaSuccess)                                                                     \
			return bail(fail(e_ == cudaErrorMemoryAllocation ? TFR_E_NOMEM : TFR_E_CUDA,       \
					 std::string(#call) + ": " + cudaGetErrorString(e_)));             \
	} while (0)
	CUH(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
	// same priority as the front stream: measured on B200, a high-priority back-end stream shortens a pipelined
	// call from 4.35 to 4.0 ms but stretches the issue-bound front-end kernel from 1.9 to 3.15 ms
	CUH(cudaStreamCreateWithFlags(&h->stream_be, cudaStreamNonBlocking));
	CUH(cudaStreamCreateWithFlags(&h->stream_early, cudaStreamNonBlocking));
	CUH(cudaStreamCreateWithFlags(&h->stream_early2, cudaStreamNonBlocking));
	h->stream_blocks.assign(cfg->n_streams, 0);
	if (const char *sp = getenv("TFR_BE_SPLIT")) h->be_split = atoi(sp);
	{   // the walk is 1 warp per stream on the call's critical path: its CTAs go first when an SM has room
		int prio_lo = 0, prio_hi = 0;
		CUH(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
		CUH(cudaStreamCreateWithPriority(&h->stream_walk, cudaStreamNonBlocking, prio_hi));
		CUH(cudaStreamCreateWithPriority(&h->stream_long, cudaStreamNonBlocking, prio_hi));
		for (auto &e : h->long_ev) CUH(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		if (getenv("TFR_NO_LONG")) h->long_split = false;   // experiments: every chain in the thread-per-chain kernel
		if (const char *bp_env = getenv("TFR_BE_PARTS")) h->be_parts = std::max(1, std::min(atoi(bp_env), (int)kMaxParts));
		for (auto &e : h->part_ev) CUH(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		for (auto &e : h->part_fm) CUH(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		for (auto &e : h->part_done) CUH(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		for (auto &e : h->part_ldone) CUH(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		for (int k = 0; k < std::max(h->be_parts, 2) && k < kMaxParts; k++) {   // (two at least: the split back-end uses one pair per slot)
			CUH(cudaStreamCreateWithFlags(&h->part_stream[k], cudaStreamNonBlocking));
			CUH(cudaStreamCreateWithPriority(&h->part_long[k], cudaStreamNonBlocking, prio_hi));
		}
		// The table walk pays where a stream's serial chain is the long pole: few, long streams.  With many streams the
		// list walk's warps hide behind the front-end anyway and the table kernel takes SM time from it (32 streams x 128 MiB
		// on B200: the same step either way, front-end 0.85 -> 0.90 ms): from 17 streams on the lists are walked.
		h->walk_table = cfg->n_streams <= 16;
		if (const char *wt = getenv("TFR_WALK_TAB")) h->walk_table = atoi(wt) != 0;
		if (const char *wk = getenv("TFR_WALK")) h->walk_cta = strcmp(wk, "warp") != 0;
		if (const char *wc = getenv("TFR_WALK_CT")) h->walk_ct = atoi(wc);
		if (const char *wd = getenv("TFR_WALK_DBG")) h->walk_dbg = atoi(wd);
		for (int k = 0; k < h->dcfg.n_demods; k++) h->t_min = std::min(h->t_min, h->dcfg.d[k].timeout);
		if (const char *mc = getenv("TFR_MIN_CHUNK")) h->min_chunk = (size_t)std::max(1, atoi(mc));
		if (const char *be = getenv("TFR_BE")) h->biq_chains = strcmp(be, "chains") == 0;
		const char *fe = getenv("TFR_FE");                  // experiments: TFR_FE=tc selects the tensor-core front-end variant
		h->use_tc = fe && !strcmp(fe, "tc") && frontend_tc_available();
		// default: the screening front-end (needs cuTensorMapEncodeTiled; TFR_FE=dense keeps the dense exact kernel everywhere)
		h->use_screen = !h->use_tc && !(fe && (!strcmp(fe, "dense") || !strcmp(fe, "old"))) && frontend_tc_available();
		if (fe && !strcmp(fe, "screen") && !h->use_screen) return bail(fail(TFR_E_CUDA, "tfr_create: TFR_FE=screen needs cuTensorMapEncodeTiled"));
		if (cfg->n_streams > 4095) h->use_screen = false;   // dense list entries: stream << 20 | block
		CUH(cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, h->device));
		for (int k = 0; k < h->dcfg.n_demods; k++)
			if (h->win_demod < 0 || h->dcfg.d[k].timeout > h->dcfg.d[h->win_demod].timeout) h->win_demod = k;
	}
	if (h->use_screen) {
		std::vector<uint8_t> blob(kScreenConstBytes);
		screen_build_consts(h->dcfg.filter, blob.data(), &h->screen_shift, &h->screen_slack);
		CUH(cudaMalloc(&h->d_screen_consts, kScreenConstBytes));
		CUH(cudaMemcpy(h->d_screen_consts, blob.data(), kScreenConstBytes, cudaMemcpyHostToDevice));
		CUH(cudaMalloc(&h->d_screen_stat, 4 * sizeof(uint32_t)));
		CUH(cudaMemset(h->d_screen_stat, 0, 4 * sizeof(uint32_t)));
		CUH(cudaMallocHost(&h->h_screen_stat, 4 * sizeof(uint32_t)));
		memset(h->h_screen_stat, 0, 4 * sizeof(uint32_t));
		CUH(cudaEventCreateWithFlags(&h->ev_screen_stat, cudaEventDisableTiming));
	}
	CUH(cudaStreamCreateWithFlags(&h->stream_fe2, cudaStreamNonBlocking));
	for (auto &e : h->chunk_ev) CUH(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	CUH(cudaEventCreateWithFlags(&h->walk_ev, cudaEventDisableTiming));
	for (auto &e : h->dbg_ev) CUH(cudaEventCreate(&e));
	h->pipelined = !(cfg->flags & TFR_FLAG_TAPS);   // the tap buffers are not slotted
	CUH(cudaEventCreate(&h->span0));
	CUH(cudaEventCreate(&h->span1));
	for (auto &sl : h->slot) {
		CUH(cudaEventCreateWithFlags(&sl.front_done, cudaEventDisableTiming));
		CUH(cudaEventCreateWithFlags(&sl.back_done, cudaEventDisableTiming));
		CUH(cudaEventCreate(&sl.fe0));
		CUH(cudaEventCreate(&sl.fe1));
		CUH(cudaMalloc(&sl.d_jobs, sizeof(StreamJob) * cfg->n_streams));
		CUH(cudaMalloc(&sl.d_tmaps, (size_t)256 * cfg->n_streams));
		CUH(cudaEventCreateWithFlags(&sl.raw_done, cudaEventDisableTiming));
		CUH(cudaEventCreateWithFlags(&sl.early_done, cudaEventDisableTiming));
		CUH(cudaMalloc(&sl.d_fin, sizeof(DemodState) * kMaxDemods * cfg->n_streams));
		CUH(cudaMemset(sl.d_fin, 0, sizeof(DemodState) * kMaxDemods * cfg->n_streams));
		CUH(cudaMalloc(&sl.d_walk_base, sizeof(int32_t) * cfg->n_streams));
		CUH(cudaMalloc(&sl.d_work_ctr, sizeof(uint32_t) * kFrontChunks));
		CUH(cudaMemset(sl.d_work_ctr, 0, sizeof(uint32_t) * kFrontChunks));
		CUH(cudaMalloc(&sl.d_hist_copy, (size_t)kHistBytes * cfg->n_streams));
		CUH(cudaMemset(sl.d_hist_copy, 128, (size_t)kHistBytes * cfg->n_streams));
		sl.h_tmaps.assign((size_t)256 * cfg->n_streams + 64, 0);
		CUH(cudaMalloc(&sl.d_wincnt, sizeof(WinCount) * cfg->n_streams));
		CUH(cudaMemset(sl.d_wincnt, 0, sizeof(WinCount) * cfg->n_streams));
		CUH(cudaMalloc(&sl.d_partcnt, sizeof(WinCount) * kMaxParts * cfg->n_streams));
		CUH(cudaMemset(sl.d_partcnt, 0, sizeof(WinCount) * kMaxParts * cfg->n_streams));
	}
	CUH(cudaEventCreate(&h->ev_h2d0));
	CUH(cudaEventCreate(&h->ev_h2d1));
	CUH(cudaMalloc(&h->d_cfg, sizeof(DevConfig)));
	CUH(cudaMemcpy(h->d_cfg, &h->dcfg, sizeof(DevConfig), cudaMemcpyHostToDevice));
	CUH(cudaMalloc(&h->d_state, sizeof(StreamState) * cfg->n_streams));
	{
		std::vector<StreamState> init(cfg->n_streams);
		for (auto &s : init) init_state(h->dcfg, s);
		CUH(cudaMemcpy(h->d_state, init.data(), sizeof(StreamState) * cfg->n_streams, cudaMemcpyHostToDevice));
	}
	CUH(cudaMalloc(&h->d_progress, sizeof(uint32_t) * cfg->n_streams));
	CUH(cudaMemset(h->d_progress, 0, sizeof(uint32_t) * cfg->n_streams));
	CUH(cudaMallocHost(&h->h_progress, sizeof(uint32_t) * cfg->n_streams));
	CUH(cudaMalloc(&h->d_counters, sizeof(Counters)));
	CUH(cudaMemset(h->d_counters, 0, sizeof(Counters)));
	CUH(cudaMalloc(&h->d_frames, sizeof(DevFrame) * h->max_frames));
	CUH(cudaMalloc(&h->d_records, sizeof(DevRecord) * h->max_records));
	if (cfg->flags & TFR_FLAG_TAPS) {
		h->tap_cap = 1u << 23;
		const size_t n = (size_t)cfg->n_streams * kMaxDemods * h->tap_cap;
		CUH(cudaMalloc(&h->d_tap_i32[0], n * sizeof(int32_t)));
		CUH(cudaMalloc(&h->d_tap_i32[1], n * sizeof(int32_t)));
		CUH(cudaMalloc(&h->d_tap_f64, n * sizeof(double)));
		CUH(cudaMalloc(&h->d_tap_cnt, (size_t)cfg->n_streams * kMaxDemods * 3 * sizeof(uint32_t)));
		CUH(cudaMemset(h->d_tap_cnt, 0, (size_t)cfg->n_streams * kMaxDemods * 3 * sizeof(uint32_t)));
	}
#undef CUH
	*out = h;
	return TFR_OK;
}

static int sync_all(tfr_handle *h)
{
	CU(cudaStreamSynchronize(h->stream));
	CU(cudaStreamSynchronize(h->stream_fe2));
	CU(cudaStreamSynchronize(h->stream_walk));
	CU(cudaStreamSynchronize(h->stream_long));
	for (auto st_ : h->part_stream) if (st_) CU(cudaStreamSynchronize(st_));
	for (auto st_ : h->part_long) if (st_) CU(cudaStreamSynchronize(st_));
	CU(cudaStreamSynchronize(h->stream_early));
	CU(cudaStreamSynchronize(h->stream_early2));
	CU(cudaStreamSynchronize(h->stream_be));
	return TFR_OK;
}

static int ensure_blocks(tfr_handle *h, tfr_handle::Slot &sl, size_t blocks, size_t wins)
{
	if (blocks > sl.cap_blocks) {
		int rc = sync_all(h);
		if (rc) return rc;
		cudaFree(sl.d_tiles); cudaFree(sl.d_dec); cudaFree(sl.d_trace); cudaFree(sl.d_events); cudaFree(sl.d_walk_tab); cudaFree(sl.d_walk_gap); cudaFree(sl.d_devfm); cudaFree(sl.d_ld);
		sl.d_tiles = nullptr; sl.d_dec = nullptr; sl.d_trace = nullptr; sl.d_events = nullptr; sl.d_walk_tab = nullptr; sl.d_walk_gap = nullptr; sl.d_devfm = nullptr; sl.d_ld = nullptr;
		sl.cap_blocks = 0;
		cudaError_t e = cudaMalloc(&sl.d_tiles, blocks * sizeof(TileDesc));
		if (e == cudaSuccess) e = cudaMalloc(&sl.d_dec, blocks * (size_t)kBlockDec * sizeof(uint32_t));
		if (e == cudaSuccess) e = cudaMalloc(&sl.d_trace, blocks * sizeof(BlockTrace));
		if (e == cudaSuccess) e = cudaMalloc(&sl.d_events, blocks * (size_t)kMaxEvt * sizeof(uint32_t));
		if (e == cudaSuccess && h->walk_table) e = cudaMalloc(&sl.d_walk_tab, blocks * (size_t)kWalkNT * 4 * sizeof(uint32_t));
		if (e == cudaSuccess && h->walk_table && h->walk_cta) e = cudaMalloc(&sl.d_walk_gap, blocks * (size_t)kWalkNT * 4 * sizeof(uint32_t));
		if (e == cudaSuccess && h->has_fm) e = cudaMalloc(&sl.d_devfm, blocks * (size_t)kBlockDec * sizeof(int32_t));
		if (e == cudaSuccess && h->has_fm && h->biq_chains) e = cudaMalloc(&sl.d_ld, (size_t)h->n_fm * blocks * (size_t)kBlockDec * sizeof(int32_t));
		if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorMemoryAllocation ? TFR_E_NOMEM : TFR_E_CUDA, std::string("work buffers: ") + cudaGetErrorString(e)); }
		sl.cap_blocks = blocks;
	}
	if (wins > sl.cap_wins) {
		int rc = sync_all(h);
		if (rc) return rc;
		cudaFree(sl.d_wins); cudaFree(sl.d_recs); cudaFree(sl.d_biq);
		sl.d_wins = nullptr; sl.d_recs = nullptr; sl.d_biq = nullptr;
		sl.cap_wins = 0;
		cudaError_t e = cudaMalloc(&sl.d_wins, wins * sizeof(WinEntry));
		if (e == cudaSuccess) e = cudaMalloc(&sl.d_recs, wins * sizeof(WinRec));
		if (e == cudaSuccess && h->has_fm && h->biq_chains) e = cudaMalloc(&sl.d_biq, wins * sizeof(BiqRec));
		if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorMemoryAllocation ? TFR_E_NOMEM : TFR_E_CUDA, std::string("window lists: ") + cudaGetErrorString(e)); }
		sl.cap_wins = wins;
	}
	return TFR_OK;
}

static int submit_common(tfr_handle *h, int stream, const uint8_t *iq, size_t nbytes, int mem, bool decimated)
{
	if (!h || !iq) return fail(TFR_E_INVAL, "tfr_submit: null argument");
	if (stream < 0 || stream >= h->cfg.n_streams) return fail(TFR_E_INVAL, "tfr_submit: stream out of range");
	const size_t blk = decimated ? (size_t)kBlockDec * 4 : (size_t)TFR_BLOCK_BYTES;
	if (nbytes == 0 || nbytes % blk) return fail(TFR_E_INVAL, decimated ? "tfr_submit_decimated: the int16 count must be a positive multiple of 16384" : "tfr_submit: nbytes must be a positive multiple of 65536");
	if (nbytes / blk > 262143ull) return fail(TFR_E_INVAL, "tfr_submit: at most 262143 blocks (16 GiB) per stream per call");
	for (auto &q : h->pend)
		if (q.pending && q.decimated != decimated) return fail(TFR_E_INVAL, "tfr_submit: raw and decimated submits cannot be mixed in one call");
	PendingSubmit &ps = h->pend[stream];
	if (ps.pending) return fail(TFR_E_BUSY, "tfr_submit: stream already has a pending submit");
	CU(cudaSetDevice(h->device));
	if (mem == TFR_MEM_DEVICE) {
		if ((uintptr_t)iq & 15) return fail(TFR_E_INVAL, "tfr_submit: device pointer must be 16-byte aligned");
		ps.dev_ptr = iq;
	} else if (mem == TFR_MEM_HOST) {
		bool idle = true;
		for (auto &q : h->pend) idle &= !q.pending;
		if (idle && h->arena.size() > 1) {   // merge the chunks of the previous round into one
			size_t tot = 0;
			for (auto &c : h->arena) tot += c.cap;
			CU(cudaStreamSynchronize(h->stream));
			CU(cudaStreamSynchronize(h->stream_be));
			for (auto &c : h->arena) cudaFree(c.ptr);
			h->arena.clear();
			uint8_t *ptr = nullptr;
			if (cudaMalloc(&ptr, tot) != cudaSuccess) { cudaGetLastError(); return fail(TFR_E_NOMEM, "tfr_submit: input arena"); }
			h->arena.push_back({ ptr, tot, 0 });
		}
		tfr_handle::Chunk *ck = nullptr;
		for (auto &c : h->arena)
			if (c.used + nbytes <= c.cap) { ck = &c; break; }
		if (!ck) {
			size_t want = std::max(nbytes, (size_t)h->cfg.max_blocks_per_submit * TFR_BLOCK_BYTES);
			if (h->arena.empty()) want *= (size_t)h->cfg.n_streams;
			if (idle && !h->arena.empty()) {   // nothing points into the old chunk: replace it
				CU(cudaStreamSynchronize(h->stream));
				CU(cudaStreamSynchronize(h->stream_be));
				for (auto &c : h->arena) cudaFree(c.ptr);
				h->arena.clear();
			}
			uint8_t *ptr = nullptr;
			cudaError_t e = cudaMalloc(&ptr, want);
			if (e != cudaSuccess) { cudaGetLastError(); return fail(TFR_E_NOMEM, std::string("tfr_submit: input arena: ") + cudaGetErrorString(e)); }
			h->arena.push_back({ ptr, want, 0 });
			ck = &h->arena.back();
		}
		uint8_t *dst = ck->ptr + ck->used;
		// the window kernel of the calls in flight (back-end stream) still reads their raw bytes out of this arena
		if (h->use_screen)
			for (auto &sl : h->slot) CU(cudaStreamWaitEvent(h->stream, sl.raw_done, 0));
		if (!h->h2d_timed) {
			CU(cudaEventRecord(h->ev_h2d0, h->stream));
			h->h2d_timed = true;
		}
		CU(cudaMemcpyAsync(dst, iq, nbytes, cudaMemcpyHostToDevice, h->stream));
		CU(cudaEventRecord(h->ev_h2d1, h->stream));
		ck->used += nbytes;
		ps.dev_ptr = dst;
	} else {
		return fail(TFR_E_INVAL, "tfr_submit: mem must be TFR_MEM_HOST or TFR_MEM_DEVICE");
	}
	ps.nbytes = nbytes;
	ps.pending = true;
	ps.decimated = decimated;
	return TFR_OK;
}

extern "C" __attribute__((visibility("default"))) int tfr_submit(tfr_handle *h, int stream, const uint8_t *iq, size_t nbytes, int mem)
{
	return submit_common(h, stream, iq, nbytes, mem, false);
}

// what fsk_demod::process(int16_t *data_iq, int len) takes (fm_demod.cpp:34): int16 I,Q already at 384 kS/s
extern "C" __attribute__((visibility("default"))) int tfr_submit_decimated(tfr_handle *h, int stream, const int16_t *iq16, size_t n_int16, int mem)
{
	return submit_common(h, stream, reinterpret_cast<const uint8_t *>(iq16), n_int16 * sizeof(int16_t), mem, true);
}

static int ensure_events(tfr_handle *h, size_t n)
{
	while (h->ev.size() < n) {
		cudaEvent_t e;
		CU(cudaEventCreate(&e));
		h->ev.push_back(e);
	}
	return TFR_OK;
}

extern "C" __attribute__((visibility("default"))) int tfr_process(tfr_handle *h)
{
	if (!h) return fail(TFR_E_INVAL, "tfr_process: null handle");
	CU(cudaSetDevice(h->device));
	cudaGetLastError();   // do not inherit a stale error from another user of the runtime
	const int ns = h->cfg.n_streams;
	std::vector<StreamJob> jobs(ns, StreamJob{ nullptr, 0, 0, 0, 0 });
	size_t total = 0, total_wins = 0;
	const int ndm = std::max(h->dcfg.n_demods, 1);
	uint32_t max_blocks = 0;
	bool decimated = false;   // the call's submits hold int16 I,Q at 384 kS/s (all of them: tfr_submit refuses a mix)
	for (int s = 0; s < ns; s++) {
		PendingSubmit &ps = h->pend[s];
		if (!ps.pending) continue;
		StreamJob &j = jobs[s];
		j.iq = ps.dev_ptr;
		decimated = ps.decimated;
		j.n_blocks = (uint32_t)(ps.nbytes / (ps.decimated ? (size_t)kBlockDec * 4 : (size_t)TFR_BLOCK_BYTES));
		j.dec_off = (uint32_t)total;
		j.win_cap = j.n_blocks * (uint32_t)kWinPerBlock + 4u;
		j.win_off = (uint32_t)total_wins;
		j.base_blocks = h->stream_blocks[s];
		h->stream_blocks[s] += j.n_blocks;
		total_wins += (size_t)j.win_cap * ndm;
		total += j.n_blocks;
		max_blocks = std::max(max_blocks, j.n_blocks);
		ps.pending = false;
	}
	for (auto &c : h->arena) c.used = 0;   // contents stay valid until the next submit overwrites them (stream ordered)
	if (total == 0) return TFR_OK;
	if (total > 0x7ffffffull || total_wins > 0xffffffffull) return fail(TFR_E_INVAL, "tfr_process: too many blocks in one call");
	const int si = h->cur ^ 1;
	tfr_handle::Slot &sl = h->slot[si];
	int rc = ensure_blocks(h, sl, total, total_wins);
	if (rc) return rc;
	h->cur = si;
	h->jobs = jobs;
	cudaStream_t sf = h->stream, sb = h->stream_be;
	// this slot's buffers are free once the back-end of the call before the previous one is done; without
	// pipelining the front-end also waits for the previous call's back-end
	CU(cudaStreamWaitEvent(sf, sl.back_done, 0));
	if (!h->pipelined) CU(cudaStreamWaitEvent(sf, h->slot[si ^ 1].back_done, 0));
	if (!h->span_open) {
		CU(cudaEventRecord(h->span0, sf));
		h->span_open = true;
	}
	// h->jobs stays untouched until the copy has run: the next tfr_process first waits for this slot's events
	CU(cudaMemcpyAsync(sl.d_jobs, h->jobs.data(), sizeof(StreamJob) * ns, cudaMemcpyHostToDevice, sf));
	// screening front-end: raw bytes, nothing that wants every decimated sample, the back-end as one part
	bool screen = h->use_screen && !decimated && !(h->cfg.flags & TFR_FLAG_KEEP_DECIM) && h->be_parts == 1 && !getenv("TFR_DEVFM_BLOCKS");
	if (screen) {
		constexpr int kDenseProbe = 16;
		if (h->screen_stat_pending && cudaEventQuery(h->ev_screen_stat) == cudaSuccess) {
			const uint32_t sp = h->h_screen_stat[0] - h->screen_seen[0], bu = h->h_screen_stat[1] - h->screen_seen[1];
			h->screen_seen[0] = h->h_screen_stat[0];
			h->screen_seen[1] = h->h_screen_stat[1];
			h->screen_stat_pending = false;
			if (sp + bu) h->dense_mode = bu > (sp + bu) / 4;
			h->dense_calls = 0;
		} else {
			cudaGetLastError();
		}
		if (h->dense_mode && !getenv("TFR_NO_DENSE_MODE")) {
			if (++h->dense_calls % kDenseProbe) screen = false;   // (every kDenseProbe-th call probes the screen again)
		}
	}
	if ((h->use_tc || screen) && !decimated) {
		// the call's tensor maps (TMA descriptors of every stream's submit), encoded on the host, 64-byte aligned
		uint8_t *tm = reinterpret_cast<uint8_t *>(((uintptr_t)sl.h_tmaps.data() + 63) & ~(uintptr_t)63);
		for (int s = 0; s < ns; s++)
			if (h->jobs[s].n_blocks && frontend_tc_encode(h->jobs[s].iq, h->jobs[s].n_blocks, tm + (size_t)256 * s))
				return fail(TFR_E_CUDA, "tfr_process: cuTensorMapEncodeTiled failed");
		CU(cudaMemcpyAsync(sl.d_tmaps, tm, (size_t)256 * ns, cudaMemcpyHostToDevice, sf));
	}

	// Auto threshold: the whole call is first run against a speculative lower bound of the threshold (see
	// spec_margin in tfr_dev.h).  The threshold kernel stops a stream where the bound fails; the rest of such a
	// stream is redone in epochs of kEpochBlocks blocks, each with a bound that provably holds for the epoch.
	const bool auto_mode = (h->dcfg.thresh_cfg == 0);

	FrontParams fp;
	memset(&fp, 0, sizeof(fp));
	fp.jobs = sl.d_jobs;
	fp.st = h->d_state;
	fp.tiles = sl.d_tiles;
	fp.dec = sl.d_dec;
	fp.t_max = h->dcfg.t_max;
	fp.keep_all = (h->cfg.flags & TFR_FLAG_KEEP_DECIM) ? 1 : 0;
	fp.events = sl.d_events;
	fp.tile0 = 0;
	fp.n_tiles = (int)max_blocks;
	fp.margin = 0;
	fp.use_progress = 0;
	fp.tmaps = sl.d_tmaps;
	fp.n_streams = ns;
	if (screen) {
		fp.screen_consts = h->d_screen_consts;
		fp.screen_shift = h->screen_shift;
		fp.screen_slack = h->screen_slack;
		fp.hist_copy = sl.d_hist_copy;
		fp.screen_stat = h->d_screen_stat;
		if (h->cfg.flags & TFR_FLAG_TAPS) {   // debug: the screen values of this call
			if (h->screen_dbg_blocks < total) {
				CU(cudaStreamSynchronize(sf));
				cudaFree(h->d_screen_dbg);
				h->d_screen_dbg = nullptr;
				h->screen_dbg_blocks = 0;
				CU(cudaMalloc(&h->d_screen_dbg, total * (size_t)kBlockDec * 2 * sizeof(int32_t)));
				h->screen_dbg_blocks = total;
			}
			CU(cudaMemsetAsync(h->d_screen_dbg, 0, total * (size_t)kBlockDec * 2 * sizeof(int32_t), sf));
			fp.screen_dbg = h->d_screen_dbg;
		}
		CU(cudaMemsetAsync(sl.d_work_ctr, 0, sizeof(uint32_t) * kFrontChunks, sf));
	}
	// one chunk of blocks through the screen: persistent CTAs, two per SM, fetching blocks from the chunk's counter
	auto launch_screen = [&](FrontParams q, int chunk, cudaStream_t st) -> cudaError_t {
		q.work_ctr = sl.d_work_ctr + chunk;
		// one persistent CTA per SM and launch: the chunk launches alternate between two streams and overlap pairwise, so two
		// CTAs are resident per SM either way (shared memory allows no more); measured on B200 launches of 296 CTAs stream no
		// faster (0.85 ms per 4 GiB), and with 148 the tail of a launch leaves room for the back-end kernels of the call in
		// flight (TFR_SCREEN_CTAS overrides, profiles/r2_screen_ctas.txt)
		static const int ctas_env = getenv("TFR_SCREEN_CTAS") ? atoi(getenv("TFR_SCREEN_CTAS")) : 0;
		return launch_frontend_screen(q, h->dcfg.filter, ctas_env > 0 ? ctas_env : h->n_sms, st);
	};
	auto launch_fe = [&](const FrontParams &q, cudaStream_t st) {
		if (decimated) return launch_frontend_i16(q, ns, st);
		return h->use_tc ? launch_frontend_tc(q, ns, h->dcfg.filter, st) : launch_frontend(q, ns, h->dcfg.filter, st);
	};
	BackParams bp;
	memset(&bp, 0, sizeof(bp));
	bp.cfg = h->d_cfg;
	bp.jobs = sl.d_jobs;
	bp.st = h->d_state;
	bp.tiles = sl.d_tiles;
	bp.dec = sl.d_dec;
	bp.trace = sl.d_trace;
	bp.frames = h->d_frames;
	bp.records = h->d_records;
	bp.counters = h->d_counters;
	bp.tap_i32[0] = h->d_tap_i32[0];
	bp.tap_i32[1] = h->d_tap_i32[1];
	bp.tap_f64 = h->d_tap_f64;
	bp.tap_cnt = h->d_tap_cnt;
	bp.tap_cap = h->tap_cap;
	bp.max_frames = h->max_frames;
	bp.max_records = h->max_records;
	bp.n_streams = ns;
	bp.events = sl.d_events;
	bp.wins = sl.d_wins;
	bp.recs = sl.d_recs;
	bp.devfm = sl.d_devfm;
	bp.wincnt = sl.d_wincnt;
	bp.ld = nullptr;   // set per launch: filter chains need the call's complete window lists (a call that runs as one part)
	bp.ld_stride = sl.cap_blocks * (size_t)kBlockDec;
	bp.biq = sl.d_biq;
	for (int k = 0; k < kMaxDemods; k++) bp.fm_slot[k] = h->fm_slot[k];
	bp.fin = sl.d_fin;
	bp.slot_tag = si + 1;
	bp.long_all = 0;
	bp.partcnt = sl.d_partcnt;
	bp.part_idx = -1;
	bp.part_lo = bp.part_hi = -1;
	bp.max_blocks = (int)max_blocks;
	bp.tile0 = 0;
	bp.n_tiles = (int)max_blocks;
	bp.margin = 0;
	bp.progress = h->d_progress;
	if (h->d_tap_cnt) CU(cudaMemsetAsync(h->d_tap_cnt, 0, (size_t)ns * kMaxDemods * 3 * sizeof(uint32_t), sf));   // taps cover one tfr_process

	// demodulators over the windows of one back-end part (BackParams::part_lo/part_hi, blocks [tile0, tile0+n_tiles))
	auto launch_demods = [&](BackParams q, int part) -> int {
		if (!h->dcfg.n_demods) return TFR_OK;
		if (screen) {
			// the decimated samples of every window (the screen stored none): exact FIR from the raw bytes
			BackParams w = q;
			w.demod = h->win_demod;
			CU(launch_decwin(w, h->dcfg.filter, sl.d_hist_copy, sl.d_dec, sb));
			CU(cudaEventRecord(sl.raw_done, sb));
			h->stats.kernel_launches += 1;
		}
		// fm_dev on the back stream, part after part (a window's filter warm-up reads the values of earlier parts).  A call
		// that runs as ONE part knows every window by now and computes fm_dev only inside them; parts work from the
		// blocks' descriptors (the demodulators' part cuts differ, so no single window list bounds a part's samples)
		if (h->has_fm) {
			if (q.part_lo < 0 && q.part_hi < 0 && !getenv("TFR_DEVFM_BLOCKS")) {
				q.demod = h->fm_demod;
				CU(launch_devfm_win(q, sb));
				if (q.ld) {
					// the low-pass of the TFA_2-family demodulators ahead of the slicers, in chains of windows, proven
					CU(launch_biq(q, h->dcfg.n_demods, sb));
					h->stats.kernel_launches += 2;
				}
			} else {
				CU(launch_devfm(q, sb));
			}
			h->stats.kernel_launches += 1;
		}
		const bool has_win = h->has_fm || (h->dcfg.d[0].kind == K_TFA1);
		if (has_win) {
			cudaStream_t sw = h->part_stream[part], sl_ = h->part_long[part];
			CU(cudaEventRecord(h->part_fm[part], sb));
			if (h->long_split) {
				// the long window chains (telegrams, retriggered noise), one warp each, beside the rest
				q.long_split = 1;
				CU(cudaStreamWaitEvent(sl_, h->part_fm[part], 0));
				CU(launch_winlong(q, h->dcfg.n_demods, sl_));
				CU(cudaEventRecord(h->part_ldone[part], sl_));
				h->stats.kernel_launches += 1;
			}
			CU(cudaStreamWaitEvent(sw, h->part_fm[part], 0));
			CU(launch_win(q, h->dcfg.n_demods, sw));
			CU(cudaEventRecord(h->part_done[part], sw));
			h->stats.kernel_launches += 1;
		}
		return TFR_OK;
	};
	// everything after the window kernels (WeatherHub walk, verifier, parsers) waits for every part
	auto join_parts = [&](int n) -> int {
		const bool has_win = h->dcfg.n_demods && (h->has_fm || (h->dcfg.d[0].kind == K_TFA1));
		if (!has_win) return TFR_OK;
		for (int k = 0; k < n; k++) {
			CU(cudaStreamWaitEvent(sb, h->part_done[k], 0));
			if (h->long_split) CU(cudaStreamWaitEvent(sb, h->part_ldone[k], 0));
		}
		return TFR_OK;
	};
	int parts_done = 0, part_tile0 = 0;   // back-end parts already queued, first block of the next one

	// ---- front: decimate + trigger, threshold walk
	{   // the pair recorded by this slot's previous call has not been read yet if no tfr_sync came in between
		float a = 0;
		if (cudaEventElapsedTime(&a, sl.fe0, sl.fe1) == cudaSuccess) h->fe_ms_acc += a;
		else cudaGetLastError();
	}
	// The call's blocks go through the front-end in kFrontChunks launches; the threshold walk of chunk k (one warp
	// per stream, a serial chain) runs on the walk stream while the front-end of chunk k+1 streams on.
	CU(cudaEventRecord(sl.fe0, sf));
	{
		// at least 8192 blocks (512 MiB, ~18 waves of CTAs) per launch: smaller launches only add tails
		const int n_chunks = (int)std::min<size_t>(kFrontChunks, std::max<size_t>(1, total / h->min_chunk));
		const int per = (int)((max_blocks + n_chunks - 1) / n_chunks);
		CU(cudaStreamWaitEvent(h->stream_walk, sl.fe0, 0));   // the walk stream starts after everything queued so far
		CU(cudaStreamWaitEvent(h->stream_fe2, sl.fe0, 0));
		int last_odd = -1;
		for (int k = 0; k < n_chunks; k++) {
			fp.tile0 = k * per;
			fp.n_tiles = std::min(per, (int)max_blocks - fp.tile0);
			if (fp.n_tiles <= 0) break;
			// the chunks are independent of each other: alternating two streams lets chunk k+1's first CTAs fill the
			// SMs that chunk k's last wave leaves idle
			cudaStream_t sk = (k & 1) ? h->stream_fe2 : sf;
			if (screen) {
				CU(launch_screen(fp, k, sk));
				h->stats.kernel_launches += 1;
			} else {
				CU(launch_fe(fp, sk));
			}
			CU(cudaEventRecord(h->chunk_ev[k], sk));
			if (k & 1) last_odd = k;
			CU(cudaStreamWaitEvent(h->stream_walk, h->chunk_ev[k], 0));
			bp.n_tiles = fp.n_tiles;
			// Back-end part j is queued as soon as the walk has passed its last chunk: it decodes the windows that are
			// complete by then while the front-end streams on (the window kernels are latency bound and leave the SMs'
			// issue slots and most of their shared memory to it).  The last part follows the whole front (below).
			const int n_parts = (h->pipelined && h->dcfg.n_demods) ? std::min(h->be_parts, n_chunks) : 1;
			const int j = parts_done;
			const bool part_end = (j + 1 < n_parts) && (k + 1 == (j + 1) * n_chunks / n_parts);
			bp.part_idx = part_end ? j : -1;
			// (table items carry positions in 31 bits; a trigger group of the burst scan must be shorter than any timeout)
			const bool use_tab = sl.d_walk_tab && max_blocks < (1u << 18) && h->t_min > 32;
			bp.walk_tab = use_tab ? sl.d_walk_tab : nullptr;
			bp.walk_gap = use_tab ? sl.d_walk_gap : nullptr;
			bp.walk_base = sl.d_walk_base;
			bp.walk_dbg = h->walk_dbg;
			bp.walk_ct = h->walk_ct;
			CU(launch_thresh2(bp, h->stream_walk));
			bp.walk_tab = nullptr;
			bp.walk_gap = nullptr;
			bp.part_idx = -1;
			h->stats.kernel_launches += use_tab ? 3 : 2;
			if (part_end) {
				CU(cudaEventRecord(h->part_ev[j], h->stream_walk));
				CU(cudaStreamWaitEvent(sb, h->part_ev[j], 0));
				BackParams q = bp;
				q.tile0 = part_tile0;
				q.n_tiles = fp.tile0 + fp.n_tiles - part_tile0;
				q.part_lo = j - 1;
				q.part_hi = j;
				rc = launch_demods(q, j);
				if (rc) return rc;
				part_tile0 = fp.tile0 + fp.n_tiles;
				parts_done = j + 1;
			}
		}
		if (last_odd >= 0) CU(cudaStreamWaitEvent(sf, h->chunk_ev[last_odd], 0));
		CU(cudaEventRecord(sl.fe1, sf));
		if (screen && !h->screen_stat_pending) {
			CU(cudaMemcpyAsync(h->h_screen_stat, h->d_screen_stat, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, sf));
			CU(cudaEventRecord(h->ev_screen_stat, sf));
			h->screen_stat_pending = true;
		}
		CU(cudaEventRecord(h->walk_ev, h->stream_walk));
		CU(cudaEventRecord(h->dbg_ev[0], h->stream_walk));
		CU(cudaStreamWaitEvent(sf, h->walk_ev, 0));   // rejoin: everything after this on the front stream sees the walk
		fp.tile0 = 0;
		bp.n_tiles = (int)max_blocks;
	}
	if (auto_mode) {
		CU(cudaMemcpyAsync(h->h_progress, h->d_progress, sizeof(uint32_t) * ns, cudaMemcpyDeviceToHost, sf));
		CU(cudaStreamSynchronize(sf));
		uint32_t remaining = 0;
		for (int s = 0; s < ns; s++)
			if (h->jobs[s].n_blocks) remaining = std::max(remaining, h->jobs[s].n_blocks - std::min(h->h_progress[s], h->jobs[s].n_blocks));
		const int n_fallback = (int)((remaining + kEpochBlocks - 1) / kEpochBlocks);
		if (n_fallback) {
			rc = ensure_events(h, 2);
			if (rc) return rc;
			fp.use_progress = 1;
			fp.n_tiles = kEpochBlocks;
			fp.margin = kEpochMargin;
			bp.n_tiles = kEpochBlocks;
			bp.margin = kEpochMargin;
			CU(cudaEventRecord(h->ev[0], sf));
			for (int e = 0; e < n_fallback; e++) {
				CU(launch_fe(fp, sf));
				CU(launch_thresh2(bp, sf));
				h->stats.kernel_launches += 2;
			}
			CU(cudaEventRecord(h->ev[1], sf));
			CU(cudaStreamSynchronize(sf));   // rare path (start-up transients): timed synchronously, threshold walk included
			float a = 0;
			CU(cudaEventElapsedTime(&a, h->ev[0], h->ev[1]));
			h->fe_ms_acc += a;
			h->stats.fallback_epochs += (uint32_t)n_fallback;
			part_tile0 = 0;   // the redone blocks hold more samples than the parts before saw: the last part's fm_dev pass covers everything
		}
	}
	CU(launch_save_history(sl.d_jobs, h->d_state, ns, decimated ? 1 : 0, sf));
	h->stats.kernel_launches += 1;
	CU(cudaEventRecord(sl.front_done, sf));

	// ---- back: demodulators over the windows, verifier, parsers - once per call, over all blocks
	CU(cudaStreamWaitEvent(sb, sl.front_done, 0));
	CU(cudaEventRecord(h->dbg_ev[1], sb));
	bp.tile0 = 0;
	bp.n_tiles = (int)max_blocks;
	// Split back-end: only the first window of every (stream, demodulator) - true carried state, the sample before position 0 -
	// and the windows whose filter warm-up reaches back to it need what the previous call's verifier leaves (thresh2_kernel
	// lists how many: partcnt row kLateRow).  Everything else of this call - decwin, fm_dev, the other windows - goes to the
	// early stream and runs beside the previous call's verifier and parsers; the back stream, call after call, then only
	// takes the late windows (a warp each), WeatherHub, the verifier and the parsers.
	bool split = screen && h->be_split && h->pipelined && h->long_split && h->dcfg.n_demods && parts_done == 0 && h->be_parts == 1 &&
		     !h->biq_chains;
	if (split && h->be_split == 1) {
		// nothing to run beside when the previous call's back-end is already done (a caller that synchronises per call): the
		// split only adds launches then
		const cudaError_t q = cudaEventQuery(h->slot[si ^ 1].back_done);
		if (q == cudaSuccess) split = false;
		else if (q != cudaErrorNotReady) CU(q);
		else cudaGetLastError();
	}
	if (h->dcfg.n_demods && split) {
		cudaStream_t se = si ? h->stream_early2 : h->stream_early;
		cudaStream_t s_win = h->part_stream[si], s_long = h->part_long[si];
		const bool has_win = h->has_fm || (h->dcfg.d[0].kind == K_TFA1);
		CU(cudaStreamWaitEvent(se, sl.front_done, 0));
		{
			BackParams w = bp;
			w.demod = h->win_demod;
			CU(launch_decwin(w, h->dcfg.filter, sl.d_hist_copy, sl.d_dec, se));
			CU(cudaEventRecord(sl.raw_done, se));
			h->stats.kernel_launches += 1;
		}
		if (h->has_fm) {
			BackParams w = bp;
			w.demod = h->fm_demod;
			CU(launch_devfm_win(w, se));
			h->stats.kernel_launches += 1;
		}
		CU(cudaEventRecord(sl.early_done, se));
		if (has_win) {
			BackParams q = bp;
			q.part_lo = kLateRow;
			q.part_hi = -1;
			q.long_split = 1;
			CU(cudaStreamWaitEvent(s_long, sl.early_done, 0));
			CU(launch_winlong(q, h->dcfg.n_demods, s_long));
			CU(cudaEventRecord(h->part_ldone[si], s_long));
			CU(cudaStreamWaitEvent(s_win, sl.early_done, 0));
			CU(launch_win(q, h->dcfg.n_demods, s_win));
			CU(cudaEventRecord(h->part_done[si], s_win));
			h->stats.kernel_launches += 2;
		}
		// late: after the previous call's back-end (stream order)
		CU(cudaStreamWaitEvent(sb, sl.early_done, 0));
		if (h->has_fm) {
			BackParams w = bp;
			w.demod = h->fm_demod;
			CU(launch_devfm_first(w, sb));
			h->stats.kernel_launches += 1;
		}
		if (has_win) {
			BackParams q = bp;
			q.part_lo = -1;
			q.part_hi = kLateRow;
			q.long_split = 1;
			q.long_all = 1;
			CU(launch_winlong(q, h->dcfg.n_demods, sb));
			h->stats.kernel_launches += 1;
			CU(cudaStreamWaitEvent(sb, h->part_done[si], 0));
			CU(cudaStreamWaitEvent(sb, h->part_ldone[si], 0));
		}
		if (h->has_whb) { CU(launch_walk(bp, h->dcfg.n_demods, sb)); h->stats.kernel_launches += 1; }
		CU(launch_verify(bp, h->dcfg.n_demods, sb));
		h->stats.kernel_launches += 1;
	} else if (h->dcfg.n_demods) {
		{   // the last part: every window the parts before it did not take
			// (a call that runs as one part knows all its windows now: filter chains, window-driven fm_dev)
			if (parts_done == 0 && h->has_fm && h->biq_chains && sl.d_ld && !getenv("TFR_DEVFM_BLOCKS")) bp.ld = sl.d_ld;
			BackParams q = bp;
			q.tile0 = part_tile0;
			q.n_tiles = (int)max_blocks - part_tile0;
			q.part_lo = parts_done - 1;
			q.part_hi = -1;
			rc = launch_demods(q, parts_done);
			if (rc) return rc;
			rc = join_parts(parts_done + 1);
			if (rc) return rc;
		}
		if (h->has_whb) { CU(launch_walk(bp, h->dcfg.n_demods, sb)); h->stats.kernel_launches += 1; }
		CU(launch_verify(bp, h->dcfg.n_demods, sb));
		h->stats.kernel_launches += 1;
	}
	CU(launch_submit_epilogue(bp, sb));
	CU(launch_parse(bp, sb));
	h->stats.kernel_launches += 2;
	CU(cudaEventRecord(sl.back_done, sb));
	CU(cudaEventRecord(h->dbg_ev[2], sb));
	CU(cudaEventRecord(h->span1, sb));
	h->stats.blocks += total;
	h->stats.raw_samples += total * (uint64_t)kBlockRaw;
	h->results_valid = false;
	return TFR_OK;
}

extern "C" __attribute__((visibility("default"))) int tfr_sync(tfr_handle *h)
{
	if (!h) return fail(TFR_E_INVAL, "tfr_sync: null handle");
	CU(cudaSetDevice(h->device));
	int rc = sync_all(h);
	if (rc) return rc;
	if (h->span_open) {
		// times cover every tfr_process since the previous tfr_sync (they overlap on the device when pipelined)
		double fe = h->fe_ms_acc;
		for (auto &sl : h->slot) {
			float a = 0;
			if (cudaEventElapsedTime(&a, sl.fe0, sl.fe1) == cudaSuccess) fe += a;
			else cudaGetLastError();
			// re-record so that a pair is never counted twice
			CU(cudaEventRecord(sl.fe0, h->stream));
			CU(cudaEventRecord(sl.fe1, h->stream));
		}
		CU(cudaStreamSynchronize(h->stream));
		h->fe_ms_acc = 0;
		float all = 0, tot = 0;
		CU(cudaEventElapsedTime(&all, h->span0, h->span1));
		CU(cudaEventElapsedTime(&tot, h->h2d_timed ? h->ev_h2d0 : h->span0, h->span1));
		h->stats.last_frontend_ms = fe;
		h->stats.last_backend_ms = all - fe;
		h->stats.last_total_ms = tot;
		h->span_open = false;
		if (getenv("TFR_DEBUG")) {
			float a = 0, b = 0, c = 0, d = 0;
			cudaEventElapsedTime(&a, h->span0, h->slot[h->cur].fe1);
			cudaEventElapsedTime(&b, h->span0, h->dbg_ev[0]);
			cudaEventElapsedTime(&c, h->span0, h->dbg_ev[1]);
			cudaEventElapsedTime(&d, h->span0, h->dbg_ev[2]);
			cudaGetLastError();
			fprintf(stderr, "[tfr] timeline of the last call (ms from its start): front-end done %.3f, walk done %.3f, back-end starts %.3f, ends %.3f\n", a, b, c, d);
		}
	}
	if (h->h2d_timed) {
		float a = 0;
		CU(cudaEventElapsedTime(&a, h->ev_h2d0, h->ev_h2d1));
		h->stats.last_h2d_ms = a;
		h->h2d_timed = false;
	} else {
		h->stats.last_h2d_ms = 0;
	}
	return TFR_OK;
}

// (int) of a double the way the reference's x86 cvttsd2si does it (tfa1.cpp:180 feeds it log10(0) = -inf)
static int d2i_x86(double v)
{
	if (!(v > -2147483649.0 && v < 2147483648.0)) return (int)0x80000000;
	return (int)v;
}

static int fetch_results(tfr_handle *h)
{
	if (h->results_valid) return TFR_OK;
	int rc = tfr_sync(h);
	if (rc) return rc;
	Counters c;
	CU(cudaMemcpy(&c, h->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
	const uint32_t nf = std::min(c.n_frames, h->max_frames);
	const uint32_t nr = std::min(c.n_records, h->max_records);
	std::vector<DevFrame> df(nf);
	std::vector<DevRecord> dr(nr);
	if (nf) CU(cudaMemcpy(df.data(), h->d_frames, sizeof(DevFrame) * nf, cudaMemcpyDeviceToHost));
	if (nr) CU(cudaMemcpy(dr.data(), h->d_records, sizeof(DevRecord) * nr, cudaMemcpyDeviceToHost));
	// reference output order: time, then registration order (SURVEY.md §8b "Threading")
	// a slot the verifier retired (a re-run with the true carry-in no longer passed the byte_cnt gate, status -2)
	// is not a flush the reference performs: it never reaches the caller
	std::vector<uint32_t> order;
	order.reserve(nf);
	for (uint32_t k = 0; k < nf; k++)
		if (df[k].status != -2) order.push_back(k);
	std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
		const DevFrame &x = df[a], &y = df[b];
		if (x.stream != y.stream) return x.stream < y.stream;
		if (x.pos != y.pos) return x.pos < y.pos;
		if (x.demod != y.demod) return x.demod < y.demod;
		return (x.status == 3) > (y.status == 3);   // a window's "Inverted SYNC" notice comes before its frame
	});
	h->frames.clear();
	h->records.clear();
	const int64_t now = (int64_t)time(nullptr);
	for (uint32_t k : order) {
		const DevFrame &x = df[k];
		tfr_frame f;
		memset(&f, 0, sizeof(f));
		f.stream = x.stream;
		f.type = x.type;
		f.status = x.status;
		f.byte_cnt = x.byte_cnt;
		f.pos = x.pos;
		f.offset = x.offset;
		f.rssi_raw = x.rssi_raw;
		// the dB conversion happens here with the host libm: tfa1.cpp:180, tfa2.cpp:434, whb.cpp:696 (as built: rssi*0.00025)
		if (x.type == TFR_TFA_WHB) f.rssi = d2i_x86(10 * log10(1 + x.rssi_raw * 0.00025));
		else f.rssi = d2i_x86(10 * log10(x.rssi_raw));
		f.n_records = 0;
		f.first_record = (int32_t)h->records.size();
		memcpy(f.rdata, x.rdata, TFR_MAX_RDATA);
		const int fidx = (int)h->frames.size();
		for (int r = 0; r < x.n_records; r++) {
			const uint32_t ri = (uint32_t)x.first_record + r;
			if (ri >= nr) break;
			const DevRecord &y = dr[ri];
			tfr_record o;
			memset(&o, 0, sizeof(o));
			o.stream = y.stream;
			o.type = y.type;
			o.id = y.id;
			o.temp = y.temp;
			o.humidity = y.humidity;
			o.alarm = y.alarm;
			o.flags = y.flags;
			o.sequence = y.sequence;
			o.rssi = f.rssi;
			o.ts = now;
			o.pos = y.pos;
			o.frame = fidx;
			h->records.push_back(o);
			f.n_records++;
		}
		h->frames.push_back(f);
	}
	h->stats.frames = h->frames.size();
	h->stats.records = h->records.size();
	h->stats.active_samples = c.active_samples;
	h->stats.windows = c.n_windows;
	h->stats.reruns = c.n_reruns;
	h->stats.reruns_sr = c.rerun_sr;
	h->stats.reruns_biquad = c.par_cheap;
	if (getenv("TFR_DEBUG")) fprintf(stderr, "[tfr] windows %llu par_cheap %u edge_par %u | verify: checked %u cheap %u full %u sr %u\n", c.n_windows, c.par_cheap, c.rerun_edge, c.ver_checked, c.ver_cheap, c.ver_full, c.rerun_sr);
	h->stats.reruns_edge = c.rerun_edge;
	h->results_valid = true;
	if (c.overflow && !h->overflow_reported) {
		h->overflow_reported = true;
		return fail(TFR_E_OVERFLOW, "a frame/record/window buffer overflowed: results are truncated; raise tfr_config.max_frames "
					    "(reported once; polling again returns the truncated results)");
	}
	return TFR_OK;
}

extern "C" __attribute__((visibility("default"))) long tfr_poll_frames(tfr_handle *h, tfr_frame *out, size_t cap)
{
	if (!h) return fail(TFR_E_INVAL, "tfr_poll_frames: null handle");
	int rc = fetch_results(h);
	if (rc) return rc;
	if (!out) return (long)h->frames.size();
	const size_t n = std::min(cap, h->frames.size());
	if (n) memcpy(out, h->frames.data(), n * sizeof(tfr_frame));
	return (long)n;
}

extern "C" __attribute__((visibility("default"))) long tfr_poll_records(tfr_handle *h, tfr_record *out, size_t cap)
{
	if (!h) return fail(TFR_E_INVAL, "tfr_poll_records: null handle");
	int rc = fetch_results(h);
	if (rc) return rc;
	if (!out) return (long)h->records.size();
	const size_t n = std::min(cap, h->records.size());
	if (n) memcpy(out, h->records.data(), n * sizeof(tfr_record));
	return (long)n;
}

extern "C" __attribute__((visibility("default"))) int tfr_clear_results(tfr_handle *h)
{
	if (!h) return fail(TFR_E_INVAL, "tfr_clear_results: null handle");
	CU(cudaSetDevice(h->device));
	{ int rc_ = sync_all(h); if (rc_) return rc_; }
	// keep active_samples running; zero the frame/record cursors
	Counters c;
	CU(cudaMemcpy(&c, h->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
	c.n_frames = c.n_records = c.overflow = 0;
	c.n_reruns = c.rerun_sr = c.rerun_biquad = c.rerun_edge = 0;
	c.par_cheap = c.ver_checked = c.ver_cheap = c.ver_full = 0;
	c.n_windows = 0;
	CU(cudaMemcpy(h->d_counters, &c, sizeof(c), cudaMemcpyHostToDevice));
	if (h->d_tap_cnt) CU(cudaMemset(h->d_tap_cnt, 0, (size_t)h->cfg.n_streams * kMaxDemods * 3 * sizeof(uint32_t)));
	h->frames.clear();
	h->records.clear();
	h->results_valid = false;
	h->overflow_reported = false;
	return TFR_OK;
}

extern "C" __attribute__((visibility("default"))) int tfr_get_thresh(tfr_handle *h, int stream, int32_t *thresh)
{
	if (!h || !thresh || stream < 0 || stream >= h->cfg.n_streams) return fail(TFR_E_INVAL, "tfr_get_thresh: bad argument");
	int rc = tfr_sync(h);
	if (rc) return rc;
	CU(cudaMemcpy(thresh, (const char *)(h->d_state + stream) + offsetof(StreamState, thresh), sizeof(int32_t), cudaMemcpyDeviceToHost));
	return TFR_OK;
}

extern "C" __attribute__((visibility("default"))) long tfr_read_block_trace(tfr_handle *h, int stream, tfr_block_trace *out, size_t cap)
{
	if (!h || stream < 0 || stream >= h->cfg.n_streams) return fail(TFR_E_INVAL, "tfr_read_block_trace: bad argument");
	int rc = tfr_sync(h);
	if (rc) return rc;
	if (h->jobs.empty()) return 0;
	const StreamJob &j = h->jobs[stream];
	if (!out) return (long)j.n_blocks;
	const size_t n = std::min<size_t>(cap, j.n_blocks);
	static_assert(sizeof(tfr_block_trace) == sizeof(BlockTrace), "trace layout");
	if (n) CU(cudaMemcpy(out, h->slot[h->cur].d_trace + j.dec_off, n * sizeof(BlockTrace), cudaMemcpyDeviceToHost));
	return (long)n;
}

extern "C" __attribute__((visibility("default"))) long tfr_read_taps(tfr_handle *h, int stream, int demod, int kind, void *out, size_t cap_elems)
{
	if (!h || stream < 0 || stream >= h->cfg.n_streams || demod < 0 || demod >= h->dcfg.n_demods || kind < 0 || kind > 2)
		return fail(TFR_E_INVAL, "tfr_read_taps: bad argument");
	if (!h->tap_cap) return fail(TFR_E_INVAL, "tfr_read_taps: handle was created without TFR_FLAG_TAPS");
	int rc = tfr_sync(h);
	if (rc) return rc;
	uint32_t cnt = 0;
	const size_t slot = (size_t)stream * kMaxDemods + demod;
	CU(cudaMemcpy(&cnt, h->d_tap_cnt + slot * 3 + kind, sizeof(cnt), cudaMemcpyDeviceToHost));
	if (cnt > h->tap_cap) return fail(TFR_E_OVERFLOW, "tap buffer overflowed");
	if (!out) return (long)cnt;
	const size_t n = std::min<size_t>(cap_elems, cnt);
	if (n) {
		if (kind == 2) CU(cudaMemcpy(out, h->d_tap_f64 + slot * h->tap_cap, n * sizeof(double), cudaMemcpyDeviceToHost));
		else CU(cudaMemcpy(out, h->d_tap_i32[kind] + slot * h->tap_cap, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
	}
	return (long)n;
}

extern "C" __attribute__((visibility("default"))) long tfr_read_decimated(tfr_handle *h, int stream, int16_t *out, size_t cap_int16)
{
	if (!h || stream < 0 || stream >= h->cfg.n_streams) return fail(TFR_E_INVAL, "tfr_read_decimated: bad argument");
	if (!(h->cfg.flags & TFR_FLAG_KEEP_DECIM)) return fail(TFR_E_INVAL, "tfr_read_decimated: handle was created without TFR_FLAG_KEEP_DECIM");
	int rc = tfr_sync(h);
	if (rc) return rc;
	if (h->jobs.empty()) return 0;
	const StreamJob &j = h->jobs[stream];
	const size_t avail = (size_t)j.n_blocks * kBlockDec * 2;
	if (!out) return (long)avail;
	const size_t n = std::min(cap_int16, avail) & ~(size_t)1;
	if (n) CU(cudaMemcpy(out, h->slot[h->cur].d_dec + (size_t)j.dec_off * kBlockDec, n * sizeof(int16_t), cudaMemcpyDeviceToHost));
	return (long)n;
}

extern "C" __attribute__((visibility("default"))) long tfr_read_screen(tfr_handle *h, int stream, int32_t *out, size_t cap_int32, int *shift, int *slack)
{
	if (!h || stream < 0 || stream >= h->cfg.n_streams) return fail(TFR_E_INVAL, "tfr_read_screen: bad argument");
	if (!(h->cfg.flags & TFR_FLAG_TAPS) || !h->use_screen) return fail(TFR_E_INVAL, "tfr_read_screen: needs TFR_FLAG_TAPS and the screening front-end");
	int rc = tfr_sync(h);
	if (rc) return rc;
	if (shift) *shift = h->screen_shift;
	if (slack) *slack = h->screen_slack;
	if (h->jobs.empty() || !h->d_screen_dbg) return 0;
	const StreamJob &j = h->jobs[stream];
	const size_t avail = (size_t)j.n_blocks * kBlockDec * 2;
	if (!out) return (long)avail;
	const size_t n = std::min(cap_int32, avail) & ~(size_t)1;
	if (n) CU(cudaMemcpy(out, h->d_screen_dbg + (size_t)j.dec_off * kBlockDec * 2, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
	return (long)n;
}

extern "C" __attribute__((visibility("default"))) long tfr_decimate(int device, const uint8_t *iq, size_t nbytes, int filter, int16_t *out, int mem)
{
	if (!iq || !out || nbytes < 4) return fail(TFR_E_INVAL, "tfr_decimate: bad argument");
	tfr_config c;
	memset(&c, 0, sizeof(c));
	c.struct_size = sizeof(c);
	c.device = device;
	c.types = 0;
	c.filter = filter;
	c.thresh = 32767;
	c.n_streams = 1;
	c.flags = TFR_FLAG_KEEP_DECIM;
	tfr_handle *h = nullptr;
	int rc = tfr_create(&c, &h);
	if (rc) return rc;
	// pad to whole blocks with zero-signal bytes; only the first nbytes/4 outputs are returned
	const size_t padded = (nbytes + TFR_BLOCK_BYTES - 1) / TFR_BLOCK_BYTES * TFR_BLOCK_BYTES;
	uint8_t *d_in = nullptr;
	long ret = 0;
	auto cuda_fail = [&](const char *what, cudaError_t e) { return (long)fail(TFR_E_CUDA, std::string("tfr_decimate: ") + what + ": " + cudaGetErrorString(e)); };
	do {
		cudaError_t e;
		if ((e = cudaMalloc(&d_in, padded)) != cudaSuccess) { ret = fail(TFR_E_NOMEM, std::string("tfr_decimate: input buffer: ") + cudaGetErrorString(e)); break; }
		if ((e = cudaMemset(d_in, 128, padded)) != cudaSuccess) { ret = cuda_fail("memset", e); break; }
		if ((e = cudaMemcpy(d_in, iq, nbytes, mem == TFR_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice)) != cudaSuccess) { ret = cuda_fail("copy in", e); break; }
		if ((rc = tfr_submit(h, 0, d_in, padded, TFR_MEM_DEVICE)) || (rc = tfr_process(h)) || (rc = tfr_sync(h))) { ret = rc; break; }
		const size_t n_out = (nbytes / 8) * 2;   // nbytes/2 raw samples -> /4 decimated -> 2 int16 each
		if ((e = cudaMemcpy(out, h->slot[h->cur].d_dec, n_out * sizeof(int16_t), mem == TFR_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost)) != cudaSuccess) { ret = cuda_fail("copy out", e); break; }
		ret = (long)n_out;
	} while (0);
	cudaFree(d_in);
	std::string keep = g_err;
	tfr_destroy(h);
	g_err = keep;
	return ret;
}

// downconvert(passes)::process_iq over a whole buffer (dsp_stuff.cpp:232-264): the decimation sweep entry point
extern "C" __attribute__((visibility("default"))) long tfr_downconvert(int device, const uint8_t *iq, size_t nbytes, int passes, int filter,
								       int16_t *out, int mem, int reps, float *kernel_ms)
{
	if (!iq || !out || nbytes < 4) return fail(TFR_E_INVAL, "tfr_downconvert: bad argument");
	if (passes < 1 || passes > 8) return fail(TFR_E_INVAL, "tfr_downconvert: passes must be 1..8");
	if (mem != TFR_MEM_HOST && mem != TFR_MEM_DEVICE) return fail(TFR_E_INVAL, "tfr_downconvert: mem must be TFR_MEM_HOST or TFR_MEM_DEVICE");
	if (mem == TFR_MEM_DEVICE && (((uintptr_t)iq & 15) || ((uintptr_t)out & 15))) return fail(TFR_E_INVAL, "tfr_downconvert: device pointers must be 16-byte aligned");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return fail(TFR_E_NODEVICE, "tfr_downconvert: no such CUDA device (there is no CPU fallback)"); }
	cudaError_t e = cudaSetDevice(device);
	if (e != cudaSuccess) return fail(TFR_E_CUDA, std::string("tfr_downconvert: ") + cudaGetErrorString(e));
	const long long n_pairs = (long long)(nbytes / 2);
	const long long n_res = n_pairs >> passes;            // every stage drops an odd trailing sample
	uint8_t *d_in = nullptr;
	int16_t *tmp[2] = { nullptr, nullptr }, *res = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	long ret = 0;
	auto cuda_fail = [&](const char *what, cudaError_t err) { cudaGetLastError(); return (long)fail(err == cudaErrorMemoryAllocation ? TFR_E_NOMEM : TFR_E_CUDA, std::string("tfr_downconvert: ") + what + ": " + cudaGetErrorString(err)); };
	do {
		const uint8_t *src = iq;
		if (mem == TFR_MEM_HOST) {
			if ((e = cudaMalloc(&d_in, nbytes)) != cudaSuccess) { ret = cuda_fail("input buffer", e); break; }
			if ((e = cudaMemcpy(d_in, iq, nbytes, cudaMemcpyHostToDevice)) != cudaSuccess) { ret = cuda_fail("copy in", e); break; }
			src = d_in;
		}
		if ((e = cudaMalloc(&tmp[0], std::max<size_t>((size_t)(n_pairs / 2) * 4, 16))) != cudaSuccess) { ret = cuda_fail("stage buffer", e); break; }
		if (passes > 1 && (e = cudaMalloc(&tmp[1], std::max<size_t>((size_t)(n_pairs / 4) * 4, 16))) != cudaSuccess) { ret = cuda_fail("stage buffer", e); break; }
		if ((e = cudaEventCreate(&ev0)) != cudaSuccess || (e = cudaEventCreate(&ev1)) != cudaSuccess) { ret = cuda_fail("events", e); break; }
		// reps > 1: timing runs (the sweep tool); the result is the same every time
		const int n_rep = reps < 1 ? 1 : reps;
		// passes 1..5 run as ONE kernel (decim_fused.cu); TFR_DC=cascade selects the launch-per-stage cascade (decim.cu),
		// which is also what passes 6..8 use
		const char *dcsel = getenv("TFR_DC");
		const bool fused = passes <= 5 && !(dcsel && !strcmp(dcsel, "cascade"));
		auto run = [&]() -> cudaError_t {
			if (fused) {
				res = tmp[0];
				return launch_downconvert_fused(src, nullptr, n_pairs, passes, filter, tmp[0], 0);
			}
			return launch_downconvert(src, n_pairs, passes, filter, tmp[0], tmp[1], &res, 0, false);
		};
		if ((e = run()) != cudaSuccess) { ret = cuda_fail("launch", e); break; }
		cudaEventRecord(ev0, 0);
		for (int r = 0; r < n_rep && e == cudaSuccess; r++) e = run();
		cudaEventRecord(ev1, 0);
		if (e != cudaSuccess) { ret = cuda_fail("launch", e); break; }
		if ((e = cudaEventSynchronize(ev1)) != cudaSuccess) { ret = cuda_fail("kernels", e); break; }
		if (kernel_ms) {
			float ms = 0;
			cudaEventElapsedTime(&ms, ev0, ev1);
			*kernel_ms = ms / (float)n_rep;
		}
		if (n_res > 0 && (e = cudaMemcpy(out, res, (size_t)n_res * 4, mem == TFR_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost)) != cudaSuccess) { ret = cuda_fail("copy out", e); break; }
		ret = (long)(n_res * 2);
	} while (0);
	if (ev0) cudaEventDestroy(ev0);
	if (ev1) cudaEventDestroy(ev1);
	cudaFree(d_in); cudaFree(tmp[0]); cudaFree(tmp[1]);
	return ret;
}


// ------------------------------------------------------------------------------------------------
// downconvert as a streaming object (dsp_stuff.h:46-56): history carried from call to call
// ------------------------------------------------------------------------------------------------
struct tfr_dc {
	int device = 0, passes = 2;
	uint8_t *d_hist[2] = { nullptr, nullptr };   // last 384 raw samples (768 B), double buffered
	int hist_cur = -1;                            // -1: nothing fed yet (zero histories)
	int16_t *d_hist16 = nullptr;                  // int16 entry: last 384 I,Q pairs
	bool hist16_valid = false;
	uint8_t *d_in = nullptr;
	int16_t *d_tmp[2] = { nullptr, nullptr };
	size_t cap_in = 0, cap_tmp = 0;
};
static constexpr int kDcHistPairs = 384;

extern "C" __attribute__((visibility("default"))) int tfr_dc_create(int device, int passes, tfr_dc **out)
{
	if (!out) return fail(TFR_E_INVAL, "tfr_dc_create: null argument");
	*out = nullptr;
	if (passes < 1 || passes > 5) return fail(TFR_E_INVAL, "tfr_dc_create: passes must be 1..5");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return fail(TFR_E_NODEVICE, "tfr_dc_create: no such CUDA device (there is no CPU fallback)"); }
	CU(cudaSetDevice(device));
	tfr_dc *h = new tfr_dc();
	h->device = device;
	h->passes = passes;
	cudaError_t e = cudaMalloc(&h->d_hist[0], 2 * kDcHistPairs);
	if (e == cudaSuccess) e = cudaMalloc(&h->d_hist[1], 2 * kDcHistPairs);
	if (e == cudaSuccess) e = cudaMalloc(&h->d_hist16, 4 * kDcHistPairs);
	if (e != cudaSuccess) { cudaGetLastError(); cudaFree(h->d_hist[0]); cudaFree(h->d_hist[1]); cudaFree(h->d_hist16); delete h; return fail(TFR_E_NOMEM, "tfr_dc_create: history buffers"); }
	*out = h;
	return TFR_OK;
}

extern "C" __attribute__((visibility("default"))) void tfr_dc_destroy(tfr_dc *h)
{
	if (!h) return;
	cudaSetDevice(h->device);
	cudaDeviceSynchronize();
	cudaFree(h->d_hist[0]); cudaFree(h->d_hist[1]); cudaFree(h->d_hist16);
	cudaFree(h->d_in); cudaFree(h->d_tmp[0]); cudaFree(h->d_tmp[1]);
	delete h;
}

static int dc_reserve(tfr_dc *h, size_t in_bytes, size_t tmp_bytes)
{
	if (in_bytes > h->cap_in) {
		cudaFree(h->d_in);
		h->d_in = nullptr;
		h->cap_in = 0;
		if (cudaMalloc(&h->d_in, in_bytes) != cudaSuccess) { cudaGetLastError(); return fail(TFR_E_NOMEM, "tfr_dc: input buffer"); }
		h->cap_in = in_bytes;
	}
	if (tmp_bytes > h->cap_tmp) {
		cudaFree(h->d_tmp[0]); cudaFree(h->d_tmp[1]);
		h->d_tmp[0] = h->d_tmp[1] = nullptr;
		h->cap_tmp = 0;
		if (cudaMalloc(&h->d_tmp[0], tmp_bytes) != cudaSuccess || cudaMalloc(&h->d_tmp[1], tmp_bytes) != cudaSuccess) { cudaGetLastError(); return fail(TFR_E_NOMEM, "tfr_dc: stage buffers"); }
		h->cap_tmp = tmp_bytes;
	}
	return TFR_OK;
}

// raw u8 IQ in, int16 I,Q out ((nbytes/2) >> passes pairs); nbytes/2 must be a multiple of 2^passes so that every
// stage's phase carries over to the next call (the reference drops an odd trailing sample of a block for good)
extern "C" __attribute__((visibility("default"))) long tfr_dc_process(tfr_dc *h, const uint8_t *iq, size_t nbytes, int filter, int16_t *out, int mem)
{
	if (!h || !iq || !out) return fail(TFR_E_INVAL, "tfr_dc_process: null argument");
	if (mem != TFR_MEM_HOST && mem != TFR_MEM_DEVICE) return fail(TFR_E_INVAL, "tfr_dc_process: mem must be TFR_MEM_HOST or TFR_MEM_DEVICE");
	const long long n_pairs = (long long)(nbytes / 2);
	if (nbytes == 0 || (nbytes & 1) || (n_pairs & ((1ll << h->passes) - 1))) return fail(TFR_E_INVAL, "tfr_dc_process: the number of IQ pairs must be a positive multiple of 2^passes");
	if (mem == TFR_MEM_DEVICE && (((uintptr_t)iq & 3) || ((uintptr_t)out & 15))) return fail(TFR_E_INVAL, "tfr_dc_process: device pointers must be aligned (input 4, output 16 bytes)");
	CU(cudaSetDevice(h->device));
	const long long n_res = n_pairs >> h->passes;
	const uint8_t *src = iq;
	int16_t *dst = out;
	if (mem == TFR_MEM_HOST) {
		int rc = dc_reserve(h, nbytes, (size_t)n_res * 4);
		if (rc) return rc;
		CU(cudaMemcpy(h->d_in, iq, nbytes, cudaMemcpyHostToDevice));
		src = h->d_in;
		dst = h->d_tmp[0];
	}
	const uint8_t *hist = h->hist_cur >= 0 ? h->d_hist[h->hist_cur] : nullptr;
	const int nxt = h->hist_cur >= 0 ? h->hist_cur ^ 1 : 0;
	CU(launch_downconvert_fused(src, hist, n_pairs, h->passes, filter, dst, 0));
	CU(launch_dc_hist(src, n_pairs, hist, h->d_hist[nxt], 0));
	h->hist_cur = nxt;
	if (mem == TFR_MEM_HOST) CU(cudaMemcpy(out, dst, (size_t)n_res * 4, cudaMemcpyDeviceToHost));
	else CU(cudaStreamSynchronize(0));
	return (long)(n_res * 2);
}

// downconvert::process_iq(int16_t *buf, int len, int filter) itself (dsp_stuff.cpp:243-264): int16 I,Q in place on the
// host, len = number of int16 (2 per pair), returns the new len; the pair count must be a multiple of 2^passes.
// The int16 path runs the stage cascade (decim.cu) over [the last 384 pairs of the previous call | this call's pairs]
// and returns the outputs that belong to this call (the 384 pairs in front absorb the cascade's zero start).
extern "C" __attribute__((visibility("default"))) long tfr_dc_process_i16(tfr_dc *h, int16_t *data_iq, int len, int filter)
{
	if (!h || !data_iq || len <= 0) return fail(TFR_E_INVAL, "tfr_dc_process_i16: bad argument");
	const long long n_pairs = len / 2;
	if ((len & 1) || (n_pairs & ((1ll << h->passes) - 1))) return fail(TFR_E_INVAL, "tfr_dc_process_i16: the number of IQ pairs must be a positive multiple of 2^passes");
	CU(cudaSetDevice(h->device));
	const long long tot = n_pairs + kDcHistPairs;
	int rc = dc_reserve(h, (size_t)tot * 4, (size_t)(tot / 2 + 8) * 4);
	if (rc) return rc;
	int16_t *d_all = reinterpret_cast<int16_t *>(h->d_in);
	if (h->hist16_valid) CU(cudaMemcpy(d_all, h->d_hist16, 4 * kDcHistPairs, cudaMemcpyDeviceToDevice));
	else CU(cudaMemset(d_all, 0, 4 * kDcHistPairs));
	CU(cudaMemcpy(d_all + 2 * kDcHistPairs, data_iq, (size_t)n_pairs * 4, cudaMemcpyHostToDevice));
	int16_t *res = nullptr;
	CU(launch_downconvert(d_all, tot, h->passes, filter, h->d_tmp[0], h->d_tmp[1], &res, 0, true));
	// the history for the next call: the last 384 pairs fed so far
	CU(cudaMemcpy(h->d_hist16, d_all + 2 * (tot - kDcHistPairs), 4 * kDcHistPairs, cudaMemcpyDeviceToDevice));
	h->hist16_valid = true;
	const long long n_res = n_pairs >> h->passes;
	CU(cudaMemcpy(data_iq, res + 2 * (kDcHistPairs >> h->passes), (size_t)n_res * 4, cudaMemcpyDeviceToHost));
	return (long)(n_res * 2);
}

extern "C" __attribute__((visibility("default"))) int tfr_parse_bytes(tfr_handle *h, int type, const uint8_t *bytes, int len, tfr_frame *frame,
			       tfr_record *recs, int max_recs)
{
	if (!h || !bytes || len < 0) return fail(TFR_E_INVAL, "tfr_parse_bytes: bad argument");
	int demod = -1;
	for (int k = 0; k < h->dcfg.n_demods; k++)
		if (h->dcfg.d[k].type == type) demod = k;
	if (demod < 0) return fail(TFR_E_INVAL, "tfr_parse_bytes: sensor type not registered in this handle");
	CU(cudaSetDevice(h->device));
	// decoder::store_bytes (decoder.cpp:35-40) then the length gate of the type's flush()
	const int kind = h->dcfg.d[demod].kind;
	const int bc = std::min(len, 256);
	bool gate;
	if (kind == K_TFA1) gate = bc >= 10;
	else if (kind == K_TX22) gate = bc >= 7 && bc < 64;
	else if (kind == K_WHB) gate = !(bc < 11 || bc > 60);
	else gate = bc >= 7;
	if (!gate) {
		// flush() does nothing below its length gate: no frame, no records (frame->status -1 says so; 0 is not an error)
		if (frame) {
			memset(frame, 0, sizeof(*frame));
			frame->type = type;
			frame->status = -1;
			frame->byte_cnt = bc;
			frame->pos = -1;
		}
		return 0;
	}
	DevFrame f;
	memset(&f, 0, sizeof(f));
	f.stream = 0;
	f.demod = demod;
	f.type = type;
	f.status = -1;
	f.byte_cnt = bc;
	f.pos = -1;
	f.rssi_raw = 1.0;   // flush(0): 10*log10(1) == 0
	memcpy(f.rdata, bytes, std::min(bc, (int)kMaxRdata));
	DevFrame *d_f = nullptr;
	DevRecord *d_r = nullptr;
	Counters *d_c = nullptr;
	Counters c;
	memset(&c, 0, sizeof(c));
	c.n_frames = 1;
	DevRecord r[8];
	{
		cudaError_t e = cudaMalloc(&d_f, sizeof(DevFrame));
		if (e == cudaSuccess) e = cudaMalloc(&d_r, sizeof(DevRecord) * 8);
		if (e == cudaSuccess) e = cudaMalloc(&d_c, sizeof(Counters));
		if (e == cudaSuccess) e = cudaMemcpy(d_f, &f, sizeof(f), cudaMemcpyHostToDevice);
		if (e == cudaSuccess) e = cudaMemcpy(d_c, &c, sizeof(c), cudaMemcpyHostToDevice);
		if (e == cudaSuccess) {
			BackParams bp;
			memset(&bp, 0, sizeof(bp));
			bp.cfg = h->d_cfg;
			bp.frames = d_f;
			bp.records = d_r;
			bp.counters = d_c;
			bp.max_frames = 1;
			bp.max_records = 8;
			e = launch_parse(bp, h->stream);
			h->stats.kernel_launches += 1;
		}
		if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
		if (e == cudaSuccess) e = cudaMemcpy(&f, d_f, sizeof(f), cudaMemcpyDeviceToHost);
		if (e == cudaSuccess) e = cudaMemcpy(r, d_r, sizeof(r), cudaMemcpyDeviceToHost);
		cudaFree(d_f); cudaFree(d_r); cudaFree(d_c);   // on every path
		if (e != cudaSuccess) { cudaGetLastError(); return fail(TFR_E_CUDA, std::string("tfr_parse_bytes: ") + cudaGetErrorString(e)); }
	}
	if (frame) {
		memset(frame, 0, sizeof(*frame));
		frame->type = type;
		frame->status = f.status;
		frame->byte_cnt = bc;
		frame->pos = -1;
		frame->rssi = 0;
		frame->rssi_raw = 1.0;
		frame->n_records = f.n_records;
		memcpy(frame->rdata, f.rdata, TFR_MAX_RDATA);
	}
	for (int k = 0; k < f.n_records && k < max_recs; k++) {
		const DevRecord &y = r[f.first_record + k];
		tfr_record &o = recs[k];
		memset(&o, 0, sizeof(o));
		o.type = y.type;
		o.id = y.id;
		o.temp = y.temp;
		o.humidity = y.humidity;
		o.alarm = y.alarm;
		o.flags = y.flags;
		o.sequence = y.sequence;
		o.rssi = 0;
		o.ts = (int64_t)time(nullptr);
		o.pos = -1;
	}
	return f.n_records;
}

extern "C" __attribute__((visibility("default"))) int tfr_get_stats(tfr_handle *h, tfr_stats *out)
{
	if (!h || !out) return fail(TFR_E_INVAL, "tfr_get_stats: null argument");
	if (h->d_screen_stat) {
		uint32_t v[4] = { 0, 0, 0, 0 };
		CU(cudaSetDevice(h->device));
		int rc = tfr_sync(h);
		if (rc) return rc;
		CU(cudaMemcpy(v, h->d_screen_stat, sizeof(v), cudaMemcpyDeviceToHost));
		h->stats.screen_blocks = v[0];
		h->stats.dense_blocks = v[1];
		h->stats.screen_candidates = v[2];
		h->stats.screen_triggers = v[3];
	}
	*out = h->stats;
	return TFR_OK;
}
