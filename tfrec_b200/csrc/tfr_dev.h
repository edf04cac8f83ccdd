// tfr_dev.h - device-side data layout shared by the front-end, the back-end and the C-ABI host code.
//
// Vocabulary follows the reference: a *stream* is one stick's IQ feed (one engine instance), a
// *block* is one 65536-byte replay block (engine.cpp:68) = 32768 raw IQ samples = 8192 decimated
// samples, a *demod* is one registered demodulator+decoder pair (main.cpp:171-218), a *window* is
// one retriggerable timeout interval of a demod (tfa1.cpp:147-148, tfa2.cpp:351-356, whb.cpp:636-642),
// a *frame* is one decoder::flush(), a *record* is one sensordata_t (decoder.h:21-31).
#pragma once
#include <stdint.h>

namespace tfr {

constexpr int kBlockBytes = 65536;      // engine.cpp:68
constexpr int kBlockRaw = 32768;        // raw IQ samples per block
constexpr int kBlockDec = 8192;         // decimated (384 kS/s) samples per block
constexpr int kIdxPerBlock = 16384;     // `len` of fsk_demod::process, index += 2 per sample (fm_demod.cpp:42)
constexpr int kHistBytes = 96;          // raw bytes of FIR history carried between submits (needs 92)
constexpr int kMaxSeg = 28;             // trigger segments per block (>= 8192/357 + 2)
constexpr int kMaxDemods = 5;
constexpr int kMaxRdata = 64;
constexpr int kMaxEvt = 128;            // trigger events kept per block; a block with more is 'dense'
constexpr int kWinPerBlock = 24;        // worst-case demod windows that can start in one block (8192/357 + 1)
constexpr int kRdataBytes = 256;        // decoder::rdata (decoder.h:52)

// demod kinds, in the reference's registration order (main.cpp:173-218)
enum Kind : int { K_TFA1 = 0, K_TFA2 = 1, K_TFA3 = 2, K_TX22 = 3, K_WHB = 4 };

struct Biquad {          // iir2 (dsp_stuff.h:18-27): direct form I, double state
	double d1, d2, y0, y1;
};

struct BiquadCoef { double b0, b1, b2, a1, a2; };

struct DemodCfg {
	int32_t kind;        // Kind
	int32_t type;        // sensor_e value reported in frames/records
	int32_t timeout;     // samples a trigger keeps the demod active: 400 / (int)(16*spb) / (int)(8*spb)
	int32_t pad;
	double spb;          // samples per bit at 384 kS/s (main.cpp:186,194,202,217)
	BiquadCoef lp;       // tfa2: 0.5/spb (tfa2.cpp:321); whb pulse filter 2.0/spb (whb.cpp:610)
	BiquadCoef lp_avg;   // whb: 0.0025/spb (whb.cpp:611)
	// TFA_2 family, the edge arithmetic of tfa2.cpp:391-398 as integers (built on the host with the same IEEE doubles):
	// `tdiff>spb/4 && tdiff<32*spb` is td_lo <= tdiff <= td_hi, and numbits = (bit_diff + est_spb/2)/est_spb, truncated,
	// is nbits[bit_diff] (bit_diff = tdiff/2 <= td_hi/2 < kNbitsTab)
	int32_t td_lo, td_hi;
	uint8_t nbits[704];
};
constexpr int kNbitsTab = 704;

struct DevConfig {
	int32_t n_demods;
	int32_t t_max;       // max timeout over the registered demods
	int32_t filter;      // 0 narrow, 1 wide
	int32_t thresh_cfg;  // -t value, 0 = auto
	int32_t flags;
	int32_t n_streams;
	DemodCfg d[kMaxDemods];
};

// everything one registered demodulator + decoder carries from sample to sample
struct DemodState {
	// demodulator base (decoder.h:61-73)
	int32_t last_bit_idx;
	int32_t timeout_cnt;
	// tfa1_demod (tfa1.h:23-33)
	int32_t mark_lvl;
	int32_t rssi_i;      // tfa1: peak hold; tfa2: power accumulator
	// tfa2_demod (tfa2.h:29-44)
	int32_t bitcnt, dmin, dmax, offset, last_bit;
	// whb_demod (whb.h:45-61)
	int32_t last_dev, avg_of;
	uint32_t step_lo;    // `step` is uint64 in the reference but reset per window; 32 bits suffice for tdiff
	uint32_t last_peak;
	double rssi_d;
	Biquad lp, lp_avg;
	// decoder side: shift register framer (tfa1.cpp:120-134, tfa2.cpp:281-314, whb.cpp:566-603)
	uint32_t sr;
	int32_t sr_cnt;
	int32_t byte_cnt;
	int32_t synced;
	int32_t invert;
	int32_t inv_cnt;     // tfa2 family: inverted sync words seen in the current window (tfa2.cpp:294-300 prints a line for each)
	int32_t w_last_bit, w_psk, w_last_psk, w_nrzs;
	uint32_t w_lfsr;
	uint8_t rdata[kRdataBytes];
};

struct alignas(128) StreamState {
	alignas(16) uint8_t hist[2][kHistBytes];  // FIR history (last raw bytes of the previous submit), double buffered
	int64_t blocks_done;          // blocks fully decoded -> position base = blocks_done * 8192
	// fsk_demod (fm_demod.h:24-30)
	int32_t thresh;
	int32_t thresh_mode;          // 1 = auto
	int32_t triggered_avg;
	int32_t runs;
	// trigger bookkeeping carried between blocks
	int32_t any_timeout;          // remaining samples for which "any demod active" holds (T_max logic)
	int32_t carry_in;             // decimated samples at the head of the next block covered by earlier triggers
	int16_t last_i, last_q;       // previous decimated sample (every demod's last_i/last_q)
	int32_t hist_parity;
	int32_t trig_age;             // samples since the last trigger as of the end of the previous call (>= 1; huge = none)
	// scratch of the current process call (threshold kernel -> window kernels)
	int32_t call_last_trig;       // position of the last trigger seen so far in this call (negative: before the call)
	uint32_t call_cursor;         // coverage cursor
	uint32_t t2_done;             // blocks of this call the threshold kernel has walked (0 between calls)
	int32_t spec_lo;              // auto threshold: the lower bound every speculative front-end launch of the current call keeps
	                              // (thresh at the start of the call - spec_margin); frozen by save_history_kernel
	uint32_t win_n[kMaxDemods];   // windows listed so far per demod
	uint32_t win_cum[kMaxDemods]; // active samples in closed windows
	uint32_t win_open[kMaxDemods];// 1 if the last listed window is still open
	uint32_t win_cont[kMaxDemods];// 1 if demod d's last window of the previous call runs on into the next call
	DemodState d[kMaxDemods];
	DemodState fin[kMaxDemods];   // state left by an unfinished (ran out of data) last window, pending verification
	Biquad lp_next[kMaxDemods];   // filter chains: the proven biquad state after the call's last window sample
};

// one entry per (stream, block) of a process() call, written by the front-end
struct TileDesc {
	uint16_t n_seg;               // trigger segments that start inside this block
	uint16_t carry_out;           // samples of the NEXT block still covered by this block's last trigger
	uint16_t seg_start[kMaxSeg];  // first covered sample (the trigger itself)
	uint16_t seg_len[kMaxSeg];    // covered samples, clipped to the block end
	uint32_t n_trig;              // samples with pwr > thresh_lo
	uint32_t pad;
};
static_assert(sizeof(TileDesc) == 2 * 2 + 4 * kMaxSeg + 8, "TileDesc layout");

struct StreamJob {
	const uint8_t *iq;            // device pointer to this stream's submitted bytes
	uint32_t n_blocks;            // blocks in this submit
	uint32_t dec_off;             // offset (in blocks) of this stream inside the sparse decimated buffer
	uint32_t win_off;             // first window-list entry of this stream (demod d starts at win_off + d*win_cap)
	uint32_t win_cap;             // window-list capacity per demod = n_blocks*kWinPerBlock + 4
	int64_t base_blocks;          // blocks of this stream decoded by earlier calls (position base = base_blocks * 8192); from the
	                              // host, so that kernels of call i+1 that run beside call i's verifier need not read StreamState
};

// windows listed for one stream by the threshold kernel of ONE call (per work-buffer slot: the back-end of call
// i reads it while the threshold kernel of call i+1 is already rewriting StreamState::win_n)
struct WinCount {
	uint32_t n[kMaxDemods];
	uint32_t cum[kMaxDemods];
};

// one demodulator window = one retriggerable timeout interval, derived from the trigger positions alone
// (a trigger at t keeps demod d active for samples t .. t+T_d-1; tfa1.cpp:147-148, tfa2.cpp:351-356)
struct WinEntry {
	uint32_t start;               // first sample (the trigger), position inside this process call
	uint32_t end;                 // sample at which timeout_cnt reaches 0 and flush() runs (may lie beyond the call)
	uint32_t cum;                 // active samples of this demod in earlier windows of this call (tap index base)
	uint32_t flags;               // kWinCont: the window was already open when the call began
};
constexpr uint32_t kWinCont = 1u;

// what a speculative window run leaves behind so that the verifier can prove (or repair) its carry-in
struct WinRec {
	int32_t frame_idx;            // frame emitted by this window, -1 if none
	uint32_t flags;               // kRec*
	// TFA_1: decoder shift register carry (tfa1.cpp:115-117 keeps sr across flush)
	uint32_t head31;              // first min(31,nbits) bits the window pushed into the framer
	uint32_t sr_final;            // shift register at window end (under the assumed carry-in)
	int32_t nbits;
	// TFA_2 family: biquad and last_bit_idx carry (tfa2.cpp:325-334 resets neither)
	int32_t first_edge;           // in-block index of the first edge candidate, valid if kRecEdge
	int32_t first_edge_block;
	int32_t lbi_end;              // last_bit_idx at window end, relative to lbi_end_block
	int32_t lbi_end_block;
	int32_t pad;
	double u_y0, u_y1;            // biquad outputs assumed at window start (after warm-up)
	double e_y0, e_y1;            // biquad outputs at window end
	unsigned long long ld_hash;   // hash of the slicer input sequence (int)y over the window
	int32_t lbi_in;               // edge repair: the predecessor chain's last_bit_idx, relative to the window's first block
	int32_t pad3;                 // 1 + frame slot of the window's "Inverted SYNC" notice (status 3), 0 = none
};
// Filter chains (biq_kernel): the TFA_2-family low-pass runs ahead of the slicers, as a chain over kBiqK consecutive
// windows from ONE speculated start state; a chain's record holds the filter outputs it assumed and the ones it left
struct BiqRec {
	double u_y0, u_y1;            // biquad outputs assumed at the chain's first sample (after warm-up)
	double e_y0, e_y1;            // biquad outputs after the chain's last sample
};
constexpr int kBiqK = 4;          // windows per filter chain
constexpr uint32_t kRecRan = 1u, kRecExact = 2u, kRecEdge = 4u, kRecUnfinished = 8u;
constexpr uint32_t kRecLbiIn = 16u;   // the run used the explicit last_bit_idx in lbi_in instead of the 'far' assumption
constexpr int32_t kPadOk = 1, kPadEdgeRepair = 2;   // WinRec::pad bits written by flag_kernel

// device-side frame / record (converted to tfr_frame / tfr_record on the host)
struct DevFrame {
	int32_t stream, demod, type, status;
	int32_t byte_cnt, offset, n_records, first_record;
	int64_t pos;
	double rssi_raw;
	uint8_t rdata[kMaxRdata];
};

struct DevRecord {
	int32_t stream, type;
	uint64_t id;
	double temp, humidity;
	int32_t alarm, flags, sequence, frame;
	int64_t pos;
};

struct BlockTrace { int32_t thresh, triggered, triggered_avg; };

// Auto threshold (fm_demod.cpp:58-73) moves by 2 every 4th block as a function of the blocks before it, so the
// front-end cannot know a block's threshold.  It keeps everything above a lower bound instead.  The whole call
// is first run against thresh_at_call_start - spec_margin(): in steady state the threshold wanders by a few
// steps only, so this speculation normally holds for every block; the threshold kernel stops a stream at the
// first block whose true threshold fell below the bound and the host re-runs the rest of that stream in epochs
// of kEpochBlocks blocks with the provable bound 2*ceil(kEpochBlocks/4).
constexpr int kEpochBlocks = 64;
constexpr int kEpochMargin = 2 * ((kEpochBlocks + 3) / 4);
__host__ __device__ inline int spec_margin(int thresh)
{
	const int m = ((thresh >> 4) + 1) & ~1;
	return m < 16 ? 16 : m;
}

struct Counters {
	uint32_t n_frames;
	uint32_t n_records;
	uint32_t overflow;
	uint32_t n_reruns;            // windows the verifier had to re-run because a speculated carry-in was wrong
	unsigned long long active_samples;
	unsigned long long n_windows;
	uint32_t rerun_sr, rerun_biquad, rerun_edge, pad2;   // why the verifier re-ran windows
	uint32_t par_cheap, ver_checked, ver_cheap, ver_full;  // parallel cheap repairs; verifier: flagged windows looked at, cheap fixes, full re-runs
};

struct FrontParams {
	const StreamJob *jobs;
	StreamState *st;
	TileDesc *tiles;
	uint32_t *dec;         // sparse decimated buffer, one uint32 (I lo16, Q hi16) per sample
	int tile0;             // first block of this epoch inside the submit
	int n_tiles;           // blocks in this epoch (grid.x)
	int t_max;
	int keep_all;          // TFR_FLAG_KEEP_DECIM: write every sample
	int margin;            // auto threshold: > 0 the launch keeps every sample with pwr > thresh - margin (thresh as of
	                       // the launch, epoch fallback); 0 = the call's frozen speculative bound StreamState::spec_lo
	int use_progress;      // 1: a stream's first block of this launch is its StreamState::t2_done (epoch fallback)
	uint32_t *events;      // [gtile][kMaxEvt]
	const void *tmaps;     // frontend_tc_kernel / frontend_screen_kernel: [stream][2] CUtensorMap (rows of the submit, and the 32 bytes in front of every row)
	// screening front-end (frontend_screen.cu)
	const uint8_t *screen_consts;  // ScreenConsts blob: the band matrix of the combined 46-tap filter, the constant operands
	int screen_shift;      // q: screen value = linear filter output * 2^q
	int screen_slack;      // what a sample's true |I|+|Q| can exceed its screen value by, in output units (rounded up)
	int n_streams;
	uint32_t *work_ctr;    // the launch's work counter: CTAs fetch (stream, block) items from it
	uint8_t *hist_copy;    // [stream][kHistBytes]: the FIR history the call started from (the window kernel runs after save_history)
	int32_t *screen_dbg;   // debug (TFR_FLAG_TAPS): [gtile][8192][2] screen values of I and Q, null otherwise
	uint32_t *screen_stat; // [0] blocks screened sparse, [1] blocks handed back dense, [2] candidates checked, [3] of them true
};

struct BackParams {
	const DevConfig *cfg;
	const StreamJob *jobs;
	StreamState *st;
	const TileDesc *tiles;
	const uint32_t *dec;
	BlockTrace *trace;       // [gtile]
	DevFrame *frames;
	DevRecord *records;
	Counters *counters;
	int32_t *tap_i32[2];     // kind 0 fm_dev, kind 1 fm_dev_nrzs: [stream][demod][tap_cap]
	double *tap_f64;         // kind 2 iir2::step outputs
	uint32_t *tap_cnt;       // [stream][demod][3]
	uint32_t tap_cap;
	uint32_t max_frames, max_records;
	int tile0, n_tiles;      // epoch range inside the submit
	int n_streams;
	int margin;              // same meaning and value as FrontParams::margin of the front-end launch before it
	uint32_t *progress;      // [stream] blocks walked so far by the threshold kernel (host reads it back)
	WinCount *wincnt;        // [stream] window counts of this call (written by the threshold kernel when a stream finishes)
	const uint32_t *events;  // [gtile][kMaxEvt] (pos<<16 | pwr) of samples with pwr > thresh_lo, in order
	WinEntry *wins;
	WinRec *recs;
	int32_t *devfm;          // fm_dev per stored sample, direct mapped like dec
	int demod;               // kernels that run per registered demod: which one
	int max_blocks;
	int long_split;          // 1: win_kernel leaves the long window chains to winlong_kernel
	// The back-end of a call runs in PARTS, each as soon as the threshold walk has passed a chunk of blocks, beside the
	// front-end of the following chunks.  partcnt[k][stream].n[d] = number of leading windows of demod d that part k and
	// the parts before it cover: all closed, ending before the blocks walked so far, and cut at a chain head (so that
	// no window chain straddles two parts).  Written by thresh2_kernel (part_idx >= 0), read by the window kernels.
	// filter chains (null / -1: the window kernels filter for themselves)
	int32_t *ld;             // [fm slot][direct mapped like dec]: (int)y of the demodulator's low-pass, per window sample
	size_t ld_stride;        // elements per fm slot
	BiqRec *biq;             // chain records, indexed like wins/recs by the chain's first window
	int fm_slot[kMaxDemods]; // demod -> slot of ld, -1 none
	DemodState *fin;         // [stream][kMaxDemods], per work-buffer slot: state left by a window the data ended in
	int slot_tag;            // 1 + work-buffer slot: frames are tagged with it until they are parsed (parse_kernel of call i
	                         // leaves the frames that call i+1's early windows are appending alone)
	int long_all;            // 1: winlong_kernel takes every chain of its window range (the few windows of a call that need
	                         // the previous call's final state run after its verifier: a warp each, not a thread)
	WinCount *partcnt;       // [kMaxParts][n_streams]
	int part_idx;            // thresh2_kernel: which row of partcnt this launch fills, -1 none
	int part_lo, part_hi;    // window kernels: rows of partcnt bounding the windows of this launch (lo -1: from window 0, hi -1: to the end)
	// threshold-walk table (walk_table_kernel, null: the walk evaluates the blocks' event lists itself)
	uint32_t *walk_tab;      // [gtile][kWalkNT][4]: what the walk needs of a block under threshold walk_base + 2k
	int32_t *walk_base;      // [stream]: threshold of column 0, written by the table kernel of the same launch
	int walk_ct;             // walk_cta_kernel: threads per CTA, 0 = by the number of streams (TFR_WALK_CT)
	int walk_dbg;            // experiments (TFR_WALK_DBG): 1 = no ladder above the table, 2 = no range below it either
	uint32_t *walk_gap;      // non-null: the walk is walk_cta_kernel; walk_tab holds [gtile][kWalkNT][2] (the chain's part of an
	                         // entry), walk_gap [gtile][kWalkNT][4] (up to four window-opening gaps)
};
constexpr int kWalkNT = 16;  // thresholds tabulated per block: the one at the start of the launch -16 .. +14 in steps of two
constexpr int kMaxParts = 8;
// partcnt row kLateRow: the leading windows of every (stream, demodulator) that depend on the previous call's final state -
// window 0 (true carried state, the sample before position 0) and the windows whose filter warm-up reaches back to it
constexpr int kLateRow = kMaxParts - 1;

}  // namespace tfr


