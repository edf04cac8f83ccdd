/* tfr.h - C ABI of the B200 IQ->telegram decode path (libtfrb200.so).
 *
 * This is the drop-in boundary for ONE path of baycom/tfrec: the two calls engine::run makes per
 * block of samples,
 *
 *     int ld = dc.process_iq(data, len, filter_type);     // engine.cpp:85  (dsp_stuff.cpp:243-264)
 *     fsk->process(data, ld);                             // engine.cpp:86  (fm_demod.cpp:34-74)
 *
 * together with everything those two reach on the device: the two-stage integer FIR decimator
 * (dsp_stuff.cpp:172-230), the power trigger and auto-threshold (fm_demod.cpp:42-73), the
 * TFA_1 / TFA_2 / TFA_3 / TX22 / WeatherHub demodulators, bit framers and telegram parsers
 * (tfa1.cpp, tfa2.cpp, whb.cpp) and the CRCs (crc8.cpp, crc32.cpp).  Results come back as POD
 * mirrors of sensordata_t (decoder.h:21-31) which the host hands to decoder::store_data
 * (decoder.cpp:46-65), so the -e exec contract (decoder.cpp:67-96) is untouched.
 *
 * Plain C types only; no exceptions, no exit(); every call returns 0 or a negative TFR_E_* code
 * and tfr_last_error() describes the last failure on the calling thread.  There is no CPU
 * fallback: without a CUDA device of compute capability 10.x tfr_create fails.
 * A handle is not thread-safe; use one per thread or lock externally (the reference hot path is
 * single-threaded too).
 */
#ifndef TFR_H
#define TFR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFR_ABI_VERSION 2
#define TFR_BLOCK_BYTES 65536   /* replay block of engine.cpp:68 (RLS); all framing is in these units */
#define TFR_SAMPLE_RATE 1536000 /* engine.cpp:15 */
#define TFR_MAX_RDATA 64

/* sensor_e, decoder.h:11-19.  The type mask (-T, main.cpp:146-148) is the OR of 1<<value. */
enum { TFR_TFA_1 = 0, TFR_TFA_2 = 1, TFR_TFA_3 = 2, TFR_TX22 = 3, TFR_TFA_WHB = 5 };

enum {
	TFR_OK = 0,
	TFR_E_INVAL = -1,     /* bad argument */
	TFR_E_NODEVICE = -2,  /* no usable sm_100 device (no CPU fallback exists) */
	TFR_E_CUDA = -3,      /* CUDA runtime error, see tfr_last_error() */
	TFR_E_NOMEM = -4,
	TFR_E_BUSY = -5,      /* stream already has a pending submit */
	TFR_E_OVERFLOW = -6   /* a result buffer overflowed; results are truncated */
};

enum { TFR_MEM_HOST = 0, TFR_MEM_DEVICE = 1 };

/* tfr_config.flags */
#define TFR_FLAG_TAPS 1u        /* keep discriminator / biquad intermediates for tfr_read_taps */
#define TFR_FLAG_KEEP_DECIM 2u  /* write every decimated sample (not only trigger windows) for tfr_read_decimated */

typedef struct {
	uint32_t struct_size;   /* = sizeof(tfr_config) */
	int32_t device;         /* CUDA ordinal */
	int32_t types;          /* -T mask, main.cpp:103,146 (default 0x07) */
	int32_t filter;         /* 0 narrow, 1 wide: -W, dsp_stuff.cpp:175-177 */
	int32_t thresh;         /* -t, 0 = auto (fm_demod.cpp:24-27) */
	int32_t n_streams;      /* independent IQ streams ("sticks"), each with its own engine state */
	uint32_t flags;
	uint32_t max_frames;    /* capacity of the per-process() frame buffer, 0 = default */
	uint64_t max_blocks_per_submit; /* sizing hint for work buffers, 0 = grow on demand */
} tfr_config;

/* One decoder::flush() that passed its minimum-length gate (tfa1.cpp:48, tfa2.cpp:76,222, whb.cpp:484).
 * rdata is what decoder::store_bytes would be given at the -X cross-check seam (decoder.cpp:35-40). */
typedef struct {
	int32_t stream;
	int32_t type;       /* sensor_e of the decoder that flushed */
	int32_t status;     /* 0 ok, 1 bad CRC, 2 failed sanity checks; 3: not a flush but a notice - the TFA_2-family window
	                     * that ended at pos saw byte_cnt inverted sync words, for each of which the reference prints
	                     * "Inverted SYNC" (tfa2.cpp:294-300); rdata is empty */
	int32_t byte_cnt;
	int64_t pos;        /* 384 kS/s sample index (per stream) at which flush() ran */
	int32_t rssi;       /* the int the reference passes to flush(): (int)(10*log10(.)) */
	int32_t offset;     /* tfa2 family: slicer offset, reported as -1536.0*offset/131072 kHz */
	double rssi_raw;    /* accumulator before the log10 (tfa1.cpp:180, tfa2.cpp:434, whb.cpp:696) */
	int32_t n_records;
	int32_t first_record;
	uint8_t rdata[TFR_MAX_RDATA];
} tfr_frame;

/* sensordata_t (decoder.h:21-31).  ts is filled with time(0) by the host at poll time. */
typedef struct {
	int32_t stream;
	int32_t type;
	uint64_t id;
	double temp;
	double humidity;
	int32_t alarm;
	int32_t flags;
	int32_t sequence;
	int32_t rssi;
	int64_t ts;
	int64_t pos;
	int32_t frame;      /* index into the frame list of the same poll */
	int32_t reserved;
} tfr_record;

typedef struct {
	int32_t thresh;         /* threshold in force for the block (fm_demod.cpp:48) */
	int32_t triggered;      /* samples with any demodulator active (fm_demod.cpp:51-52) */
	int32_t triggered_avg;  /* fm_demod.cpp:58 */
} tfr_block_trace;

typedef struct {
	uint64_t blocks;            /* 65536-byte blocks decoded since create */
	uint64_t raw_samples;       /* IQ samples consumed */
	uint64_t active_samples;    /* 384 kS/s samples that were inside a trigger window */
	uint64_t frames;
	uint64_t records;
	uint64_t kernel_launches;   /* kernels launched by this handle */
	uint64_t windows;           /* demodulator windows decoded (one GPU thread each) */
	uint64_t reruns;            /* windows the verifier re-ran because a speculated carry-in state was wrong */
	uint32_t reruns_sr, reruns_biquad, reruns_edge;   /* ... by cause */
	uint32_t fallback_epochs;   /* auto threshold: 64-block epochs re-run because the speculative threshold bound failed */
	double last_frontend_ms;    /* CUDA-event time of the decimate+trigger kernel(s) of the last tfr_process */
	double last_backend_ms;     /* ... of the demod/framer/parser kernels */
	double last_h2d_ms;
	double last_total_ms;       /* first H2D copy (or first kernel) to last kernel of the last tfr_process */
	/* screening front-end (since create; zero when it is off: TFR_FE=dense, TFR_FLAG_KEEP_DECIM, decimated submits) */
	uint64_t screen_blocks;     /* blocks whose samples the tensor-core screen proved to be no triggers, up to the candidates */
	uint64_t dense_blocks;      /* blocks the screen handed to the dense exact kernel (bursts) */
	uint64_t screen_candidates; /* samples of the screened blocks whose exact value had to be computed */
	uint64_t screen_triggers;   /* ... of which really exceeded the trigger bound */
} tfr_stats;

typedef struct tfr_handle tfr_handle;

/* replaces: new downconvert(2) + fsk_demod(&demods, thresh, dbg) + the demod registration of
 * main.cpp:171-225, once per stream */
int tfr_create(const tfr_config *cfg, tfr_handle **out);
void tfr_destroy(tfr_handle *h);

/* replaces the fread + (u8-128)<<6 conversion of engine.cpp:67-81: hand over raw rtl-sdr bytes
 * (unsigned 8-bit I,Q interleaved).  nbytes must be a multiple of TFR_BLOCK_BYTES (the reference
 * drops a trailing partial block, engine.cpp:73-76; the caller does the same).  Host memory is
 * copied to the device inside the call (pinned memory makes that asynchronous); device memory must
 * be 16-byte aligned and stay valid until tfr_sync/tfr_poll_* returns.  One submit per stream
 * between two tfr_process calls. */
int tfr_submit(tfr_handle *h, int stream, const uint8_t *iq, size_t nbytes, int mem);
/* The same for a caller that keeps its own decimator: int16 I,Q already at 384 kS/s, i.e. exactly what the reference
 * hands to fsk_demod::process(int16_t *data_iq, int len) (fm_demod.cpp:34; engine.cpp:86).  n_int16 must be a positive
 * multiple of 16384 (one reference block).  Raw and decimated submits cannot be mixed in one tfr_process call. */
int tfr_submit_decimated(tfr_handle *h, int stream, const int16_t *iq16, size_t n_int16, int mem);

/* replaces dc.process_iq + fsk->process (engine.cpp:85-86) for everything submitted: enqueues the
 * kernels on the handle's CUDA stream and returns without waiting */
int tfr_process(tfr_handle *h);
int tfr_sync(tfr_handle *h);

/* results of the tfr_process calls since the last poll, ordered by (stream, pos, registration
 * order) - the order the reference prints them in.  Both calls synchronise first.  Passing
 * out=NULL returns the number available.  A negative return is a TFR_E_* code. */
long tfr_poll_frames(tfr_handle *h, tfr_frame *out, size_t cap);
long tfr_poll_records(tfr_handle *h, tfr_record *out, size_t cap);
/* drop the results returned so far (poll calls are otherwise idempotent) */
int tfr_clear_results(tfr_handle *h);

/* fsk_demod state (fm_demod.h:24-26): current threshold of a stream, and the per-block trace the
 * reference prints with -DDD (fm_demod.cpp:60-72) for the blocks processed since the last clear */
int tfr_get_thresh(tfr_handle *h, int stream, int32_t *thresh);
long tfr_read_block_trace(tfr_handle *h, int stream, tfr_block_trace *out, size_t cap);

/* debug taps (need TFR_FLAG_TAPS): kind 0 = fm_dev results (dsp_stuff.cpp:284-292, int32),
 * kind 1 = fm_dev_nrzs results (dsp_stuff.cpp:269-279, int32), kind 2 = iir2::step outputs
 * (dsp_stuff.cpp:47-55, double); `demod` is the registration index (main.cpp:173-218 order).
 * Returns the number of elements copied (or available if out==NULL). */
long tfr_read_taps(tfr_handle *h, int stream, int demod, int kind, void *out, size_t cap_elems);

/* debug (needs TFR_FLAG_TAPS, screening front-end): the screen values of the last tfr_process for a stream, two int32
 * (I, Q) per 384 kS/s sample: the linear 46-tap filter output times 2^shift minus the centre of the floor losses.  The
 * true sample obeys |I|+|Q| <= (|vI|+|vQ|) / 2^shift + slack.  Blocks that went to the dense kernel hold zeros. */
long tfr_read_screen(tfr_handle *h, int stream, int32_t *out, size_t cap_int32, int *shift, int *slack);

/* debug (needs TFR_FLAG_KEEP_DECIM): decimated int16 I,Q of the last tfr_process for a stream,
 * i.e. the buffer process_iq leaves behind in place (dsp_stuff.cpp:197,225) */
long tfr_read_decimated(tfr_handle *h, int stream, int16_t *out, size_t cap_int16);

/* stand-alone entry points for the pieces the reference exposes as free functions / classes */
/* downconvert(2)::process_iq over a whole buffer with zero initial history; device or host pointers;
 * writes (nbytes/8)*2 int16 (I,Q interleaved at 384 kS/s) and returns that count */
long tfr_decimate(int device, const uint8_t *iq, size_t nbytes, int filter, int16_t *out, int mem);
/* downconvert(passes)::process_iq (dsp_stuff.cpp:232-264) for any passes in 1..8, i.e. decimation 2^passes
 * (BASELINE configs[4] sweeps /2 .. /32): passes-1 times decimate::process2x1 (:204-230) then process2x (:172-202,
 * `filter` 0 narrow / 1 wide), zero initial history, whole buffer.  Writes (nbytes/2 >> passes) I,Q pairs and
 * returns the number of int16 written.  `reps` > 1 repeats the stage launches for timing; *kernel_ms (may be
 * NULL) receives the CUDA-event time of one cascade. */
long tfr_downconvert(int device, const uint8_t *iq, size_t nbytes, int passes, int filter, int16_t *out, int mem,
		     int reps, float *kernel_ms);
/* downconvert as the streaming class of dsp_stuff.h:46-56 (`downconvert(int p)` + `process_iq`): the history of every
 * stage is carried from call to call, so feeding a stream block by block gives the same samples as feeding it whole.
 * passes 1..5.  tfr_dc_process takes raw rtl-sdr bytes (what engine.cpp:77-78 converts) and is the fast path (one fused
 * kernel, csrc/decim_fused.cu); tfr_dc_process_i16 is process_iq's own signature - int16 I,Q in place on the host,
 * len = number of int16, returns the new len - for callers that hold converted samples.  The number of IQ pairs of a
 * call must be a multiple of 2^passes (the reference drops an odd trailing sample of a block for good). */
typedef struct tfr_dc tfr_dc;
int tfr_dc_create(int device, int passes, tfr_dc **out);
void tfr_dc_destroy(tfr_dc *h);
long tfr_dc_process(tfr_dc *h, const uint8_t *iq, size_t nbytes, int filter, int16_t *out, int mem);
long tfr_dc_process_i16(tfr_dc *h, int16_t *data_iq, int len, int filter);
/* decoder::store_bytes + flush(0) (main.cpp:45-50, the -X seam) run through the device parser */
int tfr_parse_bytes(tfr_handle *h, int type, const uint8_t *bytes, int len, tfr_frame *frame,
		    tfr_record *recs, int max_recs);

/* (with the screening front-end active this waits for the calls in flight, like tfr_sync: the screen's counters live on
 * the device) */
int tfr_get_stats(tfr_handle *h, tfr_stats *out);
const char *tfr_last_error(void);
int tfr_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif
