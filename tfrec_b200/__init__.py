"""tfrec_b200 - B200-native IQ->telegram decode path of baycom/tfrec.

The product is the C-ABI shared library `libtfrb200.so` (include/tfr.h) built from tfrec_b200/csrc by
tfrec_b200/build.py; the C++ mirror of the reference's engine/decoder plugin surface lives in
tfrec_b200/host/.  This module is only the ctypes binding tests and bench.py use to call through that
ABI.  There is no CPU fallback: loading fails loudly if the library is missing, and tfr_create fails
without an sm_100 device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtfrb200.so")

TFA_1, TFA_2, TFA_3, TX22, TFA_WHB = 0, 1, 2, 3, 5
BLOCK_BYTES = 65536
MEM_HOST, MEM_DEVICE = 0, 1
FLAG_TAPS, FLAG_KEEP_DECIM = 1, 2
STATUS_NOTICE = 3          # tfr_frame.status of an "Inverted SYNC" notice (not a flush)

ABI_SYMBOLS = ["tfr_create", "tfr_destroy", "tfr_submit", "tfr_submit_decimated", "tfr_process", "tfr_sync", "tfr_poll_frames",
               "tfr_poll_records", "tfr_clear_results", "tfr_get_thresh", "tfr_read_block_trace", "tfr_read_taps",
               "tfr_read_decimated", "tfr_read_screen", "tfr_decimate", "tfr_downconvert", "tfr_dc_create", "tfr_dc_destroy", "tfr_dc_process",
               "tfr_dc_process_i16", "tfr_parse_bytes", "tfr_get_stats", "tfr_last_error", "tfr_abi_version"]


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("types", C.c_int32), ("filter", C.c_int32),
                ("thresh", C.c_int32), ("n_streams", C.c_int32), ("flags", C.c_uint32), ("max_frames", C.c_uint32),
                ("max_blocks_per_submit", C.c_uint64)]


class Frame(C.Structure):
    _fields_ = [("stream", C.c_int32), ("type", C.c_int32), ("status", C.c_int32), ("byte_cnt", C.c_int32),
                ("pos", C.c_int64), ("rssi", C.c_int32), ("offset", C.c_int32), ("rssi_raw", C.c_double),
                ("n_records", C.c_int32), ("first_record", C.c_int32), ("rdata", C.c_uint8 * 64)]


class Record(C.Structure):
    _fields_ = [("stream", C.c_int32), ("type", C.c_int32), ("id", C.c_uint64), ("temp", C.c_double),
                ("humidity", C.c_double), ("alarm", C.c_int32), ("flags", C.c_int32), ("sequence", C.c_int32),
                ("rssi", C.c_int32), ("ts", C.c_int64), ("pos", C.c_int64), ("frame", C.c_int32),
                ("reserved", C.c_int32)]


class BlockTrace(C.Structure):
    _fields_ = [("thresh", C.c_int32), ("triggered", C.c_int32), ("triggered_avg", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("blocks", C.c_uint64), ("raw_samples", C.c_uint64), ("active_samples", C.c_uint64),
                ("frames", C.c_uint64), ("records", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("windows", C.c_uint64), ("reruns", C.c_uint64), ("reruns_sr", C.c_uint32), ("reruns_biquad", C.c_uint32),
                ("reruns_edge", C.c_uint32), ("fallback_epochs", C.c_uint32),
                ("last_frontend_ms", C.c_double), ("last_backend_ms", C.c_double), ("last_h2d_ms", C.c_double),
                ("last_total_ms", C.c_double), ("screen_blocks", C.c_uint64), ("dense_blocks", C.c_uint64),
                ("screen_candidates", C.c_uint64), ("screen_triggers", C.c_uint64)]


class TfrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("tfr error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """dlopen the in-tree library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libtfrb200.so is not built: run `python tfrec_b200/build.py` (needs nvcc). "
                          "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    P = C.c_void_p
    L.tfr_create.argtypes = [C.POINTER(Config), C.POINTER(P)]
    L.tfr_destroy.argtypes = [P]
    L.tfr_destroy.restype = None
    L.tfr_submit.argtypes = [P, C.c_int, C.c_void_p, C.c_size_t, C.c_int]
    L.tfr_process.argtypes = [P]
    L.tfr_sync.argtypes = [P]
    L.tfr_poll_frames.argtypes = [P, C.POINTER(Frame), C.c_size_t]
    L.tfr_poll_frames.restype = C.c_long
    L.tfr_poll_records.argtypes = [P, C.POINTER(Record), C.c_size_t]
    L.tfr_poll_records.restype = C.c_long
    L.tfr_clear_results.argtypes = [P]
    L.tfr_get_thresh.argtypes = [P, C.c_int, C.POINTER(C.c_int32)]
    L.tfr_read_block_trace.argtypes = [P, C.c_int, C.POINTER(BlockTrace), C.c_size_t]
    L.tfr_read_block_trace.restype = C.c_long
    L.tfr_read_taps.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    L.tfr_read_taps.restype = C.c_long
    L.tfr_read_decimated.argtypes = [P, C.c_int, C.c_void_p, C.c_size_t]
    L.tfr_read_decimated.restype = C.c_long
    L.tfr_read_screen.argtypes = [P, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.tfr_read_screen.restype = C.c_long
    L.tfr_decimate.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int]
    L.tfr_decimate.restype = C.c_long
    L.tfr_downconvert.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                  C.POINTER(C.c_float)]
    L.tfr_downconvert.restype = C.c_long
    L.tfr_submit_decimated.argtypes = [P, C.c_int, C.c_void_p, C.c_size_t, C.c_int]
    L.tfr_dc_create.argtypes = [C.c_int, C.c_int, C.POINTER(P)]
    L.tfr_dc_destroy.argtypes = [P]
    L.tfr_dc_destroy.restype = None
    L.tfr_dc_process.argtypes = [P, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int]
    L.tfr_dc_process.restype = C.c_long
    L.tfr_dc_process_i16.argtypes = [P, C.c_void_p, C.c_int, C.c_int]
    L.tfr_dc_process_i16.restype = C.c_long
    L.tfr_parse_bytes.argtypes = [P, C.c_int, C.c_char_p, C.c_int, C.POINTER(Frame), C.POINTER(Record), C.c_int]
    L.tfr_get_stats.argtypes = [P, C.POINTER(Stats)]
    L.tfr_last_error.restype = C.c_char_p
    L.tfr_abi_version.restype = C.c_int
    _lib = L
    return L


def _check(rc):
    if rc < 0:
        raise TfrError(rc, load().tfr_last_error().decode())
    return rc


def format_exec(r) -> str:
    """decoder::execute_handler argv minus handler and ts (decoder.cpp:72-91)."""
    if r["type"] != TFA_WHB:
        nid = r["id"] | (r["type"] << 24)
        head = "%04x" % nid
    else:
        head = "%013x" % r["id"]
    return "%s %+.1f %s %d %d %d %d" % (head, r["temp"], _fmt_g(r["humidity"]), r["sequence"], r["alarm"],
                                        r["rssi"], r["flags"])


def _fmt_g(v):
    return "%g" % v


class Receiver:
    """One tfr handle = n_streams independent receivers (engine + downconvert + fsk_demod + demods)."""

    def __init__(self, types=0x07, filter=0, thresh=0, n_streams=1, device=0, flags=0, max_frames=0,
                 max_blocks_per_submit=0):
        self.L = load()
        cfg = Config(C.sizeof(Config), device, types, filter, thresh, n_streams, flags, max_frames,
                     max_blocks_per_submit)
        self.h = C.c_void_p()
        _check(self.L.tfr_create(C.byref(cfg), C.byref(self.h)))
        self.n_streams = n_streams

    def close(self):
        if getattr(self, "h", None):
            self.L.tfr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def submit(self, stream, iq, nbytes=None):
        """iq: numpy uint8 array (host) or (device_ptr:int) with nbytes."""
        if isinstance(iq, np.ndarray):
            assert iq.dtype == np.uint8 and iq.flags["C_CONTIGUOUS"]
            self._keep = getattr(self, "_keep", []) + [iq]
            _check(self.L.tfr_submit(self.h, stream, iq.ctypes.data, iq.size if nbytes is None else nbytes, MEM_HOST))
        else:
            _check(self.L.tfr_submit(self.h, stream, C.c_void_p(int(iq)), nbytes, MEM_DEVICE))

    def submit_decimated(self, stream, iq16):
        """int16 I,Q already at 384 kS/s (what fsk_demod::process(int16_t*, int) takes); numpy int16, a multiple of 16384"""
        assert isinstance(iq16, np.ndarray) and iq16.dtype == np.int16 and iq16.flags["C_CONTIGUOUS"]
        self._keep = getattr(self, "_keep", []) + [iq16]
        _check(self.L.tfr_submit_decimated(self.h, stream, iq16.ctypes.data, iq16.size, MEM_HOST))

    def submit_host_ptr(self, stream, ptr, nbytes):
        _check(self.L.tfr_submit(self.h, stream, C.c_void_p(int(ptr)), nbytes, MEM_HOST))

    def process(self):
        _check(self.L.tfr_process(self.h))

    def sync(self):
        _check(self.L.tfr_sync(self.h))
        self._keep = []

    def frames(self, notices=False):
        """the flushes of the calls since the last clear(); notices=True also returns the "Inverted SYNC" entries
        (status STATUS_NOTICE, byte_cnt = how many the window saw) in their place in the output order"""
        n = _check(self.L.tfr_poll_frames(self.h, None, 0))
        buf = (Frame * max(n, 1))()
        n = _check(self.L.tfr_poll_frames(self.h, buf, n))
        self._keep = []
        return [{"stream": f.stream, "type": f.type, "status": f.status, "byte_cnt": f.byte_cnt, "pos": f.pos,
                 "rssi": f.rssi, "offset": f.offset, "rssi_raw": f.rssi_raw, "n_records": f.n_records,
                 "rdata": bytes(f.rdata[:min(f.byte_cnt, 64)]).hex()} for f in buf[:n] if notices or f.status != STATUS_NOTICE]

    def inverted_syncs(self, stream=None):
        """how many "Inverted SYNC" lines the reference prints for what was decoded since the last clear() (tfa2.cpp:294-300)"""
        return sum(f["byte_cnt"] for f in self.frames(notices=True)
                   if f["status"] == STATUS_NOTICE and (stream is None or f["stream"] == stream))

    def records(self):
        n = _check(self.L.tfr_poll_records(self.h, None, 0))
        buf = (Record * max(n, 1))()
        n = _check(self.L.tfr_poll_records(self.h, buf, n))
        out = []
        for r in buf[:n]:
            d = {"stream": r.stream, "type": r.type, "id": r.id, "temp": r.temp, "humidity": r.humidity,
                 "alarm": r.alarm, "flags": r.flags, "sequence": r.sequence, "rssi": r.rssi, "pos": r.pos,
                 "frame": r.frame, "ts": r.ts}
            d["exec"] = format_exec(d)
            out.append(d)
        return out

    def n_records(self):
        return _check(self.L.tfr_poll_records(self.h, None, 0))

    def clear(self):
        _check(self.L.tfr_clear_results(self.h))

    def thresh(self, stream=0):
        v = C.c_int32()
        _check(self.L.tfr_get_thresh(self.h, stream, C.byref(v)))
        return v.value

    def block_trace(self, stream=0):
        n = _check(self.L.tfr_read_block_trace(self.h, stream, None, 0))
        buf = (BlockTrace * max(n, 1))()
        n = _check(self.L.tfr_read_block_trace(self.h, stream, buf, n))
        return np.array([[b.thresh, b.triggered, b.triggered_avg] for b in buf[:n]], dtype=np.int32).reshape(n, 3)

    def taps(self, stream, demod, kind):
        n = _check(self.L.tfr_read_taps(self.h, stream, demod, kind, None, 0))
        out = np.empty(n, dtype=np.float64 if kind == 2 else np.int32)
        if n:
            _check(self.L.tfr_read_taps(self.h, stream, demod, kind, out.ctypes.data, n))
        return out

    def screen(self, stream=0):
        """debug (FLAG_TAPS): (values [n, 2] int32, shift, slack) of the screening front-end for the last process()"""
        sh, sl = C.c_int(0), C.c_int(0)
        n = _check(self.L.tfr_read_screen(self.h, stream, None, 0, C.byref(sh), C.byref(sl)))
        out = np.empty(n, dtype=np.int32)
        if n:
            _check(self.L.tfr_read_screen(self.h, stream, out.ctypes.data, n, C.byref(sh), C.byref(sl)))
        return out.reshape(-1, 2), sh.value, sl.value

    def decimated(self, stream=0):
        n = _check(self.L.tfr_read_decimated(self.h, stream, None, 0))
        out = np.empty(n, dtype=np.int16)
        if n:
            _check(self.L.tfr_read_decimated(self.h, stream, out.ctypes.data, n))
        return out

    def parse_bytes(self, sensor_type, data: bytes):
        f = Frame()
        recs = (Record * 8)()
        n = _check(self.L.tfr_parse_bytes(self.h, sensor_type, data, len(data), C.byref(f), recs, 8))
        if f.status == -1:      # below the length gate of the type's flush(): the reference does nothing
            return None, []
        out = []
        for r in recs[:n]:
            d = {"stream": 0, "type": r.type, "id": r.id, "temp": r.temp, "humidity": r.humidity, "alarm": r.alarm,
                 "flags": r.flags, "sequence": r.sequence, "rssi": r.rssi, "pos": r.pos}
            d["exec"] = format_exec(d)
            out.append(d)
        return {"status": f.status, "byte_cnt": f.byte_cnt, "n_records": f.n_records}, out

    def stats(self):
        s = Stats()
        _check(self.L.tfr_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def run(self, iq_by_stream):
        """submit one array per stream, process, return (frames, records)."""
        for s, iq in enumerate(iq_by_stream):
            if iq is not None:
                self.submit(s, iq)
        self.process()
        return self.frames(), self.records()


def downconvert(iq: np.ndarray, passes=2, filter=0, device=0):
    """downconvert(passes)::process_iq (dsp_stuff.cpp:232-264): u8 IQ -> int16 I,Q at 1.536 MS/s / 2^passes"""
    iq = np.ascontiguousarray(iq, dtype=np.uint8)
    out = np.empty(((iq.size // 2) >> passes) * 2, dtype=np.int16)
    n = _check(load().tfr_downconvert(device, iq.ctypes.data, iq.size, passes, filter, out.ctypes.data, MEM_HOST, 1, None))
    return out[:n]


class Downconvert:
    """`downconvert(int p)` of dsp_stuff.h:46-56 as a streaming object: every stage's history is carried between calls"""

    def __init__(self, passes=2, device=0):
        self.passes = passes
        self._h = C.c_void_p()
        _check(load().tfr_dc_create(device, passes, C.byref(self._h)))

    def process(self, iq: np.ndarray, filter=0):
        """raw u8 IQ (a multiple of 2^passes pairs) -> int16 I,Q"""
        iq = np.ascontiguousarray(iq, dtype=np.uint8)
        out = np.empty(((iq.size // 2) >> self.passes) * 2, dtype=np.int16)
        n = _check(load().tfr_dc_process(self._h, iq.ctypes.data, iq.size, filter, out.ctypes.data, MEM_HOST))
        return out[:n]

    def process_iq(self, data_iq: np.ndarray, filter=0):
        """process_iq(int16_t *buf, int len, int filter) itself: int16 I,Q in place; returns the new len"""
        assert data_iq.dtype == np.int16 and data_iq.flags.c_contiguous
        return _check(load().tfr_dc_process_i16(self._h, data_iq.ctypes.data, data_iq.size, filter))

    def close(self):
        if self._h:
            load().tfr_dc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def downconvert_device(iq_ptr, nbytes, out_ptr, passes=2, filter=0, device=0, reps=1):
    """same with device pointers; returns (int16 written, CUDA-event ms of one cascade)"""
    ms = C.c_float(0)
    n = _check(load().tfr_downconvert(device, iq_ptr, nbytes, passes, filter, out_ptr, MEM_DEVICE, reps, C.byref(ms)))
    return n, float(ms.value)


def decimate(iq: np.ndarray, filter=0, device=0):
    iq = np.ascontiguousarray(iq, dtype=np.uint8)
    out = np.empty((iq.size // 8) * 2, dtype=np.int16)
    n = _check(load().tfr_decimate(device, iq.ctypes.data, iq.size, filter, out.ctypes.data, MEM_HOST))
    return out[:n]
