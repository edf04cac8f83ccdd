"""ctypes binding of oracle/liboracle.so -- the CPU restatement used as the parity checker.

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")


class Frame(C.Structure):
    _fields_ = [("type", C.c_int32), ("status", C.c_int32), ("pos", C.c_int64), ("byte_cnt", C.c_int32),
                ("rssi", C.c_int32), ("offset", C.c_int32), ("n_records", C.c_int32),
                ("rdata", C.c_uint8 * 64), ("line", C.c_char * 320)]


class Record(C.Structure):
    _fields_ = [("type", C.c_int32), ("alarm", C.c_int32), ("id", C.c_uint64), ("temp", C.c_double),
                ("humidity", C.c_double), ("flags", C.c_int32), ("sequence", C.c_int32), ("rssi", C.c_int32),
                ("frame", C.c_int32), ("pos", C.c_int64)]


class BlockTrace(C.Structure):
    _fields_ = [("thresh", C.c_int32), ("triggered", C.c_int32), ("triggered_avg", C.c_int32)]


_lib = None


def build():
    src = [os.path.join(ORACLE_DIR, f) for f in ("tfrec_oracle.c", "tfrec_oracle.h")]
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src):
        return
    subprocess.run(["make", "-C", ORACLE_DIR, "port"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_taps.argtypes = [C.c_void_p, C.c_int]
        L.orc_process.restype = C.c_long
        L.orc_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        for name, rt in (("orc_n_frames", C.c_size_t), ("orc_n_records", C.c_size_t), ("orc_n_blocks", C.c_size_t),
                         ("orc_frames", C.POINTER(Frame)), ("orc_records", C.POINTER(Record)),
                         ("orc_blocks", C.POINTER(BlockTrace)), ("orc_thresh", C.c_int),
                         ("orc_inverted_syncs", C.c_long)):
            getattr(L, name).restype = rt
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_clear_results.argtypes = [C.c_void_p]
        L.orc_n_tap.restype = C.c_size_t
        L.orc_n_tap.argtypes = [C.c_void_p, C.c_int]
        L.orc_tap.restype = C.c_void_p
        L.orc_tap.argtypes = [C.c_void_p, C.c_int]
        L.orc_tap_chan.restype = C.c_void_p
        L.orc_tap_chan.argtypes = [C.c_void_p, C.c_int]
        L.orc_decimate.restype = C.c_size_t
        L.orc_decimate.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.orc_downconvert.restype = C.c_size_t
        L.orc_downconvert.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        L.orc_fm_dev.argtypes = [C.c_int] * 4
        L.orc_fm_dev_nrzs.argtypes = [C.c_int] * 4
        L.orc_crc8.restype = C.c_uint8
        L.orc_crc8.argtypes = [C.c_char_p, C.c_int]
        L.orc_crc32.restype = C.c_uint32
        L.orc_crc32.argtypes = [C.c_char_p, C.c_int, C.c_uint32]
        L.orc_parse.argtypes = [C.c_int, C.c_char_p, C.c_int, C.POINTER(Frame), C.POINTER(Record), C.c_int]
        L.orc_format_exec.argtypes = [C.POINTER(Record), C.c_char_p, C.c_size_t]
        _lib = L
    return _lib


def frame_dict(f: Frame):
    n = min(f.byte_cnt, 64)
    return {"type": f.type, "status": f.status, "pos": f.pos, "byte_cnt": f.byte_cnt, "rssi": f.rssi,
            "offset": f.offset, "n_records": f.n_records, "rdata": bytes(f.rdata[:n]).hex(),
            "line": f.line.decode()}


def record_dict(r: Record):
    buf = C.create_string_buffer(256)
    lib().orc_format_exec(C.byref(r), buf, 256)
    return {"type": r.type, "id": r.id, "temp": r.temp, "humidity": r.humidity, "alarm": r.alarm,
            "flags": r.flags, "sequence": r.sequence, "rssi": r.rssi, "frame": r.frame, "pos": r.pos,
            "exec": buf.value.decode()}


class Oracle:
    """One receiver instance = engine + downconvert + fsk_demod + the registered demods."""

    def __init__(self, types=0x07, filter=0, thresh=0, taps=0):
        self.L = lib()
        self.h = self.L.orc_create(types, filter, thresh)
        if taps:
            self.L.orc_set_taps(self.h, taps)

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def process(self, iq: np.ndarray):
        iq = np.ascontiguousarray(iq, dtype=np.uint8)
        return self.L.orc_process(self.h, iq.ctypes.data, iq.size)

    def frames(self):
        n = self.L.orc_n_frames(self.h)
        p = self.L.orc_frames(self.h)
        return [frame_dict(p[i]) for i in range(n)]

    def records(self):
        n = self.L.orc_n_records(self.h)
        p = self.L.orc_records(self.h)
        return [record_dict(p[i]) for i in range(n)]

    def blocks(self):
        n = self.L.orc_n_blocks(self.h)
        p = self.L.orc_blocks(self.h)
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(n, 3)).copy() if n else np.zeros((0, 3), np.int32)
        return a

    def tap(self, kind):
        n = self.L.orc_n_tap(self.h, kind)
        p = self.L.orc_tap(self.h, kind)
        if n == 0:
            return np.zeros(0, dtype=np.float64 if kind == 2 else np.int32)
        ct = C.c_double if kind == 2 else C.c_int32
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n,)).copy()

    def tap_chan(self, kind):
        n = self.L.orc_n_tap(self.h, kind)
        p = self.L.orc_tap_chan(self.h, kind)
        if n == 0:
            return np.zeros(0, dtype=np.uint8)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n,)).copy()

    def thresh(self):
        return self.L.orc_thresh(self.h)

    def inverted_syncs(self):
        return self.L.orc_inverted_syncs(self.h)

    def clear(self):
        self.L.orc_clear_results(self.h)


def decimate(iq: np.ndarray, filter=0):
    iq = np.ascontiguousarray(iq, dtype=np.uint8)
    out = np.empty(iq.size // 4, dtype=np.int16)
    n = lib().orc_decimate(iq.ctypes.data, iq.size, filter, out.ctypes.data)
    return out[:n]


def downconvert(iq: np.ndarray, passes=2, filter=0):
    """downconvert(passes)::process_iq, dsp_stuff.cpp:232-264"""
    iq = np.ascontiguousarray(iq, dtype=np.uint8)
    out = np.empty(((iq.size // 2) >> passes) * 2 + 2, dtype=np.int16)
    n = lib().orc_downconvert(iq.ctypes.data, iq.size, passes, filter, out.ctypes.data)
    return out[:n]


def parse(sensor_type, data: bytes):
    f = Frame()
    recs = (Record * 8)()
    n = lib().orc_parse(sensor_type, data, len(data), C.byref(f), recs, 8)
    if n < 0:
        return None, []
    return frame_dict(f), [record_dict(recs[i]) for i in range(n)]


def crc8(data: bytes):
    return lib().orc_crc8(data, len(data))


def crc32(data: bytes, init):
    return lib().orc_crc32(data, len(data), init)
