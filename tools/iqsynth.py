"""Seeded synthetic rtl-sdr IQ generator (test + bench infrastructure, not part of the decode path).

Produces unsigned 8-bit offset-binary interleaved I,Q at 1.536 MS/s -- the on-disk format the
reference replays with -L (engine.cpp:67-81, sdr.cpp:233-237) -- containing telegrams for the five
decoders the reference registers (main.cpp:171-218):

  TFA_1  NRZS 38400 Bd, LSB first, sync 2d d4            (tfa1.cpp:7-31)
  TFA_2  NRZ  17240 Bd, MSB first, sync 2d d4            (tfa2.cpp:6-33)
  TFA_3  NRZ   9600 Bd                                    (tfa2.cpp:12-14)
  TX22   NRZ   8842 Bd                                    (tfa2.cpp:35-49)
  WHB    BPSK/NRZS/G3RUH 6000 Bd, sync 4b 2d d4 2b, CRC32 (whb.cpp:10-46)

Everything after the random draw is integer arithmetic (integer phase accumulator, rounded cosine
table, integer box filter) so that the same seed gives the same bytes on every machine; tests pin
the result with a sha256 stored next to the golden output.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field

import numpy as np

FS = 1536000
BLOCK_BYTES = 65536  # reference replay framing, engine.cpp:68

TFA_1, TFA_2, TFA_3, TX22, TFA_WHB = 0, 1, 2, 3, 5  # sensor_e, decoder.h:11-19
BAUD = {TFA_1: 38400, TFA_2: 17240, TFA_3: 9600, TX22: 8842, TFA_WHB: 6000}

_PH_BITS = 16
_COS = None


def _cos_table():
    global _COS
    if _COS is None:
        k = np.arange(1 << _PH_BITS, dtype=np.float64)
        # rounded to 1/2^20 so the (amp*table) product below is exact integer arithmetic
        _COS = np.rint(np.cos(2.0 * np.pi * k / (1 << _PH_BITS)) * (1 << 20)).astype(np.int64)
    return _COS


# ----------------------------------------------------------------------------- CRCs (frame builders)
def crc8(data, poly=0x31):
    """MSB-first CRC-8, init 0 (crc8.cpp:4-29)."""
    c = 0
    for b in data:
        c ^= b
        for _ in range(8):
            c = ((c << 1) ^ poly) & 0xFF if c & 0x80 else (c << 1) & 0xFF
    return c


def crc32(data, init, poly=0x04C11DB7):
    """MSB-first CRC-32, caller-supplied init, no reflection/xorout (crc32.cpp:4-30)."""
    c = init & 0xFFFFFFFF
    for b in data:
        c ^= b << 24
        for _ in range(8):
            c = ((c << 1) ^ poly) & 0xFFFFFFFF if c & 0x80000000 else (c << 1) & 0xFFFFFFFF
    return c


WHB_CRC_INIT = {0x02: 0x97D97A26, 0x03: 0xF59C5A1E, 0x04: 0x98E1D11F, 0x06: 0xA7A41254,
                0x07: 0x3303FB1D, 0x08: 0x29F0F49B, 0x09: 0xA7A41254, 0x0B: 0xE7720AE4,
                0x10: 0x62D0AFC1, 0x11: 0x8CBA0708, 0x12: 0x5A9E30AE}  # whb.cpp:50-62


def bcd3(v):
    return (v // 100) % 10, (v // 10) % 10, v % 10


def frame_tfa1(sensor_id, temp_c, hum, seq, lowbat=0):
    """2d d4 ID ID sT TT HH BB SS 56 CC (tfa1.cpp:17-31)."""
    h, t, u = bcd3(int(round((temp_c + 40.0) * 10)))
    body = [(sensor_id >> 8) & 0xFF, sensor_id & 0xFF, 0x80 | h, (t << 4) | u, hum & 0xFF,
            0x60 | ((lowbat & 1) << 7), (seq & 0xF) << 4, 0x56]
    return bytes([0x2D, 0xD4] + body + [crc8(body)])


def frame_tfa2(sensor_id, temp_c, hum):
    """2d d4 II IT TT HH CC (tfa2.cpp:19-33); sensor_id is the byte that ends up in id bits 15:8."""
    h, t, u = bcd3(int(round((temp_c + 40.0) * 10)))
    body = [0x90 | ((sensor_id >> 4) & 0xF), ((sensor_id & 0xF) << 4) | h, (t << 4) | u, hum & 0xFF]
    return bytes([0x2D, 0xD4] + body + [crc8(body)])


def frame_tx22(sensor_id, temp_c=None, hum=None):
    """2d d4 SI IQ TV VV [TV VV] CC (tfa2.cpp:38-49)."""
    words = []
    if temp_c is not None:
        h, t, u = bcd3(int(round((temp_c + 40.0) * 10)))
        words.append((0x0 << 12) | (h << 8) | (t << 4) | u)
    if hum is not None:
        h, t, u = bcd3(int(hum))
        words.append((0x1 << 12) | (h << 8) | (t << 4) | u)
    body = [0xA0 | ((sensor_id >> 2) & 0xF), ((sensor_id & 3) << 6) | 0x10 | len(words)]
    for w in words:
        body += [w >> 8, w & 0xFF]
    return bytes([0x2D, 0xD4] + body + [crc8(body)])


def frame_whb03(sensor_id48, seq, temp_c, hum, ptemp_c=None, phum=None):
    """4b 2d d4 2b LL II*6 payload CC*4, type 03 = temp/hum (whb.cpp:16-24, 143-170)."""
    def t11(v):
        r = int(round(v * 10))
        return r & 0x7FF
    ptemp_c = temp_c if ptemp_c is None else ptemp_c
    phum = hum if phum is None else phum
    ident = [(sensor_id48 >> (8 * i)) & 0xFF for i in range(5, -1, -1)]
    assert ident[0] == 0x03
    payload = [(seq >> 8) & 0x3F, seq & 0xFF, t11(temp_c) >> 8, t11(temp_c) & 0xFF, 0, hum & 0xFF,
               t11(ptemp_c) >> 8, t11(ptemp_c) & 0xFF, 0, phum & 0xFF, 0]
    plen = 4 + 1 + 6 + len(payload)  # offset of the CRC, counted from the first sync byte (whb.cpp:513-514)
    body = [plen] + ident + payload
    c = crc32(body, WHB_CRC_INIT[0x03])
    return bytes([0x4B, 0x2D, 0xD4, 0x2B] + body + [(c >> 24) & 0xFF, (c >> 16) & 0xFF, (c >> 8) & 0xFF, c & 0xFF])


def frame_tx22_words(sensor_id, words, ok=1, lowbat=0):
    """TX22 frame from raw 16-bit words (type nibble | 12-bit value): every word type of tfa2.cpp:103-151
    (0 temp, 1 hum, 2 rain, 3 wind dir/speed, 4 gust, others ignored), num = len(words) <= 7 (rdata[3]&7)."""
    assert len(words) <= 7
    body = [0xA0 | ((sensor_id >> 2) & 0xF), ((sensor_id & 3) << 6) | ((ok & 1) << 4) | ((lowbat & 1) << 3) | len(words)]
    for w in words:
        body += [(w >> 8) & 0xFF, w & 0xFF]
    return bytes([0x2D, 0xD4] + body + [crc8(body)])


# payload bytes each WeatherHub parser reads (whb.cpp:126-475): decode_02 .. decode_12
WHB_PAYLOAD_LEN = {0x02: 6, 0x03: 11, 0x04: 12, 0x06: 14, 0x07: 18, 0x08: 26, 0x09: 14, 0x0B: 27,
                   0x10: 10, 0x11: 34, 0x12: 9}


def frame_whb(stype, ident40, payload):
    """4b 2d d4 2b LL TT II*5 payload CC*4 for any WeatherHub type of whb.cpp:50-62: LL is the offset of the
    CRC-32 counted from the first sync byte (whb.cpp:497,513-514), the CRC covers LL.. with the type's init."""
    ident = [stype] + [(ident40 >> (8 * i)) & 0xFF for i in range(4, -1, -1)]
    plen = 4 + 1 + 6 + len(payload)
    body = [plen] + ident + list(payload)
    c = crc32(body, WHB_CRC_INIT[stype])
    return bytes([0x4B, 0x2D, 0xD4, 0x2B] + body + [(c >> 24) & 0xFF, (c >> 16) & 0xFF, (c >> 8) & 0xFF, c & 0xFF])


# ----------------------------------------------------------------------------- bit streams
def bits_lsb(data):
    return [(b >> i) & 1 for b in data for i in range(8)]


def bits_msb(data):
    return [(b >> (7 - i)) & 1 for b in data for i in range(8)]


def g3ruh_scramble(bits):
    """s[n] = d[n] ^ s[n-12] ^ s[n-17]; the inverse of the descrambler in whb.cpp:578-580."""
    s = []
    for n, d in enumerate(bits):
        a = s[n - 12] if n >= 12 else 0
        b = s[n - 17] if n >= 17 else 0
        s.append(d ^ a ^ b)
    return s


# ----------------------------------------------------------------------------- modulators (integer)
def fsk_iq(levels, baud, fd_hz, amp, f_off_hz=0.0):
    """Rectangular CPFSK: `levels` is a list of +1/-1 per bit; returns int64 I, Q (LSB units)."""
    nbits = len(levels)
    nsamp = (nbits * FS) // baud
    k = np.arange(nsamp, dtype=np.int64)
    bit = np.minimum((k * baud) // FS, nbits - 1)
    lv = np.asarray(levels, dtype=np.int64)[bit]
    inc_dev = int(round(fd_hz / FS * (1 << 32)))
    inc_off = int(round(f_off_hz / FS * (1 << 32)))
    ph = np.cumsum(lv * inc_dev + inc_off) & 0xFFFFFFFF
    idx = (ph >> (32 - _PH_BITS)).astype(np.int64)
    ct = _cos_table()
    i = (amp * ct[idx] + (1 << 19)) >> 20
    q = (amp * ct[(idx - (1 << (_PH_BITS - 2))) & ((1 << _PH_BITS) - 1)] + (1 << 19)) >> 20
    return i, q


def burst_tfa1(frame, amp=100, fd_hz=47000, preamble=200, tail=32):
    """NRZS: a '0' bit toggles the frequency, a '1' keeps it (tfa1.cpp:9-11); LSB first."""
    bits = [0] * preamble + bits_lsb(frame) + [0] * tail
    lv, cur = [], 1
    for b in bits:
        if b == 0:
            cur = -cur
        lv.append(cur)
    return fsk_iq(lv, BAUD[TFA_1], fd_hz, amp)


def burst_nrz(frame, sensor, amp=100, fd_hz=30000, toggles=None, f_off_hz=0.0):
    """NRZ, '1' = +fd, MSB first, preamble of 1-0 toggles (tfa2.cpp:8-17), 4x 0-1 tail."""
    if toggles is None:
        toggles = 8 if sensor == TFA_2 else 12
    bits = [1, 0] * toggles + bits_msb(frame) + [0, 1] * 4
    lv = [1 if b else -1 for b in bits]
    return fsk_iq(lv, BAUD[sensor], fd_hz, amp, f_off_hz)


def burst_whb(frame, amp=100, preamble_bytes=16, tail_bytes=1, ramp=128):
    """BPSK: data LSB first -> G3RUH scramble -> carrier sign flips on every scrambled '0'
    (SURVEY.md A.4 identity for whb.cpp:566-581); sign sequence box-filtered over `ramp` samples."""
    data = bytes(preamble_bytes) + frame + bytes(tail_bytes)
    s = g3ruh_scramble(bits_lsb(data))
    sign, cur = [], 1
    for b in s:
        if b == 0:
            cur = -cur
        sign.append(cur)
    spb = FS // BAUD[TFA_WHB]
    x = np.repeat(np.asarray(sign, dtype=np.int64), spb)
    x = np.concatenate([np.full(ramp, x[0], dtype=np.int64), x, np.full(ramp, x[-1], dtype=np.int64)])
    c = np.concatenate([[0], np.cumsum(x)])
    box = c[ramp:] - c[:-ramp]  # sum over `ramp` samples, |box| <= ramp
    i = (2 * amp * box + ramp) // (2 * ramp)
    # fixed 45 degree carrier phase so that both I and Q are exercised
    i45 = (i * 46341 + (1 << 15)) >> 16
    return i45, i45.copy()


# ----------------------------------------------------------------------------- stream assembly
@dataclass
class Burst:
    at: int                   # raw sample index where the burst starts
    sensor: int               # sensor_e value
    frame: bytes
    kwargs: dict = field(default_factory=dict)


def render_burst(b: Burst):
    if b.sensor == TFA_1:
        return burst_tfa1(b.frame, **b.kwargs)
    if b.sensor == TFA_WHB:
        return burst_whb(b.frame, **b.kwargs)
    return burst_nrz(b.frame, b.sensor, **b.kwargs)


def make_stream(n_samples, bursts, seed, sigma=1.0):
    """u8 interleaved IQ of n_samples raw samples: rounded Gaussian noise + bursts, offset 128."""
    rng = np.random.default_rng(seed)
    iq = np.rint(rng.normal(0.0, sigma, size=2 * n_samples)).astype(np.int64) if sigma > 0 \
        else np.zeros(2 * n_samples, dtype=np.int64)
    for b in bursts:
        i, q = render_burst(b)
        n = min(len(i), n_samples - b.at)
        if n <= 0:
            continue
        iq[2 * b.at:2 * (b.at + n):2] += i[:n]
        iq[2 * b.at + 1:2 * (b.at + n) + 1:2] += q[:n]
    return np.clip(iq + 128, 0, 255).astype(np.uint8)


def sha256(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ----------------------------------------------------------------------------- named fixtures
KAT_TFA1 = bytes.fromhex("2dd465b086202360e05697")  # README.md:123 of the reference
KAT_TFA2 = bytes.fromhex("2dd49905722f6b")
KAT_TX22 = bytes.fromhex("2dd4a55206131055" "45")
KAT_WHB03 = bytes.fromhex("4b2dd42b" "16" "03123456789a" "012300d5003300d4003200" "b2715d44")


def fixture_single_tfa1(seed=1):
    """BASELINE.json configs[0]: one 30.3180.IT telegram in a 1 MiB buffer."""
    return make_stream(524288, [Burst(200000, TFA_1, KAT_TFA1)], seed=seed, sigma=1.0)


def fixture_mixed5(seed=42, sigma=1.0):
    """Five-type mixed stream: 6 MiB, one burst per decoder."""
    f3 = frame_tfa2(0x9A, -3.4, 81)
    bursts = [Burst(200000, TFA_1, KAT_TFA1), Burst(600000, TFA_2, KAT_TFA2), Burst(1000000, TFA_3, f3),
              Burst(1400000, TX22, KAT_TX22), Burst(1900000, TFA_WHB, KAT_WHB03)]
    return make_stream(3145728, bursts, seed=seed, sigma=sigma)


def random_frame(sensor, rng):
    """A random but valid telegram for `sensor` (passes the reference's CRC + sanity checks)."""
    if sensor == TFA_1:
        return frame_tfa1(int(rng.integers(1, 0x7FFF)), float(rng.integers(-200, 500)) / 10.0,
                          int(rng.integers(1, 100)), int(rng.integers(0, 16)), int(rng.integers(0, 2)))
    if sensor in (TFA_2, TFA_3):
        return frame_tfa2(int(rng.integers(0, 64)) << 2, float(rng.integers(-200, 500)) / 10.0,
                          int(rng.integers(1, 100)))
    if sensor == TX22:
        return frame_tx22(int(rng.integers(0, 64)), float(rng.integers(-200, 500)) / 10.0,
                          int(rng.integers(1, 100)))
    sid = (0x03 << 40) | int(rng.integers(0, 1 << 40))
    return frame_whb03(sid, int(rng.integers(0, 0x3FFF)), float(rng.integers(-200, 500)) / 10.0,
                       int(rng.integers(1, 100)))


def fixture_continuous(n_samples, sensors, period, seed, sigma=1.0, phase=0, amp=100):
    """Continuous stream with one telegram every `period` raw samples, sensor types rotated
    (BASELINE.json configs[1..3]); returns (u8 array, list of Burst)."""
    rng = np.random.default_rng(seed ^ 0x5EED)
    bursts, at, k = [], phase + period // 4, 0
    while at + 200000 < n_samples:
        s = sensors[k % len(sensors)]
        kw = {"amp": amp}
        bursts.append(Burst(at, s, random_frame(s, rng), kw))
        at += period
        k += 1
    return make_stream(n_samples, bursts, seed=seed, sigma=sigma), bursts
