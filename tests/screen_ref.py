"""numpy restatement of the screening front-end's constants and screen values (tfrec_b200/csrc/frontend_screen.cu:
screen_build_consts and the tensor-core GEMM), for the tests.  Test infrastructure only.

The reference decimator (dsp_stuff.cpp:172-264) floors every tap product; the screen works with the plain linear filter
T = t1 (*) upsampled t2 (46 taps), rounded to 16 bits, and a bound on what the floors can change."""
import math

import numpy as np

T2 = np.array([2443, 6339, 11036, 14254, 14254, 11036, 6339, 2443], dtype=np.int64)
T1 = {0: np.array([-1087, -1082, -1065, -451, 912, 2997, 5556, 8157, 10285, 11484, 11484, 10285, 8157, 5556, 2997, 912, -451,
                   -1065, -1082, -1087], dtype=np.int64),
      1: np.array([546, 451, -317, -1844, -3198, -2817, 494, 6469, 13074, 17421, 17421, 13074, 6469, 494, -2817, -3198, -1844,
                   -317, 451, 546], dtype=np.int64)}


def consts(filt):
    t1 = T1[filt]
    T = np.zeros(46, dtype=np.int64)
    for k in range(20):
        for n in range(8):
            T[2 * k + n] += t1[k] * T2[n]
    tmax = int(np.abs(T).max())
    k = 0
    while ((tmax + (1 << k) // 2) >> k) > 32767 - 128:
        k += 1
    q = 26 - k
    T16 = np.floor(T / float(1 << k) + 0.5).astype(np.int64)
    err = int(np.abs(T - T16 * (1 << k)).sum())
    P1 = t1[t1 > 0].sum() / 65536.0
    N1 = -t1[t1 < 0].sum() / 65536.0
    e_lo, e_hi = -8.0 * N1, 8.0 * P1 + 20.0
    centre, half = 0.5 * (e_lo + e_hi), 0.5 * (e_hi - e_lo)
    round_err = 128.0 * err / 67108864.0
    centre_q = int(math.floor(centre * (1 << q) + 0.5))
    slack = int(math.ceil(2.0 * (half + round_err + 1.0 / (1 << q)))) + 1
    return {"T": T, "T16": T16, "k": k, "shift": q, "slack": slack, "centre_q": centre_q}


def screen_values(iq_u8, filt, hist=None):
    """[n_out, 2] int64: screen value of I and Q for every decimated sample of the stream (history: byte 128 = no signal)"""
    c = consts(filt)
    b = iq_u8.astype(np.int64) - 128
    pre = np.zeros(96, dtype=np.int64) if hist is None else hist.astype(np.int64) - 128
    b = np.concatenate([pre, b])                         # raw sample -48 .. ; I at even, Q at odd bytes
    n_out = iq_u8.size // 8
    out = np.empty((n_out, 2), dtype=np.int64)
    for ch in range(2):
        x = b[ch::2]                                     # x[48 + s] = raw sample s
        acc = np.zeros(n_out, dtype=np.int64)
        for i in range(46):                              # y[j] uses raw samples 4j-42+i
            acc += c["T16"][i] * x[48 - 42 + i: 48 - 42 + i + 4 * n_out: 4]
        out[:, ch] = acc - c["centre_q"]
    return out, c
