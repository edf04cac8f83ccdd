"""CPU suite: the exactness argument of fm_dev_fast (tfrec_b200/csrc/demod_dev.cuh) checked on the host.

The kernel replaces atan2 by a polynomial and sends a sample to the exact path when the scaled angle is within
1e-6 of a truncation boundary.  This test reads the polynomial's coefficients out of the CUDA source, restates the
fast path in numpy (double arithmetic, same operation order up to FMA contraction, which only moves results by
~1e-16) and checks, against the reference formula of dsp_stuff.cpp:284-292 with libm's atan2,
  * that the polynomial's error is below 1e-7 output units (10x inside the 1e-6 guard band), and
  * that every sample the fast path keeps gets exactly the reference integer.
The device code itself is covered by the GPU parity tests (hashes of every fm_dev value of the golden streams).
"""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K = 5215.189175235227   # fl(16384/pi), as built


def _coefficients():
    src = open(os.path.join(ROOT, "tfrec_b200", "csrc", "demod_dev.cuh")).read()
    body = src[src.index("int fm_dev_fast("):]
    body = body[:body.index("double a = q * p;")]
    first = re.search(r"double p = (-?0x[0-9a-f.]+p[+-]?\d+);", body).group(1)
    rest = re.findall(r"p = fma\(p, z, (-?0x[0-9a-f.]+p[+-]?\d+)\);", body)
    c = [float.fromhex(first)] + [float.fromhex(x) for x in rest]   # highest power first
    assert len(c) == 13
    return c


def _poly_atan(q, c):
    z = q * q
    p = np.full_like(z, c[0])
    for v in c[1:]:
        p = p * z + v
    return q * p


def test_polynomial_error_is_inside_the_guard_band():
    c = _coefficients()
    q = np.linspace(0.0, 1.0, 1_000_001)
    err = np.max(np.abs(_poly_atan(q, c) - np.arctan(q))) * K
    assert err < 1e-7, err


def test_fast_path_gives_the_reference_integer():
    c = _coefficients()
    rng = np.random.default_rng(11)
    kept = 0
    for scale in (50, 200, 2000, 12200):   # |I|,|Q| of noise ... full-scale decimated samples
        n = 1_000_000
        ar, aj, br, bj = [rng.integers(-scale, scale + 1, n).astype(np.int64) for _ in range(4)]
        cr = aj * bj + ar * br
        cj = br * aj - ar * bj
        ax, ay = np.abs(cr), np.abs(cj)
        special = (cj == 0) | (cr == 0) | (ax == ay)
        mx = np.maximum(ax, ay).astype(np.float64)
        mn = np.minimum(ax, ay).astype(np.float64)
        mx[special] = 1.0
        a = _poly_atan(mn / mx, c)
        a = np.where(ay > ax, np.pi / 2 - a, a)
        a = np.where(cr < 0, np.pi - a, a)
        v = a * K
        fl = np.floor(v)
        fr = v - fl
        slow = special | (fr < 1e-6) | (fr > 1.0 - 1e-6)
        ref = np.trunc(np.arctan2(cj.astype(np.float64), cr.astype(np.float64)) * K).astype(np.int64)
        mine = np.where(cj < 0, -fl, fl)
        mine = np.where(slow, 0.0, mine).astype(np.int64)   # (special cases produce inf/nan here; they are not kept)
        assert np.array_equal(mine[~slow], ref[~slow])
        assert slow.mean() < 0.02      # the exact path stays rare
        kept += int((~slow).sum())
    assert kept > 3_900_000
