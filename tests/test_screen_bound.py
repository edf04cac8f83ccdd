"""The bound the screening front-end rests on (tfrec_b200/csrc/frontend_screen.cu), checked on the CPU against the oracle's
exact decimator: for EVERY decimated sample   |I| + |Q|  <=  (|vI| + |vQ|) / 2^shift + slack,   v = the screen values the
tensor core computes.  A sample the screen discards (|vI|+|vQ| <= (thresh_lo - slack) << shift) can therefore not trigger."""
import numpy as np
import pytest

import oracle_lib as ol
import screen_ref as sr


def inputs():
    rng = np.random.default_rng(7)
    n = 1 << 18
    yield "noise4", np.clip(np.rint(rng.normal(127.5, 4.0, n)), 0, 255).astype(np.uint8)
    yield "noise40", np.clip(np.rint(rng.normal(127.5, 40.0, n)), 0, 255).astype(np.uint8)
    yield "uniform", rng.integers(0, 256, n, dtype=np.uint8)
    yield "extremes", rng.choice(np.array([0, 255], dtype=np.uint8), n)
    t = np.arange(n // 2)
    tone = 127.5 + 100 * np.exp(2j * np.pi * 0.01 * t)
    iq = np.empty(n, dtype=np.uint8)
    iq[0::2] = np.clip(np.rint(tone.real), 0, 255)
    iq[1::2] = np.clip(np.rint(tone.imag), 0, 255)
    yield "tone", iq
    # adversarial for the floors: bytes whose tap products sit just above / below integers
    yield "ramp", (np.arange(n) % 256).astype(np.uint8)
    yield "const0", np.zeros(n, dtype=np.uint8)
    yield "const255", np.full(n, 255, dtype=np.uint8)


@pytest.mark.parametrize("filt", [0, 1])
def test_screen_bound_holds_for_every_sample(filt):
    worst = -10**9
    for name, iq in inputs():
        v, c = sr.screen_values(iq, filt)
        dec = ol.decimate(iq, filt).astype(np.int64).reshape(-1, 2)
        pwr = np.abs(dec[:, 0]) + np.abs(dec[:, 1])
        bound = (np.abs(v[:, 0]) + np.abs(v[:, 1])) / float(1 << c["shift"]) + c["slack"]
        assert np.all(pwr <= bound), (name, filt, int(np.argmax(pwr - bound)))
        # and per channel: the screen value is the exact sample up to half the slack
        for ch in range(2):
            d = np.abs(dec[:, ch] - v[:, ch] / float(1 << c["shift"]))
            worst = max(worst, float(d.max()))
            assert np.all(d <= c["slack"] / 2.0), (name, filt, ch, float(d.max()))
    assert worst > 5.0   # the bound is not vacuous: the floors do move samples by a good part of it


def test_screen_constants():
    for filt, (k, shift) in {0: (14, 12), 1: (15, 11)}.items():
        c = sr.consts(filt)
        assert (c["k"], c["shift"]) == (k, shift)
        assert np.abs(c["T16"]).max() <= 32767 - 128
        assert 30 <= c["slack"] <= 36
