"""Debug probe: the tensor-core front-end (frontend_tc.cu) against the shared-memory one (frontend.cu, TFR_FE=old) on the
same bytes - decimated samples of whole buffers, several sizes and contents.  Parity proper lives in tests/."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import tfrec_b200 as tb

def dec(iq, fe, filt=0):
    os.environ["TFR_FE"] = fe
    return tb.decimate(iq, filt)

rng = np.random.default_rng(5)
bad = 0
for name, iq in (("ramp", (np.arange(4 * 65536) % 256).astype(np.uint8)),
                 ("const128", np.full(2 * 65536, 128, np.uint8)),
                 ("rand", rng.integers(0, 256, 5 * 65536, dtype=np.uint8)),
                 ("extremes", rng.choice(np.array([0, 255], np.uint8), 3 * 65536)),
                 ("rand_big", rng.integers(0, 256, 64 * 65536, dtype=np.uint8))):
    for filt in (0, 1):
        a = dec(iq, "old", filt)
        b = dec(iq, "tc", filt)
        ok = np.array_equal(a, b)
        msg = ""
        if not ok:
            bad += 1
            d = np.nonzero(a != b)[0]
            msg = " first diff at int16 index %d (sample %d, thread %d, output %d): old %s tc %s; %d of %d differ" % (
                d[0], d[0] // 2, (d[0] // 2 % 8192) // 64, (d[0] // 2) % 64, a[d[0]:d[0] + 6], b[d[0]:d[0] + 6], d.size, a.size)
        print("%-9s filter %d: %s%s" % (name, filt, "equal" if ok else "DIFFERENT", msg), flush=True)
print("FAILED" if bad else "all equal")
sys.exit(1 if bad else 0)
