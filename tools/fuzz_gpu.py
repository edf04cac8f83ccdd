"""Randomised differential run of the CUDA path against the oracle (test infrastructure; needs a GPU).

Every case draws a stream (noise level, telegram mix and spacing, amplitudes), a decoder mask, filter, threshold (auto or
fixed, also inside the noise), a call pattern (block counts per submit, synchronised or in flight) and the back-end split
mode, decodes it through the C ABI and compares frames, records, "Inverted SYNC" count, per-block threshold trace of the
last call and the final threshold with the oracle.   python tools/fuzz_gpu.py [cases] [seed]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np

import iqsynth as g
import oracle_lib as ol


def frame_key(f):
    return (f["type"], f["status"], f["pos"], f["byte_cnt"], f["rssi"], f["offset"], f["n_records"], f["rdata"])


def one_case(tb, rng, k):
    n_blocks = int(rng.integers(12, 72))
    n = n_blocks * 32768
    sigma = float(rng.choice([0.6, 1.0, 2.0, 4.0, 8.0]))
    sensors = [g.TFA_1, g.TFA_2, g.TFA_3, g.TX22, g.TFA_WHB]
    rng.shuffle(sensors)
    period = int(rng.integers(150000, 900000))
    amp = int(rng.choice([12, 30, 60, 100]))
    iq, bursts = g.fixture_continuous(n, sensors[:int(rng.integers(1, 6))], period, seed=int(rng.integers(1, 1 << 30)), sigma=sigma, amp=amp)
    types = int(rng.choice([0x01, 0x07, 0x0E, 0x2F, 0x21, 0x06]))
    filt = int(rng.integers(0, 2))
    thresh = int(rng.choice([0, 0, 0, 500, 300, 150, 60]))
    split = str(rng.choice(["0", "1", "2"]))
    in_flight = bool(rng.integers(0, 2))
    os.environ["TFR_BE_SPLIT"] = split
    os.environ["TFR_MIN_CHUNK"] = str(int(rng.choice([1, 4, 8192])))
    desc = "case %d: blocks %d sigma %.1f amp %d types %#x filter %d thresh %d split %s in_flight %d bursts %d min_chunk %s" % (
        k, n_blocks, sigma, amp, types, filt, thresh, split, in_flight, len(bursts), os.environ["TFR_MIN_CHUNK"])
    if os.environ.get("FUZZ_VERBOSE"):
        print("start", desc, flush=True)
    rx = tb.Receiver(types=types, filter=filt, thresh=thresh)
    off = 0
    while off < n_blocks:
        nb = int(min(n_blocks - off, rng.integers(1, 24)))
        rx.submit(0, iq[off * 65536:(off + nb) * 65536].copy())
        rx.process()
        if not in_flight:
            rx.sync()
        off += nb
    o = ol.Oracle(types=types, filter=filt, thresh=thresh)
    o.process(iq)
    ok = True
    if [frame_key(f) for f in rx.frames()] != [frame_key(f) for f in o.frames()]:
        ok = False
        print("FRAMES differ:", desc)
    if [r["exec"] for r in rx.records()] != [r["exec"] for r in o.records()]:
        ok = False
        print("RECORDS differ:", desc)
    if rx.inverted_syncs() != o.inverted_syncs():
        ok = False
        print("INVERTED SYNC count differs (%d / %d):" % (rx.inverted_syncs(), o.inverted_syncs()), desc)
    tr = rx.block_trace(0)
    if not np.array_equal(tr, o.blocks()[-len(tr):]) or rx.thresh(0) != o.thresh():
        ok = False
        print("TRACE / threshold differ:", desc)
    st = rx.stats()
    nf, nr = len(o.frames()), len(o.records())
    rx.close()
    o.close()
    print("%s  %s  frames %d records %d screen/dense %d/%d" % ("ok  " if ok else "FAIL", desc, nf, nr, st["screen_blocks"], st["dense_blocks"]),
          flush=True)
    return ok


def main():
    import tfrec_b200 as tb
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    bad = 0
    for k in range(cases):
        bad += 0 if one_case(tb, rng, k) else 1
    print("fuzz: %d cases, %d failed" % (cases, bad))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
