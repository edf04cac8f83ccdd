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
    S = int(rng.choice([1, 1, 2, 3]))                       # streams in the handle, each with its own data and oracle
    sigma = float(rng.choice([0.6, 1.0, 2.0, 4.0, 8.0]))
    amp = int(rng.choice([12, 30, 60, 100]))
    types = int(rng.choice([0x01, 0x07, 0x0E, 0x2F, 0x21, 0x06]))
    filt = int(rng.integers(0, 2))
    thresh = int(rng.choice([0, 0, 0, 500, 300, 150, 60]))
    split = str(rng.choice(["0", "1", "2"]))
    in_flight = bool(rng.integers(0, 2))
    os.environ["TFR_BE_SPLIT"] = split
    os.environ["TFR_MIN_CHUNK"] = str(int(rng.choice([1, 4, 8192])))
    walk = str(rng.choice(["lists", "cta64", "cta256", "warp"]))   # the threshold walk's variants
    for key in ("TFR_WALK_TAB", "TFR_WALK", "TFR_WALK_CT"):
        os.environ.pop(key, None)
    os.environ.update({"lists": {"TFR_WALK_TAB": "0"}, "cta64": {"TFR_WALK_TAB": "1", "TFR_WALK_CT": "64"},
                       "cta256": {"TFR_WALK_TAB": "1", "TFR_WALK_CT": "256"}, "warp": {"TFR_WALK_TAB": "1", "TFR_WALK": "warp"}}[walk])
    streams = []
    for s_ in range(S):
        n_blocks = int(rng.integers(12, 72))
        sensors = [g.TFA_1, g.TFA_2, g.TFA_3, g.TX22, g.TFA_WHB]
        rng.shuffle(sensors)
        period = int(rng.integers(150000, 900000))
        iq, bursts = g.fixture_continuous(n_blocks * 32768, sensors[:int(rng.integers(1, 6))], period, seed=int(rng.integers(1, 1 << 30)),
                                          sigma=sigma, amp=amp)
        streams.append([iq, n_blocks, 0])
    desc = "case %d: streams %d blocks %s sigma %.1f amp %d types %#x filter %d thresh %d split %s in_flight %d min_chunk %s walk %s" % (
        k, S, [x[1] for x in streams], sigma, amp, types, filt, thresh, split, in_flight, os.environ["TFR_MIN_CHUNK"], walk)
    if os.environ.get("FUZZ_VERBOSE"):
        print("start", desc, flush=True)
    rx = tb.Receiver(types=types, filter=filt, thresh=thresh, n_streams=S)
    last_nb = [0] * S
    while any(x[2] < x[1] for x in streams):
        for s_, x in enumerate(streams):                    # ragged: a stream may sit a call out, lengths differ
            if x[2] < x[1] and rng.integers(0, 4) > 0:
                nb = int(min(x[1] - x[2], rng.integers(1, 24)))
                rx.submit(s_, x[0][x[2] * 65536:(x[2] + nb) * 65536].copy())
                x[2] += nb
                last_nb[s_] = nb
            else:
                last_nb[s_] = 0 if x[2] < x[1] or last_nb[s_] == 0 else 0
        rx.process()
        if not in_flight:
            rx.sync()
    ok = True
    frames, records = rx.frames(), rx.records()
    nf = nr = 0
    for s_, x in enumerate(streams):
        o = ol.Oracle(types=types, filter=filt, thresh=thresh)
        o.process(x[0])
        if [frame_key(f) for f in frames if f["stream"] == s_] != [frame_key(f) for f in o.frames()]:
            ok = False
            print("FRAMES differ (stream %d):" % s_, desc)
        if [r["exec"] for r in records if r["stream"] == s_] != [r["exec"] for r in o.records()]:
            ok = False
            print("RECORDS differ (stream %d):" % s_, desc)
        if rx.inverted_syncs(s_) != o.inverted_syncs():
            ok = False
            print("INVERTED SYNC count differs (stream %d: %d / %d):" % (s_, rx.inverted_syncs(s_), o.inverted_syncs()), desc)
        if rx.thresh(s_) != o.thresh():
            ok = False
            print("THRESHOLD differs (stream %d):" % s_, desc)
        tr = rx.block_trace(s_)                              # of the last call (empty if the stream sat it out)
        if len(tr) and not np.array_equal(tr, o.blocks()[-len(tr):]):
            ok = False
            print("TRACE differs (stream %d):" % s_, desc)
        nf += len(o.frames())
        nr += len(o.records())
        o.close()
    st = rx.stats()
    rx.close()
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
