"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the oracle on the same seeded
inputs and against the golden vectors the unmodified reference produced.  Integer/byte results are
compared bit-exactly; the only floating-point intermediates (iir2::step outputs) are compared bitwise
too and, failing that, within 1e-5 relative as BASELINE.json's north_star allows."""
import hashlib

import numpy as np
import pytest

import iqsynth as g
import make_golden
import oracle_lib as ol

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def tb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import tfrec_b200
    tfrec_b200.load()
    return tfrec_b200


def frame_key(f):
    return (f["type"], f["status"], f["pos"], f["byte_cnt"], f["rssi"], f["offset"], f["n_records"], f["rdata"])


def record_key(r):
    return (r["type"], r["id"], r["temp"], r["humidity"], r["alarm"], r["sequence"], r["rssi"], r["pos"])


def compare_with_oracle(rx, iq, types, filt, thresh, stream=0, check_taps=True):
    o = ol.Oracle(types=types, filter=filt, thresh=thresh, taps=7 if check_taps else 0)
    o.process(iq)
    gf = [f for f in rx.frames() if f["stream"] == stream]
    gr = [r for r in rx.records() if r["stream"] == stream]
    of, orr = o.frames(), o.records()
    assert [frame_key(f) for f in gf] == [frame_key(f) for f in of]
    assert [record_key(r) for r in gr] == [record_key(r) for r in orr]
    assert [r["exec"] for r in gr] == [r["exec"] for r in orr]
    assert rx.inverted_syncs(stream) == o.inverted_syncs(), "'Inverted SYNC' lines (tfa2.cpp:294-300)"
    tr = rx.block_trace(stream)
    assert np.array_equal(tr, o.blocks()), "per-block threshold/trigger trace differs"
    assert rx.thresh(stream) == o.thresh()
    if check_taps:
        nd = bin(types & 0x2F).count("1")
        for kind in (0, 1, 2):
            ov, oc = o.tap(kind), o.tap_chan(kind)
            for d in range(nd):
                gv = rx.taps(stream, d, kind)
                ref = ov[oc == d]
                assert gv.size == ref.size, (kind, d, gv.size, ref.size)
                if kind < 2:
                    bad = int((gv != ref).sum())
                    assert bad == 0, "kind %d demod %d: %d of %d discriminator values differ" % (kind, d, bad, ref.size)
                else:
                    same = gv.view(np.uint64) == ref.view(np.uint64)
                    if not same.all():
                        rel = np.abs(gv - ref) / np.maximum(np.abs(ref), 1.0)
                        assert rel.max() < 1e-5, "iir2 outputs differ by %g relative" % rel.max()
                        assert (gv.astype(np.int64) == ref.astype(np.int64)).all()
    return o


def test_decimator_bit_exact(tb, golden):
    rng = np.random.default_rng(5)
    fixtures = {
        "uniform_bytes": rng.integers(0, 256, size=4 * 65536, dtype=np.uint8),
        "extremes": np.tile(np.array([0, 255, 255, 0, 0, 0, 255, 255], dtype=np.uint8), 65536 // 8 * 2),
        "single_tfa1": g.fixture_single_tfa1(seed=1),
    }
    for name, iq in fixtures.items():
        for filt, key in ((0, "narrow"), (1, "wide")):
            d = tb.decimate(iq, filt)
            ref = ol.decimate(iq, filt)
            assert d.size == ref.size
            assert np.array_equal(d, ref), "%s/%s: %d samples differ" % (name, key, int((d != ref).sum()))
            assert sha(d.astype("<i2")) == golden["decimator"][name][key]["sha256"]


@pytest.mark.parametrize("passes", [1, 2, 3, 4, 5])
def test_downconvert_streaming_object(tb, passes):
    """dsp_stuff.h:46-56 `downconvert(p)` as an object that is fed block by block: raw bytes through the fused kernel,
    int16 I,Q in place through process_iq's own signature - both against the oracle run over the whole stream"""
    rng = np.random.default_rng(70 + passes)
    iq = rng.integers(0, 256, size=2 * 37 * 4096, dtype=np.uint8)
    for filt in (0, 1):
        want = ol.downconvert(iq, passes, filt)
        cuts = [0, 2 * 32, 2 * 32 + 2 * 4096, 2 * 9 * 4096 + 2 * 64, 2 * 9 * 4096 + 2 * 96, 2 * 30 * 4096, iq.size]   # tiny, tile-sized, long
        dc = tb.Downconvert(passes)
        got = np.concatenate([dc.process(iq[a:b], filt) for a, b in zip(cuts[:-1], cuts[1:])])
        dc.close()
        assert np.array_equal(got, want), "u8 streaming, passes %d filter %d: %d differ" % (passes, filt, int((got != want).sum()))
        x = ((iq.astype(np.int16) - 128) << 6).astype(np.int16)            # engine.cpp:77-78
        dc = tb.Downconvert(passes)
        parts = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            buf = x[a:b].copy()
            n = dc.process_iq(buf, filt)
            parts.append(buf[:n].copy())
        dc.close()
        got = np.concatenate(parts)
        assert np.array_equal(got, want), "int16 streaming, passes %d filter %d" % (passes, filt)


@pytest.mark.parametrize("name,types", [("mixed5", 0x2F), ("strong_t7", 0x07)])
def test_decimated_int16_entry(tb, hot_fixture, name, types):
    """tfr_submit_decimated: the samples fsk_demod::process(int16_t *data_iq, int len) gets (fm_demod.cpp:34) - a caller
    that keeps its own decimator.  Fed with the reference decimator's output, in two calls, the decode is the oracle's"""
    iq = hot_fixture(name)
    nb = iq.size // 65536
    dec = ol.decimate(iq[:nb * 65536], 0)                 # int16 I,Q at 384 kS/s, 16384 int16 per block
    rx = tb.Receiver(types=types, thresh=0)
    cut = (nb // 3) * 16384
    for part in (dec[:cut], dec[cut:]):
        rx.submit_decimated(0, np.ascontiguousarray(part))
        rx.process()
    o = ol.Oracle(types=types, thresh=0)
    o.process(iq[:nb * 65536])
    assert [frame_key(f) for f in rx.frames()] == [frame_key(f) for f in o.frames()]
    assert [r["exec"] for r in rx.records()] == [r["exec"] for r in o.records()]
    tr = rx.block_trace(0)                                # of the last call
    assert np.array_equal(tr, o.blocks()[-len(tr):])
    assert rx.thresh(0) == o.thresh()
    rx.close()


def test_downconvert_cascade_variant(tb, monkeypatch):
    """TFR_DC=cascade: one launch per stage (decim.cu, what passes 6..8 always use) instead of the fused kernel"""
    monkeypatch.setenv("TFR_DC", "cascade")
    rng = np.random.default_rng(9)
    iq = rng.integers(0, 256, size=3 * 65536 + 4096 + 20, dtype=np.uint8)
    for passes in (1, 3, 5):
        assert np.array_equal(tb.downconvert(iq, passes, 0), ol.downconvert(iq, passes, 0))


def test_tensor_core_front_end_variant_bit_exact(tb, golden, hot_fixture, monkeypatch):
    """frontend_tc.cu (TFR_FE=tc: TMA tensor boxes + tcgen05.mma byte->float conversion + tensor-memory loads) is an
    opt-in variant of the front-end; it has to produce the reference's samples and records like the default kernel"""
    monkeypatch.setenv("TFR_FE", "tc")
    rng = np.random.default_rng(5)
    fixtures = {
        "uniform_bytes": rng.integers(0, 256, size=4 * 65536, dtype=np.uint8),
        "extremes": np.tile(np.array([0, 255, 255, 0, 0, 0, 255, 255], dtype=np.uint8), 65536 // 8 * 2),
        "single_tfa1": g.fixture_single_tfa1(seed=1),
    }
    for name, iq in fixtures.items():
        for filt, key in ((0, "narrow"), (1, "wide")):
            d = tb.decimate(iq, filt)
            assert np.array_equal(d, ol.decimate(iq, filt)), "%s/%s" % (name, key)
            assert sha(d.astype("<i2")) == golden["decimator"][name][key]["sha256"]
    # and a whole decode through it, in several submits (history carried between calls)
    iq = hot_fixture("mixed5")
    rx = tb.Receiver(types=0x2F, thresh=0)
    for off in range(0, iq.size, 7 * 65536):
        rx.submit(0, iq[off:off + 7 * 65536].copy())
        rx.process()
    o = ol.Oracle(types=0x2F)
    o.process(iq)
    assert [r["exec"] for r in rx.records()] == [r["exec"] for r in o.records()]
    assert rx.thresh(0) == o.thresh()
    rx.close()


def test_screen_values_are_the_linear_filter(tb, hot_fixture, monkeypatch):
    """Screening front-end (frontend_screen.cu): the tensor-core GEMM of the raw bytes with the 16-bit combined filter is
    an exact integer computation - every screen value equals the numpy restatement, across submits (carried history) and
    for both filters; the bound itself is checked against the oracle in tests/test_screen_bound.py"""
    import screen_ref as sr
    monkeypatch.setenv("TFR_NO_DENSE_MODE", "1")          # random bytes are all bursts: keep them going through the screen
    rng = np.random.default_rng(11)
    cases = {"mixed5": hot_fixture("mixed5")[:24 * 65536],
             "uniform": rng.integers(0, 256, size=6 * 65536, dtype=np.uint8),
             "extremes": np.tile(np.array([0, 255, 255, 0, 0, 0, 255, 255], dtype=np.uint8), 65536 // 8 * 3)}
    for name, iq in cases.items():
        for filt in (0, 1):
            rx = tb.Receiver(types=0x07, thresh=0, filter=filt, flags=tb.FLAG_TAPS)
            cut = (iq.size // 65536 // 2) * 65536
            want, c = sr.screen_values(iq, filt)
            got = []
            for part in (iq[:cut], iq[cut:]):
                rx.submit(0, part.copy())
                rx.process()
                v, shift, slack = rx.screen(0)
                assert (shift, slack) == (c["shift"], c["slack"])
                got.append(v)
            got = np.concatenate(got)
            assert got.shape == want.shape
            bad = np.nonzero((got != want).any(axis=1))[0]
            assert bad.size == 0, "%s filter %d: %d screen values differ, first at sample %d: got %s want %s" % (
                name, filt, bad.size, bad[0], got[bad[0]], want[bad[0]])
            st = rx.stats()
            assert st["screen_blocks"] + st["dense_blocks"] == iq.size // 65536
            rx.close()


def test_dense_front_end_variant(tb, hot_fixture, monkeypatch):
    """TFR_FE=dense: every block through the dense exact kernel of frontend.cu (what the screen hands bursts to)"""
    monkeypatch.setenv("TFR_FE", "dense")
    iq = hot_fixture("mixed5")
    rx = tb.Receiver(types=0x2F, thresh=0)
    for off in range(0, iq.size, 7 * 65536):
        rx.submit(0, iq[off:off + 7 * 65536].copy())
        rx.process()
    o = ol.Oracle(types=0x2F)
    o.process(iq)
    assert [r["exec"] for r in rx.records()] == [r["exec"] for r in o.records()]
    assert rx.thresh(0) == o.thresh()
    assert rx.stats()["screen_blocks"] == 0
    rx.close()


def test_decimator_ragged_length(tb):
    rng = np.random.default_rng(8)
    iq = rng.integers(0, 256, size=65536 + 4096 + 12, dtype=np.uint8)
    assert np.array_equal(tb.decimate(iq, 0), ol.decimate(iq, 0))


def test_downconvert_all_passes_bit_exact(tb, golden):
    """BASELINE configs[4]: downconvert(p), p = 1..5 (/2 .. /32), narrow and wide, against the oracle and the
    compiled reference's hashes"""
    rng = np.random.default_rng(5)
    fixtures = {
        "uniform_bytes": rng.integers(0, 256, size=4 * 65536, dtype=np.uint8),
        "extremes": np.tile(np.array([0, 255, 255, 0, 0, 0, 255, 255], dtype=np.uint8), 65536 // 8 * 2),
        "single_tfa1": g.fixture_single_tfa1(seed=1),
    }
    for name, iq in fixtures.items():
        for passes in (1, 2, 3, 4, 5):
            for filt, key in ((0, "narrow"), (1, "wide")):
                d = tb.downconvert(iq, passes, filt)
                ref = ol.downconvert(iq, passes, filt)
                assert d.size == ref.size == iq.size >> passes
                assert np.array_equal(d, ref), "%s p=%d %s: %d samples differ" % (name, passes, key, int((d != ref).sum()))
                assert sha(d.astype("<i2")) == golden["decimator"][name]["passes"][str(passes)][key]["sha256"]
        assert np.array_equal(tb.downconvert(iq, 2, 0), tb.decimate(iq, 0))   # the fused front-end is the same cascade


def test_downconvert_ragged_lengths_and_errors(tb):
    rng = np.random.default_rng(9)
    for size in (4, 6, 64, 4096 + 6, 65536 + 4096 + 12, 3 * 65536 + 2):
        iq = rng.integers(0, 256, size=size, dtype=np.uint8)
        for passes in (1, 2, 3, 5):
            assert np.array_equal(tb.downconvert(iq, passes, 1), ol.downconvert(iq, passes, 1)), (size, passes)
    with pytest.raises(tb.TfrError):
        tb.downconvert(np.zeros(64, np.uint8), 0, 0)
    with pytest.raises(tb.TfrError):
        tb.downconvert(np.zeros(64, np.uint8), 9, 0)


def test_downconvert_device_pointers_large(tb):
    """full-size property check (no oracle at this size): 2^26 raw samples on the device, /8; a causal filter
    cascade with zero history gives, for a prefix of the input, a prefix of the output"""
    import torch
    n = 1 << 27
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    iq = torch.randint(0, 256, (n,), device="cuda", dtype=torch.uint8, generator=gen)
    out = torch.empty(((n // 2) >> 3) * 2, device="cuda", dtype=torch.int16)
    cnt, ms = tb.downconvert_device(iq.data_ptr(), n, out.data_ptr(), passes=3, filter=0, reps=3)
    assert cnt == out.numel() and ms > 0
    head = tb.downconvert(iq[:1 << 20].cpu().numpy(), 3, 0)
    assert np.array_equal(out[:head.size].cpu().numpy(), head)
    assert np.array_equal(head, ol.downconvert(iq[:1 << 20].cpu().numpy(), 3, 0))


def test_parser_seam_against_reference(tb, golden):
    rx = tb.Receiver(types=0x2F, thresh=500)
    for k in golden["kat_frames"]:
        frame, recs = rx.parse_bytes(k["sensor"], bytes.fromhex(k["hex"]))
        of, orecs = ol.parse(k["sensor"], bytes.fromhex(k["hex"]))
        assert [r["exec"] for r in recs] == k["exec"], k
        if of is None:
            assert frame is None
        else:
            assert frame["status"] == of["status"] and frame["n_records"] == of["n_records"]
            assert [(r["id"], r["temp"], r["humidity"]) for r in recs] == [(r["id"], r["temp"], r["humidity"]) for r in orecs]
    rx.close()


def executed(lines):
    """the -e lines decoder::store_data lets through in mode 0: first appearance of an id always; a WeatherHub
    (13-digit id) repeat only when its sequence number differs from the last one stored for that id"""
    seen, out = {}, []
    for ln in lines:
        f = ln.split()
        if len(f[0]) == 13:
            if f[0] in seen and seen[f[0]] == f[3]:
                continue
            seen[f[0]] = f[3]
        out.append(ln)
    return out


CASES = [(name, label) for name, (_, cases) in make_golden.hotpath_fixtures().items() for (label, _, _) in cases]


@pytest.mark.parametrize("name,label", CASES)
def test_hot_path_against_oracle_and_reference(tb, golden, hot_fixture, name, label):
    c = golden["hotpath"][name]["cases"][label]
    iq = hot_fixture(name)
    assert sha(iq) == golden["hotpath"][name]["input_sha256"]
    kw = c["oracle"]
    rx = tb.Receiver(types=kw["types"], filter=kw["filter"], thresh=kw["thresh"], flags=tb.FLAG_TAPS)
    rx.submit(0, iq)
    rx.process()
    compare_with_oracle(rx, iq, kw["types"], kw["filter"], kw["thresh"])
    # and directly against what the unmodified reference's `-q -e /bin/echo` printed: every record in output
    # order, minus the WeatherHub repeats decoder::store_data does not execute (decoder.cpp:46-65)
    assert executed([r["exec"] for r in rx.records()]) == c["exec"]
    tr = rx.block_trace(0)
    assert sha(tr.astype("<i4")) == c["trace_sha256"]
    rx.close()


@pytest.mark.parametrize("name", ["mixed5", "cont_noisy"])
def test_filter_chains_ahead_of_the_slicers(tb, hot_fixture, monkeypatch, name):
    """TFR_BE=chains: the TFA_2-family low-pass runs ahead of the slicers as chains over four windows (biq_kernel),
    proven and repaired by biq_verify_kernel; the window kernels then only slice.  An opt-in variant of the back-end
    (slower on B200) that has to give the reference's results like the default: frames, records, block
    trace, threshold and every fm_dev / iir2::step value"""
    monkeypatch.setenv("TFR_BE", "chains")
    iq = hot_fixture(name)
    rx = tb.Receiver(types=0x2F, thresh=0, flags=tb.FLAG_TAPS)
    rx.submit(0, iq)
    rx.process()
    compare_with_oracle(rx, iq, 0x2F, 0, 0)
    rx.close()


def test_streaming_submits_carry_state(tb, hot_fixture):
    iq = hot_fixture("mixed5")
    rx = tb.Receiver(types=0x2F, thresh=0)
    for off in range(0, iq.size, 7 * 65536):
        rx.submit(0, iq[off:off + 7 * 65536].copy())
        rx.process()
    o = ol.Oracle(types=0x2F)
    o.process(iq)
    assert [frame_key(f) for f in rx.frames()] == [frame_key(f) for f in o.frames()]
    assert [r["exec"] for r in rx.records()] == [r["exec"] for r in o.records()]
    assert rx.thresh(0) == o.thresh()
    rx.close()


@pytest.mark.parametrize("parts", [1, 2, 4, 8])
def test_front_end_chunks_and_back_end_parts(tb, hot_fixture, monkeypatch, parts):
    """A large call goes through the front-end in up to 8 chunk launches and through the demodulators in parts that
    start while later chunks are still in the front-end.  TFR_MIN_CHUNK brings that machinery down to
    the size of the fixtures: whatever the cut, frames, records, block trace and threshold are the oracle's."""
    monkeypatch.setenv("TFR_MIN_CHUNK", "24")
    monkeypatch.setenv("TFR_BE_PARTS", str(parts))
    names = ["mixed5", "cont_noisy", "strong_t7"]
    iqs = [hot_fixture(n) for n in names]
    rx = tb.Receiver(types=0x0F, thresh=0, n_streams=len(iqs))
    nb = min(x.size for x in iqs) // 65536
    half = (nb // 2) * 65536
    for lo, hi in ((0, half), (half, nb * 65536)):          # two calls: state carried across them, parts in both
        for s, iq in enumerate(iqs):
            rx.submit(s, iq[lo:hi].copy())
        rx.process()
    frames, records = rx.frames(), rx.records()
    for s, iq in enumerate(iqs):
        o = ol.Oracle(types=0x0F)
        o.process(iq[:nb * 65536])
        assert [frame_key(f) for f in frames if f["stream"] == s] == [frame_key(f) for f in o.frames()], names[s]
        assert [r["exec"] for r in records if r["stream"] == s] == [r["exec"] for r in o.records()], names[s]
        assert rx.thresh(s) == o.thresh()
    rx.close()


@pytest.mark.parametrize("walk", ["cta64", "cta256", "warp", "lists"])
def test_threshold_walk_variants(tb, hot_fixture, monkeypatch, walk):
    """The threshold walk from the per-block table (walk_table_kernel) by a CTA per stream of two or eight warps
    (walk_cta_kernel), by a warp per stream (thresh2_kernel), and from the event lists: the same block trace, thresholds,
    windows (hence frames and records) as the oracle, over several launches per call (TFR_MIN_CHUNK), two calls, signals
    that push the threshold out of the table's range, and a start-up transient (noise far below the start threshold)."""
    env = {"cta64": {"TFR_WALK_TAB": "1", "TFR_WALK_CT": "64"}, "cta256": {"TFR_WALK_TAB": "1", "TFR_WALK_CT": "256"},
           "warp": {"TFR_WALK_TAB": "1", "TFR_WALK": "warp"}, "lists": {"TFR_WALK_TAB": "0"}}[walk]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    monkeypatch.setenv("TFR_MIN_CHUNK", "24")
    names = ["mixed5", "cont_noisy", "strong_t7"]
    iqs = [hot_fixture(n) for n in names] + [g.make_stream(6 * 1024 * 1024, [], seed=33, sigma=4.0)]
    rx = tb.Receiver(types=0x0F, thresh=0, n_streams=len(iqs))
    nb = min(x.size for x in iqs) // 65536
    half = (nb // 2) * 65536
    for lo, hi in ((0, half), (half, nb * 65536)):
        for s, iq in enumerate(iqs):
            rx.submit(s, iq[lo:hi].copy())
        rx.process()
    frames, records = rx.frames(), rx.records()
    for s, iq in enumerate(iqs):
        o = ol.Oracle(types=0x0F)
        o.process(iq[:half])
        o.process(iq[half:nb * 65536])
        assert np.array_equal(rx.block_trace(s), o.blocks()[half // 65536:]), (walk, s)
        assert [frame_key(f) for f in frames if f["stream"] == s] == [frame_key(f) for f in o.frames()], (walk, s)
        assert [r["exec"] for r in records if r["stream"] == s] == [r["exec"] for r in o.records()], (walk, s)
        assert rx.thresh(s) == o.thresh()
    rx.close()


@pytest.mark.parametrize("mode", ["0", "2"])
def test_split_back_end_across_calls(tb, hot_fixture, monkeypatch, mode):
    """TFR_BE_SPLIT=2: every call's decwin, fm_dev and the windows that do not need the previous call's final state run on
    the early stream beside the previous call's verifier; only window 0 and the windows whose warm-up reaches it wait
    (one warp each).  =0: the whole back-end call after call.  Several small calls in flight, all five decoders."""
    monkeypatch.setenv("TFR_BE_SPLIT", mode)
    monkeypatch.setenv("TFR_MIN_CHUNK", "4")
    for name, types in (("mixed5", 0x2F), ("cont_noisy", 0x07), ("strong_t7", 0x07)):
        iq = hot_fixture(name)
        rx = tb.Receiver(types=types, thresh=0)
        step = 5 * 65536
        for off in range(0, iq.size - iq.size % 65536, step):
            rx.submit(0, iq[off:off + step].copy())
            rx.process()                                   # no sync in between: calls in flight
        o = ol.Oracle(types=types, thresh=0)
        o.process(iq)
        assert [frame_key(f) for f in rx.frames()] == [frame_key(f) for f in o.frames()], (name, mode)
        assert [r["exec"] for r in rx.records()] == [r["exec"] for r in o.records()], (name, mode)
        assert rx.inverted_syncs() == o.inverted_syncs()
        assert rx.thresh(0) == o.thresh()
        rx.close()


def test_bursty_input_switches_to_the_dense_front_end(tb):
    """A fixed threshold inside the noise makes every block a burst: the screening front-end runs them in place (exact),
    reports it, and the following calls take the dense kernel (every 16th probes the screen again).  Results stay the
    oracle's across the switches."""
    rng = np.random.default_rng(23)
    iq = np.clip(np.rint(rng.normal(127.5, 6.0, 40 * 65536)), 0, 255).astype(np.uint8)
    rx = tb.Receiver(types=0x07, thresh=120)
    o = ol.Oracle(types=0x07, thresh=120)
    step = 2 * 65536
    for off in range(0, iq.size, step):
        rx.submit(0, iq[off:off + step].copy())
        rx.process()
        rx.sync()
    o.process(iq)
    assert [frame_key(f) for f in rx.frames()] == [frame_key(f) for f in o.frames()]
    assert [r["exec"] for r in rx.records()] == [r["exec"] for r in o.records()]
    assert rx.inverted_syncs() == o.inverted_syncs()
    st = rx.stats()
    # 20 calls of 2 blocks: the first one (and the probes, calls 17 ..) went through the screen as bursts, the rest dense
    assert 2 <= st["dense_blocks"] <= 8 and st["screen_blocks"] == 0, st
    rx.close()


def test_multi_stream_batch(tb, hot_fixture):
    names = ["single_tfa1", "mixed5", "strong_t7", "noise_only"]
    iqs = [hot_fixture(n) for n in names]
    rx = tb.Receiver(types=0x2F, thresh=0, n_streams=len(iqs))
    for s, iq in enumerate(iqs):
        rx.submit(s, iq)
    rx.process()
    frames, records = rx.frames(), rx.records()
    for s, iq in enumerate(iqs):
        o = ol.Oracle(types=0x2F)
        o.process(iq)
        assert [frame_key(f) for f in frames if f["stream"] == s] == [frame_key(f) for f in o.frames()]
        assert [r["exec"] for r in records if r["stream"] == s] == [r["exec"] for r in o.records()]
        assert np.array_equal(rx.block_trace(s), o.blocks())
    rx.close()


def test_device_pointer_submit(tb, hot_fixture):
    import torch
    iq = hot_fixture("mixed5")
    t = torch.from_numpy(iq).cuda()
    rx = tb.Receiver(types=0x07, thresh=500)
    rx.submit(0, t.data_ptr(), nbytes=t.numel())
    rx.process()
    o = ol.Oracle(types=0x07, thresh=500)
    o.process(iq)
    assert [r["exec"] for r in rx.records()] == [r["exec"] for r in o.records()]
    rx.close()


def test_submit_rejects_partial_blocks(tb):
    rx = tb.Receiver()
    with pytest.raises(tb.TfrError):
        rx.submit(0, np.zeros(1000, dtype=np.uint8))
    rx.submit(0, np.full(65536, 128, dtype=np.uint8))
    with pytest.raises(tb.TfrError):    # one submit per stream between process calls
        rx.submit(0, np.full(65536, 128, dtype=np.uint8))
    rx.process()
    assert rx.frames() == []
    rx.close()


# ---------------------------------------------------------------------------------------------------
# auto-threshold speculation, the two-slot pipeline and window chains (round-1 session 2 additions)
# ---------------------------------------------------------------------------------------------------
def test_auto_threshold_speculation_fallback_and_steady_state(tb):
    """Noise below the start threshold: the threshold falls from 500 by 2 every 4th block, so the whole-call speculative bound
    (thresh - spec_margin) fails inside the first call and the rest of it is redone in 64-block epochs; the
    second call starts near equilibrium and the speculation holds.  Both must match the oracle block by block."""
    n_raw = 24 * 1024 * 1024 // 2
    iq = g.make_stream(n_raw, [], seed=21, sigma=4.0)
    half = (iq.size // 2) // 65536 * 65536
    rx = tb.Receiver(types=0x07, thresh=0)
    o = ol.Oracle(types=0x07, thresh=0)
    rx.submit(0, iq[:half].copy())
    rx.process()
    rx.sync()
    fb1 = rx.stats()["fallback_epochs"]
    assert fb1 > 0, "the start-up transient must exercise the epoch fallback"
    o.process(iq[:half])
    assert np.array_equal(rx.block_trace(0), o.blocks())
    assert rx.thresh(0) == o.thresh()
    rx.submit(0, iq[half:].copy())
    rx.process()
    rx.sync()
    assert rx.stats()["fallback_epochs"] == fb1, "at equilibrium the whole-call speculation must hold"
    o2 = ol.Oracle(types=0x07, thresh=0)
    o2.process(iq)
    assert np.array_equal(rx.block_trace(0), o2.blocks()[half // 65536:])
    assert rx.thresh(0) == o2.thresh()
    assert [frame_key(f) for f in rx.frames()] == [frame_key(f) for f in o2.frames()]
    rx.close()


def test_pipelined_ragged_multi_stream_calls(tb, hot_fixture):
    """Several tfr_process calls in flight without a sync in between (front-end of call i+1 overlaps the
    back-end of call i), streams of different lengths, one stream skipping calls: every stream must still see
    exactly what one uninterrupted reference run over its bytes produces."""
    names = ["mixed5", "cont_noisy", "strong_t7"]
    iqs = [hot_fixture(n) for n in names]
    step = [5 * 65536, 9 * 65536, 13 * 65536]
    rx = tb.Receiver(types=0x2F, thresh=0, n_streams=3)
    offs = [0, 0, 0]
    call = 0
    while any(offs[s] < iqs[s].size for s in range(3)):
        for s in range(3):
            if s == 2 and call % 3 == 1:
                continue                      # this stick delivers nothing this time
            if offs[s] < iqs[s].size:
                n = min(step[s], iqs[s].size - offs[s])
                rx.submit(s, iqs[s][offs[s]:offs[s] + n].copy())
                offs[s] += n
        rx.process()
        call += 1
    frames, records = rx.frames(), rx.records()
    for s, iq in enumerate(iqs):
        o = ol.Oracle(types=0x2F)
        o.process(iq)
        assert [frame_key(f) for f in frames if f["stream"] == s] == [frame_key(f) for f in o.frames()], names[s]
        assert [r["exec"] for r in records if r["stream"] == s] == [r["exec"] for r in o.records()], names[s]
        assert rx.thresh(s) == o.thresh()
    rx.close()


@pytest.mark.parametrize("thresh,types", [(260, 0x0E), (320, 0x07), (0, 0x0E)])
def test_dense_near_windows_chains(tb, thresh, types):
    """A threshold inside the noise: windows follow each other within a timeout (`near` windows are run as
    chains by one thread, with carried biquad state and last_bit_idx) and many speculated biquad start states
    need the verifier.  Discriminator values, all biquad outputs and the frames are compared with the oracle."""
    iq = g.make_stream(3 * 1024 * 1024, [], seed=33 + thresh, sigma=4.0)
    rx = tb.Receiver(types=types, thresh=thresh, flags=tb.FLAG_TAPS)
    rx.submit(0, iq)
    rx.process()
    compare_with_oracle(rx, iq, types, 0, thresh)
    st = rx.stats()
    assert st["windows"] > 20, st
    rx.close()


def test_weak_signal_many_retriggers(tb):
    """Telegrams just above the noise: bursts where triggers come and go, so window chains get long and the
    block event lists overflow into the dense path of the threshold kernel."""
    sensors = [g.TFA_1, g.TFA_2, g.TFA_3]
    iq = g.fixture_continuous(6 * 1024 * 1024, sensors, 500000, seed=17, sigma=5.0, amp=14)[0]
    for thresh in (0, 280):
        rx = tb.Receiver(types=0x07, thresh=thresh, flags=tb.FLAG_TAPS)
        rx.submit(0, iq)
        rx.process()
        compare_with_oracle(rx, iq, 0x07, 0, thresh)
        rx.close()


# ---------------------------------------------------------------------------------------------------
# round 2: the benchmarked workload itself, the verifier's retired frames, long TFA_1 runs
# ---------------------------------------------------------------------------------------------------
def test_bench_stream_three_submits_against_oracle(tb):
    """One stream of bench.py's workload exactly as the bench drives it: a 128 MiB buffer (2048 blocks per submit,
    telegram every 15.36 M samples) submitted three times in a row with carried state, `-T 7`, auto threshold.
    Frames, records, the per-block threshold trace of every call and the final threshold against the oracle."""
    import torch
    import bench
    nbytes = 128 << 20
    buf, n_bursts = bench.make_stream_gpu(5, nbytes, 4.0, torch.device("cuda", 0))
    assert n_bursts >= 4
    iq = buf.cpu().numpy()
    rx = tb.Receiver(types=0x07, thresh=0, max_blocks_per_submit=nbytes // 65536)
    o = ol.Oracle(types=0x07, thresh=0)
    traces = []
    for _ in range(3):
        rx.submit(0, buf.data_ptr(), nbytes=nbytes)
        rx.process()
        traces.append(rx.block_trace(0))
        o.process(iq)
    assert np.array_equal(np.concatenate(traces), o.blocks()), "per-block threshold/trigger trace differs"
    assert rx.thresh(0) == o.thresh()
    gf, of = rx.frames(), o.frames()
    assert [frame_key(f) for f in gf] == [frame_key(f) for f in of]
    assert [record_key(r) for r in rx.records()] == [record_key(r) for r in o.records()]
    assert [r["exec"] for r in rx.records()] == [r["exec"] for r in o.records()]
    assert len(o.records()) >= 6, "the workload must actually decode telegrams"
    rx.close()


def test_tfa1_long_carrier_run(tb):
    """A TFA_1 telegram followed by an unmodulated carrier that keeps the window open for 0.9 s, then one phase
    flip: the gap between the two dips exceeds 655k index units, i.e. MORE THAN 32767 one-bits in a single run of
    `for (n = 22; n <= gap; n += 20) store_bit(1)` (tfa1.cpp:168-173).  The decoder is synced by then, so every
    eight of them add a byte: byte_cnt of the frame (4702) pins the exact count."""
    n = 52 * 32768
    rng = np.random.default_rng(4)
    iq = np.rint(rng.normal(0, 1.0, size=2 * n)).astype(np.int64)
    bi, bq = g.burst_tfa1(g.KAT_TFA1)
    at = 100000
    iq[2 * at:2 * (at + len(bi)):2] += bi
    iq[2 * at + 1:2 * (at + len(bi)) + 1:2] += bq
    e = at + len(bi)
    ci = np.full(1500000, bi[-1], dtype=np.int64)
    cq = np.full(1500000, bq[-1], dtype=np.int64)
    ci[1400000:] *= -1
    cq[1400000:] *= -1
    iq[2 * e:2 * (e + len(ci)):2] += ci
    iq[2 * e + 1:2 * (e + len(ci)) + 1:2] += cq
    iq = np.clip(iq + 128, 0, 255).astype(np.uint8)
    rx = tb.Receiver(types=0x01, thresh=500, flags=tb.FLAG_TAPS)
    rx.submit(0, iq)
    rx.process()
    o = compare_with_oracle(rx, iq, 0x01, 0, 500)
    assert [f["byte_cnt"] for f in o.frames()] == [4702]
    rx.close()
