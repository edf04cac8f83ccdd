"""CPU suite: the oracle port (oracle/tfrec_oracle.c) against the golden vectors the unmodified
reference produced (tools/make_golden.py), plus the reference's own known-answer telegram."""
import hashlib
import struct

import numpy as np
import pytest

import iqsynth as g
import make_golden
import oracle_lib as ol


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_readme_known_answer():
    # README.md:123 of the reference: 2d d4 65 b0 86 20 23 60 e0 56 97 -> ID 65b0 +22.0 35% seq e lowbat 0
    frame, recs = ol.parse(g.TFA_1, bytes.fromhex("2dd465b086202360e05697"))
    assert frame["status"] == 0 and len(recs) == 1
    r = recs[0]
    assert (r["id"], r["temp"], r["humidity"], r["sequence"], r["alarm"]) == (0x65B0, 22.0, 35.0, 0xE, 0)
    assert ol.crc8(bytes.fromhex("65b086202360e056")) == 0x97
    assert frame["line"] == "TFA1 ID 65b0 +22.0 35% seq e lowbat 0 RSSI 0"


def test_crc_parameters():
    # crc8.cpp / crc32.cpp: MSB-first, no reflection, no xorout
    assert ol.crc8(b"") == 0 and ol.crc8(b"\x01") == 0x31
    assert ol.crc32(b"", 0x12345678) == 0x12345678
    assert ol.crc32(b"\x00\x00\x00\x01", 0) == 0x04C11DB7
    assert ol.crc32(b"123456789", 0xFFFFFFFF) == 0x0376E6E7  # CRC-32/MPEG-2 check value


def test_parser_seam_against_reference(golden):
    for k in golden["kat_frames"]:
        frame, recs = ol.parse(k["sensor"], bytes.fromhex(k["hex"]))
        lines = [frame["line"]] if frame and frame["line"] else []
        assert lines == k["lines"], k
        assert [r["exec"] for r in recs] == k["exec"], k


def test_decimator_against_reference(golden):
    rng = np.random.default_rng(5)
    fixtures = {
        "uniform_bytes": rng.integers(0, 256, size=4 * 65536, dtype=np.uint8),
        "extremes": np.tile(np.array([0, 255, 255, 0, 0, 0, 255, 255], dtype=np.uint8), 65536 // 8 * 2),
        "single_tfa1": g.fixture_single_tfa1(seed=1),
    }
    for name, iq in fixtures.items():
        e = golden["decimator"][name]
        assert sha(iq) == e["input_sha256"]
        for filt, key in ((0, "narrow"), (1, "wide")):
            d = ol.decimate(iq, filt)
            assert d.size == e[key]["n"] and d[:48].tolist() == e[key]["head"]
            assert sha(d.astype("<i2")) == e[key]["sha256"]


def test_downconvert_all_passes_against_reference(golden):
    """BASELINE configs[4]: every decimation /2 .. /32 of downconvert(p) (dsp_stuff.cpp:232-264) against the
    compiled reference's output (oracle/_ref/ref_decim, hashes in tests/golden/decimator.json)"""
    rng = np.random.default_rng(5)
    fixtures = {
        "uniform_bytes": rng.integers(0, 256, size=4 * 65536, dtype=np.uint8),
        "extremes": np.tile(np.array([0, 255, 255, 0, 0, 0, 255, 255], dtype=np.uint8), 65536 // 8 * 2),
        "single_tfa1": g.fixture_single_tfa1(seed=1),
    }
    for name, iq in fixtures.items():
        for passes in (1, 2, 3, 4, 5):
            for filt, key in ((0, "narrow"), (1, "wide")):
                e = golden["decimator"][name]["passes"][str(passes)][key]
                d = ol.downconvert(iq, passes, filt)
                assert d.size == e["n"] == iq.size >> passes
                assert d[:16].tolist() == e["head"] and int(d.min()) == e["min"] and int(d.max()) == e["max"]
                assert sha(d.astype("<i2")) == e["sha256"]
        # passes = 2 is the cascade the decode path uses
        assert np.array_equal(ol.downconvert(iq, 2, 0), ol.decimate(iq, 0))


def test_downconvert_ragged_and_tiny():
    rng = np.random.default_rng(9)
    iq = rng.integers(0, 256, size=4096 + 6, dtype=np.uint8)      # odd number of pairs at several stages
    for passes in (1, 2, 3):
        d = ol.downconvert(iq, passes, 0)
        assert d.size == ((iq.size // 2) >> passes) * 2
        # a prefix of the input gives a prefix of the output (causal filters, zero history)
        d2 = ol.downconvert(iq[:2048], passes, 0)
        assert np.array_equal(d[:d2.size], d2)
    assert ol.downconvert(iq[:2], 1, 0).size == 0


def test_biquad_coefficients_as_built(golden):
    import ctypes as C
    for k, row in enumerate(golden["biquad_coeffs"]):
        out = (C.c_double * 5)()
        ol.lib().orc_biquad_coeffs(k, out)
        assert [struct.pack(">d", v).hex() for v in out] == row


def test_fm_dev_signed_zero_and_exact_angles():
    L = ol.lib()
    assert L.orc_fm_dev(-13, -39, 0, 0) == 16384      # cr=-0, cj=+0 -> +pi
    assert L.orc_fm_dev(0, -27, 0, 61) == -16384      # cj=-0, cr<0  -> -pi
    assert L.orc_fm_dev(100, 0, 100, 100) == -4096
    assert L.orc_fm_dev_nrzs(3, 4, 5, 6) == 39


CASES = [(name, label) for name, (_, cases) in make_golden.hotpath_fixtures().items() for (label, _, _) in cases]


@pytest.mark.parametrize("name,label", CASES)
def test_hot_path_against_reference(golden, hot_fixture, name, label):
    e = golden["hotpath"][name]
    c = e["cases"][label]
    iq = hot_fixture(name)
    assert sha(iq) == e["input_sha256"], "synthetic input differs from the one the golden was made from"
    o = ol.Oracle(taps=7, **c["oracle"])
    assert o.process(iq) == c["n_blocks"]
    assert [f["line"] for f in o.frames() if f["line"]] == c["lines"]
    assert [r["exec"] for r in o.records() if not (r["flags"] & 0x100)] == c["exec"]
    assert o.inverted_syncs() == c["inverted_syncs"]
    tr = o.blocks()
    assert tr[:8].tolist() == c["trace_head"] and tr[-4:].tolist() == c["trace_tail"]
    assert sha(tr.astype("<i4")) == c["trace_sha256"] and o.thresh() == c["final_thresh"]
    for kind, key in ((0, "fm_dev"), (1, "fm_dev_nrzs"), (2, "iir2_step")):
        v = o.tap(kind)
        t = c["taps"][key]
        assert v.size == t["n"]
        assert sha(v.astype("<i4") if kind < 2 else v.astype("<f8")) == t["sha256"], key


def test_block_framing_drops_partial_tail():
    # engine.cpp:70-76: a short read ends the replay, the partial block is never processed
    iq = g.fixture_single_tfa1(seed=1)
    o = ol.Oracle(types=1)
    assert o.process(iq[: 3 * 65536 + 1000]) == 3
    o2 = ol.Oracle(types=1)
    assert o2.process(iq[:1000]) == 0 and o2.frames() == []


def test_streaming_equals_one_shot():
    iq = g.fixture_mixed5(seed=42)
    a = ol.Oracle(types=0x2F)
    a.process(iq)
    b = ol.Oracle(types=0x2F)
    for off in range(0, iq.size, 5 * 65536):
        b.process(iq[off:off + 5 * 65536])
    assert a.frames() == b.frames() and a.records() == b.records()
