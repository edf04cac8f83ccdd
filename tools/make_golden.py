#!/usr/bin/env python
"""Generate tests/golden/*.json by running the UNMODIFIED reference (oracle/_ref, built by
`make -C oracle ref` from /root/reference) on seeded synthetic inputs from tools/iqsynth.py.

Only runs where /root/reference (and therefore oracle/_ref) exists.  The JSON files are committed;
tests re-generate the inputs from the same seeds, check their sha256, and compare the oracle port
(CPU tests) and the CUDA path (GPU tests) against what the reference printed / computed here.
"""
from __future__ import annotations

import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import iqsynth as g  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")
TAP_DT = np.dtype([("kind", "<i4"), ("a", "<i4"), ("b", "<i4"), ("c", "<i4"), ("d", "<i4"), ("res", "<i4"),
                   ("din", "<f8"), ("dout", "<f8"), ("obj", "<u8")])
DECODE_PREFIX = ("TFA1 ", "TFA2 ", "TFA3 ", "TX22 ID", "WHB0", "WHB1", "WHB: Probably")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# --------------------------------------------------------------------------- fixture catalogue
def hotpath_fixtures():
    """name -> (builder, list of (label, reference argv, oracle kwargs))"""
    all5 = [g.TFA_1, g.TFA_2, g.TFA_3, g.TX22, g.TFA_WHB]
    return {
        "single_tfa1": (lambda: g.fixture_single_tfa1(seed=1), [
            ("T1_auto", ["-T", "1"], dict(types=0x01, filter=0, thresh=0)),
            ("T1_t500", ["-T", "1", "-t", "500"], dict(types=0x01, filter=0, thresh=500)),
        ]),
        "mixed5": (lambda: g.fixture_mixed5(seed=42), [
            ("T2f_auto", ["-T", "2f"], dict(types=0x2F, filter=0, thresh=0)),
            ("T2f_t500", ["-T", "2f", "-t", "500"], dict(types=0x2F, filter=0, thresh=500)),
            ("T2f_wide", ["-T", "2f", "-W"], dict(types=0x2F, filter=1, thresh=0)),
            ("T7_auto", ["-T", "7"], dict(types=0x07, filter=0, thresh=0)),
            ("Te_t300", ["-T", "e", "-t", "300"], dict(types=0x0E, filter=0, thresh=300)),
        ]),
        "cont_noisy": (lambda: g.fixture_continuous(8 * 1024 * 1024, all5, 1500000, seed=7, sigma=4.0, amp=60)[0], [
            ("T2f_auto", ["-T", "2f"], dict(types=0x2F, filter=0, thresh=0)),
            ("T7_auto", ["-T", "7"], dict(types=0x07, filter=0, thresh=0)),
        ]),
        "noise_only": (lambda: g.make_stream(4 * 1024 * 1024, [], seed=11, sigma=6.0), [
            ("T2f_auto", ["-T", "2f"], dict(types=0x2F, filter=0, thresh=0)),
            ("T2f_t400", ["-T", "2f", "-t", "400"], dict(types=0x2F, filter=0, thresh=400)),
        ]),
        "strong_t7": (lambda: g.fixture_continuous(6 * 1024 * 1024, [g.TFA_1, g.TFA_2, g.TFA_3], 700000, seed=3,
                                                    sigma=2.0, amp=110)[0], [
            ("T7_auto", ["-T", "7"], dict(types=0x07, filter=0, thresh=0)),
        ]),
    }


def kat_frames():
    """(sensor_e, bytes) for the -X parser seam: the five known answers, random valid frames,
    and single-bit corruptions of them."""
    rng = np.random.default_rng(2024)
    out = [(g.TFA_1, g.KAT_TFA1), (g.TFA_2, g.KAT_TFA2), (g.TFA_3, g.KAT_TFA2), (g.TX22, g.KAT_TX22),
           (g.TFA_WHB, g.KAT_WHB03 + bytes(4))]
    for s in (g.TFA_1, g.TFA_2, g.TFA_3, g.TX22, g.TFA_WHB):
        for _ in range(6):
            f = g.random_frame(s, rng)
            if s == g.TFA_WHB:
                f = f + bytes(3)
            out.append((s, f))
            bad = bytearray(f)
            bad[int(rng.integers(2, len(f)))] ^= 1 << int(rng.integers(0, 8))
            out.append((s, bytes(bad)))
    # TFA_1 special cases: 30.3181 (hum 0x6a), sensor fail (hum 0x7f / temp 0xaa), lowbat
    out.append((g.TFA_1, g.frame_tfa1(0x1234, 21.5, 0x6A, 3)))
    out.append((g.TFA_1, g.frame_tfa1(0x1234, 21.5, 0x7F, 3)))
    out.append((g.TFA_1, g.frame_tfa1(0x7FFF, -12.3, 99, 15, lowbat=1)))
    out.append((g.TFA_2, g.frame_tfa2(0xFC, 0.0, 0x7D)))
    out.append((g.TFA_2, g.frame_tfa2(0x04, 59.9, 0x6A)))
    out.append((g.TX22, g.frame_tx22(63, temp_c=-5.5)))
    out.append((g.TX22, g.frame_tx22(1, hum=40)))
    # every WeatherHub payload parser and CRC-32 init value (whb.cpp:50-62, 126-475): random payloads of the
    # length each parser reads, plus one corrupted copy per type; an unknown type; a length byte beyond 60
    rng = np.random.default_rng(2025)
    for stype, plen in g.WHB_PAYLOAD_LEN.items():
        for k in range(4):
            f = g.frame_whb(stype, int(rng.integers(0, 1 << 40)), rng.integers(0, 256, size=plen).tolist())
            f = f + bytes(int(rng.integers(0, 4)))   # byte count usually 2-3 bytes longer than the payload (whb.cpp:487)
            out.append((g.TFA_WHB, f))
        bad = bytearray(f)
        bad[int(rng.integers(4, len(f) - 3))] ^= 1 << int(rng.integers(0, 8))
        out.append((g.TFA_WHB, bytes(bad)))
    # temperatures at the 11-bit sign boundary and extremes (cvt_temp, whb.cpp:109-123), type 09's 12-bit second probe
    for t in (0x000, 0x3FF, 0x400, 0x7FF):
        out.append((g.TFA_WHB, g.frame_whb(0x02, 0x1122334455, [0x3F, 0xFF, t >> 8, t & 0xFF, (t ^ 0x7FF) >> 8, (t ^ 0x7FF) & 0xFF])))
    out.append((g.TFA_WHB, g.frame_whb(0x09, 0xA1B2C3D4E5, [0, 1, 0x07, 0xFF, 0x0F, 0xFF, 0, 50, 0x04, 0x00, 0x08, 0x00, 0, 99])))
    unk = bytearray(g.frame_whb(0x03, 0x0102030405, [0] * 11))
    unk[5] = 0x05                                       # not in crc_initvals -> "Probably unsupported sensor type"
    out.append((g.TFA_WHB, bytes(unk)))
    long_len = bytearray(g.frame_whb(0x03, 0x0102030405, [0] * 11))
    long_len[4] = 61                                    # plen > 60 (whb.cpp:499-500)
    out.append((g.TFA_WHB, bytes(long_len)))
    out.append((g.TFA_WHB, g.frame_whb(0x11, 0x0F0E0D0C0B, rng.integers(0, 256, size=34).tolist()) + bytes(8)))   # 61 bytes: > 60 gate
    # TX22: every word type, several records per frame (sub-ids 2/3/4), num up to 7, flag bits, unknown word types
    W = lambda t, v: ((t & 0xF) << 12) | (v & 0xFFF)
    bcd = lambda v: (v // 100 % 10) << 8 | (v // 10 % 10) << 4 | v % 10
    out.append((g.TX22, g.frame_tx22_words(5, [W(0, bcd(613)), W(1, bcd(55)), W(2, 0x123), W(3, 0x7FE), W(4, 0x0A5)])))
    out.append((g.TX22, g.frame_tx22_words(63, [W(2, 0xFFF)])))
    out.append((g.TX22, g.frame_tx22_words(0, [W(3, 0xF00), W(4, 0xFFF)], ok=0)))
    out.append((g.TX22, g.frame_tx22_words(17, [W(4, 0x001), W(2, 0), W(0, bcd(0))], lowbat=1)))
    out.append((g.TX22, g.frame_tx22_words(33, [W(7, 0x123), W(0xF, 0xFFF), W(1, bcd(100))], ok=0, lowbat=1)))
    out.append((g.TX22, g.frame_tx22_words(9, [W(0, bcd(400)), W(0, bcd(999)), W(1, bcd(1)), W(1, bcd(99)), W(2, 1), W(3, 0x10A), W(4, 0x3E8)])))
    out.append((g.TX22, g.frame_tx22_words(21, [])))
    for _ in range(8):
        n = int(rng.integers(1, 8))
        words = [W(int(rng.integers(0, 6)), int(rng.integers(0, 4096))) for _ in range(n)]
        f = g.frame_tx22_words(int(rng.integers(0, 64)), words, ok=int(rng.integers(0, 2)), lowbat=int(rng.integers(0, 2)))
        out.append((g.TX22, f))
    bad = bytearray(f)
    bad[4] ^= 0x10
    out.append((g.TX22, bytes(bad)))
    return out


# --------------------------------------------------------------------------- reference runners
def run_ref(argv, env=None):
    e = dict(os.environ)
    if env:
        e.update(env)
    r = subprocess.run(argv, capture_output=True, env=e)
    return r.stdout.decode("latin1")


def decode_lines(stdout):
    return [ln for ln in stdout.splitlines() if ln.startswith(DECODE_PREFIX)]


def exec_lines(stdout):
    """`-q -e /bin/echo` output: drop the trailing ts (decoder.cpp:72-94), keep argv order."""
    out = []
    for ln in stdout.splitlines():
        m = re.fullmatch(r"([0-9a-f]+ [+-][0-9.]+ \S+ -?\d+ -?\d+ -?\d+ -?\d+) \d+", ln.strip())
        if m:
            out.append(m.group(1))
    return out


def block_trace(stdout):
    """-DDD: per process() call 'Trigger ratio t/n, avg a' and threshold moves (fm_demod.cpp:60-72)."""
    rows, thresh = [], None
    for ln in stdout.splitlines():
        m = re.search(r"Trigger ratio (\d+)/(\d+), avg (\d+)", ln)
        if m:
            rows.append([int(m.group(1)), int(m.group(3))])
            continue
        m = re.search(r"(Increased|Decreased) trigger level to (\d+)", ln)
        if m:
            rows[-1].append(int(m.group(2)))
    return rows


def gen_hotpath():
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for name, (builder, cases) in hotpath_fixtures().items():
            iq = builder()
            path = os.path.join(td, name + ".iq")
            iq.tofile(path)
            entry = {"input_sha256": sha(iq), "n_bytes": int(iq.size), "cases": {}}
            for label, argv, okw in cases:
                tapf = os.path.join(td, "taps.bin")
                so = run_ref([os.path.join(REF, "tfrec_taps"), *argv, "-L", path], {"TFR_TAP_FILE": tapf})
                taps = np.fromfile(tapf, dtype=TAP_DT)
                se = run_ref([os.path.join(REF, "tfrec"), *argv, "-q", "-e", "/bin/echo", "-L", path])
                sd = run_ref([os.path.join(REF, "tfrec"), *argv, "-DDD", "-L", path])
                tr = block_trace(sd)
                thresh0 = okw["thresh"] if okw["thresh"] else 500
                th, trace = thresh0, []
                for row in tr:
                    trace.append([th, row[0], row[1]])
                    if len(row) > 2:
                        th = row[2]
                c = {"argv": argv, "oracle": okw, "lines": decode_lines(so), "exec": exec_lines(se),
                     "inverted_syncs": so.count("Inverted SYNC"),
                     "n_blocks": len(trace), "final_thresh": th,
                     "trace_sha256": sha(np.asarray(trace, dtype=np.int32)),
                     "trace_head": trace[:8], "trace_tail": trace[-4:], "taps": {}}
                for kind, key, field in ((0, "fm_dev", "res"), (1, "fm_dev_nrzs", "res"), (2, "iir2_step", "dout")):
                    v = taps[field][taps["kind"] == kind]
                    v = v.astype("<i4") if kind < 2 else v.astype("<f8")
                    c["taps"][key] = {"n": int(v.size), "sha256": sha(v),
                                      "head": [int(x) if kind < 2 else float(x).hex() for x in v[:16]]}
                entry["cases"][label] = c
                print(name, label, len(c["lines"]), "lines", len(c["exec"]), "exec", len(trace), "blocks",
                      {k: t["n"] for k, t in c["taps"].items()})
            out[name] = entry
    return out


def gen_decimator():
    out = {}
    rng = np.random.default_rng(5)
    fixtures = {
        "uniform_bytes": rng.integers(0, 256, size=4 * 65536, dtype=np.uint8),
        "extremes": np.tile(np.array([0, 255, 255, 0, 0, 0, 255, 255], dtype=np.uint8), 65536 // 8 * 2),
        "single_tfa1": g.fixture_single_tfa1(seed=1),
    }
    with tempfile.TemporaryDirectory() as td:
        for name, iq in fixtures.items():
            p = os.path.join(td, name)
            iq.tofile(p)
            e = {"input_sha256": sha(iq), "n_bytes": int(iq.size)}
            for filt in (0, 1):
                q = p + ".s16"
                subprocess.run([os.path.join(REF, "ref_decim"), p, q, str(filt)], check=True)
                d = np.fromfile(q, dtype="<i2")
                # framing independence: 4096-byte blocks must give the same stream
                subprocess.run([os.path.join(REF, "ref_decim"), p, q, str(filt), "4096"], check=True)
                d2 = np.fromfile(q, dtype="<i2")
                assert np.array_equal(d, d2)
                e["narrow" if filt == 0 else "wide"] = {"n": int(d.size), "sha256": sha(d), "head": d[:48].tolist(),
                                                         "min": int(d.min()), "max": int(d.max())}
            # downconvert(passes) for every decimation of BASELINE configs[4] (dsp_stuff.cpp:232-264): /2 .. /32
            e["passes"] = {}
            for passes in (1, 2, 3, 4, 5):
                ep = {}
                for filt in (0, 1):
                    q = p + ".s16"
                    subprocess.run([os.path.join(REF, "ref_decim"), p, q, str(filt), "65536", str(passes)], check=True)
                    d = np.fromfile(q, dtype="<i2")
                    subprocess.run([os.path.join(REF, "ref_decim"), p, q, str(filt), "4096", str(passes)], check=True)
                    assert np.array_equal(d, np.fromfile(q, dtype="<i2"))
                    ep["narrow" if filt == 0 else "wide"] = {"n": int(d.size), "sha256": sha(d), "head": d[:16].tolist(),
                                                             "min": int(d.min()), "max": int(d.max())}
                e["passes"][str(passes)] = ep
            assert e["passes"]["2"]["narrow"]["sha256"] == e["narrow"]["sha256"]
            out[name] = e
            print("decimator", name, e["narrow"]["n"])
    return out


def gen_kat():
    out = []
    with tempfile.TemporaryDirectory() as td:
        for sensor, frame in kat_frames():
            p = os.path.join(td, "x.txt")
            with open(p, "w") as f:
                f.write(" ".join("%02x" % b for b in frame) + "\n")
            mask = "%x" % (1 << sensor)
            so = run_ref([os.path.join(REF, "tfrec"), "-T", mask, "-X", p])
            se = run_ref([os.path.join(REF, "tfrec"), "-T", mask, "-q", "-e", "/bin/echo", "-X", p])
            out.append({"sensor": sensor, "hex": frame.hex(), "lines": decode_lines(so), "exec": exec_lines(se)})
    print("kat frames", len(out), "decoded", sum(1 for k in out if k["lines"]))
    return out


def gen_biquad():
    """as-built iir2 coefficients (5 doubles each) read out of the reference object"""
    src = r'''
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "dsp_stuff.h"
int main(){ double cut[]={0.5/((1536000/4.0)/17240),0.5/((1536000/4.0)/9600),0.5/((1536000/4.0)/8842),2.0/64.0,0.0025/64.0};
 for(int k=0;k<5;k++){ iir2 f(cut[k]); double m[10]; memcpy(m,(void*)&f,sizeof(m));
  for(int j=5;j<10;j++){ uint64_t u; memcpy(&u,&m[j],8); printf("%016lx ",(unsigned long)u);} printf("\n"); } }
'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "c.cpp")
        open(c, "w").write(src)
        exe = os.path.join(td, "c")
        subprocess.run(["g++", "-O2", "-include", "stdint.h", "-I/root/reference", "-o", exe, c,
                        os.path.join(REF, "dsp_stuff.o"), "-lm"], check=True)
        rows = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split("\n")
    return [r.split() for r in rows if r.strip()]


def main():
    if not os.path.exists(os.path.join(REF, "tfrec_taps")):
        sys.exit("oracle/_ref missing: run `make -C oracle ref` where /root/reference exists")
    os.makedirs(GOLD, exist_ok=True)
    only = sys.argv[1:]   # e.g. `make_golden.py decimator` regenerates one file
    for name, fn in (("kat_frames", gen_kat), ("decimator", gen_decimator), ("biquad_coeffs", gen_biquad),
                     ("hotpath", gen_hotpath)):
        if only and name not in only:
            continue
        with open(os.path.join(GOLD, name + ".json"), "w") as f:
            json.dump(fn(), f, indent=1)
            f.write("\n")


if __name__ == "__main__":
    main()
