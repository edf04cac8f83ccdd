"""GPU suite: the tfrec-compatible command line (C++ host mirror of engine/fsk_demod/decoder over the C ABI)
against what the unmodified reference printed for the same inputs (tests/golden/hotpath.json, kat_frames.json)."""
import os
import re
import subprocess

import pytest

import make_golden

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "tfrec_b200", "tfrec_b200_cli")


@pytest.fixture(scope="module")
def cli():
    from tfrec_b200 import build
    build.build()
    build.build_cli()
    return CLI


def run(argv):
    return subprocess.run(argv, capture_output=True, timeout=300).stdout.decode("latin1")


@pytest.mark.parametrize("name,label", [("mixed5", "T2f_auto"), ("mixed5", "T2f_wide"), ("mixed5", "Te_t300"),
                                        ("single_tfa1", "T1_auto"), ("strong_t7", "T7_auto")])
def test_cli_replay_matches_reference(cli, golden, hot_fixture, tmp_path, name, label):
    c = golden["hotpath"][name]["cases"][label]
    path = tmp_path / (name + ".iq")
    hot_fixture(name).tofile(str(path))
    out = run([cli, *c["argv"], "-L", str(path)])
    assert make_golden.decode_lines(out) == c["lines"]
    assert out.count("Inverted SYNC\n") == c["inverted_syncs"]      # tfa2.cpp:294-300, printed whatever -D / -q say
    assert "done reading dump" in out
    out = run([cli, *c["argv"], "-q", "-e", "/bin/echo", "-L", str(path)])
    assert make_golden.exec_lines(out) == c["exec"]


def test_cli_hex_replay_matches_reference(cli, golden, tmp_path):
    # one -X file per sensor type (one process start each), frames in golden order
    by_sensor = {}
    for k in golden["kat_frames"]:
        by_sensor.setdefault(k["sensor"], []).append(k)
    for sensor, ks in by_sensor.items():
        p = tmp_path / ("x%d.txt" % sensor)
        p.write_text("".join(" ".join(re.findall("..", k["hex"])) + "\n" for k in ks))
        mask = "%x" % (1 << sensor)
        want_lines = [ln for k in ks for ln in k["lines"]]
        want_exec = []
        seen = {}   # decoder::store_data (decoder.cpp:46-65): a WeatherHub repeat with the same sequence is not exec'd again
        for k in ks:
            for ln in k["exec"]:
                f = ln.split()
                if sensor == 5:
                    if f[0] in seen and seen[f[0]] == f[3]:
                        continue
                    seen[f[0]] = f[3]
                want_exec.append(ln)
        assert make_golden.decode_lines(run([cli, "-T", mask, "-X", str(p)])) == want_lines
        assert make_golden.exec_lines(run([cli, "-T", mask, "-q", "-e", "/bin/echo", "-X", str(p)])) == want_exec


def test_cli_trigger_trace_matches_reference(cli, golden, hot_fixture, tmp_path):
    c = golden["hotpath"]["noise_only"]["cases"]["T2f_auto"]
    path = tmp_path / "n.iq"
    hot_fixture("noise_only").tofile(str(path))
    out = run([cli, *c["argv"], "-DDD", "-L", str(path)])
    tr = make_golden.block_trace(out)
    th, trace = 500, []
    for row in tr:
        trace.append([th, row[0], row[1]])
        if len(row) > 2:
            th = row[2]
    assert trace[:8] == c["trace_head"] and trace[-4:] == c["trace_tail"] and len(trace) == c["n_blocks"]


def test_cli_live_stdin_feed_and_save(cli, golden, hot_fixture, tmp_path):
    """live mode: raw u8 IQ on stdin in place of librtlsdr (sdr.cpp:228-271), -S saving what was consumed
    (sdr.cpp:38-44, 233-234); the saved file replayed with -L gives the same telegrams, and both equal the reference's -L run"""
    c = golden["hotpath"]["mixed5"]["cases"]["T2f_auto"]
    iq = hot_fixture("mixed5")
    saved = tmp_path / "saved.iq"
    tail = bytes(1000)   # a partial trailing block is dropped (engine.cpp:73-76)
    r = subprocess.run([cli, *c["argv"], "-S", str(saved)], input=iq.tobytes() + tail, capture_output=True, timeout=300)
    out = r.stdout.decode("latin1")
    assert make_golden.decode_lines(out) == c["lines"]
    assert saved.read_bytes() == iq.tobytes()
    assert make_golden.decode_lines(run([cli, *c["argv"], "-L", str(saved)])) == c["lines"]
    r = subprocess.run([cli, *c["argv"], "-q", "-e", "/bin/echo", "-L", "-"], input=iq.tobytes(), capture_output=True, timeout=300)
    assert make_golden.exec_lines(r.stdout.decode("latin1")) == c["exec"]
