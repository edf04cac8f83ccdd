"""CPU suite: the -e sink of the host mirror (tfrec_b200/host/decoder.cpp).  Default = the reference's contract, one
system() per telegram (decoder.cpp:67-96); TFREC_EXEC=async = the same command lines, in the same order, through one
shell that the decoder does not wait for.  No device is touched: the test drives decoder::store_data directly."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "tfrec_b200/host/decoder.h"
#include <stdio.h>
int main(int argc, char **argv)
{
	decoder tfa(TFA_2), whb(TFA_WHB);
	tfa.set_params(argv[1], 0, 0);
	whb.set_params(argv[1], 0, 0);
	for (int k = 0; k < 40; k++) {
		sensordata_t d;
		d.type = (k & 1) ? TFA_WHB : TFA_2;
		d.id = (k & 1) ? 0x0b1234567890ull + k : 0x1000 + k;
		d.temp = 20.5 + k;
		d.humidity = 40 + k;
		d.sequence = k;
		d.alarm = k & 1;
		d.rssi = 70 + k;
		d.flags = 0;
		d.ts = 1700000000 + k;
		((k & 1) ? whb : tfa).store_data(d);
		if (k == 20) decoder::flush_exec();
	}
	// a WeatherHub repeat with the same sequence number is not executed again (decoder.cpp:55-61)
	sensordata_t r;
	r.type = TFA_WHB; r.id = 0x0b1234567890ull + 39; r.temp = 1; r.humidity = 2; r.sequence = 39; r.alarm = 0; r.rssi = 3; r.flags = 0; r.ts = 4;
	whb.store_data(r);
	return 0;
}
'''


@pytest.fixture(scope="module")
def sink_exe(tmp_path_factory):
    from tfrec_b200 import build
    build.build()
    d = tmp_path_factory.mktemp("sink")
    src = d / "sink.cpp"
    src.write_text(SRC)
    exe = d / "sink"
    lib = os.path.join(ROOT, "tfrec_b200")
    subprocess.run(["g++", "-O1", "-std=c++11", "-I", ROOT, "-o", str(exe), str(src), os.path.join(lib, "host", "decoder.cpp"),
                    "-L" + lib, "-ltfrb200", "-Wl,-rpath," + lib], check=True)
    return str(exe)


def test_exec_sink_default_and_async_run_the_same_commands(sink_exe):
    env = dict(os.environ)
    env.pop("TFREC_EXEC", None)
    sync = subprocess.run([sink_exe, "/bin/echo"], capture_output=True, env=env, check=True).stdout.decode().splitlines()
    env["TFREC_EXEC"] = "async"
    asyn = subprocess.run([sink_exe, "/bin/echo"], capture_output=True, env=env, check=True).stdout.decode().splitlines()
    assert len(sync) == 40 and sync == asyn
    assert sync[0] == "1001000 +20.5 40 0 0 70 0 1700000000"          # id | type << 24, decoder.cpp:73-80
    assert sync[1] == "00b1234567891 +21.5 41 1 1 71 0 1700000001"    # WeatherHub: 13 hex digits, decoder.cpp:82-90
