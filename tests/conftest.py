import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json
    out = {}
    gdir = os.path.join(ROOT, "tests", "golden")
    for name in ("kat_frames", "decimator", "biquad_coeffs", "hotpath"):
        with open(os.path.join(gdir, name + ".json")) as f:
            out[name] = json.load(f)
    return out


_FIXTURE_CACHE = {}


@pytest.fixture(scope="session")
def hot_fixture():
    """name -> u8 IQ array, regenerated from the seeds in tools/make_golden.py and cached"""
    import make_golden

    def get(name):
        if name not in _FIXTURE_CACHE:
            _FIXTURE_CACHE[name] = make_golden.hotpath_fixtures()[name][0]()
        return _FIXTURE_CACHE[name]
    return get
