"""GPU suite: a short randomised differential run against the oracle (tools/fuzz_gpu.py): random streams, decoder masks,
filters, thresholds (auto, fixed, inside the noise), call patterns (ragged submits, synchronised or in flight), front-end
chunking and back-end split modes - frames, records, "Inverted SYNC" count, block trace and threshold all equal the
oracle's.  (1980 cases of the same generator passed on the B200 while the round-2 kernels were written.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_random_cases_against_the_oracle(monkeypatch):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import fuzz_gpu
    import tfrec_b200 as tb
    tb.load()
    monkeypatch.setenv("TFR_BE_SPLIT", "1")       # (the cases set these per case; restored afterwards)
    monkeypatch.setenv("TFR_MIN_CHUNK", "8192")
    rng = np.random.default_rng(2026)
    bad = [k for k in range(24) if not fuzz_gpu.one_case(tb, rng, k)]
    assert not bad, "cases %s differ from the oracle" % bad
