"""CPU: the numpy restatement oracle/features_ref.py against outputs of the reference's own process_traces
(tests/golden/features.npz, made by oracle/make_golden_features.py from /root/reference) -- bit for bit."""
import hashlib
import os

import numpy as np
import pytest

from oracle import features_ref

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "features.npz")
SYNTHETIC = ["empty", "one", "two", "repeats", "unsorted", "exact_cap", "cap_plus_one", "cap_small", "cap_two"]


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("name", SYNTHETIC)
def test_synthetic_cases_bit_exact(golden, name):
    got = features_ref.process_points(golden[f"{name}_points"], int(golden[f"{name}_maxlen"]))
    want = golden[f"{name}_feats"]
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("k", [0, 1, 2])
def test_real_traces_bit_exact(golden, k):
    got = features_ref.process_points(golden[f"real{k}_points"], int(golden[f"real{k}_maxlen"]))
    assert tuple(got.shape) == tuple(golden[f"real{k}_shape"])
    assert np.array_equal(got[::7].view(np.uint32), golden[f"real{k}_rows7"].view(np.uint32))
    assert hashlib.sha256(got.tobytes()).digest() == golden[f"real{k}_sha256"].tobytes()


def test_downsample_index_matches_linspace():
    for n, cap in [(51, 50), (3145, 3000), (41130, 3000), (1000, 17), (33, 2), (7, 3)]:
        assert np.array_equal(features_ref.downsample_index(n, cap), np.linspace(0, n - 1, cap, dtype=int))


def test_collate_pads_with_zero_rows():
    a, b = np.ones((3, 11), np.float32), np.ones((5, 11), np.float32)
    batch, mask = features_ref.collate([a, b])
    assert batch.shape == (2, 5, 11) and mask.sum() == 8 and batch[0, 3:].sum() == 0 and not mask[0, 3:].any()
