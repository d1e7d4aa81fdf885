"""GPU: rs_trace_features (through roomslam_b200.preprocess) against the reference's own outputs (golden) and the
numpy oracle, bit for bit."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import features_ref

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "features.npz")
SYNTHETIC = ["empty", "one", "two", "repeats", "unsorted", "exact_cap", "cap_plus_one", "cap_small", "cap_two"]


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def presorted(points):
    """numpy's argsort order (ties included), so the kernel sees exactly the rows the reference differenced."""
    p = np.asarray(points, np.float32).reshape(-1, 4)
    return p[np.argsort(p[:, 3])] if len(p) else p


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("name", SYNTHETIC)
def test_golden_synthetic(golden, name):
    from roomslam_b200 import preprocess
    cap = int(golden[f"{name}_maxlen"])
    out = preprocess.trace_features([presorted(golden[f"{name}_points"])], max_len=cap, sort=False)
    want = golden[f"{name}_feats"]
    assert tuple(out["traces"].shape) == (1,) + want.shape
    assert np.array_equal(bits(out["traces"][0].cpu().numpy()), bits(want))
    assert out["trace_mask"].all() and int(out["lengths"][0]) == want.shape[0]


@pytest.mark.parametrize("k", [0, 1, 2])
def test_golden_real_traces(golden, k):
    from roomslam_b200 import preprocess
    out = preprocess.trace_features([golden[f"real{k}_points"]], max_len=int(golden[f"real{k}_maxlen"]), sort=True)
    got = out["traces"][0].cpu().numpy()
    assert got.shape == tuple(golden[f"real{k}_shape"])
    assert np.array_equal(bits(got[::7]), bits(golden[f"real{k}_rows7"]))
    assert hashlib.sha256(got.tobytes()).digest() == golden[f"real{k}_sha256"].tobytes()


def test_batch_ragged_matches_oracle_collate(golden):
    from roomslam_b200 import preprocess
    rng = np.random.default_rng(5)
    traces = [golden[f"real{k}_points"] for k in range(3)] + [np.zeros((0, 4), np.float32), golden["two_points"]]
    for n in (1, 3, 257, 1999, 2000, 2001, 9000):
        t = np.cumsum(rng.uniform(1e-4, 0.1, n)) + rng.uniform(0, 1e4)
        traces.append(np.stack([rng.normal(0, 3, n), rng.normal(1.6, 0.1, n), rng.normal(0, 3, n), t], 1).astype(np.float32))
    for cap in (2000, 3000):
        out = preprocess.trace_features(traces, max_len=cap, sort=True, check_sorted=True)
        want, wmask = features_ref.collate([features_ref.process_points(t, cap) for t in traces])
        assert np.array_equal(bits(out["traces"].cpu().numpy()), bits(want))
        assert np.array_equal(out["trace_mask"].cpu().numpy(), wmask)
        assert np.array_equal(out["lengths"].cpu().numpy(), wmask.sum(1))


def test_unsorted_input_is_flagged():
    from roomslam_b200 import preprocess
    pts = np.array([[0, 0, 0, 2.0], [1, 0, 0, 1.0], [2, 0, 0, 3.0]], np.float32)
    with pytest.raises(ValueError):
        preprocess.trace_features([pts], sort=False, check_sorted=True)
    preprocess.trace_features([pts], sort=True, check_sorted=True)


def test_long_unsorted_trace_is_detected_between_kept_rows():
    """A trace LONGER than max_len whose only out-of-order timestamps sit between two kept (down-sampled) rows: the
    default sort="auto" must still sort it (the reference always argsorts, inference.py:38-39) and check_sorted must raise."""
    from roomslam_b200 import preprocess
    rng = np.random.default_rng(11)
    n, cap = 5000, 100
    t = np.cumsum(rng.uniform(0.01, 0.1, n)) + 7.0
    pts = np.stack([rng.normal(0, 3, n), rng.normal(1.6, 0.1, n), rng.normal(0, 3, n), t], 1).astype(np.float32)
    kept = set(features_ref.downsample_index(n, cap).tolist())
    k = next(i for i in range(1000, n - 3) if not ({i - 1, i, i + 1, i + 2} & kept))
    pts[[k, k + 1]] = pts[[k + 1, k]]                       # one swapped pair, neither row nor its neighbours is kept
    with pytest.raises(ValueError):
        preprocess.trace_features([pts], max_len=cap, sort=False, check_sorted=True)
    auto = preprocess.trace_features([pts, pts[:50]], max_len=cap)
    want, wmask = features_ref.collate([features_ref.process_points(pts, cap), features_ref.process_points(pts[:50], cap)])
    assert np.array_equal(bits(auto["traces"].cpu().numpy()), bits(want))
    assert np.array_equal(auto["trace_mask"].cpu().numpy(), wmask)


def test_large_batch_property():
    """Full-size property: 4096 traces x 3000 points; speed^2 == |v|^2 and a == diff(v) recomputed in torch."""
    from roomslam_b200 import preprocess
    g = torch.Generator(device="cuda").manual_seed(3)
    B, N = 4096, 3000
    pts = torch.randn(B, N, 4, device="cuda", generator=g)
    pts[..., 3] = torch.cumsum(torch.rand(B, N, device="cuda", generator=g) * 0.1 + 1e-3, 1)
    offsets = torch.arange(B + 1, dtype=torch.int64) * N
    out = preprocess.trace_features(pts.reshape(-1, 4), offsets, max_len=3000, sort=False, check_sorted=True)
    f = out["traces"]
    assert out["trace_mask"].all() and f.shape == (B, N, 11)
    a = pts.clone(); a[..., 3] = pts[..., 3] - pts[:, :1, 3]
    assert torch.equal(f[..., :4], a)
    d = torch.zeros_like(a); d[:, 1:] = a[:, 1:] - a[:, :-1]
    v = d[..., :3] / d[..., 3:].clamp_min(1e-3)
    assert torch.equal(f[..., 4:7], v)
    acc = torch.zeros_like(v); acc[:, 1:] = v[:, 1:] - v[:, :-1]
    assert torch.equal(f[..., 7:10], acc)
    torch.testing.assert_close(f[..., 10], v.norm(dim=-1), rtol=1e-6, atol=0)


def test_auto_sort_only_when_needed(golden):
    from roomslam_b200 import preprocess
    pts = golden["unsorted_points"]
    auto = preprocess.trace_features([pts, golden["two_points"]], max_len=3000)            # default sort="auto"
    want, wmask = features_ref.collate([features_ref.process_points(pts), features_ref.process_points(golden["two_points"])])
    assert np.array_equal(bits(auto["traces"].cpu().numpy()), bits(want))
    assert np.array_equal(auto["trace_mask"].cpu().numpy(), wmask)


def test_resample_windows_bit_exact(golden):
    """Recorded traces -> 10 Hz, T = 500 floor-plane windows for the GRU: GPU kernel against numpy.arange + numpy.interp."""
    from oracle import resample_ref
    from roomslam_b200 import preprocess
    rng = np.random.default_rng(2)
    traces = [golden[f"real{k}_points"].astype(np.float64) for k in range(3)]
    t = np.cumsum(rng.uniform(0.001, 0.3, 4000)) + 1234.5678
    traces.append(np.stack([rng.normal(0, 3, 4000), rng.normal(1.6, .1, 4000), rng.normal(0, 3, 4000), t], 1)[rng.permutation(4000)])
    traces.append(np.zeros((1, 4)))                                        # too short: no windows
    for seq_len, hz in ((500, 10.0), (64, 7.3), (50, 30.0)):
        out = preprocess.resample_windows(traces, seq_len=seq_len, hz=hz)
        want = [resample_ref.resample_windows(tr, seq_len, hz) for tr in traces]
        idx = np.concatenate([np.full(len(w), b) for b, w in enumerate(want)])
        want = np.concatenate(want)
        assert want.shape[0] > 0 and tuple(out["windows"].shape) == want.shape
        assert np.array_equal(out["trace"].numpy(), idx)
        assert np.array_equal(bits(out["windows"].cpu().numpy()), bits(want))
