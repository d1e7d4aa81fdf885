"""CPU self-checks of the oracles: the GRU oracle's analytic gradients
against finite differences in fp64 (torch.autograd.gradcheck), so that the contract the CUDA path is held to is
self-consistent, and hypothesis property tests of the two binning restatements (numpy == C) on random bounds,
resolutions and point clouds including NaN / Inf / cell-edge / out-of-range points."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import baseline_ref, heatmap_ref_c
from oracle.room_slam_ref import RoomSLAM as Ref


def _target(B, N, g):
    return {"classes": torch.randint(0, 4, (B, N), generator=g), "positions": torch.rand(B, N, 2, generator=g, dtype=torch.float64) * 10,
            "sizes": torch.rand(B, N, 2, generator=g, dtype=torch.float64) + 0.2,
            "orientations": (torch.rand(B, N, generator=g, dtype=torch.float64) - 0.5) * 6,
            "valid": (torch.rand(B, N, generator=g) < 0.6).double()}


@pytest.mark.parametrize("layers,use_mask", [(1, False), (2, False), (2, True)])
def test_oracle_gradcheck_fp64(layers, use_mask):
    """d(total loss)/d(input) and d/d(every parameter) of the oracle in fp64 against central differences."""
    g = torch.Generator().manual_seed(7)
    B, T, H, N = 2, 5, 4, 3
    ref = Ref(hidden_size=H, num_layers=layers, max_objects=N, decoder_hidden=8, dropout=0.5 if use_mask else 0.0).double()
    ref.train(use_mask)
    mask = ref.make_dropout_mask(B, T, g).double() if use_mask else None
    tgt = _target(B, N, g)
    x = torch.randn(B, T, 2, generator=g, dtype=torch.float64, requires_grad=True)
    names = [n for n, _ in ref.named_parameters()]
    params = [p for _, p in ref.named_parameters()]

    def loss_of(xx, *ps):
        out = torch.func.functional_call(ref, dict(zip(names, ps)), (xx, mask))
        return ref.compute_loss(out, tgt)["total"]

    assert torch.autograd.gradcheck(loss_of, (x, *params), eps=1e-6, atol=1e-6, rtol=1e-4, nondet_tol=0.0)


def test_oracle_each_loss_component_gradcheck():
    g = torch.Generator().manual_seed(3)
    B, T, N = 2, 4, 3
    ref = Ref(hidden_size=4, num_layers=1, max_objects=N, decoder_hidden=8, dropout=0.0).double().eval()
    tgt = _target(B, N, g)
    x = torch.randn(B, T, 2, generator=g, dtype=torch.float64, requires_grad=True)
    for key in ("class", "position", "size", "orientation", "validity"):
        assert torch.autograd.gradcheck(lambda xx: ref.compute_loss(ref(xx), tgt)[key], (x,), eps=1e-6, atol=1e-6, rtol=1e-4), key


finite = st.floats(min_value=-40.0, max_value=40.0, allow_nan=False, width=32)
weird = st.sampled_from([float("nan"), float("inf"), float("-inf"), 0.0, -0.0, 1e-30, -1e-30, 3.0e38])


@settings(max_examples=60, deadline=None)
@given(x_min=st.floats(-20, 5), y_min=st.floats(-20, 5), w=st.floats(0.25, 12), h=st.floats(0.25, 12),
       res=st.sampled_from([0.05, 0.1, 0.013, 0.25, 1.0]), v=st.sampled_from([0.1, 0.5, 0.01]),
       n=st.integers(1, 6), t=st.integers(1, 40), seed=st.integers(0, 2 ** 31 - 1),
       specials=st.lists(st.tuples(st.integers(0, 5), st.integers(0, 39), st.one_of(weird, finite), st.one_of(weird, finite)), max_size=6),
       on_edges=st.booleans())
def test_binning_numpy_equals_c_property(x_min, y_min, w, h, res, v, n, t, seed, specials, on_edges):
    b = baseline_ref.OccupancyHeatmapBaseline(bounds=(x_min, x_min + w, y_min, y_min + h), resolution=res, stationary_speed=v)
    rng = np.random.default_rng(seed)
    pts = np.stack([rng.uniform(x_min - 1, x_min + w + 1, (n, t)), rng.uniform(y_min - 1, y_min + h + 1, (n, t))], -1).astype(np.float32)
    if on_edges:        # points exactly on cell edges k * res (as fp32 computes them) and repeated points (stationary samples)
        k = rng.integers(0, max(1, b.gx), (n, t))
        pts[..., 0] = (np.float32(x_min) + k.astype(np.float32) * np.float32(res)).astype(np.float32)
        pts[:, 1::2] = pts[:, 0:-1:2] if t > 1 and t % 2 == 0 else pts[:, 1::2]
    for (i, j, px, py) in specials:
        if i < n and j < t:
            pts[i, j] = (px, py)
    o1 = baseline_ref.bin_points(pts, b.bounds[0], b.bounds[2], b.resolution, b.gx, b.gy, b.thr2)
    o2 = heatmap_ref_c.bin_points(pts, b.bounds[0], b.bounds[2], b.resolution, b.gx, b.gy, b.thr2)
    assert np.array_equal(o1[0], o2[0]) and np.array_equal(o1[1], o2[1]) and o1[2] == o2[2]
    assert int(o1[0].sum()) + o1[2] == n * t                  # every point is binned or dropped, never both
    assert (o1[1] <= o1[0]).all()                             # a stationary sample is also a visit of that cell
