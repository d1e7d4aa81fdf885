"""GPU parity: CUDA binning (through the C ABI) == oracle, bit for bit."""
import numpy as np
import pytest
import torch

from oracle import baseline_ref, heatmap_ref_c
from roomslam_b200 import OccupancyHeatmapBaseline, synth

pytestmark = pytest.mark.gpu
VARIANTS = [0, 1, 2, 3, 4, 5]


def _gpu_bin(pts, variant=0, **kw):
    b = OccupancyHeatmapBaseline(**kw)
    b._variant = variant
    occ, stat, nd = b.bin(torch.as_tensor(pts).cuda())
    return occ.cpu().numpy(), stat.cpu().numpy(), nd


@pytest.mark.parametrize("variant", VARIANTS)
def test_edge_points_bit_exact(golden_heatmap, variant):
    occ, stat, nd = _gpu_bin(golden_heatmap["edge_points"], variant)
    assert np.array_equal(occ, golden_heatmap["edge_occ"])
    assert np.array_equal(stat, golden_heatmap["edge_stat"])
    assert nd == int(golden_heatmap["edge_dropped"])


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("name", ["synth_a", "synth_b", "synth_c", "synth_d", "synth_e"])
def test_synth_bit_exact_vs_golden(golden_heatmap, name, variant):
    n, t, seed = (int(v) for v in golden_heatmap[f"{name}_shape"])
    if variant in (1, 2, 4, 5) and t % 2:
        pytest.skip("TMA variants need an even seq_len (16-byte aligned rows)")
    occ, stat, nd = _gpu_bin(synth.make_traces(n, t, seed=seed), variant)
    assert np.array_equal(occ, golden_heatmap[f"{name}_occ"])
    assert np.array_equal(stat, golden_heatmap[f"{name}_stat"])
    assert nd == int(golden_heatmap[f"{name}_dropped"])


def test_real_traces_bit_exact(golden_heatmap, real_traces):
    b = OccupancyHeatmapBaseline(bounds=tuple(golden_heatmap["real_bounds"]), resolution=0.05)
    occ, stat, nd = b.bin(torch.as_tensor(real_traces["windows"]).cuda())
    assert np.array_equal(occ.cpu().numpy(), golden_heatmap["real_occ"])
    assert np.array_equal(stat.cpu().numpy(), golden_heatmap["real_stat"])
    assert nd == int(golden_heatmap["real_dropped"])
    assert np.array_equal(b.stationary_cells(5.0).cpu().numpy(), golden_heatmap["real_cells_5s"])


@pytest.mark.parametrize("variant", [1, 2, 4, 5])
def test_counter_folding_one_hot_cell(variant):
    """Every point in ONE cell and always stationary: both 16-bit shared fields cross 0x8000 many times."""
    pts = torch.full((4096, 500, 2), 3.3, dtype=torch.float32)
    occ, stat, nd = _gpu_bin(pts, variant)
    c = int(np.floor(np.float32(3.3) / np.float32(0.05)))
    assert nd == 0 and occ[c, c] == 4096 * 500 and stat[c, c] == 4096 * 499
    assert occ.sum() == occ[c, c] and stat.sum() == stat[c, c]


@pytest.mark.parametrize("shape", [(0, 500), (5, 0), (1, 1), (1, 2), (31, 8), (33, 10), (513, 16), (40, 1000)])
def test_ragged_and_empty_shapes(shape):
    n, t = shape
    pts = synth.make_traces(n, t, seed=n + t) if n and t else torch.zeros(n, t, 2)
    b = baseline_ref.OccupancyHeatmapBaseline()
    ref = baseline_ref.bin_points(pts.numpy(), 0, 0, 0.05, 200, 200, b.thr2)
    occ, stat, nd = _gpu_bin(pts)
    assert np.array_equal(occ, ref[0]) and np.array_equal(stat, ref[1]) and nd == ref[2]


def test_unaligned_view_takes_generic_path():
    base = synth.make_traces(9, 65, seed=2).cuda()
    view = base[1:, :, :]                       # odd seq_len -> rows are only 8-byte aligned
    b = OccupancyHeatmapBaseline()
    occ, stat, nd = b.bin(view)
    ref = baseline_ref.bin_points(view.cpu().numpy(), 0, 0, 0.05, 200, 200, np.float32(b.thr2))
    assert np.array_equal(occ.cpu().numpy(), ref[0]) and np.array_equal(stat.cpu().numpy(), ref[1]) and nd == ref[2]


def test_other_grids_and_bounds():
    pts = synth.make_traces(500, 200, seed=4)
    for bounds, res in (((0, 10, 0, 10), 0.1), ((2, 7.3, 1, 9), 0.07), ((0, 10, 0, 10), 0.025), ((-5, 15, -5, 15), 0.05)):
        ob = baseline_ref.OccupancyHeatmapBaseline(bounds=bounds, resolution=res)
        ref = ob.bin(pts.numpy())
        occ, stat, nd = _gpu_bin(pts, bounds=bounds, resolution=res)   # the last two exceed 40960 cells: generic path
        assert np.array_equal(occ, ref[0]) and np.array_equal(stat, ref[1]) and nd == ref[2]


def test_host_buffer_entry_matches():
    pts = synth.make_traces(3000, 500, seed=6)
    b = OccupancyHeatmapBaseline()
    occ, stat, nd = b.bin(pts)                  # CPU tensor -> rs_heatmap_bin_host
    assert not occ.is_cuda
    ref = heatmap_ref_c.bin_points(pts.numpy(), 0, 0, 0.05, 200, 200, np.float32(b.thr2))
    assert np.array_equal(occ.numpy(), ref[0]) and np.array_equal(stat.numpy(), ref[1]) and nd == ref[2]


def test_large_config_properties_and_c_oracle():
    """1/8 of BASELINE config 2 (125k traces x 500): conservation + bit-exact against the C restatement."""
    n, t = 125_000, 500
    pts = synth.make_traces(n, t, seed=0, device="cuda")
    b = OccupancyHeatmapBaseline()
    occ, stat, nd = b.bin(pts)
    assert int(occ.sum().item()) + nd == n * t
    assert bool((stat <= occ).all())
    ref = heatmap_ref_c.bin_points(pts.cpu().numpy(), 0, 0, 0.05, 200, 200, np.float32(b.thr2))
    assert np.array_equal(occ.cpu().numpy(), ref[0]) and np.array_equal(stat.cpu().numpy(), ref[1]) and nd == ref[2]
    # linearity: binning two halves with accumulate == binning the whole
    occ2 = torch.zeros_like(occ); stat2 = torch.zeros_like(stat); d2 = torch.zeros(1, dtype=torch.int64, device="cuda")
    b.bin_into(pts[: n // 2], occ2, stat2, d2, accumulate=True)
    b.bin_into(pts[n // 2:], occ2, stat2, d2, accumulate=True)
    assert torch.equal(occ2, occ) and torch.equal(stat2, stat) and int(d2.item()) == nd


def test_property_random_rooms_cuda_equals_oracles():
    """Hypothesis property test: random bounds / resolutions / stationary speeds, points on
    cell edges, outside the room, NaN, +-Inf, denormals, T = 1, ragged batch sizes: CUDA == numpy oracle == C oracle,
    every point is binned or dropped, and a stationary sample is also a visit."""
    from hypothesis import given, settings, strategies as st
    finite = st.floats(min_value=-40.0, max_value=40.0, allow_nan=False, width=32)
    weird = st.sampled_from([float("nan"), float("inf"), float("-inf"), 0.0, -0.0, 1e-30, -1e-30, 3.0e38])

    @settings(max_examples=40, deadline=None)
    @given(x_min=st.floats(-20, 5), y_min=st.floats(-20, 5), w=st.floats(0.25, 12), h=st.floats(0.25, 12),
           res=st.sampled_from([0.05, 0.1, 0.013, 0.25, 1.0]), v=st.sampled_from([0.1, 0.5, 0.01]),
           n=st.sampled_from([1, 3, 31, 32, 33, 130]), t=st.sampled_from([1, 2, 7, 8, 16, 50]), seed=st.integers(0, 2 ** 31 - 1),
           specials=st.lists(st.tuples(st.integers(0, 129), st.integers(0, 49), st.one_of(weird, finite), st.one_of(weird, finite)), max_size=8),
           on_edges=st.booleans(), variant=st.sampled_from(VARIANTS))
    def check(x_min, y_min, w, h, res, v, n, t, seed, specials, on_edges, variant):
        kw = dict(bounds=(x_min, x_min + w, y_min, y_min + h), resolution=res, stationary_speed=v)
        ref = baseline_ref.OccupancyHeatmapBaseline(**kw)
        if variant in (1, 2, 4, 5) and (t % 2 or ref.gx * ref.gy > 40960):
            variant = 0
        rng = np.random.default_rng(seed)
        pts = np.stack([rng.uniform(x_min - 1, x_min + w + 1, (n, t)), rng.uniform(y_min - 1, y_min + h + 1, (n, t))], -1).astype(np.float32)
        if on_edges:
            k = rng.integers(0, max(1, ref.gx), (n, t))
            pts[..., 0] = (np.float32(x_min) + k.astype(np.float32) * np.float32(res)).astype(np.float32)
            if t > 1:
                pts[:, 1:] = np.where(rng.random((n, t - 1, 1)) < 0.5, pts[:, :-1], pts[:, 1:])      # repeated points: stationary
        for (i, j, px, py) in specials:
            if i < n and j < t:
                pts[i, j] = (px, py)
        o_np = baseline_ref.bin_points(pts, ref.bounds[0], ref.bounds[2], ref.resolution, ref.gx, ref.gy, ref.thr2)
        o_c = heatmap_ref_c.bin_points(pts, ref.bounds[0], ref.bounds[2], ref.resolution, ref.gx, ref.gy, ref.thr2)
        occ, stat, nd = _gpu_bin(pts, variant, **kw)
        assert np.array_equal(occ, o_np[0]) and np.array_equal(stat, o_np[1]) and nd == o_np[2]
        assert np.array_equal(occ, o_c[0]) and np.array_equal(stat, o_c[1]) and nd == o_c[2]
        assert int(occ.sum()) + nd == n * t and (stat <= occ).all()

    check()
