"""CPU: the binning oracle against the committed golden fixtures, and numpy restatement == C restatement."""
import numpy as np
import pytest

from oracle import baseline_ref, heatmap_ref_c
from roomslam_b200 import synth

SYNTH = ["synth_a", "synth_b", "synth_c", "synth_d", "synth_e"]


def _bin_both(pts, b):
    o1 = baseline_ref.bin_points(pts, b.bounds[0], b.bounds[2], b.resolution, b.gx, b.gy, b.thr2)
    o2 = heatmap_ref_c.bin_points(pts, b.bounds[0], b.bounds[2], b.resolution, b.gx, b.gy, b.thr2)
    return o1, o2


def test_grid_shape_is_robust_to_fp_division():
    assert baseline_ref.grid_shape((0, 10, 0, 10), 0.05) == (200, 200)
    assert baseline_ref.grid_shape((0, 1, 0, 0.3), 0.1) == (3, 10)
    assert baseline_ref.grid_shape((-2.0, 2.5, -6.5, 3.0), 0.05) == (190, 90)


def test_edge_points_match_golden_numpy_and_c(golden_heatmap):
    b = baseline_ref.OccupancyHeatmapBaseline()
    (occ, stat, nd), (occ_c, stat_c, nd_c) = _bin_both(golden_heatmap["edge_points"], b)
    for o, s, n in ((occ, stat, nd), (occ_c, stat_c, nd_c)):
        assert np.array_equal(o, golden_heatmap["edge_occ"])
        assert np.array_equal(s, golden_heatmap["edge_stat"])
        assert n == int(golden_heatmap["edge_dropped"])
    assert occ.sum() + nd == golden_heatmap["edge_points"].shape[0] * golden_heatmap["edge_points"].shape[1]


@pytest.mark.parametrize("name", SYNTH)
def test_synth_matches_golden(golden_heatmap, name):
    n, t, seed = (int(v) for v in golden_heatmap[f"{name}_shape"])
    tr = synth.make_traces(n, t, seed=seed).numpy()
    assert float(tr.astype(np.float64).sum()) == float(golden_heatmap[f"{name}_input_sum"])  # generator is stable
    b = baseline_ref.OccupancyHeatmapBaseline()
    (occ, stat, nd), (occ_c, stat_c, nd_c) = _bin_both(tr, b)
    assert np.array_equal(occ, golden_heatmap[f"{name}_occ"]) and np.array_equal(occ_c, occ)
    assert np.array_equal(stat, golden_heatmap[f"{name}_stat"]) and np.array_equal(stat_c, stat)
    assert nd == nd_c == int(golden_heatmap[f"{name}_dropped"])
    assert occ.sum() + nd == n * t and stat.sum() <= occ.sum()


def test_real_traces_match_golden(golden_heatmap, real_traces):
    b = baseline_ref.OccupancyHeatmapBaseline(bounds=tuple(golden_heatmap["real_bounds"]), resolution=0.05)
    occ, stat, nd = b.bin(real_traces["windows"])
    assert np.array_equal(occ, golden_heatmap["real_occ"]) and np.array_equal(stat, golden_heatmap["real_stat"])
    assert nd == int(golden_heatmap["real_dropped"])
    assert np.array_equal(b.stationary_cells(5.0), golden_heatmap["real_cells_5s"])


def test_stationary_rules():
    b = baseline_ref.OccupancyHeatmapBaseline()
    tr = np.zeros((1, 5, 2), np.float32) + np.float32(2.02)
    tr[0, 3] = (2.5, 2.5)                      # jump: not stationary at t=3; t=4 jumps back: not stationary
    occ, stat, nd = b.bin(tr)
    cell = (int(np.floor(np.float32(2.02) / np.float32(0.05))),) * 2
    assert occ[cell] == 4 and stat[cell] == 2 and nd == 0       # t=0 never stationary; t=1,2 are
    tr[0, 1] = np.nan                           # NaN breaks the chain on both sides
    occ, stat, nd = b.bin(tr)
    assert nd == 1 and stat[cell] == 0
    assert list(b.stationary_cells(0.1)) == [] and b.stationary_cells(0.0).size == 0


def test_empty_and_degenerate_inputs():
    b = baseline_ref.OccupancyHeatmapBaseline()
    for shape in ((0, 500, 2), (3, 0, 2)):
        (occ, stat, nd), (occ_c, stat_c, nd_c) = _bin_both(np.zeros(shape, np.float32), b)
        assert occ.sum() == stat.sum() == nd == 0 and occ_c.sum() == stat_c.sum() == nd_c == 0


def test_c_restatement_threads_agree():
    tr = synth.make_traces(300, 64, seed=9).numpy()
    b = baseline_ref.OccupancyHeatmapBaseline()
    a = heatmap_ref_c.bin_points(tr, 0, 0, 0.05, 200, 200, b.thr2, n_threads=1)
    c = heatmap_ref_c.bin_points(tr, 0, 0, 0.05, 200, 200, b.thr2, n_threads=4)
    assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1]) and a[2] == c[2]
