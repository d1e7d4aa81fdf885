"""CPU: host-side logic that needs no GPU (flat parameter buffers, bucket order, distributed reducers on gloo)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.room_slam_ref import RoomSLAM as RefRoomSLAM
from roomslam_b200 import RoomSLAM, synth
from roomslam_b200.baseline import grid_shape
from roomslam_b200.train_utils import FlatParams, GradReducer, default_bucket


def test_flat_params_are_views_and_state_dict_roundtrip():
    torch.manual_seed(0)
    m = RoomSLAM()
    ref = RefRoomSLAM()
    flat = FlatParams(m)
    assert flat.numel == sum(p.numel() for p in ref.parameters()) == 555108
    m.load_state_dict(ref.state_dict())                   # copies through the views
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        assert torch.equal(p, q)
        assert flat.flat.data_ptr() <= p.data_ptr() < flat.flat.data_ptr() + 4 * flat.numel
    flat.flat.zero_()
    assert all(float(p.abs().max()) == 0.0 for p in m.parameters())
    # backward order: decoder, then layer 1, then layer 0
    order = [default_bucket(n) for n in flat.names]
    assert order == sorted(order)
    assert flat.names[0].startswith("decoder.") and flat.names[-1].startswith("encoder.") and "_l0" in flat.names[-1]
    assert len(flat.bucket_ranges) == 3 and flat.bucket_ranges[-1][2] == flat.numel


def test_grid_shape_matches_oracle_rule():
    from oracle import baseline_ref
    for b, r in (((0, 10, 0, 10), 0.05), ((-2.0, 2.5, -6.5, 3.0), 0.05), ((0, 1, 0, 0.3), 0.1), ((0, 7.77, 0, 3.21), 0.07)):
        assert grid_shape(b, r) == baseline_ref.grid_shape(b, r)


def test_synth_is_seeded_and_has_pauses():
    a = synth.make_traces(64, 200, seed=3)
    b = synth.make_traces(64, 200, seed=3)
    c = synth.make_traces(64, 200, seed=4)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert float(a.min()) >= 0.0 and float(a.max()) <= 10.0
    paused = (a[:, 1:] == a[:, :-1]).all(-1).float().mean().item()
    assert 0.15 < paused < 0.45
    t = synth.make_targets(64, 10, 4, seed=3)
    assert t["valid"].sum(1).min() >= 1 and set(t["classes"].unique().tolist()) <= {0, 1, 2, 3}


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    m = RoomSLAM()
    flat = FlatParams(m)
    red = GradReducer(flat)
    assert red.world == world and len(red.ranges) == 2
    flat.zero_grad()
    red.prepare()
    # emulate backward: write rank-dependent gradients in backward order and fire the hooks by hand
    for i, p in enumerate(flat.params):
        p.grad.fill_(float(rank + 1))
        red._make_hook(i)(p)
    assert all(red.launched)
    red.finish()
    assert torch.allclose(flat.grad, torch.full_like(flat.grad, sum(range(1, world + 1))))
    # heatmap-style exact integer reduction
    grid = torch.arange(12, dtype=torch.int32).reshape(3, 4) * (rank + 1)
    dist.all_reduce(grid, op=dist.ReduceOp.SUM)
    assert torch.equal(grid, torch.arange(12, dtype=torch.int32).reshape(3, 4) * sum(range(1, world + 1)))
    dist.destroy_process_group()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")


def test_grad_reducer_two_ranks_gloo(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_collider_targets_match_the_reference_loader():
    """roomslam_b200.data.colliders_to_targets against the reference's own _process_colliders on the real collider file
    and on edge cases (tests/golden/colliders.npz, oracle/make_golden_colliders.py)."""
    import json
    import os
    import numpy as np
    from roomslam_b200 import data
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "colliders.npz"))
    for name in ("train", "edge", "overflow", "none"):
        t = data.colliders_to_targets(json.loads(str(g[f"{name}_json"])))
        assert np.array_equal(t["boxes"].numpy(), g[f"{name}_boxes"]), name
        assert np.array_equal(t["labels"].numpy(), g[f"{name}_labels"]) and np.array_equal(t["valid_mask"].numpy(), g[f"{name}_valid"])
