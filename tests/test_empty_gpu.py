"""GPU: empty batches go through every public entry point without a launch error and give empty / zero results."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_gru_model_empty_batch(precision):
    from roomslam_b200 import RoomSLAM
    m = RoomSLAM(precision=precision).cuda().eval()
    with torch.no_grad():
        out = m(torch.zeros(0, 500, 2, device="cuda"))
    assert out["class_logits"].shape == (0, 10, 4) and out["positions"].shape == (0, 10, 2)


def test_heatmap_empty_batch():
    from roomslam_b200 import OccupancyHeatmapBaseline
    b = OccupancyHeatmapBaseline()
    occ, stat, dropped = b.bin(torch.zeros(0, 500, 2, device="cuda"))
    assert int(occ.sum()) == 0 and int(stat.sum()) == 0 and dropped == 0


def test_next_rows_empty_batch():
    from roomslam_b200 import preprocess
    from roomslam_b200.evaluation import MetricAccumulator, nms_batch
    from roomslam_b200.lstm_model import build_model
    from roomslam_b200.set_loss import SetCriterion
    out = preprocess.trace_features([], max_len=100)
    assert out["traces"].shape[0] == 0
    model = build_model(num_queries=30, d_model=128).cuda().eval()
    with torch.no_grad():
        pred = model(torch.zeros(0, 40, 11, device="cuda"), torch.zeros(0, 40, dtype=torch.bool, device="cuda"))
    assert pred["pred_boxes"].shape == (0, 30, 6) and pred["pred_classes"].shape == (0, 30, 4)
    tg = {"boxes": torch.zeros(0, 50, 6, device="cuda"), "labels": torch.zeros(0, 50, dtype=torch.long, device="cuda"),
          "valid_mask": torch.zeros(0, 50, dtype=torch.bool, device="cuda")}
    losses = SetCriterion({"class_loss": 2.0})(pred, tg)
    assert float(losses["total_loss"]) == 0.0
    acc = MetricAccumulator("cuda")
    acc.update(pred, tg)
    assert acc.compute()["tp"] == 0
    keep, n, _, _ = nms_batch(pred["pred_boxes"], pred["pred_classes"])
    assert keep.shape == (0, 30) and n.numel() == 0
