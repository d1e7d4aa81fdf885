"""GPU: the upstream training recipe end to end on the library (src/benchmark/train.py:190-232,440-444): raw points ->
rs_trace_features -> BiLSTM query decoder -> Hungarian set loss -> backward -> clip -> AdamW; then evaluation."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_training_loop_reduces_the_set_loss_and_evaluates():
    from roomslam_b200 import preprocess
    from roomslam_b200.evaluation import MetricAccumulator, mean_average_precision, nms_batch
    from roomslam_b200.lstm_model import build_model
    from roomslam_b200.set_loss import SetCriterion
    rng = np.random.default_rng(0)
    traces = []
    for n in (400, 777, 650, 512):                                    # ragged raw traces (x, y, z, timestamp)
        t = np.cumsum(rng.uniform(0.01, 0.05, n))
        traces.append(np.stack([np.cumsum(rng.normal(0, 0.03, n)), 1.6 + rng.normal(0, 0.01, n), np.cumsum(rng.normal(0, 0.03, n)), t], 1).astype(np.float32))
    batch = preprocess.trace_features(traces, max_len=600)
    x, mask = batch["traces"], batch["trace_mask"]
    assert x.shape == (4, 600, 11) and mask.sum(1).tolist() == [400, 600, 600, 512]
    B, M = 4, 50
    g = torch.Generator().manual_seed(1)
    gt = torch.cat([torch.randn(B, M, 3, generator=g), torch.rand(B, M, 3, generator=g) * 0.5 + 0.2], -1)
    valid = torch.zeros(B, M, dtype=torch.bool); valid[:, :6] = True
    targets = {"boxes": (gt * valid[..., None]).cuda(), "labels": torch.randint(0, 4, (B, M), generator=g).cuda(), "valid_mask": valid.cuda()}
    torch.manual_seed(0)
    model = build_model(num_queries=30, d_model=128, model_type="lstm", dropout=0.0).cuda().train()
    crit = SetCriterion({"class_loss": 2.0, "l1_loss": 5.0, "giou_loss": 2.0})
    opt = torch.optim.AdamW(model.parameters(), lr=2e-3, weight_decay=1e-4)
    first = last = None
    for it in range(40):
        opt.zero_grad()
        losses = crit(model(x, mask), targets)
        losses["total_loss"].backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        last = float(losses["total_loss"].detach())
        first = last if first is None else first
    assert np.isfinite(last) and last < 0.85 * first, (first, last)
    model.eval()
    with torch.no_grad():
        out = model(x, mask)
    acc = MetricAccumulator("cuda")
    acc.update(out, targets)
    m = acc.compute()
    assert m["tp"] + m["fp"] == 24 and m["fn"] == 0 and 0.0 <= m["mIoU"] <= 1.0
    keep, n, conf, label = nms_batch(out["pred_boxes"], out["pred_classes"])
    assert keep.shape == (4, 30) and int(n.max()) <= 30
    mAP, aps = mean_average_precision(out["pred_boxes"], out["pred_classes"], targets["boxes"], targets["labels"], targets["valid_mask"])
    assert 0.0 <= mAP <= 1.0


def test_real_traces_and_real_colliders_overfit():
    """The upstream data end to end: three recorded traces (tests/golden/features.npz) and the recorded collider set
    (tests/golden/colliders.npz) -> features -> BiLSTM -> Hungarian set loss; 60 AdamW steps must fit the scene."""
    import json
    import os
    from roomslam_b200 import data, preprocess
    from roomslam_b200.evaluation import evaluate_metrics
    from roomslam_b200.lstm_model import build_model
    from roomslam_b200.set_loss import SetCriterion
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    f = np.load(os.path.join(gdir, "features.npz"))
    c = np.load(os.path.join(gdir, "colliders.npz"))
    batch = preprocess.trace_features([f[f"real{k}_points"] for k in range(3)], max_len=800)
    x, mask = batch["traces"], batch["trace_mask"]
    assert x.shape == (3, 800, 11) and bool(mask.all())
    tgt = data.colliders_to_targets(json.loads(str(c["train_json"])))
    targets = {k: v.unsqueeze(0).expand(3, *v.shape).contiguous().cuda() for k, v in tgt.items()}
    torch.manual_seed(1)
    model = build_model(num_queries=30, d_model=128, model_type="lstm", dropout=0.0).cuda().train()
    crit = SetCriterion({"class_loss": 2.0, "l1_loss": 5.0, "giou_loss": 2.0})
    opt = torch.optim.AdamW(model.parameters(), lr=3e-3, weight_decay=1e-4)
    hist = []
    for it in range(60):
        opt.zero_grad()
        losses = crit(model(x, mask), targets)
        losses["total_loss"].backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        hist.append(float(losses["total_loss"].detach()))
    assert np.isfinite(hist[-1]) and hist[-1] < 0.75 * hist[0], (hist[0], hist[-1])
    loader = [{"traces": x.cpu(), "trace_mask": mask.cpu(), **{k: v.cpu() for k, v in targets.items()}}]
    m = evaluate_metrics(model, loader, "cuda")
    assert m["tp"] + m["fp"] == 33 and m["fn"] == 0 and m["cls_acc"] > 0.5


def test_train_benchmark_driver_on_files(tmp_path):
    """train_benchmark.py end to end on trace / collider JSON files in the upstream dataset format."""
    import json
    import os
    import subprocess
    import sys
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    f = np.load(os.path.join(gdir, "features.npz"))
    cols = json.loads(str(np.load(os.path.join(gdir, "colliders.npz"))["train_json"]))
    for split, ks in (("train", (0, 1)), ("val", (2,))):
        d = tmp_path / split
        d.mkdir()
        json.dump({"colliders": cols}, open(d / "colliders.json", "w"))
        for k in ks:
            pts = f[f"real{k}_points"][:900]
            json.dump([{"timestamp": float(p[3]), "x": float(p[0]), "y": float(p[1]), "z": float(p[2])} for p in pts],
                      open(d / f"human_data_{k}.json", "w"))
    root = os.path.dirname(os.path.dirname(__file__))
    out = subprocess.run([sys.executable, os.path.join(root, "train_benchmark.py"), "--data_dir", str(tmp_path / "train"),
                          "--val_dir", str(tmp_path / "val"), "--save_dir", str(tmp_path / "ckpt"), "--epochs", "3",
                          "--max_trace_len", "600"], capture_output=True, text=True, cwd=root, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Epoch 2:" in out.stdout and "Training completed!" in out.stdout
    ck = torch.load(tmp_path / "ckpt" / "best_model.pth", map_location="cpu", weights_only=False)
    assert {"epoch", "model_state_dict", "optimizer_state_dict", "val_loss", "metrics", "config"} <= set(ck)
    assert "encoder.lstm.weight_hh_l1_reverse" in ck["model_state_dict"] and ck["config"]["model_type"] == "lstm"
    # the shipped inference CLI on that checkpoint (threshold 0 keeps every query that survives NMS)
    res = tmp_path / "pred.json"
    out = subprocess.run([sys.executable, os.path.join(root, "inference_benchmark.py"), "--checkpoint", str(tmp_path / "ckpt" / "best_model.pth"),
                          "--input", str(tmp_path / "val" / "human_data_2.json"), "--output", str(res), "--threshold", "0.0"],
                         capture_output=True, text=True, cwd=root, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    pred = json.load(open(res))
    assert pred["metadata"]["num_colliders"] == len(pred["colliders"]) >= 1
    assert set(pred["colliders"][0]) == {"type", "label", "confidence", "center", "size", "radius", "height"}
