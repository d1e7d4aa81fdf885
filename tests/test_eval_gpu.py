"""GPU: batched evaluation kernels against the reference's own evaluate_metrics / post_process_predictions outputs
(tests/golden/eval.npz) and the oracle (mAP, larger batches)."""
import os

import numpy as np
import pytest
import torch

from oracle import eval_ref
from oracle.make_golden_eval import CASES, case_inputs

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "eval.npz")


def cuda(d):
    return {k: v.cuda() for k, v in d.items()}


@pytest.mark.parametrize("name", list(CASES))
def test_metrics_against_reference_golden(name):
    from roomslam_b200.evaluation import MetricAccumulator, evaluate_metrics
    golden = np.load(GOLDEN)
    boxes, logits, targets = case_inputs(name)
    B = boxes.shape[0]
    acc = MetricAccumulator("cuda", 0.5)
    for lo, hi in ((0, B // 2), (B // 2, B)):
        acc.update({"pred_boxes": boxes[lo:hi].cuda(), "pred_classes": logits[lo:hi].cuda()}, cuda({k: v[lo:hi] for k, v in targets.items()}))
    m = acc.compute()
    for k in ("tp", "fp", "fn"):
        assert m[k] == int(golden[f"{name}_metric_{k}"]), k
    for k in ("mIoU", "precision", "recall", "f1", "cls_acc"):
        assert abs(m[k] - float(golden[f"{name}_metric_{k}"])) < 1e-6, k

    class Stub(torch.nn.Module):
        def forward(self, traces, mask):
            lo, hi = int(traces[0, 0, 0]), int(traces[0, 0, 1])
            return {"pred_boxes": boxes[lo:hi].cuda(), "pred_classes": logits[lo:hi].cuda()}
    loader = [{"traces": torch.tensor([[[float(lo), float(hi)]]]), "trace_mask": torch.ones(1, 1, dtype=torch.bool),
               **{k: v[lo:hi] for k, v in targets.items()}} for lo, hi in ((0, B // 2), (B // 2, B))]
    assert evaluate_metrics(Stub(), loader, "cuda") == m


@pytest.mark.parametrize("name", list(CASES))
def test_nms_against_reference_golden(name):
    from roomslam_b200.evaluation import nms_batch, post_process_predictions
    golden = np.load(GOLDEN)
    boxes, logits, _ = case_inputs(name)
    keep, n, conf, label = nms_batch(boxes.cuda(), logits.cuda(), 0.7, 0.3)
    want = golden[f"{name}_nms_keep"]
    assert np.array_equal(keep.cpu().numpy(), want)
    assert np.array_equal(n.cpu().numpy(), (want >= 0).sum(1))
    preds = post_process_predictions(boxes[1].cuda(), logits[1].cuda())
    assert len(preds) == int((want[1] >= 0).sum()) and all(p["type"] == "BoxCollider" for p in preds)
    if preds:
        q = int(want[1][0])
        assert preds[0]["center"]["x"] == pytest.approx(float(boxes[1, q, 0])) and preds[0]["label"] in ("BLOCK", "LOW", "MID", "HIGH")


def test_map_against_oracle():
    from roomslam_b200.evaluation import ap_flags, mean_average_precision
    for name in CASES:
        boxes, logits, t = case_inputs(name)
        conf_r, label_r, flags_r, n_gt_r = eval_ref.map_flags(boxes, logits, t["boxes"], t["labels"], t["valid_mask"])
        conf, label, flags, n_gt = ap_flags(boxes.cuda(), logits.cuda(), t["boxes"].cuda(), t["labels"].cuda(), t["valid_mask"].cuda())
        assert np.array_equal(flags.cpu().numpy(), flags_r) and np.array_equal(label.cpu().numpy(), label_r)
        assert np.array_equal(n_gt.cpu().numpy(), n_gt_r)
        np.testing.assert_allclose(conf.cpu().numpy(), conf_r, rtol=1e-6)
        want, aps_r = eval_ref.mean_average_precision(boxes, logits, t["boxes"], t["labels"], t["valid_mask"])
        got, aps = mean_average_precision(boxes.cuda(), logits.cuda(), t["boxes"].cuda(), t["labels"].cuda(), t["valid_mask"].cuda())
        assert got == pytest.approx(want, abs=1e-9)
        for a, r in zip(aps, aps_r):
            assert (a != a and r != r) or a == pytest.approx(r, abs=1e-9)


def test_large_batch_counts_against_oracle():
    """4096 scenes x 30 queries x 50 slots: integer counts equal the oracle's (scipy matching + torch IoU)."""
    from roomslam_b200.evaluation import MetricAccumulator
    g = torch.Generator().manual_seed(9)
    B, Q, M = 4096, 30, 50
    gt = torch.cat([torch.randn(B, M, 3, generator=g) * 3, torch.rand(B, M, 3, generator=g) * 2 + 0.3], -1)
    valid = torch.rand(B, M, generator=g) < 0.4
    labels = torch.randint(0, 4, (B, M), generator=g)
    boxes = gt[:, :Q] + torch.randn(B, Q, 6, generator=g) * 0.12
    boxes[..., 3:] = boxes[..., 3:].clamp_min(0.05)
    logits = torch.randn(B, Q, 4, generator=g)
    targets = {"boxes": gt, "labels": labels, "valid_mask": valid}
    acc = MetricAccumulator("cuda")
    acc.update({"pred_boxes": boxes.cuda(), "pred_classes": logits.cuda()}, cuda(targets))
    m = acc.compute()
    want = eval_ref.metrics_from_counts(eval_ref.batch_counts({"pred_boxes": boxes, "pred_classes": logits}, targets))
    for k in ("tp", "fp", "fn"):
        assert m[k] == want[k], k
    for k in ("mIoU", "cls_acc", "f1"):
        assert abs(m[k] - want[k]) < 1e-6, k


def test_slot_evaluator_against_oracle_c5_style():
    """BASELINE config 5 in miniature: chunked inference of the GRU model + slot evaluation, against the torch oracle."""
    from roomslam_b200 import RoomSLAM, synth
    from roomslam_b200.evaluation import SlotEvaluator
    torch.manual_seed(0)
    model = RoomSLAM(precision="bf16").cuda().eval()
    x, tgt = synth.make_sample(3000, 500, 10, seed=4, device="cuda")
    ev = SlotEvaluator(4, 0.5)
    preds = []
    with torch.no_grad():
        for s in range(0, 3000, 1024):
            pred = model(x[s:s + 1024])
            # make the problem non-trivial: pull half of the predictions onto their targets
            pred["positions"] = torch.where(torch.rand_like(pred["positions"]) < 0.5, tgt["positions"][s:s + 1024], pred["positions"])
            pred["sizes"] = torch.where(torch.rand_like(pred["sizes"]) < 0.7, tgt["sizes"][s:s + 1024] * 1.05, pred["sizes"])
            ev.update(pred, {k: v[s:s + 1024] for k, v in tgt.items()})
            preds.append({k: v.float().cpu() for k, v in pred.items()})
    got = ev.compute()
    cat = {k: torch.cat([p[k] for p in preds]) for k in preds[0]}
    want = eval_ref.slot_eval(cat, {k: v.cpu() for k, v in tgt.items()})
    assert got["n_slots"] == 30000
    for k in ("mean_iou", "class_accuracy", "validity_accuracy", "precision", "recall", "mAP"):
        assert got[k] == pytest.approx(want[k], abs=2e-6), k
