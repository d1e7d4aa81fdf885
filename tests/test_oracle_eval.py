"""CPU: oracle/eval_ref.py against the reference's own evaluate_metrics / post_process_predictions outputs
(tests/golden/eval.npz, oracle/make_golden_eval.py)."""
import os

import numpy as np
import pytest

from oracle import eval_ref
from oracle.make_golden_eval import CASES, case_inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "eval.npz")


@pytest.mark.parametrize("name", list(CASES))
def test_metrics_and_nms_match_reference(name):
    golden = np.load(GOLDEN)
    boxes, logits, targets = case_inputs(name)
    B = boxes.shape[0]
    total = None
    for lo, hi in ((0, B // 2), (B // 2, B)):
        c = eval_ref.batch_counts({"pred_boxes": boxes[lo:hi], "pred_classes": logits[lo:hi]},
                                  {k: v[lo:hi] for k, v in targets.items()})
        total = c if total is None else {k: total[k] + c[k] for k in c}
    m = eval_ref.metrics_from_counts(total)
    for k in ("tp", "fp", "fn"):
        assert m[k] == int(golden[f"{name}_metric_{k}"])
    for k in ("mIoU", "precision", "recall", "f1", "cls_acc"):
        assert abs(m[k] - float(golden[f"{name}_metric_{k}"])) < 1e-6
    for b in range(B):
        keep, _, _ = eval_ref.nms_order(boxes[b], logits[b])
        want = golden[f"{name}_nms_keep"][b]
        assert np.array_equal(keep, want[want >= 0]), b


def test_average_precision_known_values():
    assert eval_ref.average_precision([0.9, 0.8, 0.7], [1, 1, 1], 3) == pytest.approx(1.0)
    assert eval_ref.average_precision([0.9, 0.8, 0.7, 0.6], [1, 0, 1, 0], 2) == pytest.approx(0.5 * 1.0 + 0.5 * (2 / 3))
    assert eval_ref.average_precision([0.9], [0], 4) == 0.0
    assert np.isnan(eval_ref.average_precision([0.9], [0], 0))
