"""CPU: oracle/set_loss_ref.py against the reference's own HungarianMatcher + SetCriterion outputs
(tests/golden/set_loss.npz, oracle/make_golden_set_loss.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import set_loss_ref
from oracle.make_golden_set_loss import CASES, case_inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "set_loss.npz")


@pytest.mark.parametrize("name", list(CASES))
def test_restatement_matches_reference(name):
    golden = np.load(GOLDEN)
    boxes, logits, targets = case_inputs(name)
    boxes.requires_grad_(True); logits.requires_grad_(True)
    losses, pairs = set_loss_ref.set_loss({"pred_boxes": boxes, "pred_classes": logits}, targets)
    for b, (p, q) in enumerate(pairs):
        assert np.array_equal(p, golden[f"{name}_pred_idx"][b, : len(p)]) and np.array_equal(q, golden[f"{name}_gt_idx"][b, : len(q)])
        assert (golden[f"{name}_pred_idx"][b, len(p):] == -1).all()
    for k in ("class_loss", "l1_loss", "giou_loss", "total_loss"):
        assert abs(float(losses[k]) - float(golden[f"{name}_{k}"])) <= 1e-6 * max(1.0, abs(float(golden[f"{name}_{k}"])))
    if losses["total_loss"].requires_grad:
        losses["total_loss"].backward()
        np.testing.assert_allclose(boxes.grad.numpy(), golden[f"{name}_dboxes"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(logits.grad.numpy(), golden[f"{name}_dlogits"], rtol=1e-5, atol=1e-7)
