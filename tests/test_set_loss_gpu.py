"""GPU: batched Hungarian matcher + set loss against the reference's own outputs (tests/golden/set_loss.npz) and, on
larger random batches, against scipy.optimize.linear_sum_assignment + the torch oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import set_loss_ref
from oracle.make_golden_set_loss import CASES, case_inputs

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "set_loss.npz")


def to_cuda(targets):
    return {k: v.cuda() for k, v in targets.items()}


@pytest.mark.parametrize("name", list(CASES))
def test_against_reference_golden(name):
    from roomslam_b200.set_loss import SetCriterion
    golden = np.load(GOLDEN)
    boxes, logits, targets = case_inputs(name)
    boxes, logits = boxes.cuda().requires_grad_(True), logits.cuda().requires_grad_(True)
    crit = SetCriterion(dict(set_loss_ref.WEIGHTS))
    tg = to_cuda(targets)
    pairs = crit.matcher(boxes, logits, tg["boxes"], tg["labels"], tg["valid_mask"])
    for b, (p, q) in enumerate(pairs):                         # index-exact
        want_p = golden[f"{name}_pred_idx"][b]
        want_q = golden[f"{name}_gt_idx"][b]
        assert np.array_equal(p, want_p[want_p >= 0]) and np.array_equal(q, want_q[want_q >= 0]), b
    losses = crit({"pred_boxes": boxes, "pred_classes": logits}, tg)
    for k in ("class_loss", "l1_loss", "giou_loss", "total_loss"):
        want = float(golden[f"{name}_{k}"])
        assert abs(float(losses[k].detach()) - want) <= 1e-5 * max(1.0, abs(want)), k
    losses["total_loss"].backward()
    np.testing.assert_allclose(boxes.grad.cpu().numpy(), golden[f"{name}_dboxes"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(logits.grad.cpu().numpy(), golden[f"{name}_dlogits"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("B,Q,M", [(512, 30, 50), (64, 80, 50), (64, 128, 64), (33, 1, 1), (100, 50, 7)])
def test_matches_scipy_on_random_batches(B, Q, M):
    from roomslam_b200.set_loss import HungarianMatcher, SetCriterion
    g = torch.Generator().manual_seed(B + Q)
    boxes = torch.cat([torch.randn(B, Q, 3, generator=g) * 3, torch.rand(B, Q, 3, generator=g) * 2 + 0.05], -1)
    logits = torch.randn(B, Q, 4, generator=g) * 2
    gt_boxes = torch.cat([torch.randn(B, M, 3, generator=g) * 3, torch.rand(B, M, 3, generator=g) * 2 + 0.05], -1)
    labels = torch.randint(0, 4, (B, M), generator=g)
    valid = torch.rand(B, M, generator=g) < torch.rand(B, 1, generator=g)
    targets = {"boxes": gt_boxes, "labels": labels, "valid_mask": valid}
    want = set_loss_ref.match(boxes, logits, gt_boxes, labels, valid)
    got = HungarianMatcher()(boxes.cuda(), logits.cuda(), gt_boxes.cuda(), labels.cuda(), valid.cuda())
    for b in range(B):
        assert np.array_equal(got[b][0], want[b][0]) and np.array_equal(got[b][1], want[b][1]), b
    bx, lg = boxes.clone().requires_grad_(True), logits.clone().requires_grad_(True)
    ref_losses, _ = set_loss_ref.set_loss({"pred_boxes": bx, "pred_classes": lg}, targets, pairs=want)
    ref_losses["total_loss"].backward()
    cb, cl = boxes.cuda().requires_grad_(True), logits.cuda().requires_grad_(True)
    losses = SetCriterion(dict(set_loss_ref.WEIGHTS))({"pred_boxes": cb, "pred_classes": cl}, to_cuda(targets))
    losses["total_loss"].backward()
    for k in ref_losses:
        assert abs(float(losses[k]) - float(ref_losses[k])) <= 1e-5 * max(1.0, abs(float(ref_losses[k]))), k
    np.testing.assert_allclose(cb.grad.cpu().numpy(), bx.grad.numpy(), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(cl.grad.cpu().numpy(), lg.grad.numpy(), rtol=1e-4, atol=1e-7)


def test_individual_losses_backpropagate():
    from roomslam_b200.set_loss import SetCriterion
    boxes, logits, targets = case_inputs("trained")
    bx, lg = boxes.clone().requires_grad_(True), logits.clone().requires_grad_(True)
    ref, pairs = set_loss_ref.set_loss({"pred_boxes": bx, "pred_classes": lg}, targets)
    (ref["giou_loss"] * 3 + ref["class_loss"]).backward()
    cb, cl = boxes.cuda().requires_grad_(True), logits.cuda().requires_grad_(True)
    out = SetCriterion(dict(set_loss_ref.WEIGHTS))({"pred_boxes": cb, "pred_classes": cl}, to_cuda(targets))
    (out["giou_loss"] * 3 + out["class_loss"]).backward()
    np.testing.assert_allclose(cb.grad.cpu().numpy(), bx.grad.numpy(), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(cl.grad.cpu().numpy(), lg.grad.numpy(), rtol=1e-4, atol=1e-7)


def test_non_finite_predictions_do_not_hang():
    from roomslam_b200.set_loss import HungarianMatcher
    boxes, logits, targets = case_inputs("narrow")
    boxes[0] = float("nan")
    tg = to_cuda(targets)
    pairs = HungarianMatcher()(boxes.cuda(), logits.cuda(), tg["boxes"], tg["labels"], tg["valid_mask"])
    assert len(pairs[0][0]) == 0 and len(pairs[1][0]) > 0
