"""Generates tests/golden/lstm.npz from the REFERENCE's own model (src/benchmark/model.py build_model(model_type='lstm')),
imported from /root/reference in the authoring container.

    python -m oracle.make_golden_lstm [--reference /root/reference]

Weights come from oracle.lstm_ref.seeded_state (a pure function of (seed, sorted state_dict keys)), so the GPU box can
rebuild them without the reference.  Inputs: real trace features (tests/golden/features.npz through the feature
oracle) cut into ragged windows, and seeded noise.  The scalar objective is sum(pred_boxes * Wb) + sum(pred_classes * Wc)
with seeded Wb, Wc, so every parameter receives a gradient without needing the Hungarian matcher.
Stored: outputs, the objective, and per-parameter gradients (complete for the small masked case; L2 norm + first 32
entries otherwise).  Dropout is off (model.eval() leaves nn.LSTM's inter-layer dropout inactive).
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import features_ref  # noqa: E402
from oracle.lstm_ref import TraceToColliderLSTMRef, seeded_state  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = {  # name: (d_model, num_queries, B, N, seed, input kind)
    "small": (64, 7, 3, 50, 11, "noise"),
    "full": (128, 30, 2, 300, 12, "real"),
}


def case_inputs(name):
    d_model, Q, B, N, seed, kind = CASES[name]
    g = torch.Generator().manual_seed(seed)
    if kind == "noise":
        traces = torch.randn(B, N, 11, generator=g)
        lengths = [N, N - 13, 1][:B]
    else:
        pts = np.load(os.path.join(GOLDEN, "features.npz"))["real0_points"]
        feats = [features_ref.process_points(pts[0:N]), features_ref.process_points(pts[N:2 * N - 50])]
        batch, _ = features_ref.collate(feats)
        traces = torch.from_numpy(batch)
        lengths = [f.shape[0] for f in feats]
    mask = torch.zeros(B, N, dtype=torch.bool)
    for b, L in enumerate(lengths):
        mask[b, :L] = True
        traces[b, L:] = 0
    wb = torch.randn(B, Q, 6, generator=g)
    wc = torch.randn(B, Q, 4, generator=g)
    return traces, mask, wb, wc


def run(model, traces, mask, wb, wc):
    model.zero_grad()
    out = model(traces, mask)
    obj = (out["pred_boxes"] * wb).sum() + (out["pred_classes"] * wc).sum()
    obj.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    return out["pred_boxes"].detach(), out["pred_classes"].detach(), obj.detach(), grads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    sys.path.insert(0, os.path.join(args.reference, "src", "benchmark"))
    import model as ref_model

    out = {}
    for name, (d_model, Q, B, N, seed, kind) in CASES.items():
        with contextlib.redirect_stdout(io.StringIO()):
            ref = ref_model.build_model(num_queries=Q, d_model=d_model, model_type="lstm")
        mine = TraceToColliderLSTMRef(d_model, Q)
        assert sorted(ref.state_dict().keys()) == sorted(mine.state_dict().keys()), "state_dict keys differ from the reference"
        state = seeded_state(mine, seed)
        ref.load_state_dict(state, strict=True)
        ref.eval()
        traces, mask, wb, wc = case_inputs(name)
        for tag, m in (("mask", mask), ("nomask", None)):
            boxes, classes, obj, grads = run(ref, traces, m, wb, wc)
            out[f"{name}_{tag}_boxes"] = boxes.numpy()
            out[f"{name}_{tag}_classes"] = classes.numpy()
            out[f"{name}_{tag}_objective"] = obj.numpy()
            for k, gval in grads.items():
                if name == "small" and tag == "mask":
                    out[f"{name}_{tag}_grad/{k}"] = gval.numpy()
                else:
                    out[f"{name}_{tag}_gradnorm/{k}"] = np.float64(gval.double().norm())
                    out[f"{name}_{tag}_gradhead/{k}"] = gval.flatten()[:32].numpy()
            print(name, tag, "objective", float(obj), "boxes", tuple(boxes.shape))
    np.savez_compressed(os.path.join(GOLDEN, "lstm.npz"), **out)
    print("wrote lstm.npz", os.path.getsize(os.path.join(GOLDEN, "lstm.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
