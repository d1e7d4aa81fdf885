"""CPU: oracle/lstm_ref.py (restatement of the shipped BiLSTM + query-decoder model) against outputs and gradients of
the reference's own build_model(model_type='lstm') (tests/golden/lstm.npz, oracle/make_golden_lstm.py)."""
import os

import numpy as np
import pytest
import torch

from oracle.lstm_ref import TraceToColliderLSTMRef, seeded_state
from oracle.make_golden_lstm import CASES, case_inputs, run

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "lstm.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("tag", ["mask", "nomask"])
def test_restatement_matches_reference(golden, name, tag):
    d_model, Q, B, N, seed, _ = CASES[name]
    model = TraceToColliderLSTMRef(d_model, Q).eval()
    model.load_state_dict(seeded_state(model, seed))
    traces, mask, wb, wc = case_inputs(name)
    boxes, classes, obj, grads = run(model, traces, mask if tag == "mask" else None, wb, wc)
    np.testing.assert_allclose(boxes.numpy(), golden[f"{name}_{tag}_boxes"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(classes.numpy(), golden[f"{name}_{tag}_classes"], rtol=1e-5, atol=1e-6)
    for k, g in grads.items():
        if f"{name}_{tag}_grad/{k}" in golden.files:
            want = golden[f"{name}_{tag}_grad/{k}"]
            assert np.abs(g.numpy() - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), k
        else:
            assert abs(float(g.double().norm()) - float(golden[f"{name}_{tag}_gradnorm/{k}"])) <= 1e-5 * max(1.0, float(golden[f"{name}_{tag}_gradnorm/{k}"])), k
            np.testing.assert_allclose(g.flatten()[:32].numpy(), golden[f"{name}_{tag}_gradhead/{k}"], rtol=1e-4, atol=1e-5, err_msg=k)


def test_state_dict_keys_are_the_reference_names():
    keys = set(TraceToColliderLSTMRef(64, 5).state_dict().keys())
    for k in ("encoder.input_proj.weight", "encoder.lstm.weight_hh_l1_reverse", "encoder.out_proj.bias",
              "decoder.query_embed.weight", "decoder.center_delta_head.layers.2.weight", "decoder.gamma_mlp.0.bias",
              "decoder.inv_temp", "decoder.class_head.weight"):
        assert k in keys
