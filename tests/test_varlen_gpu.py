"""GPU: variable-length traces (SURVEY.md 8(f) rank 1 "masked / packed variable-length GRU") against torch.nn.GRU on a
packed sequence (torch.nn.utils.rnn.pack_padded_sequence): outputs past a trace's end are zero, h_n is the state at its
last valid step, padded steps carry no gradient.  fp32 1e-4; bf16 2e-2 (L2-relative on the gradients; batches of >= 128
traces: with a handful of traces the bf16 rounding noise of the saved gate gradients does not average out, with or without
lengths -- tools/varlen_probe.py)."""
import numpy as np
import pytest
import torch

from oracle.room_slam_ref import RoomSLAM as Ref

pytestmark = pytest.mark.gpu


def build(precision, hidden=128, seed=0):
    from roomslam_b200 import RoomSLAM
    torch.manual_seed(seed)
    ref = Ref(hidden_size=hidden, dropout=0.0).train()
    dev = RoomSLAM(hidden_size=hidden, dropout=0.0, precision=precision)
    dev.load_state_dict(ref.state_dict())
    return ref, dev.cuda().train()


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / max(1.0, float(b.abs().max())))


def l2rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / max(1e-12, float(b.norm())))


@pytest.mark.parametrize("precision,B,T,hidden", [("fp32", 5, 40, 128), ("fp32", 9, 33, 64), ("bf16", 300, 37, 128), ("bf16", 130, 64, 128)])
def test_packed_sequence_semantics(precision, B, T, hidden):
    from roomslam_b200 import synth
    ref, dev = build(precision, hidden)
    x, tgt = synth.make_sample(B, T, 10, seed=3)
    g = torch.Generator().manual_seed(B)
    lengths = torch.randint(1, T + 1, (B,), generator=g)
    lengths[0], lengths[-1] = T, 1
    out_r, hn_r = ref.encode(x, lengths=lengths)
    out_d, hn_d = dev.encode(x.cuda(), lengths=lengths)
    tol = 1e-4 if precision == "fp32" else 2e-2
    assert rel(out_d, out_r) < tol and rel(hn_d, hn_r) < tol
    pad = torch.arange(T)[None, :] >= lengths[:, None]
    assert float(out_d.cpu()[pad].abs().max()) == 0.0                       # exactly zero past the end
    # loss + gradients through the padded batch
    loss_r = ref.compute_loss(ref(x, lengths=lengths), tgt)["total"]
    loss_r.backward()
    loss_d = dev.compute_loss(dev(x.cuda(), lengths=lengths), {k: v.cuda() for k, v in tgt.items()})["total"]
    loss_d.backward()
    assert abs(float(loss_d.detach()) - float(loss_r.detach())) < tol * max(1.0, abs(float(loss_r.detach())))
    for (k, pr), (_, pd) in zip(ref.named_parameters(), dev.named_parameters()):
        if precision == "fp32":
            assert rel(pd.grad, pr.grad) < tol, k
        else:
            assert l2rel(pd.grad, pr.grad) < tol, k


def test_full_lengths_equal_no_lengths():
    from roomslam_b200 import synth
    _, dev = build("bf16")
    x, _ = synth.make_sample(64, 50, 10, seed=1, device="cuda")
    with torch.no_grad():
        a = dev.encode(x)[1]
        b = dev.encode(x, lengths=torch.full((64,), 50))[1]
    assert torch.equal(a, b)


def test_lengths_are_validated():
    from roomslam_b200 import synth
    _, dev = build("fp32")
    x, _ = synth.make_sample(4, 10, 10, seed=1, device="cuda")
    for bad in (torch.tensor([10, 0, 3, 4]), torch.tensor([11, 1, 3, 4]), torch.tensor([1, 2, 3])):
        with pytest.raises(ValueError):
            dev(x, lengths=bad)
