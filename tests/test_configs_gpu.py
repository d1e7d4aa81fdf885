"""GPU: the other BASELINE.json configurations at sizes the CPU oracle finishes in seconds.
C4 long-trace variant (H=256, T=4000) in fp32; C5 batched inference in bf16; data-parallel pieces on one GPU."""
import pytest
import torch

from oracle.room_slam_ref import RoomSLAM as RefRoomSLAM
from roomslam_b200 import OccupancyHeatmapBaseline, RoomSLAM, synth
from roomslam_b200.train_utils import FlatParams, FusedAdamW, GradReducer

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


def test_c4_long_trace_h256_t4000_fp32():
    torch.manual_seed(0)
    ref = RefRoomSLAM(hidden_size=256, dropout=0.0).eval()
    dev = RoomSLAM(hidden_size=256, dropout=0.0).cuda().eval()
    dev.load_state_dict(ref.state_dict())
    x, tgt = synth.make_sample(2, 4000, 10, seed=1)
    lr = ref.compute_loss(ref(x), tgt); lr["total"].backward()
    ld = dev.compute_loss(dev(x.cuda()), {k: v.cuda() for k, v in tgt.items()}); ld["total"].backward()
    assert abs(ld["total"].item() - lr["total"].item()) <= 1e-4 * abs(lr["total"].item())
    g = dict(ref.named_parameters())
    for n, p in dev.named_parameters():
        assert rel(p.grad, g[n].grad) < 1e-4, n


def test_c5_batched_inference_bf16():
    torch.manual_seed(1)
    ref = RefRoomSLAM(dropout=0.0).eval()
    dev = RoomSLAM(dropout=0.0, precision="bf16").cuda()
    dev.load_state_dict(ref.state_dict())
    x = synth.make_traces(1000, 500, seed=2)                 # host tensor, ragged last chunk (1000 = 3*300 + 100)
    pred = dev.predict(x, batch_size=300)
    assert not pred["positions"].is_cuda and pred["class_logits"].shape == (1000, 10, 4)
    with torch.no_grad():
        want = ref(x[:64])
    for k in want:
        assert rel(pred[k][:64], want[k]) < 2e-2, k
    again = dev.predict(x[600:700], batch_size=64)           # chunking must not change results
    for k in again:
        assert torch.allclose(again[k], pred[k][600:700], rtol=0, atol=0) or rel(again[k], pred[k][600:700]) < 1e-6, k


def test_fused_adamw_matches_torch_adamw_with_clipping():
    """Same gradients in, same parameters out: clip_grad_norm_(1.0) + torch AdamW vs csrc/optim.cu (two steps)."""
    torch.manual_seed(2)
    ref = RefRoomSLAM(dropout=0.0).train()
    dev = RoomSLAM(dropout=0.0).cuda().train()
    dev.load_state_dict(ref.state_dict())
    flat = FlatParams(dev)
    opt = FusedAdamW(flat, lr=1e-3, weight_decay=1e-2, max_grad_norm=1.0)
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-3, weight_decay=1e-2)
    x, tgt = synth.make_sample(16, 64, 10, seed=3)
    rp = dict(ref.named_parameters())
    for _ in range(2):
        ropt.zero_grad(); ref.compute_loss(ref(x), tgt)["total"].backward()
        flat.zero_grad()
        for n, p in dev.named_parameters():          # identical gradients on both sides (Adam is sign-like: tiny
            p.grad.copy_(rp[n].grad)                 # gradient differences would be amplified to +-lr)
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0); ropt.step()
        opt.step()
    for n, p in dev.named_parameters():
        assert (p.detach().cpu() - rp[n].detach()).abs().max() <= 2e-6, n
    assert flat.grad.data_ptr() == dev.decoder.trunk[0].weight.grad.data_ptr()      # grads still live in the flat buffer


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_training_loop_reduces_the_loss(precision):
    """FlatParams + GradReducer (world 1) + FusedAdamW drive a real optimisation: the loss goes down."""
    torch.manual_seed(5)
    dev = RoomSLAM(dropout=0.0, precision=precision).cuda().train()
    flat = FlatParams(dev)
    red = GradReducer(flat)
    opt = FusedAdamW(flat, lr=3e-3, max_grad_norm=1.0)
    x, tgt = synth.make_sample(64, 50, 10, seed=6, device="cuda")
    losses = []
    for _ in range(25):
        flat.zero_grad(); red.prepare()
        loss = dev.compute_loss(dev(x), tgt)["total"]
        loss.backward(); red.finish(); opt.step()
        losses.append(loss.item())
    assert losses[-1] < 0.8 * losses[0], losses


def test_heatmap_shards_sum_to_the_whole():
    """What bin_distributed does across ranks, emulated on one GPU: per-shard grids add up bit-exactly."""
    pts = synth.make_traces(4000, 200, seed=4, device="cuda")
    b = OccupancyHeatmapBaseline()
    occ, stat, nd = b.bin(pts)
    acc_o, acc_s, acc_d = torch.zeros_like(occ), torch.zeros_like(stat), 0
    for shard in pts.chunk(8):
        o, s, d = b.bin(shard.contiguous())
        acc_o += o; acc_s += s; acc_d += d
    assert torch.equal(acc_o, occ) and torch.equal(acc_s, stat) and acc_d == nd


def test_host_batch_prefetcher_round_trip():
    from roomslam_b200 import synth
    from roomslam_b200.train_utils import HostBatchPrefetcher
    pf = HostBatchPrefetcher("cuda")
    batches = []
    for seed in range(3):
        x, tgt = synth.make_sample(64, 500, 10, seed=seed)
        batches.append((x.pin_memory(), {k: v.pin_memory() for k, v in tgt.items()}))
    pf.submit(*batches[0])
    for k in range(3):
        x, tgt = pf.get()
        if k + 1 < 3:
            pf.submit(*batches[k + 1])
        y = (x * 2).sum()                                   # consumer work on the current stream
        assert torch.equal(x.cpu(), batches[k][0]) and all(torch.equal(tgt[n].cpu(), batches[k][1][n]) for n in tgt)
        assert torch.isfinite(y)
    assert not pf.has_pending
    with pytest.raises(RuntimeError):
        pf.get()


def test_precision_auto_picks_kernels_by_batch():
    from roomslam_b200 import RoomSLAM, functional as Fn, synth
    torch.manual_seed(0)
    m = RoomSLAM(dropout=0.0, precision="auto").cuda().train()
    for B, want in ((32, "gru_fwd_f32_kernel"), (256, "rec_fwd_pair_kernel")):
        x, tgt = synth.make_sample(B, 60, 10, seed=1, device="cuda")
        Fn.enable_kernel_timing(True)
        m.compute_loss(m(x), tgt)["total"].backward()
        names = set(Fn.collect_kernel_timing())
        Fn.enable_kernel_timing(False)
        assert want in names, (B, names)
    with pytest.raises(ValueError):
        RoomSLAM(precision="fp16")
