"""H = 256 bf16 path against the fp32 CUDA path (itself 1e-4-pinned to the oracle): quick error table (scratch tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roomslam_b200 import RoomSLAM, synth
def l2(a, b): a, b = a.double(), b.double(); return float((a - b).norm() / b.norm().clamp_min(1e-12))
shapes = [(3, 5, 1), (130, 40, 2), (256, 200, 2)] if len(sys.argv) < 2 else [tuple(int(v) for v in sys.argv[1:4])]
for B, T, L in shapes:
    torch.manual_seed(0)
    ref = RoomSLAM(hidden_size=256, num_layers=L, dropout=0.0, precision="fp32").cuda().train()
    dev = RoomSLAM(hidden_size=256, num_layers=L, dropout=0.0, precision="bf16").cuda().train()
    dev.load_state_dict(ref.state_dict())
    x, tgt = synth.make_sample(B, T, 10, seed=1, device="cuda")
    er, hr = ref.encode(x); ed, hd = dev.encode(x)
    print(f"B={B} T={T} L={L}: out {l2(ed, er):.4f} h_n {l2(hd, hr):.4f}", flush=True)
    lr = ref.compute_loss(ref(x), tgt)["total"]; lr.backward()
    ld = dev.compute_loss(dev(x), tgt)["total"]; ld.backward()
    g = dict(ref.named_parameters())
    worst = sorted(((l2(p.grad, g[n].grad), n) for n, p in dev.named_parameters()), reverse=True)[:4]
    print(f"   loss {float(ld):.6f} vs {float(lr):.6f}; worst grads", ["%s %.4f" % (n, e) for e, n in worst], flush=True)
