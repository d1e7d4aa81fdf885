"""How much of the bf16-mode gradient error is the bf16 rounding of the WEIGHTS (a perturbed network) and how much the
bf16 activations / gate gradients?  Oracle evaluated at the original fp32 weights vs at the bf16-rounded weights."""
import sys, torch
sys.path.insert(0, ".")
from oracle.room_slam_ref import RoomSLAM as Ref
from roomslam_b200 import RoomSLAM, synth
def l2rel(a, b): a, b = a.double().cpu(), b.double().cpu(); return float((a - b).norm() / max(1e-12, float(b.norm())))
for B, T in ((32, 500), (5, 40), (64, 100)):
    torch.manual_seed(0)
    ref = Ref(hidden_size=128, dropout=0.0).train()
    dev = RoomSLAM(hidden_size=128, dropout=0.0, precision="bf16"); dev.load_state_dict(ref.state_dict()); dev = dev.cuda().train()
    refq = Ref(hidden_size=128, dropout=0.0).train()
    sd = {k: (v.bfloat16().float() if k.startswith("encoder.weight") else v) for k, v in ref.state_dict().items()}
    refq.load_state_dict(sd)
    x, tgt = synth.make_sample(B, T, 10, seed=3)
    for m in (ref, refq): m.compute_loss(m(x), tgt)["total"].backward()
    dev.compute_loss(dev(x.cuda()), {k: v.cuda() for k, v in tgt.items()})["total"].backward()
    g, gq, gd = dict(ref.named_parameters()), dict(refq.named_parameters()), dict(dev.named_parameters())
    enc = [k for k in g if k.startswith("encoder")]
    print(f"B={B} T={T}: worst encoder grad L2-rel  vs fp32-weight oracle {max(l2rel(gd[k].grad, g[k].grad) for k in enc):.4f}   "
          f"vs bf16-weight oracle {max(l2rel(gd[k].grad, gq[k].grad) for k in enc):.4f}   (oracle@bf16 weights vs oracle@fp32 weights: {max(l2rel(gq[k].grad, g[k].grad) for k in enc):.4f})")
