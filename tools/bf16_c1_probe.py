import sys, torch
sys.path.insert(0, ".")
from oracle.room_slam_ref import RoomSLAM as Ref
from roomslam_b200 import RoomSLAM, synth
def l2rel(a, b): a, b = a.double().cpu(), b.double().cpu(); return float((a - b).norm() / max(1e-12, float(b.norm())))
def mx(a, b): a, b = a.double().cpu(), b.double().cpu(); return float((a - b).abs().max() / max(1e-12, float(b.abs().max())))
for B, T in ((32, 500), (5, 500), (5, 40), (64, 100)):
    torch.manual_seed(0)
    ref = Ref(hidden_size=128, dropout=0.0).train(); dev = RoomSLAM(hidden_size=128, dropout=0.0, precision="bf16"); dev.load_state_dict(ref.state_dict()); dev = dev.cuda().train()
    x, tgt = synth.make_sample(B, T, 10, seed=3)
    lr = ref.compute_loss(ref(x), tgt); lr["total"].backward()
    ld = dev.compute_loss(dev(x.cuda()), {k: v.cuda() for k, v in tgt.items()}); ld["total"].backward()
    rows = sorted(((l2rel(pd.grad, pr.grad), mx(pd.grad, pr.grad), k) for (k, pr), (_, pd) in zip(ref.named_parameters(), dev.named_parameters())), reverse=True)[:3]
    print(f"B={B} T={T} loss rel {abs(float(ld['total'])-float(lr['total']))/float(lr['total']):.2e}", ["%s l2 %.3f max %.3f" % (k, a, b) for a, b, k in rows])
