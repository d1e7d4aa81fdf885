import sys, torch
sys.path.insert(0, ".")
from oracle.room_slam_ref import RoomSLAM as Ref
from roomslam_b200 import RoomSLAM, synth
def l2rel(a, b): a, b = a.double().cpu(), b.double().cpu(); return float((a - b).norm() / max(1e-12, float(b.norm())))
for B, T in ((5, 40), (130, 64)):
    for use_len in (False, True):
        torch.manual_seed(0)
        ref = Ref(hidden_size=128, dropout=0.0).train(); dev = RoomSLAM(hidden_size=128, dropout=0.0, precision="bf16"); dev.load_state_dict(ref.state_dict()); dev = dev.cuda().train()
        x, tgt = synth.make_sample(B, T, 10, seed=3)
        g = torch.Generator().manual_seed(B); lengths = torch.randint(1, T + 1, (B,), generator=g); lengths[0], lengths[-1] = T, 1
        L = lengths if use_len else None
        ref.compute_loss(ref(x, lengths=L), tgt)["total"].backward()
        dev.compute_loss(dev(x.cuda(), lengths=L), {k: v.cuda() for k, v in tgt.items()})["total"].backward()
        worst = max((l2rel(pd.grad, pr.grad), k) for (k, pr), (_, pd) in zip(ref.named_parameters(), dev.named_parameters()))
        print(B, T, "lengths" if use_len else "full", "worst grad L2-rel %.4f %s" % worst)
