import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.room_slam_ref import RoomSLAM as Ref
from roomslam_b200 import RoomSLAM, synth
def errs(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a-b).abs().max()/b.abs().max().clamp_min(1e-9)), float((a-b).norm()/b.norm().clamp_min(1e-12))
for (B,T,L,um) in [(3,2,2,False),(1,1,2,False),(4,12,1,False)]:
    torch.manual_seed(B+T)
    ref = Ref(num_layers=L, dropout=0.1 if um else 0.0); dev = RoomSLAM(num_layers=L, dropout=ref.dropout, precision="bf16").cuda()
    dev.load_state_dict(ref.state_dict()); ref.train(um); dev.train(um)
    x, tgt = synth.make_sample(B, T, 10, seed=B)
    mask = ref.make_dropout_mask(B, T, torch.Generator().manual_seed(1)) if um else None
    er, hr = ref.encode(x, mask); lr = ref.compute_loss(ref(x, mask), tgt); lr["total"].backward()
    xm = mask.cuda() if um else None
    ed, hd = dev.encode(x.cuda(), xm); ld = dev.compute_loss(dev(x.cuda(), xm), {k: v.cuda() for k, v in tgt.items()}); ld["total"].backward()
    print(f"case B={B} T={T} L={L} mask={um}: out {errs(ed, er)} h_n {errs(hd, hr)} loss {ld['total'].item():.5f} vs {lr['total'].item():.5f}")
    rg = dict(ref.named_parameters())
    for n, p in dev.named_parameters():
        m, l2 = errs(p.grad, rg[n].grad)
        if l2 > 1.2e-2: print(f"   {n:34s} max-rel {m:.4f}  l2-rel {l2:.4f}")
