"""Two bf16 training steps at a given batch (for ncu captures at full scale; scratch tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roomslam_b200 import RoomSLAM, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(0)
m = RoomSLAM(dropout=0.0, precision="bf16").cuda().train()
x, tgt = synth.make_sample(B, 500, 10, seed=0, device="cuda")
for _ in range(2):
    m.zero_grad(); l = m.compute_loss(m(x), tgt)["total"]; l.backward()
torch.cuda.synchronize(); print("loss", l.item())
