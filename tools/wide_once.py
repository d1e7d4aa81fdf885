"""One H = 256 training step at a reduced T (for ncu captures of the wide recurrence kernels; scratch tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roomslam_b200 import RoomSLAM, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
torch.manual_seed(0)
m = RoomSLAM(hidden_size=256, dropout=0.0, precision="bf16").cuda().train()
x, tgt = synth.make_sample(B, T, 10, seed=0, device="cuda")
for i in range(2):
    m.zero_grad(); l = m.compute_loss(m(x), tgt)["total"]; l.backward()
torch.cuda.synchronize(); print("loss", l.item())
