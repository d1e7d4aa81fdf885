"""BASELINE config 4 (H = 256, T = 4000, batch 1024) on the bf16 tensor-core path: step time + per-kernel breakdown (scratch tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roomslam_b200 import RoomSLAM, synth, functional as F_
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
torch.manual_seed(0)
m = RoomSLAM(hidden_size=256, dropout=0.0, precision="bf16").cuda().train()
x, tgt = synth.make_sample(B, T, 10, seed=0, device="cuda")
def step():
    m.zero_grad(); l = m.compute_loss(m(x), tgt)["total"]; l.backward(); return l
step(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2): step()
e1.record(); torch.cuda.synchronize()
F_.enable_kernel_timing(True); step(); k = F_.collect_kernel_timing(); F_.enable_kernel_timing(False)
print("B", B, "T", T, "ms/step %.2f" % (e0.elapsed_time(e1) / 2), {n: round(v[0], 2) for n, v in k.items()}, "peak GB %.1f" % (torch.cuda.max_memory_allocated() / 1e9))
