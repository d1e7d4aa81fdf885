"""BASELINE config 4 (H=256, T=4000, batch 1024) in fp32 mode, as micro-batches with gradient accumulation (scratch)."""
import sys, time, torch
sys.path.insert(0, ".")
from roomslam_b200 import RoomSLAM, synth
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 128
m = RoomSLAM(hidden_size=256, dropout=0.0, precision="fp32").cuda().train()
x, tgt = synth.make_sample(mb, 4000, 10, seed=0, device="cuda")
def micro():
    m.compute_loss(m(x), tgt)["total"].backward()
micro(); torch.cuda.synchronize()
m.zero_grad()
t0 = time.time(); micro(); micro(); torch.cuda.synchronize(); dt = (time.time() - t0) / 2
print(f"C4 micro-batch {mb}: {dt*1e3:.0f} ms fwd+bwd -> {mb/dt:.0f} traces/s; batch 1024 = {1024//mb} micro-batches = {dt*1024/mb:.1f} s/step; peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
from roomslam_b200 import functional as Fn
Fn.enable_kernel_timing(True); micro(); kt = Fn.collect_kernel_timing(); Fn.enable_kernel_timing(False)
for k, (t, n, fl) in sorted(kt.items(), key=lambda kv: -kv[1][0]): print(f"  {k:36s} {t:9.1f} ms x{n:2d}  {fl/t/1e9 if t else 0:8.2f} TFLOP/s")
