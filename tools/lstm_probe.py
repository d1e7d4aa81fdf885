"""Times the BiLSTM + query-decoder model fwd+bwd with the per-kernel breakdown (scratch)."""
import sys, time, torch
sys.path.insert(0, ".")
from roomslam_b200 import functional as Fn
from roomslam_b200.lstm_model import TraceToColliderLSTM
B, N, D, Q = [int(a) for a in sys.argv[1:5]] if len(sys.argv) > 4 else (256, 3000, 128, 30)
m = TraceToColliderLSTM(D, Q).cuda().eval()
x = torch.randn(B, N, 11, device="cuda"); mask = torch.ones(B, N, dtype=torch.bool, device="cuda")
def step():
    m.zero_grad(set_to_none=True)
    out = m(x, mask)
    (out["pred_boxes"].sum() + out["pred_classes"].sum()).backward()
for _ in range(2): step()
torch.cuda.synchronize(); t0 = time.time()
for _ in range(3): step()
torch.cuda.synchronize(); ms = (time.time() - t0) / 3 * 1e3
print(f"B={B} N={N} D={D} Q={Q}: {ms:.1f} ms/step fwd+bwd  {B/ms*1e3:.0f} traces/s  {B*N/ms/1e3:.1f} Mpts/s  peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
Fn.enable_kernel_timing(True); step(); kt = Fn.collect_kernel_timing(); Fn.enable_kernel_timing(False)
for k, (t, n, fl) in sorted(kt.items(), key=lambda kv: -kv[1][0]): print(f"  {k:36s} {t:8.2f} ms x{n:2d}  {fl/t/1e9 if t else 0:8.2f} TFLOP/s")
with torch.no_grad():
    for _ in range(2): m(x, mask)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(3): m(x, mask)
    torch.cuda.synchronize(); print(f"  inference: {(time.time()-t0)/3*1e3:.1f} ms")
