"""Times rs_trace_features on B x N points (scratch)."""
import sys, torch
sys.path.insert(0, ".")
from roomslam_b200 import preprocess
B, N = 4096, 3000
pts = torch.randn(B, N, 4, device="cuda"); pts[..., 3] = torch.cumsum(torch.rand(B, N, device="cuda") * 0.1, 1)
off = torch.arange(B + 1, dtype=torch.int64) * N
flat = pts.reshape(-1, 4)
for cap in (3000, 1000):
    for _ in range(3): preprocess.trace_features(flat, off, max_len=cap, sort=False)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): out = preprocess.trace_features(flat, off, max_len=cap, sort=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    rows = out["traces"].shape[0] * out["traces"].shape[1]
    byts = rows * 45 + (B * N * 16 if cap >= N else rows * 48)
    print(f"cap {cap}: {ms:.3f} ms  {B*N/ms/1e6:.1f} Gpts/s  {byts/ms/1e6:.0f} GB/s (algorithmic)")
