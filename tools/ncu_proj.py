"""Layer-1 input projection at the benchmark size (8192 traces x 500 steps, K = 256 -> N = 768): the launch ncu captures."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roomslam_b200 import _lib, layout as L

B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 500
tiles = (B + 127) // 128
xt = torch.randn(tiles, T + 2, 32, 128, 8, device="cuda").bfloat16()
w = torch.randn(768, 256, device="cuda") * 0.1
b = torch.randn(768, device="cuda")
ct = torch.empty(tiles, T + 2, 96, 128, 8, device="cuda", dtype=torch.bfloat16)
kch = L.int_array([0, 8, 16, 24])
wt = L.tile_weight_nt(w)
st = torch.cuda.current_stream().cuda_stream
def run():
    _lib.call("rs_blk_gemm_nt", xt.data_ptr(), 256, ctypes.addressof(kch), 4, wt.data_ptr(), 6, ct.data_ptr(), 768, 0, b.data_ptr(),
              tiles * (T + 2), st)
for _ in range(3): run()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
torch.cuda.synchronize(); e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = 2.0 * tiles * 128 * (T + 2) * 256 * 768
by = tiles * (T + 2) * (256 + 768) * 256
print(f"projection B={B}: {ms:.3f} ms  {fl/ms/1e9:.0f} TFLOP/s  {by/ms/1e6:.0f} GB/s")
