"""tcgen05 GEMM probe: NT and TN modes against torch.matmul (scratch tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roomslam_b200 import _lib

def st(): return torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
dev = "cuda"
def nt(M, N, K, bias=True):
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    C = torch.full((M, N), 7.0, device=dev).bfloat16()
    _lib.call("rs_gemm_bf16_nt", A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N, b.data_ptr() if bias else 0, M, N, K, 0, st())
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t() + (b if bias else 0)
    err = (C.float() - ref).abs().max().item() / ref.abs().max().item()
    print(f"NT M={M} N={N} K={K} rel_err={err:.3e}", flush=True)
    return err
def tn(rows, M, N, a_shift=0, b_shift=0, a_col0=0, b_col0=0, lda=None, ldb=None):
    lda = lda or (a_col0 + M); ldb = ldb or (b_col0 + N)
    A = torch.randn(rows + 3, lda, device=dev).bfloat16(); B = torch.randn(rows + 3, ldb, device=dev).bfloat16()
    C = torch.ones(M, N, device=dev)
    n = rows
    _lib.call("rs_gemm_bf16_tn_acc", A.data_ptr(), lda, A.shape[0], a_col0, a_shift, B.data_ptr(), ldb, B.shape[0], b_col0, b_shift,
              C.data_ptr(), N, M, N, n, st())
    torch.cuda.synchronize()
    Ar = A[a_shift:a_shift + n, a_col0:a_col0 + M].float(); Br = B[b_shift:b_shift + n, b_col0:b_col0 + N].float()
    ref = 1.0 + Ar.t() @ Br
    err = (C - ref).abs().max().item() / ref.abs().max().item()
    print(f"TN rows={rows} M={M} N={N} shifts=({a_shift},{b_shift}) cols=({a_col0},{b_col0}) rel_err={err:.3e}", flush=True)
    return err
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("nt", "all"):
    nt(128, 128, 64, bias=False); nt(128, 128, 256); nt(1000, 256, 128); nt(4096 * 5 + 17, 768, 256)
if which in ("tn", "all"):
    tn(64, 128, 128); tn(1000, 128, 128); tn(5000, 256, 128, 1, 0); tn(70000, 768, 256, 0, 1, 0, 0); tn(3000, 128, 128, 0, 0, 128, 256, 512, 512)
