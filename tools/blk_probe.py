"""Block-GEMM probe (tile-major layout, no-swizzle descriptors) against torch (scratch tool)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roomslam_b200 import _lib, layout as L

torch.manual_seed(0)
dev = "cuda"
st = lambda: torch.cuda.current_stream().cuda_stream

def nt(B, T, Ca, kcols, N, Cc, c_col0, bias=True):
    x = torch.randn(B, T, Ca, device=dev)
    xt = L.to_tile_major(x)
    K = len(kcols) * 64
    w = torch.randn(N, K, device=dev) * 0.1
    b = torch.randn(N, device=dev) if bias else None
    ct = torch.zeros(L.n_tiles(B), T + 2, Cc // 8, 128, 8, device=dev, dtype=torch.bfloat16)
    kch = L.int_array([c // 8 for c in kcols])
    _lib.call("rs_blk_gemm_nt", xt.data_ptr(), Ca, ctypes.addressof(kch), len(kcols), L.tile_weight_nt(w).data_ptr(), N // 128,
              ct.data_ptr(), Cc, c_col0 // 8, b.data_ptr() if bias else 0, xt.shape[0] * xt.shape[1], st())
    torch.cuda.synchronize()
    xa = torch.cat([x[..., c:c + 64] for c in kcols], -1).bfloat16().float()
    ref = xa @ w.bfloat16().float().t() + (b if bias else 0)
    got = L.from_tile_major(ct, B, T).float()[..., c_col0:c_col0 + N]
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    other = L.from_tile_major(ct, B, T).float()
    other[..., c_col0:c_col0 + N] = 0
    print(f"NT B={B} T={T} K={K} N={N} rel_err={err:.3e} stray={other.abs().max().item():.1e}", flush=True)

def tn(B, T, Ca, mcols, Cb, b_col0, n_cols, shift):
    a = torch.randn(B, T, Ca, device=dev); b = torch.randn(B, T, Cb, device=dev)
    at, bt = L.to_tile_major(a), L.to_tile_major(b)
    M = len(mcols) * 128
    C = torch.ones(M, n_cols, device=dev)
    mch = L.int_array([c // 8 for c in mcols]); rows = L.int_array([i * 128 for i in range(len(mcols))])
    _lib.call("rs_blk_gemm_tn_acc", at.data_ptr(), Ca, ctypes.addressof(mch), ctypes.addressof(rows), len(mcols), bt.data_ptr(), Cb,
              b_col0 // 8, n_cols, shift, 0, C.data_ptr(), n_cols, at.shape[0], T, st())
    torch.cuda.synchronize()
    aa = torch.cat([a[..., c:c + 128] for c in mcols], -1).bfloat16().float()          # (B,T,M)
    bb = torch.zeros(B, T, n_cols, device=dev)
    src = b[..., b_col0:b_col0 + n_cols].bfloat16().float()
    if shift == 0: bb = src
    elif shift == -1: bb[:, 1:] = src[:, :-1]
    else: bb[:, :-1] = src[:, 1:]
    ref = 1.0 + torch.einsum("btm,btn->mn", aa, bb)
    err = (C - ref).abs().max().item() / ref.abs().max().item()
    print(f"TN B={B} T={T} M={M} N={n_cols} shift={shift} rel_err={err:.3e}", flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("nt", "all"):
    nt(128, 1, 64, [0], 128, 128, 0, bias=False)
    nt(200, 5, 256, [0, 64, 128, 192], 768, 768, 0)
    nt(300, 7, 1024, [0, 64, 128, 192, 256, 320, 512, 576, 640, 704, 768, 832], 256, 256, 0)
    nt(128, 3, 128, [64], 128, 512, 256)
if which in ("tn", "all"):
    tn(128, 1, 128, [0], 128, 0, 128, 0)
    tn(300, 9, 1024, [0, 128, 256, 512, 640, 768], 256, 0, 256, 0)
    tn(300, 9, 1024, [0, 128, 384], 256, 128, 128, -1)
    tn(300, 9, 1024, [512, 640, 896], 256, 0, 128, 1)
    tn(500, 33, 256, [0, 128], 16, 0, 16, 0)
