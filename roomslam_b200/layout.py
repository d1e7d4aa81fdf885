"""Tile-major activation layout helpers (bf16 mode).  See csrc/gemm_blk.cu for the layout's rationale.

(B, T, C) float  <->  (tiles, T+2, C/8, 128, 8) bf16, trace b = tile*128 + row, time row t' = t + 1, zero pad rows.
These conversions are plumbing (torch ops); the hot path keeps everything in the tile-major form.
"""
from __future__ import annotations

import ctypes

import torch

TILE = 128


def n_tiles(B: int) -> int:
    return (B + TILE - 1) // TILE


def empty_tm(B: int, T: int, C: int, device, zero_pads: bool = True) -> torch.Tensor:
    buf = torch.empty(n_tiles(B), T + 2, C // 8, TILE, 8, device=device, dtype=torch.bfloat16)
    if zero_pads:
        buf[:, 0].zero_()
        buf[:, T + 1].zero_()
    return buf


def to_tile_major(x: torch.Tensor) -> torch.Tensor:
    B, T, C = x.shape
    assert C % 8 == 0
    nt = n_tiles(B)
    xp = torch.zeros(nt * TILE, T + 2, C, device=x.device, dtype=torch.bfloat16)
    xp[:B, 1:T + 1] = x.to(torch.bfloat16)
    return xp.view(nt, TILE, T + 2, C // 8, 8).permute(0, 2, 3, 1, 4).contiguous()


def from_tile_major(xt: torch.Tensor, B: int, T: int) -> torch.Tensor:
    nt, Tp, C8, _, _ = xt.shape
    x = xt.permute(0, 3, 1, 2, 4).reshape(nt * TILE, Tp, C8 * 8)
    return x[:B, 1:T + 1]


def tile_weight_nt(w: torch.Tensor) -> torch.Tensor:
    """W [N, K] (N % 128 == 0, K % 64 == 0) -> [N/128][K/64][8][128][8] bf16: B operand pieces of rs_blk_gemm_nt."""
    N, K = w.shape
    return w.to(torch.bfloat16).view(N // 128, 128, K // 64, 8, 8).permute(0, 2, 3, 1, 4).contiguous()


def int_array(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])
