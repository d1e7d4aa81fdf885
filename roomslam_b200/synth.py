"""Seeded synthetic traces and object targets (`train.py --create_sample_data`, upstream README.md:64-68).

Upstream ships no generator; this is decision D13 of SURVEY.md 8(a): room [0,10]^2 m,
p_0 ~ U(1,9)^2, alternating move / pause segments with geometric lengths (about 30 % of samples paused),
moving step N(0, 0.05^2) per axis reflected at the walls, fp32, all randomness from one torch.Generator.
Pauses repeat the previous point exactly, so the stationary-time rule (D11) is exercised.

Works on CPU and CUDA tensors (the generator's device decides); streams differ between devices, so
parity fixtures are always generated on the CPU and copied.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

ROOM = (0.0, 10.0, 0.0, 10.0)
P_STOP, P_GO = 0.03, 0.07     # moving->paused, paused->moving per step: paused fraction 0.3
STEP_SIGMA = 0.05


def make_traces(n_traces: int, seq_len: int, seed: int = 0, device: str | torch.device = "cpu",
                out: Optional[torch.Tensor] = None, chunk: int = 131072) -> torch.Tensor:
    """(n_traces, seq_len, 2) float32 traces in metres."""
    device = torch.device(device)
    if out is None:
        out = torch.empty(n_traces, seq_len, 2, dtype=torch.float32, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    lo, hi = ROOM[0], ROOM[1]
    for s in range(0, n_traces, chunk):
        n = min(chunk, n_traces - s)
        p = torch.rand(n, 2, generator=gen, device=device) * 8.0 + 1.0
        paused = torch.rand(n, generator=gen, device=device) < 0.3
        steps = torch.randn(n, seq_len, 2, generator=gen, device=device) * STEP_SIGMA
        flips = torch.rand(n, seq_len, generator=gen, device=device)
        view = out[s:s + n]
        view[:, 0] = p
        for t in range(1, seq_len):
            u = flips[:, t]
            paused = torch.where(paused, u >= P_GO, u < P_STOP)
            q = p + steps[:, t]
            q = q.abs()                              # reflect at 0
            q = hi - (hi - q).abs()                  # reflect at 10
            q = q.clamp_(lo, hi)
            p = torch.where(paused[:, None], p, q)
            view[:, t] = p
    return out


def make_targets(n_traces: int, max_objects: int = 10, num_classes: int = 4, seed: int = 0,
                 device: str | torch.device = "cpu") -> Dict[str, torch.Tensor]:
    """Random object layouts: K ~ U{1..N} valid slots first, classes U{0..C-1}, pos U(0,10)^2,
    size U(0.2,2)^2, orientation U(-pi,pi) (D13; output format README.md:93-108)."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 0x5EED)
    N = max_objects
    k = torch.randint(1, N + 1, (n_traces,), generator=gen, device=device)
    valid = (torch.arange(N, device=device)[None, :] < k[:, None])
    return {
        "classes": torch.randint(0, num_classes, (n_traces, N), generator=gen, device=device),
        "positions": torch.rand(n_traces, N, 2, generator=gen, device=device) * 10.0,
        "sizes": torch.rand(n_traces, N, 2, generator=gen, device=device) * 1.8 + 0.2,
        "orientations": (torch.rand(n_traces, N, generator=gen, device=device) * 2.0 - 1.0) * math.pi,
        "valid": valid.to(torch.float32),
    }


def make_sample(n_traces: int, seq_len: int = 500, max_objects: int = 10, seed: int = 0,
                device: str | torch.device = "cpu"):
    return make_traces(n_traces, seq_len, seed, device), make_targets(n_traces, max_objects, 4, seed, device)
