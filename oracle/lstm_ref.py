"""TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline may import this; the product path
never does).  torch-CPU restatement of the model the upstream repository actually trains:

  * TraceEncoderRef   <- LSTMTraceEncoder      src/benchmark/model.py:6-57
  * QueryDecoderRef   <- SimpleQueryDecoder    src/benchmark/model.py:60-137
  * TraceToColliderLSTMRef <- TraceToColliderLSTM src/benchmark/model.py:140-153, build_model(model_type='lstm') :406-443

Parameter names equal the reference's state_dict keys, so checkpoints load either way.

PARITY PINNED: tests/golden/lstm.npz holds outputs and parameter gradients of the reference's own build_model(...)
(imported from /root/reference by oracle/make_golden_lstm.py) on seeded weights and inputs; tests/test_oracle_lstm.py
checks this restatement against them.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class _MLP2(nn.Module):
    def __init__(self, d_in, d_hidden, d_out):
        super().__init__()
        self.layers = nn.Sequential(nn.Linear(d_in, d_hidden), nn.ReLU(), nn.Linear(d_hidden, d_out))

    def forward(self, x):
        return self.layers(x)


class TraceEncoderRef(nn.Module):
    def __init__(self, input_dim=11, d_model=128, num_layers=2, dropout=0.1):
        super().__init__()
        self.input_proj = nn.Linear(input_dim, d_model)                                   # model.py:15
        self.lstm = nn.LSTM(d_model, d_model // 2, num_layers=num_layers, batch_first=True, bidirectional=True,
                            dropout=dropout if num_layers > 1 else 0.0)                   # :16-23
        self.out_proj = nn.Linear(d_model, d_model)                                       # :24

    @staticmethod
    def stats(traces, valid):
        """Masked mean of (x, y, z) and RMS of the centred (x, z), floor 1e-3 (model.py:38-46)."""
        xyz = traces[..., :3]
        w = valid.unsqueeze(-1).to(xyz.dtype)
        count = valid.sum(1, keepdim=True).clamp_min(1).unsqueeze(-1)
        mean = (xyz * w).sum(1, keepdim=True) / count
        dev = (xyz - mean) * w
        rms = torch.sqrt((dev[..., [0, 2]] ** 2).sum(dim=(1, 2), keepdim=True) / count).clamp_min(1e-3)
        return mean, rms

    def forward(self, traces, mask=None):
        B, N, _ = traces.shape
        valid = mask if mask is not None else torch.ones(B, N, dtype=torch.bool, device=traces.device)
        mean, rms = self.stats(traces, valid)
        seq, _ = self.lstm(self.input_proj(traces))          # padded steps are NOT skipped (:48-50)
        return self.out_proj(seq), traces[..., :3].contiguous(), mean, rms


class QueryDecoderRef(nn.Module):
    def __init__(self, d_model=128, num_queries=30):
        super().__init__()
        self.num_queries = num_queries
        self.query_embed = nn.Embedding(num_queries, d_model)
        self.q_proj = nn.Linear(d_model, d_model)
        self.k_proj = nn.Linear(d_model, d_model)
        self.v_proj = nn.Linear(d_model, d_model)
        self.scale = d_model ** 0.5
        self.center_delta_head = _MLP2(d_model, d_model, 3)
        self.size_head = _MLP2(d_model, d_model, 3)
        self.class_head = nn.Linear(d_model, 4)
        self.gamma_mlp = nn.Sequential(nn.Linear(d_model, d_model), nn.ReLU(), nn.Linear(d_model, d_model))
        self.beta_mlp = nn.Sequential(nn.Linear(d_model, d_model), nn.ReLU(), nn.Linear(d_model, d_model))
        self.inv_temp = nn.Parameter(torch.tensor(1.0))

    def forward(self, memory, coords, mean, rms, memory_mask=None):
        B = memory.shape[0]
        if memory_mask is None:
            summary = memory.mean(1, keepdim=True)                                        # model.py:103-104
        else:
            cnt = memory_mask.sum(1, keepdim=True).clamp_min(1).unsqueeze(-1)
            summary = (memory * memory_mask.unsqueeze(-1)).sum(1, keepdim=True) / cnt     # :99-101
        gamma, beta = self.gamma_mlp(summary), self.beta_mlp(summary)                     # :106-107
        q = self.q_proj(self.query_embed.weight).unsqueeze(0).expand(B, -1, -1)           # :96,110
        k, v = self.k_proj(memory), self.v_proj(memory)                                   # :111-112
        logits = torch.bmm(q, k.transpose(1, 2)) * self.inv_temp / self.scale             # :113
        if memory_mask is not None:
            logits = logits.masked_fill(~memory_mask.unsqueeze(1), float("-inf"))         # :115-117
        attn = torch.softmax(logits, -1)                                                  # :119
        feat = torch.bmm(attn, v) * (1.0 + gamma) + beta                                  # :120-123
        anchor = torch.bmm(attn, (coords - mean) / rms)                                   # :126-127
        centre = (anchor + self.center_delta_head(feat)) * rms + mean                     # :129,133
        size = (F.softplus(self.size_head(feat)) + 1e-4) * rms                            # :130-131,134
        return torch.cat([centre, size], -1), self.class_head(feat)                       # :136-138


class TraceToColliderLSTMRef(nn.Module):
    def __init__(self, d_model=128, num_queries=30, lstm_layers=2, dropout=0.1):
        super().__init__()
        self.encoder = TraceEncoderRef(11, d_model, lstm_layers, dropout)
        self.decoder = QueryDecoderRef(d_model, num_queries)

    def forward(self, traces, mask=None):
        memory, coords, mean, rms = self.encoder(traces, mask)
        boxes, classes = self.decoder(memory, coords, mean, rms, mask)
        return {"pred_boxes": boxes, "pred_classes": classes}


def seeded_state(model: nn.Module, seed: int, scale: float = 0.15) -> dict:
    """Deterministic weights independent of construction order: uniform(-scale, scale) drawn per parameter from a
    generator seeded with (seed, index of the sorted key)."""
    state = {}
    for idx, key in enumerate(sorted(model.state_dict().keys())):
        ref = model.state_dict()[key]
        g = torch.Generator().manual_seed(seed * 1000 + idx)
        val = (torch.rand(ref.shape, generator=g, dtype=torch.float32) * 2 - 1) * scale
        state[key] = val + 1.0 if key.endswith("inv_temp") else val
    return state
