"""CPU oracle for the Room-SLAM GRU path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this file.  The product (roomslam_b200/) never does.

PARITY UNPINNED: the upstream repository documents this model (README.md:110-125 architecture + loss,
README.md:147-157 hyper-parameters) but ships no code, tests or golden vectors for it.
The arithmetic therefore lives in third-party torch (requirements.txt:1, `torch>=2.0.0`; pinned here to
the installed 2.11.0): torch.nn.GRU (equations torch/nn/modules/rnn.py:1221-1224, gate order r|z|n
rnn.py:1290-1297), nn.Linear, cross_entropy, l1, binary_cross_entropy_with_logits.  Every free
parameter the README leaves open is fixed by SURVEY.md section 8(a) decisions D1-D9, cited inline.
The nearest shipped analogue is src/benchmark/model.py:6-153 (2-layer bidirectional recurrent
encoder, batch_first, MLP helper :351-369, softplus sizes :129) and src/benchmark/train.py:72-73,
:132, :433-437 (CE + L1 conventions and loss weights).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

LOSS_WEIGHTS = {"class": 2.0, "position": 5.0, "size": 5.0, "orientation": 1.0, "validity": 1.0}  # D9
CLASS_NAMES = ("GROUND", "LOW", "MID", "BLOCK")  # D7, README.md:20-23


class Decoder(nn.Module):
    """MLP decoder with object heads (README.md:117-120; D6; analogue model.py:351-369)."""

    def __init__(self, in_features: int, hidden: int, max_objects: int, num_classes: int):
        super().__init__()
        self.max_objects = max_objects
        self.num_classes = num_classes
        self.trunk = nn.Sequential(
            nn.Linear(in_features, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU()
        )
        self.class_head = nn.Linear(hidden, max_objects * num_classes)
        self.pos_head = nn.Linear(hidden, max_objects * 2)
        self.size_head = nn.Linear(hidden, max_objects * 2)
        self.orient_head = nn.Linear(hidden, max_objects)
        self.valid_head = nn.Linear(hidden, max_objects)

    def forward(self, latent: torch.Tensor) -> Dict[str, torch.Tensor]:
        B, N, C = latent.shape[0], self.max_objects, self.num_classes
        f = self.trunk(latent)
        return {
            "class_logits": self.class_head(f).view(B, N, C),
            "positions": self.pos_head(f).view(B, N, 2),
            "sizes": F.softplus(self.size_head(f)).view(B, N, 2) + 1e-4,  # analogue model.py:129
            "orientations": self.orient_head(f).view(B, N),
            "validity_logits": self.valid_head(f).view(B, N),
        }


class RoomSLAM(nn.Module):
    """Bidirectional-GRU encoder + MLP decoder (README.md:110-120; D1-D6)."""

    def __init__(self, input_size: int = 2, hidden_size: int = 128, num_layers: int = 2,
                 max_objects: int = 10, num_classes: int = 4, dropout: float = 0.1,
                 decoder_hidden: int = 256):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        self.max_objects, self.num_classes, self.dropout = max_objects, num_classes, dropout
        # dropout=0 in the stock module: inter-layer dropout is applied with an explicit mask (D4)
        self.encoder = nn.GRU(input_size, hidden_size, num_layers=num_layers, batch_first=True,
                              bidirectional=True, dropout=0.0)
        self.decoder = Decoder(2 * hidden_size, decoder_hidden, max_objects, num_classes)

    # -- dropout mask (D4) -------------------------------------------------------------------
    def make_dropout_mask(self, batch: int, seq_len: int, generator: Optional[torch.Generator] = None,
                          device=None) -> Optional[torch.Tensor]:
        """Bernoulli keep-mask scaled by 1/(1-p), shape (L-1, B, T, 2H); None when no dropout applies."""
        if self.num_layers < 2 or self.dropout <= 0.0:
            return None
        keep = 1.0 - self.dropout
        shape = (self.num_layers - 1, batch, seq_len, 2 * self.hidden_size)
        m = (torch.rand(shape, generator=generator) < keep).to(torch.float32) / keep
        return m.to(device) if device is not None else m

    def _layer_weights(self, layer: int):
        names = []
        for sfx in ("", "_reverse"):
            names += [f"weight_ih_l{layer}{sfx}", f"weight_hh_l{layer}{sfx}",
                      f"bias_ih_l{layer}{sfx}", f"bias_hh_l{layer}{sfx}"]
        return [getattr(self.encoder, n) for n in names]

    def encode(self, x: torch.Tensor, dropout_mask: Optional[torch.Tensor] = None, lengths: Optional[torch.Tensor] = None):
        """Returns (out (B,T,2H) of the top layer, h_n (2L,B,H)) exactly as nn.GRU would; with `lengths`, exactly as nn.GRU
        on torch.nn.utils.rnn.pack_padded_sequence(x, lengths) would (variable-length traces)."""
        if lengths is not None:
            if dropout_mask is not None or (self.training and self.dropout > 0 and self.num_layers > 1):
                raise NotImplementedError("oracle: lengths are supported without inter-layer dropout")
            packed = torch.nn.utils.rnn.pack_padded_sequence(x, torch.as_tensor(lengths).cpu(), batch_first=True, enforce_sorted=False)
            out, h_n = self.encoder(packed)
            out, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=x.shape[1])
            return out, h_n
        if dropout_mask is None and self.training and self.dropout > 0 and self.num_layers > 1:
            dropout_mask = self.make_dropout_mask(x.shape[0], x.shape[1], device=x.device)
        if dropout_mask is None:
            return self.encoder(x)
        if dropout_mask.dim() == 3:
            dropout_mask = dropout_mask.unsqueeze(0)
        h0 = x.new_zeros(2, x.shape[0], self.hidden_size)  # D3: zero initial state
        inp, h_all = x, []
        for layer in range(self.num_layers):
            # stock ATen GRU on this layer's weights: same kernel nn.GRU dispatches to
            out, h_n = torch._VF.gru(inp, h0, self._layer_weights(layer), True, 1, 0.0,
                                     self.training, True, True)
            h_all.append(h_n)
            inp = out * dropout_mask[layer] if layer < self.num_layers - 1 else out
        return inp, torch.cat(h_all, dim=0)

    def forward(self, x: torch.Tensor, dropout_mask: Optional[torch.Tensor] = None,
                lengths: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        _, h_n = self.encode(x, dropout_mask, lengths)
        latent = torch.cat([h_n[-2], h_n[-1]], dim=-1)  # D5, README.md:115
        return self.decoder(latent)

    # -- loss (README.md:122-125; D8, D9) ----------------------------------------------------
    def compute_loss(self, pred: Dict[str, torch.Tensor], target: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        valid = target["valid"].to(pred["class_logits"].dtype)      # fp32 in use; fp64 under the oracle's own gradcheck
        nv = valid.sum().clamp_min(1.0)
        C = pred["class_logits"].shape[-1]
        ce = F.cross_entropy(pred["class_logits"].reshape(-1, C), target["classes"].reshape(-1).long(),
                             reduction="none").view_as(valid)
        loss_class = (ce * valid).sum() / nv
        loss_pos = ((pred["positions"] - target["positions"]).abs() * valid[..., None]).sum() / (nv * 2.0)
        loss_size = ((pred["sizes"] - target["sizes"]).abs() * valid[..., None]).sum() / (nv * 2.0)
        loss_orient = ((pred["orientations"] - target["orientations"]).abs() * valid).sum() / nv
        loss_valid = F.binary_cross_entropy_with_logits(pred["validity_logits"], valid)
        w = LOSS_WEIGHTS
        total = (w["class"] * loss_class + w["position"] * loss_pos + w["size"] * loss_size
                 + w["orientation"] * loss_orient + w["validity"] * loss_valid)
        return {"total": total, "class": loss_class, "position": loss_pos, "size": loss_size,
                "orientation": loss_orient, "validity": loss_valid}


def param_count(model: nn.Module) -> int:
    return sum(p.numel() for p in model.parameters())


def gru_flops_per_trace(T: int, H: int, L: int, I: int = 2) -> float:
    """Algorithmic forward FLOPs of the encoder per trace."""
    macs = 0
    for layer in range(L):
        il = I if layer == 0 else 2 * H
        macs += 3 * H * il + 3 * H * H
    return 2.0 * 2 * T * macs


if __name__ == "__main__":
    m = RoomSLAM()
    print("params", param_count(m), "enc", param_count(m.encoder), "dec", param_count(m.decoder))
    print("fwd MFLOP/trace", gru_flops_per_trace(500, 128, 2) / 1e6, math.isclose(gru_flops_per_trace(500, 128, 2), 394.752e6))
