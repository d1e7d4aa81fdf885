"""One launch of each SURVEY.md 8(f) kernel at its bench size (for ncu captures; scratch tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roomslam_b200 import preprocess
from roomslam_b200.lstm_model import TraceToColliderLSTM
from roomslam_b200.set_loss import SetCriterion
g = torch.Generator().manual_seed(0)
B, N = 4096, 3000
pts = torch.randn(B, N, 4, generator=g); pts[..., 3] = torch.cumsum(torch.rand(B, N, generator=g) * 0.1 + 1e-3, 1)
preprocess.trace_features(pts.reshape(-1, 4).cuda(), torch.arange(B + 1, dtype=torch.int64) * N, max_len=3000, sort=False)
del pts
B, N, Q, M = 20, 3000, 30, 50
model = TraceToColliderLSTM(128, Q).cuda().train(); model.encoder.dropout = 0.0
x = torch.randn(B, N, 11, generator=g).cuda(); mask = torch.ones(B, N, dtype=torch.bool, device="cuda")
tg = {"boxes": torch.cat([torch.randn(B, M, 3, generator=g), torch.rand(B, M, 3, generator=g) + 0.2], -1).cuda(),
      "labels": torch.randint(0, 4, (B, M), generator=g).cuda(), "valid_mask": (torch.rand(B, M, generator=g) < 0.3).cuda()}
SetCriterion({"class_loss": 2.0, "l1_loss": 5.0, "giou_loss": 2.0})(model(x, mask), tg)["total_loss"].backward()
torch.cuda.synchronize(); print("ok")
