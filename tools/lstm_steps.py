"""Three training steps of the BiLSTM pipeline at the reference batch (for the ncu launch list; scratch tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roomslam_b200.lstm_model import TraceToColliderLSTM
from roomslam_b200.set_loss import SetCriterion
g = torch.Generator().manual_seed(0)
B, N, Q, M = int(sys.argv[1]) if len(sys.argv) > 1 else 20, 3000, 30, 50
model = TraceToColliderLSTM(128, Q).cuda().train(); model.encoder.dropout = 0.0
x = torch.randn(B, N, 11, generator=g).cuda(); mask = torch.ones(B, N, dtype=torch.bool, device="cuda")
tg = {"boxes": torch.cat([torch.randn(B, M, 3, generator=g), torch.rand(B, M, 3, generator=g) + 0.2], -1).cuda(),
      "labels": torch.randint(0, 4, (B, M), generator=g).cuda(), "valid_mask": (torch.rand(B, M, generator=g) < 0.3).cuda()}
crit = SetCriterion({"class_loss": 2.0, "l1_loss": 5.0, "giou_loss": 2.0})
for _ in range(3):
    model.zero_grad(set_to_none=True)
    crit(model(x, mask), tg)["total_loss"].backward()
torch.cuda.synchronize(); print("ok")
