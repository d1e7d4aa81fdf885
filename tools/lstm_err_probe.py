"""fp64 CPU oracle vs fp32 CPU oracle vs GPU kernels for a large-batch case (scratch)."""
import sys, torch
sys.path.insert(0, ".")
from oracle.lstm_ref import TraceToColliderLSTMRef, seeded_state
from oracle.make_golden_lstm import run
from roomslam_b200.lstm_model import TraceToColliderLSTM
d_model, Q, B, N = [int(a) for a in sys.argv[1:5]] if len(sys.argv) > 4 else (128, 80, 300, 33)
ref = TraceToColliderLSTMRef(d_model, Q).eval(); ref.load_state_dict(seeded_state(ref, 77))
ref64 = TraceToColliderLSTMRef(d_model, Q).eval().double(); ref64.load_state_dict({k: v.double() for k, v in seeded_state(ref, 77).items()})
m = TraceToColliderLSTM(d_model, Q).eval(); m.load_state_dict(seeded_state(m, 77)); m = m.cuda()
g = torch.Generator().manual_seed(N)
traces = torch.randn(B, N, 11, generator=g)
lengths = torch.randint(min(8, N), N + 1, (B,), generator=g); lengths[0] = N
mask = torch.arange(N)[None, :] < lengths[:, None]
traces = traces * mask[..., None]
wb, wc = torch.randn(B, Q, 6, generator=g), torch.randn(B, Q, 4, generator=g)
r32 = run(ref, traces, mask, wb, wc)
r64 = run(ref64, traces.double(), mask, wb.double(), wc.double())
gg = run(m, traces.cuda(), mask.cuda(), wb.cuda(), wc.cuda())
def err(a, b): return float((a.double().cpu() - b).abs().max() / max(1.0, float(b.abs().max())))
print("boxes  cpu32 %.2e gpu %.2e" % (err(r32[0], r64[0]), err(gg[0], r64[0])))
print("class  cpu32 %.2e gpu %.2e" % (err(r32[1], r64[1]), err(gg[1], r64[1])))
worst = sorted(((err(gg[3][k], r64[3][k]), err(r32[3][k], r64[3][k]), k) for k in r64[3]), reverse=True)[:6]
for e_gpu, e_cpu, k in worst: print("grad %-45s gpu %.2e cpu32 %.2e" % (k, e_gpu, e_cpu))
