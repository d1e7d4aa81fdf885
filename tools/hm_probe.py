"""Times the binning kernel variants with CUDA events (scratch tool; bench.py is the contract)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roomslam_b200 import OccupancyHeatmapBaseline, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
T = 500
pts = synth.make_traces(n, T, seed=0, device="cuda")
torch.cuda.synchronize()
b = OccupancyHeatmapBaseline()
occ = torch.empty(b.gy, b.gx, dtype=torch.int32, device="cuda"); stat = torch.empty_like(occ)
dr = torch.empty(1, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for variant in (1, 4, 5, 2):
    b._variant = variant
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.bin_into(pts, occ, stat, dr); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    best = min(ts[2:]); 
    print(json.dumps({"variant": variant, "n_traces": n, "ms": ts, "gpts": n * T / best / 1e6,
                      "GBs": n * T * 8 / best / 1e6, "frac_6544": n * T * 8 / best / 1e6 / 6544.7,
                      "occ_sum": int(occ.sum()), "stat_sum": int(stat.sum()), "dropped": int(dr)}))
