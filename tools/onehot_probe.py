import sys; sys.path.insert(0,'/root/repo')
import torch, numpy as np
from roomslam_b200 import OccupancyHeatmapBaseline
for variant in (1,2,4,5,3):
    for n in (64, 1024, 4096):
        pts = torch.full((n, 500, 2), 3.3, dtype=torch.float32).cuda()
        b = OccupancyHeatmapBaseline(); b._variant = variant
        occ, stat, nd = b.bin(pts)
        c = int(np.floor(np.float32(3.3) / np.float32(0.05)))
        print(variant, n, nd, int(occ[c,c]) - n*500, int(stat[c,c]) - n*499, int(occ.sum()) - n*500, int(stat.sum())-n*499)
