"""roomslam_b200: B200-native (sm_100a) implementation of Room-SLAM's data-parallel hot path.

Public surface mirrors the upstream README's `src/models/room_slam.py` and `src/models/baseline.py`:
`RoomSLAM` (nn.Module: bi-GRU encoder + MLP decoder heads + multi-task loss) and `OccupancyHeatmapBaseline`.
The callers either side of that path mirror the shipped `src/benchmark/` code:
`preprocess.trace_features`, `lstm_model.build_model`, `set_loss.SetCriterion` / `HungarianMatcher`,
`evaluation.evaluate_metrics` / `post_process_predictions` / `mean_average_precision`.
All arithmetic runs in hand-written CUDA behind the C ABI in include/roomslam_b200.h; no CPU fallback.
"""
from .baseline import OccupancyHeatmapBaseline  # noqa: F401
from . import synth  # noqa: F401

from .model import RoomSLAM  # noqa: F401

__all__ = ["OccupancyHeatmapBaseline", "RoomSLAM", "synth"]
