"""Drop-in for the reference's neural_rrt.py (2D CNN-guided RRT*): `RRTStar(occ_map, heat_map, start, goal,
max_iterations, goal_threshold, neural_bias)` (neural_rrt.py:50-64,215-249) on the B200 engine.
occ_map: (H, W, 3) float image, converted to gray like the reference (:59); heat_map: (1, H, W, 1) clamp(-3, 1) network
output, stored as `heat_map[0] + 10` (:64).  The sampler CDF is built once per instance instead of once per draw."""
import random  # noqa: F401
import sys
from pathlib import Path

import numpy as np

from ._planner_base import Node, RRTStarBase, calculate_length, get_from_string  # noqa: F401
from .train import UNet_cooler  # noqa: F401  (the reference module re-exports the model class)

BASE_PATH = Path('/home/czarek/mgr/eval_data/')
MODEL_PATH = "/home/czarek/mgr/models/sampling_cnn_vol3_32.pth"
MAX_ITERATIONS = 5000
GOAL_THRESHOLD = 5.0


def get_start_finish_coordinates(path: str) -> tuple:
    x_start = int(get_from_string(path, "_sx", "_sy"))
    y_start = int(get_from_string(path, "_sy", "_fx"))
    x_finish = int(get_from_string(path, "_fx", "_fy"))
    y_finish = int(get_from_string(path, "_fy", ".png"))
    return (y_start, x_start), (y_finish, x_finish)


class RRTStar(RRTStarBase):
    _dim, _neural = 2, True
    _module = sys.modules[__name__]

    def __init__(self, occ_map, heat_map, start, goal, max_iterations: int, goal_threshold: float, neural_bias: float):
        gray = np.dot(np.asarray(occ_map)[..., :3], [0.2989, 0.5870, 0.1140])
        self.map_height, self.map_width = gray.shape
        self._init_common(gray, np.asarray(heat_map)[0] + 10, start, goal, max_iterations, goal_threshold, neural_bias)

    def search_space_volume(self) -> float:
        return self.map_width * self.map_height

    def generate_neural_sample(self):
        return self._sample_via_kernel(True)
