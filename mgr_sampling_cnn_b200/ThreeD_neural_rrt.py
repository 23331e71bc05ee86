"""Drop-in for the reference's ThreeD_neural_rrt.py (3D CNN-guided RRT*; ThreeD_neural_rrt.py:54-68,233-265) on the
B200 engine.  heat_map: normalised + thresholded network output, stored as `heat_map - 250` (:68); neural samples are
redrawn until they hit a free voxel (:80-108) — bounded here (NRRT_SAMPLER_STUCK) where the reference would spin."""
import random  # noqa: F401
import sys
from pathlib import Path

import numpy as np

from ._planner_base import Node, RRTStarBase, calculate_length, get_from_string  # noqa: F401
from .ThreeD_train import ThreeD_UNet_cooler  # noqa: F401

BASE_PATH = Path('/home/czarek/mgr/3D_eval_data/test/')
MODEL_PATH = "/home/czarek/mgr/models/3D_sampling_cnn_vol2_47.pth"
MAX_ITERATIONS = 5000
GOAL_THRESHOLD = 3.0


def get_start_finish_coordinates(path: str) -> tuple:
    v = [int(get_from_string(path, a, b)) for a, b in (("_sx", "_sy"), ("_sy", "_sz"), ("_sz", "_fx"), ("_fx", "_fy"),
                                                       ("_fy", "_fz"), ("_fz", ".npy"))]
    return tuple(v[:3]), tuple(v[3:])


class RRTStar(RRTStarBase):
    _dim, _neural = 3, True
    _module = sys.modules[__name__]

    def __init__(self, occ_map, heat_map, start, goal, max_iterations: int, goal_threshold: float, neural_bias: float):
        self.map_height, self.map_width, self.map_depth = occ_map.shape
        self._init_common(occ_map, np.asarray(heat_map) - 250, start, goal, max_iterations, goal_threshold, neural_bias)

    def search_space_volume(self) -> float:
        return self.map_width * self.map_height * self.map_width

    def generate_neural_sample(self):
        return self._sample_via_kernel(True)
