"""Drop-in for the reference's ThreeD_rrt_star.py (3D uniform RRT*; ThreeD_rrt_star.py:50-62,193-224) on the B200
engine.  occ_map: (80, 80, 80)-like array, 0 = FREE (polarity is flipped w.r.t. 2D, :69,128); positions are (x, y, z)."""
import glob
import random  # noqa: F401
import sys

import numpy as np

from ._planner_base import Node, RRTStarBase, calculate_length, get_from_string  # noqa: F401

MAPS_DIRECTORY = '/home/czarek/mgr/3D_maps/start_finish/*.npy'
MAX_ITERATIONS = 5000
GOAL_THRESHOLD = 3.0


def get_blank_maps_list() -> list:
    return sorted(glob.glob(MAPS_DIRECTORY))


def get_start_finish_coordinates(path: str) -> tuple:
    v = [int(get_from_string(path, a, b)) for a, b in (("_sx", "_sy"), ("_sy", "_sz"), ("_sz", "_fx"), ("_fx", "_fy"),
                                                       ("_fy", "_fz"), ("_fz", ".npy"))]
    return tuple(v[:3]), tuple(v[3:])


class RRTStar(RRTStarBase):
    _dim, _neural = 3, False
    _module = sys.modules[__name__]

    def __init__(self, occ_map, start, goal, max_iterations: int, goal_threshold: float):
        self.map_height, self.map_width, self.map_depth = occ_map.shape
        self._init_common(occ_map, None, start, goal, max_iterations, goal_threshold, 0.0)

    def search_space_volume(self) -> float:
        return self.map_width * self.map_height * self.map_width  # sic (ThreeD_rrt_star.py:187)


def generate_paths(maps=None):
    """The reference's demo (ThreeD_rrt_star.py:262-280): the first task file of MAPS_DIRECTORY -> one plan.  Plotting is a no-op
    (out of scope); returns (map_path, path, iterations) of the task it ran, None when the directory is empty."""
    maps = get_blank_maps_list() if maps is None else maps
    for map_path in maps:
        print(map_path)
        occ_map = np.load(map_path)
        start, finish = get_start_finish_coordinates(map_path)
        rrt = RRTStar(occ_map=occ_map, start=start, goal=finish, max_iterations=MAX_ITERATIONS,
                      goal_threshold=GOAL_THRESHOLD)
        path, iterations = rrt.rrt_star()
        if path:
            print(path)
            rrt.visualize_path(path)
        else:
            print("COULDN'T FIND A PATH FOR THIS EXAMPLE:", map_path)
        return map_path, path, iterations
    return None


if __name__ == '__main__':
    generate_paths()
