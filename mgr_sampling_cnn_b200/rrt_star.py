"""Drop-in for the reference's rrt_star.py (2D uniform RRT*): `RRTStar(occ_map, start, goal, max_iterations,
goal_threshold).rrt_star() -> (path, iteration_no)` (rrt_star.py:49-61,184-215), executed on the B200 engine.
occ_map: 2-D array, 0 = obstacle (rrt_star.py:67,120); positions are (y, x)."""
import glob
import random  # noqa: F401  (module attribute = the randomness injection point, as in the reference)
import sys

import cv2

from ._planner_base import Node, RRTStarBase, calculate_length, get_from_string  # noqa: F401

MAPS_DIRECTORY = '/home/czarek/mgr/maps/start_finish/*.png'
MAX_ITERATIONS = 5000
GOAL_THRESHOLD = 5.0


def get_blank_maps_list() -> list:
    return sorted(glob.glob(MAPS_DIRECTORY))


def get_start_finish_coordinates(path: str) -> tuple:
    x_start = int(get_from_string(path, "_sx", "_sy"))
    y_start = int(get_from_string(path, "_sy", "_fx"))
    x_finish = int(get_from_string(path, "_fx", "_fy"))
    y_finish = int(get_from_string(path, "_fy", ".png"))
    return (y_start, x_start), (y_finish, x_finish)


class RRTStar(RRTStarBase):
    _dim, _neural = 2, False
    _module = sys.modules[__name__]

    def __init__(self, occ_map, start, goal, max_iterations: int, goal_threshold: float):
        self.map_height, self.map_width = occ_map.shape
        self._init_common(occ_map, None, start, goal, max_iterations, goal_threshold, 0.0)

    def search_space_volume(self) -> float:
        return self.map_width * self.map_height


def generate_paths(maps=None):
    """The reference's demo (rrt_star.py:277-297): the first task file of MAPS_DIRECTORY -> one plan.  Plotting is a no-op
    (out of scope); returns (map_path, path, iterations) of the task it ran, None when the directory is empty."""
    maps = get_blank_maps_list() if maps is None else maps
    for map_path in maps:
        print(map_path)
        occ_map = cv2.imread(map_path, 0)
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
