"""Drop-in for the reference's generate_maps.py (2D occupancy maps: 255 = free, 0 = obstacle).

Same function names, arguments and `random` call order as /root/reference/generate_maps.py:7-44, so a given
`random.seed(s)` produces the same map.  cv2.rectangle(..., thickness=-1) fills the inclusive box pt1..pt2 clipped to the
image, which is a NumPy slice assignment here (no OpenCV dependency on the planning path).
"""
import random

import numpy as np


def create_map(width: int, height: int) -> np.ndarray:
    occ_map = np.zeros((height, width, 1), np.uint8)
    occ_map.fill(255)
    occ_map[0, :] = 0
    occ_map[:, 0] = 0
    occ_map[-1, :] = 0
    occ_map[:, -1] = 0
    for _ in range(0, random.randint(2, 7)):
        occ_map = add_obstacle(occ_map, 75, 350, 25, 75)     # long in x
    for _ in range(0, random.randint(2, 7)):
        occ_map = add_obstacle(occ_map, 25, 75, 75, 350)     # long in y
    for _ in range(0, random.randint(1, 3)):
        occ_map = add_obstacle(occ_map, 75, 175, 75, 175)    # symmetrical big
    for _ in range(0, random.randint(5, 15)):
        occ_map = add_obstacle(occ_map, 25, 125, 25, 125)    # symmetrical small
    return occ_map


def add_obstacle(occ_map: np.ndarray, x_min: int, x_max: int, y_min: int, y_max: int) -> np.ndarray:
    x_start = random.randint(0, 512)   # placement is drawn from 0..512 whatever the map size (generate_maps.py:35-36)
    y_start = random.randint(0, 512)
    x_end = random.randint(x_start + x_min, x_start + x_max)
    y_end = random.randint(y_start + y_min, y_start + y_max)
    occ_map[max(y_start, 0):y_end + 1, max(x_start, 0):x_end + 1] = 0
    return occ_map
