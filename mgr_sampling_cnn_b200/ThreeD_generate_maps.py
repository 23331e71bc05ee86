"""Drop-in for the reference's ThreeD_generate_maps.py (3D voxel maps: 0 = free, 255 = obstacle), following
/root/reference/ThreeD_generate_maps.py:5-41 call for call."""
import random

import numpy as np


def create_map(width: int, height: int, depth: int) -> np.ndarray:
    occ_map = np.zeros((height, width, depth), dtype=float)
    for _ in range(random.randint(1, 3)):
        occ_map = add_obstacle(occ_map, 20, 55, 10, 20, 10, 20)   # long in x
    for _ in range(random.randint(1, 3)):
        occ_map = add_obstacle(occ_map, 10, 20, 20, 55, 10, 20)   # long in y
    for _ in range(random.randint(1, 3)):
        occ_map = add_obstacle(occ_map, 10, 20, 10, 20, 20, 55)   # long in z
    for _ in range(random.randint(1, 3)):
        occ_map = add_obstacle(occ_map, 20, 30, 20, 30, 20, 30)   # symmetrical big
    for _ in range(random.randint(5, 15)):
        occ_map = add_obstacle(occ_map, 10, 20, 10, 20, 10, 20)   # symmetrical small
    return occ_map


def add_obstacle(occ_map: np.ndarray, x_min: int, x_max: int, y_min: int, y_max: int, z_min: int, z_max: int) -> np.ndarray:
    x_start = random.randint(0, occ_map.shape[0] - 1)
    y_start = random.randint(0, occ_map.shape[1] - 1)
    z_start = random.randint(0, occ_map.shape[2] - 1)
    x_end = random.randint(x_start + x_min, x_start + x_max)
    y_end = random.randint(y_start + y_min, y_start + y_max)
    z_end = random.randint(z_start + z_min, z_start + z_max)
    occ_map[x_start:x_end, y_start:y_end, z_start:z_end] = 255
    return occ_map
