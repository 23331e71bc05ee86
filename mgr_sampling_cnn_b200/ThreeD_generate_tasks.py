"""Drop-in for the reference's ThreeD_generate_tasks.py: free start + free goal at distance >= 20 voxels, following
/root/reference/ThreeD_generate_tasks.py:25-62 and its file-name format `..._sx_sy_sz_fx_fy_fz.npy`."""
import math
import random

import numpy as np

TASK_NUMBER_FOR_SINGLE_MAP = 10


def draw_start_and_finish_coords(occ_map: np.ndarray):
    dimensions = occ_map.shape
    x_start = random.randint(0, dimensions[0] - 1)
    y_start = random.randint(0, dimensions[1] - 1)
    z_start = random.randint(0, dimensions[2] - 1)
    while occ_map[x_start, y_start, z_start] == 255:
        x_start = random.randint(0, dimensions[0] - 1)
        y_start = random.randint(0, dimensions[1] - 1)
        z_start = random.randint(0, dimensions[2] - 1)
        if occ_map[x_start, y_start, z_start] != 255:
            break
    x_finish = random.randint(0, dimensions[0] - 1)
    y_finish = random.randint(0, dimensions[1] - 1)
    z_finish = random.randint(0, dimensions[2] - 1)
    dst = 0.0
    while occ_map[x_finish, y_finish, z_finish] == 255 or dst < 20.0:
        x_finish = random.randint(0, dimensions[0] - 1)
        y_finish = random.randint(0, dimensions[1] - 1)
        z_finish = random.randint(0, dimensions[2] - 1)
        dx, dy, dz = x_start - x_finish, y_start - y_finish, z_start - z_finish
        dst = math.sqrt(float(dx * dx + dy * dy + dz * dz))
        if occ_map[x_finish, y_finish, z_finish] != 255 and dst >= 20.0:
            break
    return (x_start, y_start, z_start), (x_finish, y_finish, z_finish)


def draw_start_and_finish(occ_map: np.ndarray, path: str, number: int) -> str:
    (x_start, y_start, z_start), (x_finish, y_finish, z_finish) = draw_start_and_finish_coords(occ_map)
    new_path = add_to_string(path, 'sx', str(x_start))
    new_path = add_to_string(new_path, 'sy', str(y_start))
    new_path = add_to_string(new_path, 'sz', str(z_start))
    new_path = add_to_string(new_path, 'fx', str(x_finish))
    new_path = add_to_string(new_path, 'fy', str(y_finish))
    new_path = add_to_string(new_path, 'fz', str(z_finish))
    new_path = add_to_string(new_path, 'path', str(number))
    new_path = new_path.replace('blanks', 'start_finish')
    return new_path


def add_to_string(path: str, where: str, to_add: str) -> str:
    start_index = path.find(where)
    end_index = start_index + len(where)
    return path[:end_index] + to_add + path[end_index:]
