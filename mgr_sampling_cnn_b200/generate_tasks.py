"""Drop-in for the reference's generate_tasks.py: random free start + free goal at distance >= 100 px, encoded into
the file name (`map_{m}_path{k}_sx{..}_sy{..}_fx{..}_fy{..}.png`).  Follows /root/reference/generate_tasks.py:27-58,76-82
including its `random` call order and its quirk of drawing x from dimensions[0] and y from dimensions[1]."""
import math
import random

import numpy as np

TASK_NUMBER_FOR_SINGLE_MAP = 10


def draw_start_and_finish_coords(occ_map: np.ndarray):
    """The drawing part of draw_start_and_finish: returns (x_start, y_start, x_finish, y_finish)."""
    dimensions = occ_map.shape
    x_start = 0
    y_start = 0
    while occ_map[y_start, x_start] == 0:
        x_start = random.randint(0, dimensions[0] - 1)
        y_start = random.randint(0, dimensions[1] - 1)
        if occ_map[y_start, x_start] != 0:
            break
    x_finish = 0
    y_finish = 0
    dst = 0.0
    while occ_map[y_finish, x_finish] == 0 or dst < 100.0:
        x_finish = random.randint(0, dimensions[0] - 1)
        y_finish = random.randint(0, dimensions[1] - 1)
        dx, dy = x_start - x_finish, y_start - y_finish
        dst = math.sqrt(float(dx * dx + dy * dy))  # scipy euclidean on ints: exact squares, one sqrt
        if occ_map[y_finish, x_finish] != 0 and dst >= 100.0:
            break
    return x_start, y_start, x_finish, y_finish


def draw_start_and_finish(occ_map: np.ndarray, path: str, number: int) -> str:
    x_start, y_start, x_finish, y_finish = draw_start_and_finish_coords(occ_map)
    new_path = add_to_string(path, 'sx', str(x_start))
    new_path = add_to_string(new_path, 'sy', str(y_start))
    new_path = add_to_string(new_path, 'fx', str(x_finish))
    new_path = add_to_string(new_path, 'fy', str(y_finish))
    new_path = add_to_string(new_path, 'path', str(number))
    new_path = new_path.replace('blanks', 'start_finish')
    return new_path


def add_to_string(path: str, where: str, to_add: str) -> str:
    start_index = path.find(where)
    end_index = start_index + len(where)
    return path[:end_index] + to_add + path[end_index:]
