"""Synthetic benchmark workloads from the repo's own generate_maps / generate_tasks drop-ins (BASELINE.json configs).

Config 2/3: 4096 2D tasks = 410 maps x 10 tasks; map m uses random.seed(1000+m), its tasks random.seed(2000+m)
(SURVEY.md §8d).  Config 4/5: 3D maps of 80^3 voxels, 10 tasks per map.  Tasks whose start and goal lie in different
free components are redrawn (the reference pipeline deletes tasks A* cannot solve, generate_paths.py:111-113); the
4-/6-connected component test is the stand-in SURVEY §8d describes.
"""
import random
from dataclasses import dataclass

import numpy as np

from . import ThreeD_generate_maps, ThreeD_generate_tasks, generate_maps, generate_tasks


@dataclass
class Workload:
    dim: int
    shape: tuple
    occ_maps: np.ndarray      # [n_maps, *shape] uint8: 2D 255 = free / 0 = obstacle; 3D 0 = free / 255 = obstacle
    task_to_map: np.ndarray   # [B] int32
    starts: np.ndarray        # [B, dim] int32, reference order: (y, x) in 2D, (x, y, z) in 3D
    goals: np.ndarray         # [B, dim] int32
    coords: np.ndarray        # [B, 1, 2, dim] float32 = the CNN's `coords` input ((x, y) order in 2D, train.py:57-60)
    filtered: bool
    redrawn: int


def _labels(free: np.ndarray):
    from scipy import ndimage
    return ndimage.label(free)[0]  # default structure = 4-connected (2D) / 6-connected (3D)


def make_2d(n_tasks: int, first_map: int = 0, size: int = 512, filter_unreachable: bool = True) -> Workload:
    per = generate_tasks.TASK_NUMBER_FOR_SINGLE_MAP
    n_maps = (n_tasks + per - 1) // per
    maps, t2m, starts, goals, coords, redrawn = [], [], [], [], [], 0
    for m in range(first_map, first_map + n_maps):
        random.seed(1000 + m)
        occ = generate_maps.create_map(size, size)[:, :, 0]
        lab = _labels(occ != 0) if filter_unreachable else None
        random.seed(2000 + m)
        k = 0
        while k < per and len(starts) < n_tasks:
            sx, sy, fx, fy = generate_tasks.draw_start_and_finish_coords(occ)
            if lab is not None and lab[sy, sx] != lab[fy, fx]:
                redrawn += 1
                continue
            starts.append((sy, sx)), goals.append((fy, fx)), coords.append([[sx, sy], [fx, fy]]), t2m.append(m - first_map)
            k += 1
        maps.append(occ)
    return Workload(2, (size, size), np.stack(maps), np.array(t2m, np.int32), np.array(starts, np.int32),
                    np.array(goals, np.int32), np.array(coords, np.float32)[:, None], filter_unreachable, redrawn)


def make_3d(n_tasks: int, first_map: int = 0, size: int = 80, filter_unreachable: bool = True) -> Workload:
    per = ThreeD_generate_tasks.TASK_NUMBER_FOR_SINGLE_MAP
    n_maps = (n_tasks + per - 1) // per
    maps, t2m, starts, goals, coords, redrawn = [], [], [], [], [], 0
    for m in range(first_map, first_map + n_maps):
        random.seed(1000 + m)
        occ = ThreeD_generate_maps.create_map(size, size, size)
        lab = _labels(occ == 0) if filter_unreachable else None
        random.seed(2000 + m)
        k = 0
        while k < per and len(starts) < n_tasks:
            s, f = ThreeD_generate_tasks.draw_start_and_finish_coords(occ)
            if lab is not None and lab[s] != lab[f]:
                redrawn += 1
                continue
            starts.append(s), goals.append(f), coords.append([list(s), list(f)]), t2m.append(m - first_map)
            k += 1
        maps.append(occ.astype(np.uint8))
    return Workload(3, (size, size, size), np.stack(maps), np.array(t2m, np.int32), np.array(starts, np.int32),
                    np.array(goals, np.int32), np.array(coords, np.float32)[:, None], filter_unreachable, redrawn)
