"""TEST INFRASTRUCTURE — the whole hot path on host cores, built from the oracle pieces (bench.py's cpu_baseline and
`--impl reference` legs; never imported by the product).

The reference is pure Python (CPython + NumPy + SciPy + torch-CPU) and cannot travel to the GPU box, so the CPU arm is
the oracle PORT: torch-CPU fp32 forward (oracle/cnn_oracle.py, same ATen kernels the reference's model would run),
then the C restatement of the sampler + planner (oracle/planner_oracle.c) fanned out over host threads.
"""
import os
import time

import numpy as np
import torch

from . import cnn_oracle
from . import planner_oracle as po


def heat_from_logits(logits: np.ndarray, dim: int) -> np.ndarray:
    """Planner-side heat map from raw network output (neural_rrt.py:351,:64; ThreeD_neural_rrt.py:339-346,:68)."""
    if dim == 2:
        return (np.clip(logits, -3, 1) + 10).astype(np.float32)
    v = logits
    with np.errstate(invalid="ignore", divide="ignore"):
        v = ((v - v.min()) / (v.max() - v.min())) * 1.0
    masked = v.copy()
    masked[~(v > 0.96)] = 0
    return (masked - 250).astype(np.float32)


def run(state_dict, dim, shape, occ_maps, task_to_map, starts, goals, coords, seed, task_ids, max_iterations=5000,
        goal_threshold=None, neural_bias=0.75, threads=None, uniform_only=False):
    """Returns (seconds_total, seconds_cnn, seconds_plan, results dict)."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    goal_threshold = goal_threshold if goal_threshold is not None else (5.0 if dim == 2 else 3.0)
    B = len(task_ids)
    t0 = time.perf_counter()
    cdfs = None
    if not uniform_only:
        free01 = (occ_maps[task_to_map] != 0).astype(np.float32)
        x = torch.from_numpy(free01)
        if dim == 2:
            x = x[:, None].repeat(1, 3, 1, 1)
        else:
            x = x.permute(0, 3, 1, 2).contiguous()  # ToTensorV2: the file's last axis first (ThreeD_train.py:59-66)
        with torch.no_grad():
            logits = cnn_oracle.forward(state_dict, x, torch.from_numpy(coords), dim).numpy()
        cdfs = np.stack([po.build_cdf(heat_from_logits(logits[b, 0], dim))[0] for b in range(B)])
    t1 = time.perf_counter()
    free = np.stack([(po.occ_free_2d(m) if dim == 2 else po.occ_free_3d(m)).reshape(-1) for m in occ_maps])
    rad = po.radius_table(dim, shape, max_iterations)
    res = po.plan_batch(dim, shape, free, task_to_map, cdfs, neural_bias, rad, starts, goals, seed, task_ids,
                        max_iterations, goal_threshold, pow_mode=1, n_threads=threads)
    t2 = time.perf_counter()
    return t2 - t0, t1 - t0, t2 - t1, res
