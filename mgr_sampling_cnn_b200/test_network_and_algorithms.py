"""Batched counterpart of the reference's evaluation harness (test_network_and_algorithms.py:148-183,
ThreeD_test_network_and_algorithms.py:162-194): the table behind the published results slide — time, path length and
iterations of RRT* and Neural+RRT* per task, with mean and median — produced by the batched engine.

Differences by design: tasks come from this package's generate_maps / generate_tasks drop-ins (the reference reads a
private evaluation set), the model is whatever the caller passes (random-init by default: no trained weights ship with
the reference), and `Time` is the batch's wall time divided by the number of tasks (the GPU plans thousands of tasks at
once; there is no per-task timer to read).  Length is calculate_length(path) (test_network_and_algorithms.py:52-56),
Iterations the planner's iteration_no.  The A* columns (astar_pathfinding, test_network_and_algorithms.py:58-70) come from
engine.grid_path_lengths: the length of the reference's A* path (a shortest path) for every task at once, NaN where the
reference's astar returns None; its Iterations column is '-' as in the reference.
"""
from time import perf_counter

import numpy as np
import pandas as pd
import torch

from . import engine, workload

MAX_ITERATIONS = 5000
GOAL_THRESHOLD = 5.0
GOAL_THRESHOLD_3D = 3.0


def calculate_length(path):
    """test_network_and_algorithms.py:52-56 (the engine computes the same sum on the device: PlanResult.path_len)."""
    p = np.asarray(path, dtype=np.float64)
    total = 0.0
    for i in range(len(p) - 1):  # left-to-right like the reference's loop over distance.euclidean
        d = p[i] - p[i + 1]
        s = d[0] * d[0] + d[1] * d[1]
        if len(d) == 3:
            s = s + d[2] * d[2]
        total += float(np.sqrt(s))
    return total


def _timed(fn, device):
    torch.cuda.synchronize(device)
    t0 = perf_counter()
    out = fn()
    torch.cuda.synchronize(device)
    return perf_counter() - t0, out


def astar_pathfinding(wl, device):
    """astar(occ_map_gray, start, finish) + calculate_length for every task (test_network_and_algorithms.py:58-70; 3D twin:
    ThreeD_generate_paths.astar)."""
    occ = torch.from_numpy(wl.occ_maps).to(device)
    bits = engine.pack_occupancy(occ, wl.dim, device)
    engine.grid_path_lengths(bits, wl.task_to_map[:1], wl.starts[:1], wl.goals[:1], wl.shape)  # warm-up
    secs, (length, _) = _timed(lambda: engine.grid_path_lengths(bits, wl.task_to_map, wl.starts, wl.goals, wl.shape), device)
    return secs, length


def rrt_star_pathfinding(wl, device, seed=0):
    """rrt_star.RRTStar(...).rrt_star() for every task of the workload (test_network_and_algorithms.py:90-104)."""
    occ = torch.from_numpy(wl.occ_maps).to(device)
    bits = engine.pack_occupancy(occ, wl.dim, device)
    secs, res = _timed(lambda: engine.plan_batch(bits, wl.task_to_map, wl.starts, wl.goals, dim=wl.dim, shape=wl.shape,
                                                 max_iterations=MAX_ITERATIONS, seed=seed), device)
    return secs, res


def neural_rrt_star_pathfinding(model, wl, device, seed=0):
    """CNN forward + clamp / normalise + neural_rrt.RRTStar(..., neural_bias=0.75).rrt_star() for every task
    (test_network_and_algorithms.py:107-145): the timer wraps the forward and the plan, as in the reference."""
    gp = engine.GuidedPlanner(model, wl.dim, wl.shape, max_iterations=MAX_ITERATIONS)
    occ = torch.from_numpy(wl.occ_maps).to(device)
    args = (occ, torch.from_numpy(wl.task_to_map).to(device), wl.starts, wl.goals, torch.from_numpy(wl.coords).to(device))
    gp.plan(*args, seed=seed)  # warm-up: buffers, tensor maps
    secs, res = _timed(lambda: gp.plan(*args, seed=seed), device)
    return secs, res


def main(n_tasks=1000, three_d=False, model=None, device="cuda:0", seed=0):
    """Returns (per-task DataFrame, summary DataFrame with mean / median), columns as in the reference's xlsx."""
    from . import ThreeD_train, train
    device = torch.device(device)
    wl = workload.make_3d(n_tasks) if three_d else workload.make_2d(n_tasks)
    if model is None:
        torch.manual_seed(0)
        model = (ThreeD_train.ThreeD_UNet_cooler() if three_d else train.UNet_cooler()).eval().to(device)
    n = len(wl.task_to_map)
    columns = pd.MultiIndex.from_product([["A*", "RRT*", "Neural+RRT*"], ["Time", "Length", "Iterations"]])
    df = pd.DataFrame(index=range(n), columns=columns)
    df[("A*", "Time")] = df[("A*", "Length")] = df[("A*", "Iterations")] = "-"
    secs, length = astar_pathfinding(wl, device)
    df[("A*", "Time")] = secs / n
    df[("A*", "Length")] = length.cpu().numpy()
    for name, (secs, res) in (("RRT*", rrt_star_pathfinding(wl, device, seed)),
                              ("Neural+RRT*", neural_rrt_star_pathfinding(model, wl, device, seed))):
        df[(name, "Time")] = secs / n
        df[(name, "Length")] = res.path_len.cpu().numpy()
        df[(name, "Iterations")] = res.iterations.cpu().numpy()
    num = df[[c for c in df.columns if not (df[c] == "-").all()]].astype(float)
    summary = pd.DataFrame({"mean": num.mean(), "median": num.median()}).T
    return df, summary


if __name__ == "__main__":
    import sys
    frame, summ = main(int(sys.argv[1]) if len(sys.argv) > 1 else 1000, three_d="--3d" in sys.argv)
    print(frame)
    print(summ)
