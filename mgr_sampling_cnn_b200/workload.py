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


class _BoundedRandom(object):
    """Stands in for the `random` module attribute of the task-generator drop-ins while a workload is built: same stream
    (it forwards to CPython's generator), but a bound on the draws of ONE task.  The reference's loops
    (generate_tasks.py:33-46) pick a free start and then redraw the goal until it is free and 100 px / 20 voxels away from
    that start: they never return once a start is drawn that has no such cell (any 128 x 128 map has such starts).  The
    builder raises instead of spinning."""

    def __init__(self, limit: int = 2_000_000):
        self.limit, self.count = limit, 0

    def randint(self, a, b):
        self.count += 1
        if self.count > self.limit:
            raise ValueError("the reference's task generator would never return on this map: the drawn start has no free "
                             "cell far enough away (use the reference's map size, or a larger one)")
        return random.randint(a, b)


def _draw_bounded(module, occ):
    saved, module.random = module.random, _BoundedRandom()
    try:
        return module.draw_start_and_finish_coords(occ)
    finally:
        module.random = saved


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
            sx, sy, fx, fy = _draw_bounded(generate_tasks, occ)
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
            s, f = _draw_bounded(ThreeD_generate_tasks, occ)
            if lab is not None and lab[s] != lab[f]:
                redrawn += 1
                continue
            starts.append(s), goals.append(f), coords.append([list(s), list(f)]), t2m.append(m - first_map)
            k += 1
        maps.append(occ.astype(np.uint8))
    return Workload(3, (size, size, size), np.stack(maps), np.array(t2m, np.int32), np.array(starts, np.int32),
                    np.array(goals, np.int32), np.array(coords, np.float32)[:, None], filter_unreachable, redrawn)


# ---- the same workloads generated on the GPU (csrc/generate.cu) ------------------------------------------
@dataclass
class DeviceWorkload:
    """Workload whose arrays are torch tensors resident on the GPU (same fields and conventions as Workload)."""
    dim: int
    shape: tuple
    occ_maps: "torch.Tensor"      # uint8 [n_maps, *shape]
    task_to_map: "torch.Tensor"   # int32 [B]
    starts: "torch.Tensor"        # int32 [B, dim]
    goals: "torch.Tensor"
    coords: "torch.Tensor"        # float32 [B, 1, 2, dim]
    filtered: bool
    redrawn: int

    def to_host(self) -> Workload:
        return Workload(self.dim, self.shape, self.occ_maps.cpu().numpy(), self.task_to_map.cpu().numpy(), self.starts.cpu().numpy(),
                        self.goals.cpu().numpy(), self.coords.cpu().numpy(), self.filtered, self.redrawn)


def make_on_device(dim: int, n_tasks: int, first_map: int = 0, size: int = None, filter_unreachable: bool = True, device=None,
                   map_seed: int = 1000, task_seed: int = 2000) -> DeviceWorkload:
    """make_2d / make_3d without the host: maps, tasks and the reachability filter run on the GPU and produce the same
    arrays bit for bit (CPython's MT19937 / random.seed / randint are reproduced in csrc/generate.cu; map m is drawn after
    random.seed(map_seed + m), its tasks after random.seed(task_seed + m): SURVEY 8d)."""
    import ctypes as C

    import torch

    from . import _lib, engine
    lib = _lib.load()
    dev = engine._require_cuda(device)
    size = size or (512 if dim == 2 else 80)
    shape = (size,) * dim
    per = generate_tasks.TASK_NUMBER_FOR_SINGLE_MAP
    n_maps = (n_tasks + per - 1) // per
    s = list(shape) + [1] * (3 - dim)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    m = torch.arange(first_map, first_map + n_maps, dtype=torch.int64, device=dev)
    occ = torch.empty((n_maps,) + shape, dtype=torch.uint8, device=dev)
    _lib.count_launch()
    _lib.check(lib.nrrt_generate_maps(dim, p(m + map_seed), n_maps, s[0], s[1], s[2], p(occ), st))
    seeds_t = m + task_seed
    t2m = torch.empty(n_tasks, dtype=torch.int32, device=dev)
    starts = torch.empty((n_tasks, dim), dtype=torch.int32, device=dev)
    goals = torch.empty((n_tasks, dim), dtype=torch.int32, device=dev)
    coords = torch.empty((n_tasks, 1, 2, dim), dtype=torch.float32, device=dev)
    counters = torch.zeros(2, dtype=torch.int32, device=dev)
    bits = engine.pack_occupancy(occ, dim, dev) if filter_unreachable else None
    n_cand = per if not filter_unreachable else per + 6
    while True:
        cs = torch.empty((n_maps, n_cand, dim), dtype=torch.int32, device=dev)
        cg = torch.empty_like(cs)
        done = torch.empty(n_maps, dtype=torch.int32, device=dev)
        _lib.count_launch()
        _lib.check(lib.nrrt_generate_task_candidates(dim, p(seeds_t), n_maps, s[0], s[1], s[2], p(occ), n_cand, p(cs), p(cg), p(done), st))
        steps = None
        if filter_unreachable:
            c2m = torch.arange(n_maps, dtype=torch.int32, device=dev).repeat_interleave(n_cand)
            _, steps = engine.grid_path_lengths(bits, c2m, cs.reshape(-1, dim), cg.reshape(-1, dim), shape, axis_only=True)
        _lib.count_launch()
        _lib.check(lib.nrrt_select_tasks(dim, n_maps, n_cand, per, n_tasks, p(cs), p(cg), None if steps is None else p(steps), p(t2m),
                                         p(starts), p(goals), p(coords), p(counters[0:1]), p(counters[1:2]), st))
        redrawn, short = (int(v) for v in counters.cpu().tolist())
        if int(done.min().item()) < 0:
            raise _lib.NrrtError("a generated map has no admissible start / goal cells (the reference would loop forever)")
        if short == 0:
            break
        n_cand *= 2  # a map with many unreachable candidates: draw more (the stream does not depend on the filter)
    return DeviceWorkload(dim, shape, occ, t2m, starts, goals, coords, filter_unreachable, redrawn)


def task_file_names(wl, first_map: int = 0, directory: str = "") -> list:
    """The reference's wire format: start / goal encoded into the task file's name (generate_tasks.py:48-55,76-82;
    ThreeD_generate_tasks.py:50-60), parsed back by MapsDataset (train.py:55-59) and get_start_finish_coordinates."""
    t2m = np.asarray(wl.task_to_map.cpu() if hasattr(wl.task_to_map, "cpu") else wl.task_to_map)
    st = np.asarray(wl.starts.cpu() if hasattr(wl.starts, "cpu") else wl.starts)
    gl = np.asarray(wl.goals.cpu() if hasattr(wl.goals, "cpu") else wl.goals)
    names, k, prev = [], 0, -1
    for b in range(len(t2m)):
        k = k + 1 if t2m[b] == prev else 0
        prev = t2m[b]
        if wl.dim == 2:
            (sy, sx), (fy, fx) = st[b], gl[b]
            names.append(f"{directory}map_{first_map + t2m[b]}_path{k}_sx{sx}_sy{sy}_fx{fx}_fy{fy}.png")
        else:
            names.append(f"{directory}map_{first_map + t2m[b]}_path{k}_sx{st[b][0]}_sy{st[b][1]}_sz{st[b][2]}"
                         f"_fx{gl[b][0]}_fy{gl[b][1]}_fz{gl[b][2]}.npy")
    return names
