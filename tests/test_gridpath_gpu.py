"""GPU: nrrt_grid_paths (batched shortest grid paths = the reference's A*, 2D and 3D) against paths recorded from the reference's
generate_paths.astar, and at batch size against the size-independent properties of the problem."""
import numpy as np
import pytest
import torch

from mgr_sampling_cnn_b200 import engine, workload
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim", [2, 3])
def test_lengths_match_the_reference_astar(dim, cuda_device):
    g = H.load(f"astar{dim}d")
    n = int(g["n_cases"])
    by_shape = {}
    for i in range(n):
        by_shape.setdefault(g[f"occ_{i}"].shape, []).append(i)
    for shape, idx in by_shape.items():
        maps = np.stack([g[f"occ_{i}"] for i in idx])
        bits = engine.pack_occupancy(maps, dim, cuda_device)
        starts = np.stack([g[f"start_{i}"] for i in idx])
        goals = np.stack([g[f"goal_{i}"] for i in idx])
        length, steps = engine.grid_path_lengths(bits, np.arange(len(idx), dtype=np.int32), starts, goals, shape)
        torch.cuda.synchronize()
        length, steps = length.cpu().numpy(), steps.cpu().numpy()
        for j, i in enumerate(idx):
            if not bool(g[f"reached_{i}"]):
                assert np.isnan(length[j]) and (steps[j] == -1).all(), f"case {i}"
                continue
            path = g[f"path_{i}"]
            kind = np.abs(np.diff(path, axis=0)).sum(axis=1)          # 1 = axis step, 2 = plane diagonal, 3 = space diagonal
            assert [int((kind == k).sum()) for k in (1, 2, 3)] == steps[j].tolist(), f"case {i}"
            ref = float(g[f"length_{i}"])
            assert abs(length[j] - ref) <= 1e-9 * max(ref, 1.0), f"case {i}: {length[j]} vs {ref}"  # order of additions only


def test_batch_properties_at_full_size(cuda_device):
    """512 x 512 maps from the repo's generator, 10 tasks each: reachability = same 4-connected free component (a diagonal
    step needs both orthogonal neighbours free, so it decomposes into two axis steps); length is symmetric in start / goal
    and bounded below by the octile distance; tasks sharing a map and a small slot count reuse the staged grid / work slots."""
    wl = workload.make_2d(40, filter_unreachable=False)
    bits = engine.pack_occupancy(wl.occ_maps, 2, cuda_device)
    fwd = [t.cpu().numpy() for t in engine.grid_path_lengths(bits, wl.task_to_map, wl.starts, wl.goals, wl.shape, n_slots=7)]
    bwd = [t.cpu().numpy() for t in engine.grid_path_lengths(bits, wl.task_to_map, wl.goals, wl.starts, wl.shape)]
    torch.cuda.synchronize()
    assert np.array_equal(fwd[1], bwd[1])
    from scipy import ndimage
    for b in range(40):
        lab, _ = ndimage.label(wl.occ_maps[wl.task_to_map[b]] != 0)
        sy, sx = wl.starts[b]
        gy, gx = wl.goals[b]
        same = lab[sy, sx] != 0 and lab[sy, sx] == lab[gy, gx]
        assert same == (fwd[1][b, 0] >= 0), f"task {b}"
        if same:
            dy, dx = abs(int(sy) - int(gy)), abs(int(sx) - int(gx))
            octile = max(dy, dx) - min(dy, dx) + min(dy, dx) * np.sqrt(2)
            assert fwd[0][b] >= octile - 1e-9


def test_3d_batch_symmetry_and_bounds(cuda_device):
    """80^3 maps from the repo's generator: symmetric in start / goal, bounded below by the 26-connected free-space distance,
    start == goal gives 0."""
    wl = workload.make_3d(20, filter_unreachable=False)
    bits = engine.pack_occupancy(wl.occ_maps, 3, cuda_device)
    fwd = [t.cpu().numpy() for t in engine.grid_path_lengths(bits, wl.task_to_map, wl.starts, wl.goals, wl.shape, n_slots=5)]
    bwd = [t.cpu().numpy() for t in engine.grid_path_lengths(bits, wl.task_to_map, wl.goals, wl.starts, wl.shape)]
    same = [t.cpu().numpy() for t in engine.grid_path_lengths(bits, wl.task_to_map, wl.starts, wl.starts, wl.shape)]
    torch.cuda.synchronize()
    assert np.array_equal(fwd[1], bwd[1]) and (same[0] == 0).all()
    for b in range(20):
        if fwd[1][b, 0] < 0:
            continue
        d = np.sort(np.abs(wl.starts[b].astype(int) - wl.goals[b].astype(int)))
        free = (d[2] - d[1]) + (d[1] - d[0]) * np.sqrt(2) + d[0] * np.sqrt(3)
        assert fwd[0][b] >= free - 1e-9


def _serpentine(h, w):
    """Walls on every other row with a one-cell door alternating left / right: one corridor of ~h/2 * w cells."""
    occ = np.full((h, w), 255, np.uint8)
    for k, r in enumerate(range(1, h, 2)):
        occ[r, :] = 0
        occ[r, (w - 1) if k % 2 == 0 else 0] = 255
    return occ


def _diagonal_maze(n=256, period=8, width=4):
    """Diagonal stripes (free iff (r - c) mod 8 < 4) joined by doors alternately near their two ends: shortest paths run
    thousands of DIAGONAL steps (a diagonal move needs both orthogonal neighbours free: the stripes are 4 wide)."""
    r, c = np.mgrid[0:n, 0:n]
    occ = np.where((r - c) % period < width, 255, 0).astype(np.uint8)
    for j in range(-n // period, n // period):
        lo = period * j + width
        cells = [(rr, rr - d) for d in range(lo, lo + period - width) for rr in range(n) if 0 <= rr - d < n]
        if not cells:
            continue
        s = np.array([rr + cc for rr, cc in cells])
        target = (s.min() + 6) if j % 2 == 0 else (s.max() - 6)
        for (rr, cc), sv in zip(cells, s):
            if abs(sv - target) <= 1:
                occ[rr, cc] = 255
    return occ


def test_thousands_of_diagonal_steps_exact(cuda_device):
    """Candidate costs n1 + n2 sqrt(2) with n2 in the thousands are ordered exactly (38 fraction bits): step counts to 64
    goals spread over the maze equal the oracle's heap Dijkstra."""
    from oracle import astar_oracle as ao
    occ = _diagonal_maze()
    ns, nd = ao.dial_lengths(occ, (1, 0))
    assert nd.max() > 4000
    rng = np.random.default_rng(5)
    cand = np.argwhere(ns >= 0)
    goals = np.concatenate([cand[rng.choice(len(cand), 63, replace=False)], [[0, 245]]]).astype(np.int32)
    bits = engine.pack_occupancy(occ[None], 2, cuda_device)
    length, steps = engine.grid_path_lengths(bits, np.zeros(64, np.int32), np.tile(np.array([[1, 0]], np.int32), (64, 1)),
                                             goals, occ.shape)
    torch.cuda.synchronize()
    steps = steps.cpu().numpy()
    for b, (gy, gx) in enumerate(goals):
        assert (int(ns[gy, gx]), int(nd[gy, gx]), 0) == tuple(steps[b].tolist()), f"goal {gy, gx}"


def test_long_paths_and_the_cost_limit(cuda_device):
    """Paths of thousands of steps (far beyond the ~1000 diagonal steps where a 20-bit fraction could misorder candidates):
    exact step counts vs the oracle's Dijkstra; a path longer than the label format's cost limit (8191) is reported as
    "limit" (-2, +inf), never as "no path"."""
    from oracle import astar_oracle as ao
    small, big = _serpentine(128, 96), _serpentine(512, 512)
    ns, nd = ao.dial_lengths(small, (0, 0))
    goal = (126, 95)
    assert 5000 < ns[goal] + nd[goal] < 8000
    for occ, g, expect in ((small, goal, (int(ns[goal]), int(nd[goal]), 0)), (big, (510, 511), (-2, -2, -2))):
        bits = engine.pack_occupancy(occ[None], 2, cuda_device)
        length, steps = engine.grid_path_lengths(bits, np.zeros(1, np.int32), np.array([[0, 0]], np.int32),
                                                 np.array([g], np.int32), occ.shape)
        torch.cuda.synchronize()
        assert tuple(steps[0].cpu().tolist()) == expect
        if expect[0] >= 0:
            assert abs(float(length[0]) - (expect[0] + expect[1] * np.sqrt(2))) <= 1e-9 * float(length[0])
        else:
            assert np.isposinf(float(length[0]))


def _legal_path(occ, path, dim):
    """Every cell free and every move one the reference's A* may make (generate_paths.py:40-62, ThreeD_generate_paths.py:44-76)."""
    free = (occ != 0) if dim == 2 else (occ == 0)
    for p in path:
        if not free[tuple(p)]:
            return False
    for a, b in zip(path[:-1], path[1:]):
        d = b - a
        if np.abs(d).max() != 1:
            return False
        if dim == 2:
            if np.abs(d).sum() == 2 and not (free[a[0] + d[0], a[1]] and free[a[0], a[1] + d[1]]):
                return False
        else:
            if not (free[b[0], a[1], a[2]] and free[a[0], b[1], a[2]] and free[a[0], a[1], b[2]]):
                return False
    return True


@pytest.mark.parametrize("dim", [2, 3])
def test_extracted_paths_are_shortest_paths(dim, cuda_device):
    """nrrt_grid_paths_ex: the returned cells form a legal start -> goal path with exactly the reference path's step kinds."""
    g = H.load(f"astar{dim}d")
    by_shape = {}
    for i in range(int(g["n_cases"])):
        by_shape.setdefault(g[f"occ_{i}"].shape, []).append(i)
    for shape, idx in by_shape.items():
        maps = np.stack([g[f"occ_{i}"] for i in idx])
        bits = engine.pack_occupancy(maps, dim, cuda_device)
        starts = np.stack([g[f"start_{i}"] for i in idx])
        goals = np.stack([g[f"goal_{i}"] for i in idx])
        length, steps, paths, path_n = [t.cpu().numpy() for t in
                                        engine.grid_paths(bits, np.arange(len(idx), dtype=np.int32), starts, goals, shape, path_cap=1024)]
        for j, i in enumerate(idx):
            ref = g[f"path_{i}"]
            if not bool(g[f"reached_{i}"]):
                assert path_n[j] == 0
                continue
            assert path_n[j] == len(ref), f"case {i}"
            p = paths[j, :path_n[j]].astype(np.int64)
            assert np.array_equal(p[0], starts[j]) and np.array_equal(p[-1], goals[j])
            assert _legal_path(maps[j], p, dim), f"case {i}"
            kind = np.abs(np.diff(p, axis=0)).sum(axis=1)
            assert [int((kind == k).sum()) for k in (1, 2, 3)] == steps[j].tolist()
    # a cap that is too small is reported, not overrun
    wl = workload.make_2d(10, filter_unreachable=True)
    bits = engine.pack_occupancy(wl.occ_maps, 2, cuda_device)
    _, st, paths, pn = [t.cpu().numpy() for t in engine.grid_paths(bits, wl.task_to_map, wl.starts, wl.goals, wl.shape, path_cap=8)]
    assert (pn == -(st.sum(1) + 1)).all() and not paths.any()
