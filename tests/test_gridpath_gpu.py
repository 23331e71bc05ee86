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
