"""Mask makers on the device (csrc/masks.cu) vs the masks the unmodified reference wrote (tests/golden/masks2d.npz,
masks3d.npz) and vs the oracle restatement at BASELINE sizes: bit-exact."""
import os

import numpy as np
import pytest
import torch

from mgr_sampling_cnn_b200 import ThreeD_enlarge_path as ep
from mgr_sampling_cnn_b200 import generate_paths as gp
from oracle import mask_oracle as mo

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_masks_2d_match_reference_goldens():
    g = np.load(os.path.join(GOLDEN, "masks2d.npz"))
    for i in range(int(g["n_cases"])):
        got = gp.path_masks(g[f"occ_{i}"][None], [0], [g[f"path_{i}"]])[0].cpu().numpy()
        want = g[f"mask_{i}"]
        assert np.array_equal(got, want), "case %d: %d pixels differ" % (i, (got != want).sum())


def test_masks_2d_batched_vs_oracle_512():
    g = np.load(os.path.join(GOLDEN, "masks2d.npz"))
    occs = np.stack([g["occ_0"], np.full((512, 512), 255, np.uint8)])
    rng = np.random.default_rng(4)
    paths, t2m = [], []
    for b in range(12):
        n = int(rng.integers(1, 900))
        walk = np.cumsum(rng.integers(-1, 2, (n, 2)), 0) + rng.integers(0, 512, 2)
        paths.append(np.clip(walk, 0, 511).astype(np.int32))
        t2m.append(b % 2)
    paths.append(g["path_0"])
    t2m.append(0)
    got = gp.path_masks(occs, t2m, paths).cpu().numpy()
    for b, p in enumerate(paths):
        assert np.array_equal(got[b], mo.mask_2d(occs[t2m[b]], p)), "task %d" % b


def test_masks_2d_odd_sizes_and_empty_path():
    rng = np.random.default_rng(9)
    occ = (rng.random((77, 150)) > 0.2).astype(np.uint8) * 255
    paths = [np.zeros((0, 2), np.int32), np.array([(0, 0), (76, 149), (40, 75)], np.int32),
             np.stack([np.arange(77), np.arange(77) * 149 // 76], 1).astype(np.int32)]
    got = gp.path_masks(occ[None], [0, 0, 0], paths).cpu().numpy()
    assert not got[0].any()
    for b in (1, 2):
        assert np.array_equal(got[b], mo.mask_2d(occ, paths[b]))


def test_enlarge_3d_matches_reference_goldens():
    g = np.load(os.path.join(GOLDEN, "masks3d.npz"))
    assert ep.layer_values().tolist() == mo.layer_values().tolist()
    for i in range(int(g["n_cases"])):
        occ, path = g[f"occ_{i}"], g[f"path_{i}"]
        vols = ep.path_volumes(path[None], [len(path)], occ.shape)
        assert np.array_equal(vols[0].cpu().numpy(), mo.path_volume(occ.shape, path))
        got = ep.enlarge_paths(occ[None], [0], vols)[0].cpu().numpy()
        assert np.array_equal(got, g[f"mask_{i}"]), "case %d" % i


def test_enlarge_3d_batched_vs_oracle_80():
    rng = np.random.default_rng(2)
    occs = (rng.random((2, 80, 80, 80)) < 0.15).astype(np.uint8) * 255
    vols = np.zeros((6, 80, 80, 80), np.uint8)
    t2m = [0, 1, 0, 1, 0, 1]
    for b in range(6):
        walk = np.clip(np.cumsum(rng.integers(-1, 2, (60, 3)), 0) + rng.integers(0, 80, 3), 0, 79)
        vols[b][walk[:, 0], walk[:, 1], walk[:, 2]] = 255
    vols[5][3, 3, 3] = 100           # a value that is not 255 is not a path voxel but survives the max
    for layers in (5, 3):
        got = ep.enlarge_paths(occs, t2m, vols, layers=layers).cpu().numpy()
        for b in range(6):
            assert np.array_equal(got[b], mo.enlarge_3d(occs[t2m[b]], vols[b], layers)), "task %d layers %d" % (b, layers)


def test_generate_paths_file_pipeline(tmp_path):
    """generate_paths.generate_paths() / ThreeD_generate_paths.generate_paths() on task files: A* path -> mask / path volume next
    to the task file, unreachable tasks deleted (generate_paths.py:100-115, ThreeD_generate_paths.py:125-138)."""
    import cv2
    from mgr_sampling_cnn_b200 import ThreeD_generate_paths as gp3
    g = np.load(os.path.join(GOLDEN, "astar2d.npz"))
    d2 = tmp_path / "maps" / "start_finish"
    d2.mkdir(parents=True)
    names = {}
    for i in (0, 2, 20, 22):          # two 512 x 512 tasks, an unreachable pocket, a 16 x 16 wall
        occ, s, f = g[f"occ_{i}"], g[f"start_{i}"], g[f"goal_{i}"]
        if occ.shape[0] < 64:
            big = np.zeros((64, 64), np.uint8)
            big[:occ.shape[0], :occ.shape[1]] = occ
            occ = big
        name = str(d2 / f"map_{i}_path0_sx{s[1]}_sy{s[0]}_fx{f[1]}_fy{f[0]}.png")
        cv2.imwrite(name, occ)
        names[i] = (name, occ, tuple(int(v) for v in s), tuple(int(v) for v in f), bool(g[f"reached_{i}"]))
    gp.MAPS_DIRECTORY = str(d2 / "*.png")
    gp.generate_paths(batch=3)
    for i, (name, occ, s, f, reached) in names.items():
        out = name.replace("start_finish", "paths")
        if not reached:
            assert not os.path.exists(name) and not os.path.exists(out)
            continue
        path = gp.astar(occ, s, f)
        assert path[0] == s and path[-1] == f and len(path) == len(g[f"path_{i}"])
        assert np.array_equal(cv2.imread(out, 0), mo.mask_2d(occ, np.array(path)))
    g3 = np.load(os.path.join(GOLDEN, "astar3d.npz"))
    d3 = tmp_path / "3D_maps" / "start_finish"
    d3.mkdir(parents=True)
    for i in (0, 12):                 # a task on an 80^3 map, the walled-in goal
        s, f = g3[f"start_{i}"], g3[f"goal_{i}"]
        np.save(str(d3 / f"map_{i}_path0_sx{s[0]}_sy{s[1]}_sz{s[2]}_fx{f[0]}_fy{f[1]}_fz{f[2]}.npy"), g3[f"occ_{i}"])
    gp3.MAPS_DIRECTORY = str(d3 / "*.npy")
    gp3.generate_paths()
    left = sorted(os.listdir(d3))
    assert len(left) == 1 and left[0].startswith("map_0_")
    vol = np.load(str(tmp_path / "3D_maps" / "paths" / left[0]))
    assert vol.dtype == g3["occ_0"].dtype and int((vol == 255).sum()) == len(g3["path_0"]) and set(np.unique(vol)) <= {0, 255}
    assert gp3.astar(g3["occ_12"], tuple(g3["start_12"]), tuple(g3["goal_12"])) is None
