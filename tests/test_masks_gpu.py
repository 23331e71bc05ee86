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
