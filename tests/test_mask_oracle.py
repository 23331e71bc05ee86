"""Mask-maker oracle (oracle/mask_oracle.py) pinned to masks written by the unmodified reference functions
(generate_paths.visualize_path, ThreeD_enlarge_path.enlarge_the_path; fixtures from oracle/gen_golden_masks.py) and, where
OpenCV is installed, to cv2.circle / cv2.GaussianBlur themselves."""
import os

import numpy as np
import pytest

from oracle import mask_oracle as mo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_mask_2d_matches_reference_goldens():
    g = np.load(os.path.join(GOLDEN, "masks2d.npz"))
    n = int(g["n_cases"])
    assert n >= 12
    for i in range(n):
        got = mo.mask_2d(g[f"occ_{i}"], g[f"path_{i}"])
        assert np.array_equal(got, g[f"mask_{i}"]), "case %d differs in %d pixels" % (i, (got != g[f"mask_{i}"]).sum())


def test_enlarge_3d_matches_reference_goldens():
    g = np.load(os.path.join(GOLDEN, "masks3d.npz"))
    n = int(g["n_cases"])
    assert n >= 12
    assert mo.layer_values().tolist() == [206, 157, 108, 59, 10]
    for i in range(n):
        occ = g[f"occ_{i}"]
        got = mo.enlarge_3d(occ, mo.path_volume(occ.shape, g[f"path_{i}"]))
        assert np.array_equal(got, g[f"mask_{i}"]), "case %d" % i


def test_fixed_point_kernel_and_disk():
    kf = mo.gaussian_kernel_fixed()
    assert kf.sum() == 256 and kf[16] == 34 and np.count_nonzero(kf) == 19 and np.array_equal(kf, kf[::-1])
    assert mo.disk_half_widths().tolist() == [10, 9, 9, 9, 9, 8, 8, 7, 6, 4, 0]


def test_against_opencv_when_available():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for t in range(4):
        a = rng.integers(0, 256, (70, 90)).astype(np.uint8) if t < 2 else (rng.random((70, 90)) > 0.9).astype(np.uint8) * 255
        assert np.array_equal(mo.gaussian_blur_u8(a), cv2.GaussianBlur(a, (33, 33), cv2.BORDER_WRAP))
    pts = np.array([(0, 0), (5, 88), (69, 40), (30, 30), (35, 36)])
    img = np.zeros((70, 90), np.uint8)
    for (y, x) in pts:
        img = cv2.circle(img, (int(x), int(y)), 10, 255, -1)
    assert np.array_equal(mo.draw_disks((70, 90), pts), img)
