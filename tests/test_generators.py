"""CPU: the drop-in map/task generators reproduce the reference's outputs for the same random.seed (golden maps_tasks)."""
import random

import numpy as np

from mgr_sampling_cnn_b200 import ThreeD_generate_maps, ThreeD_generate_tasks, generate_maps, generate_tasks
from tests import helpers as H


def test_generators_match_reference():
    g = H.load("maps_tasks")
    for i in range(2):
        ms, ts = (int(v) for v in g[f"seeds2d_{i}"])
        random.seed(ms)
        occ = generate_maps.create_map(512, 512)
        assert occ.shape == (512, 512, 1) and occ.dtype == np.uint8
        assert np.array_equal(occ[:, :, 0], g[f"occ2d_{i}"])
        random.seed(ts)
        names = [generate_tasks.draw_start_and_finish(occ[:, :, 0], "blanks/map_0_path_sx_sy_fx_fy.png", k) for k in range(10)]
        assert names == [str(n) for n in g[f"names2d_{i}"]]
        ms, ts = (int(v) for v in g[f"seeds3d_{i}"])
        random.seed(ms)
        occ3 = ThreeD_generate_maps.create_map(80, 80, 80)
        assert occ3.dtype == np.float64 and np.array_equal(occ3.astype(np.uint8), g[f"occ3d_{i}"])
        random.seed(ts)
        names = [ThreeD_generate_tasks.draw_start_and_finish(occ3, "blanks/map_0_path_sx_sy_sz_fx_fy_fz.npy", k) for k in range(10)]
        assert names == [str(n) for n in g[f"names3d_{i}"]]
