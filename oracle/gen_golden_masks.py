"""TEST INFRASTRUCTURE. Golden training masks from the UNMODIFIED reference mask makers:
generate_paths.visualize_path (generate_paths.py:118-169, through cv2 4.13 and real files in a temp directory, as the
reference works) -> tests/golden/masks2d.npz, and ThreeD_enlarge_path.enlarge_the_path (ThreeD_enlarge_path.py:25-65) ->
tests/golden/masks3d.npz.

Run in the authoring container only (needs /root/reference and cv2):   python oracle/gen_golden_masks.py
Inputs: the reference-recorded A* paths of tests/golden/astar2d.npz / astar3d.npz (maps >= 64 cells wide), plus paths that
hug the image border (clipped circles, reflected blur border) and cross obstacles (the free-space mask).
"""
import importlib.util
import os
import sys
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
from gen_golden import GOLDEN, save  # noqa: E402


def load_reference(fname):
    shim = os.path.join(_HERE, "ref_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    spec = importlib.util.spec_from_file_location("_ref_" + fname[:-3], "/root/reference/" + fname)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_mask_2d(mod, occ, path):
    import cv2
    with tempfile.TemporaryDirectory(prefix="nrrtgold") as td:
        for d in ("start_finish", "start_finish_visualized", "paths", "paths_with_points"):
            os.makedirs(os.path.join(td, d))
        name = "map_0_path0_sx1_sy1_fx2_fy2.png"
        src = os.path.join(td, "start_finish", name)
        cv2.imwrite(src, occ)
        cv2.imwrite(os.path.join(td, "start_finish_visualized", name), cv2.cvtColor(occ, cv2.COLOR_GRAY2BGR))
        mod.visualize_path(occ_map=occ, path=[tuple(int(v) for v in p) for p in path], directory=src)
        return cv2.imread(os.path.join(td, "paths", name), 0)


def main2d():
    import cv2
    mod = load_reference("generate_paths.py")
    a = np.load(os.path.join(GOLDEN, "astar2d.npz"))
    cases = []
    for i in range(int(a["n_cases"])):
        if bool(a[f"reached_{i}"]) and a[f"occ_{i}"].shape[0] >= 64 and len(a[f"path_{i}"]) > 0:
            cases.append((a[f"occ_{i}"], a[f"path_{i}"]))
    big = [c for c in cases if c[0].shape[0] == 512]
    small = [c for c in cases if c[0].shape[0] < 512]
    cases = big[:6] + small[:4]
    rng = np.random.default_rng(11)
    # border huggers on a 128 x 192 map with a few obstacles (gray levels around the 240 threshold)
    occ = np.full((128, 192), 255, np.uint8)
    occ[30:60, 50:80] = 0
    occ[90:100, 0:40] = 240
    occ[100:110, 0:40] = 241
    edge = [(0, x) for x in range(0, 192)] + [(y, 191) for y in range(1, 128)]
    cases.append((occ, np.array(edge, np.int32)))
    cases.append((occ, np.array([(127, 0)], np.int32)))
    cases.append((occ, np.array([(y, y) for y in range(0, 128)], np.int32)))          # crosses the obstacle
    walk = np.cumsum(rng.integers(-1, 2, (400, 2)), 0) + np.array([64, 96])
    walk[:, 0] = np.clip(walk[:, 0], 0, 127)
    walk[:, 1] = np.clip(walk[:, 1], 0, 191)
    cases.append((occ, walk.astype(np.int32)))
    out = {}
    for i, (occ_i, path) in enumerate(cases):
        m = reference_mask_2d(mod, occ_i, path)
        out[f"occ_{i}"], out[f"path_{i}"], out[f"mask_{i}"] = occ_i, np.asarray(path, np.int32), m
        print(i, occ_i.shape, len(path), "mask sum", int(m.astype(np.int64).sum()), flush=True)
    save("masks2d", n_cases=np.array(len(cases)), cv2_version=np.array(cv2.__version__), **out)


def main3d():
    mod = load_reference("ThreeD_enlarge_path.py")
    a = np.load(os.path.join(GOLDEN, "astar3d.npz"))
    cases = []
    for i in range(int(a["n_cases"])):
        if bool(a[f"reached_{i}"]) and len(a[f"path_{i}"]) > 0:
            cases.append((a[f"occ_{i}"], a[f"path_{i}"]))
    occ = np.zeros((24, 16, 40), np.uint8)
    occ[5:9, 3:12, 10:30] = 255
    corner = np.array([(0, 0, 0), (23, 15, 39), (23, 0, 39), (4, 3, 10), (9, 7, 20)], np.int32)   # clipped cubes, next to a block
    cases.append((occ, corner))
    out = {}
    with tempfile.TemporaryDirectory(prefix="nrrtgold") as td:
        for i, (occ_i, path) in enumerate(cases):
            vol = np.zeros_like(occ_i)
            for p in path:                                          # ThreeD_generate_paths.py:142-150
                vol[tuple(int(v) for v in p)] = 255
            op, pp = os.path.join(td, "occ.npy"), os.path.join(td, "path.npy")
            np.save(op, occ_i)
            np.save(pp, vol)
            mod.enlarge_the_path(op, pp)
            m = np.load(pp)
            out[f"occ_{i}"], out[f"path_{i}"], out[f"mask_{i}"] = occ_i, np.asarray(path, np.int32), m
            print(i, occ_i.shape, len(path), "mask sum", int(m.astype(np.int64).sum()), flush=True)
    save("masks3d", n_cases=np.array(len(cases)), **out)


if __name__ == "__main__":
    if "--3d" in sys.argv:
        main3d()
    elif "--2d" in sys.argv:
        main2d()
    else:
        main2d()
        main3d()
