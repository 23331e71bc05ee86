"""TEST INFRASTRUCTURE. Golden A* paths from the UNMODIFIED reference (generate_paths.py:11-74, ThreeD_generate_paths.py:14-97)
for tests/golden/astar2d.npz and astar3d.npz.

Run in the authoring container only (needs /root/reference):   python oracle/gen_golden_astar.py
Cases: the 20 tasks of the two 512x512 maps of tests/golden/maps_tasks.npz (generate_maps / generate_tasks output of the
reference) and small hand-made maps: a corner that must not be cut (generate_paths.py:55-58), an unreachable goal,
start == goal.  Stored per case: map, start, goal ((y, x) as generate_paths.py:81-87 returns them), the reference's path,
its length (test_network_and_algorithms.py:52-56, calculate_length) and the reachable flag.
"""
import importlib.util
import os
import sys

import numpy as np
from scipy.spatial import distance

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
from gen_golden import GOLDEN, parse_name_2d, save  # noqa: E402


def load_reference_astar(fname="generate_paths.py"):
    shim = os.path.join(_HERE, "ref_shim")  # no-op matplotlib etc. (the 3D file imports pyplot at module scope)
    if shim not in sys.path:
        sys.path.insert(0, shim)
    spec = importlib.util.spec_from_file_location("_ref_" + fname[:-3], "/root/reference/" + fname)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.astar


def calculate_length(path):  # test_network_and_algorithms.py:52-56
    length = 0
    for i in range(len(path) - 1):
        length += distance.euclidean(path[i], path[i + 1])
    return length


def small_cases():
    cases = []
    m = np.full((12, 12), 255, np.uint8)          # corner: the diagonal past an obstacle corner is forbidden
    m[5, 0:6] = 0
    m[0:6, 5] = 0
    m[5, 5] = 0
    cases.append((m.copy(), (4, 4), (6, 6)))       # enclosed pocket: unreachable
    m2 = np.full((16, 16), 255, np.uint8)
    m2[8, 3:13] = 0
    cases.append((m2.copy(), (6, 8), (10, 8)))     # around a wall
    m3 = np.full((16, 16), 255, np.uint8)
    m3[7, 8] = 0
    m3[8, 7] = 0
    cases.append((m3.copy(), (7, 7), (8, 8)))      # both orthogonal neighbours blocked: no diagonal step
    m4 = np.full((16, 16), 255, np.uint8)
    m4[7, 8] = 0
    cases.append((m4.copy(), (7, 7), (8, 8)))      # one orthogonal neighbour blocked: still no diagonal step
    cases.append((np.full((8, 8), 255, np.uint8), (3, 3), (3, 3)))   # start == goal
    rng = np.random.default_rng(3)
    for _ in range(6):                              # random 64x64 mazes, 30 % obstacles
        m5 = (rng.random((64, 64)) > 0.3).astype(np.uint8) * 255
        free = np.argwhere(m5 != 0)
        s, g = free[rng.integers(len(free))], free[rng.integers(len(free))]
        cases.append((m5, (int(s[0]), int(s[1])), (int(g[0]), int(g[1]))))
    return cases


def main():
    astar = load_reference_astar()
    mt = dict(np.load(os.path.join(GOLDEN, "maps_tasks.npz"), allow_pickle=False))
    cases = []
    for mi in (0, 1):
        occ = mt[f"occ2d_{mi}"]
        for name in mt[f"names2d_{mi}"]:
            s, g = parse_name_2d(str(name))
            cases.append((occ, s, g))
    cases += small_cases()
    out = {}
    for i, (occ, s, g) in enumerate(cases):
        path = astar(occ, tuple(s), tuple(g))
        out[f"occ_{i}"] = occ
        out[f"start_{i}"] = np.array(s, np.int32)
        out[f"goal_{i}"] = np.array(g, np.int32)
        out[f"reached_{i}"] = np.array(path is not None)
        out[f"path_{i}"] = np.array(path if path else np.zeros((0, 2)), np.int32).reshape(-1, 2)
        out[f"length_{i}"] = np.array(float(calculate_length(path)) if path else np.nan)
        print(i, occ.shape, s, g, "len", out[f"length_{i}"], "points", len(out[f"path_{i}"]), flush=True)
    save("astar2d", n_cases=np.array(len(cases)), **out)


def main3d():
    """3D: a few tasks of the two generated 80^3 maps (the reference's Python A* needs minutes on the long ones, so the three
    shortest of each map) + random 16^3 mazes + a walled-in goal.  Maps hold {0 = free, 255 = obstacle}."""
    from gen_golden import parse_name_3d
    astar = load_reference_astar("ThreeD_generate_paths.py")
    mt = dict(np.load(os.path.join(GOLDEN, "maps_tasks.npz"), allow_pickle=False))
    cases = []
    for mi in (0, 1):
        occ = mt[f"occ3d_{mi}"]
        tasks = sorted((parse_name_3d(str(n)) for n in mt[f"names3d_{mi}"]),
                       key=lambda sg: sum((a - b) ** 2 for a, b in zip(*sg)))
        cases += [(occ, s, g) for s, g in tasks[:3]]
    rng = np.random.default_rng(5)
    for dens in (0.2, 0.35, 0.45, 0.2, 0.35, 0.45):
        m = (rng.random((16, 16, 16)) < dens).astype(np.uint8) * 255
        free = np.argwhere(m == 0)
        s, g = free[rng.integers(len(free))], free[rng.integers(len(free))]
        cases.append((m, tuple(int(v) for v in s), tuple(int(v) for v in g)))
    m = np.zeros((10, 10, 10), np.uint8)
    m[3:8, 3:8, 3:8] = 255
    m[5, 5, 5] = 0                                   # free pocket inside a solid block: unreachable
    cases.append((m, (0, 0, 0), (5, 5, 5)))
    cases.append((np.zeros((6, 6, 6), np.uint8), (2, 2, 2), (2, 2, 2)))
    out = {}
    for i, (occ, s, g) in enumerate(cases):
        path = astar(occ, tuple(s), tuple(g))
        out[f"occ_{i}"] = occ
        out[f"start_{i}"] = np.array(s, np.int32)
        out[f"goal_{i}"] = np.array(g, np.int32)
        out[f"reached_{i}"] = np.array(path is not None)
        out[f"path_{i}"] = np.array(path if path else np.zeros((0, 3)), np.int32).reshape(-1, 3)
        out[f"length_{i}"] = np.array(float(calculate_length(path)) if path else np.nan)
        print(i, occ.shape, s, g, "len", out[f"length_{i}"], "points", len(out[f"path_{i}"]), flush=True)
    save("astar3d", n_cases=np.array(len(cases)), **out)


if __name__ == "__main__":
    if "--3d" in sys.argv:
        main3d()
    else:
        main()
