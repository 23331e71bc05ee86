"""CPU: the A* restatement (oracle/astar_oracle.py) against paths recorded from the reference's generate_paths.astar."""
import numpy as np

from oracle import astar_oracle as ao
from tests import helpers as H


def test_astar_restatement_reproduces_the_reference_paths():
    g = H.load("astar2d")
    n = int(g["n_cases"])
    assert n >= 30
    for i in range(n):
        path = ao.astar(g[f"occ_{i}"], g[f"start_{i}"], g[f"goal_{i}"])
        if not bool(g[f"reached_{i}"]):
            assert path is None
            continue
        assert np.array_equal(np.array(path, np.int32).reshape(-1, 2), g[f"path_{i}"]), f"case {i}"
        assert ao.path_length(path) == float(g[f"length_{i}"])


def test_shortest_path_step_counts_give_the_astar_length():
    """A* with the consistent Euclidean heuristic returns a shortest path; its length is n_straight + n_diagonal * sqrt(2)
    up to the order of the additions (the property the GPU kernel is checked with at full size)."""
    g = H.load("astar2d")
    for i in range(20, int(g["n_cases"])):  # the small maps: the heap Dijkstra in pure Python is slow at 512 x 512
        ns, nd = ao.dial_lengths(g[f"occ_{i}"], g[f"start_{i}"])
        gy, gx = (int(v) for v in g[f"goal_{i}"])
        if not bool(g[f"reached_{i}"]):
            assert ns[gy, gx] < 0
            continue
        assert abs(ns[gy, gx] + nd[gy, gx] * np.sqrt(2) - float(g[f"length_{i}"])) <= 1e-9 * max(1.0, float(g[f"length_{i}"]))


def test_astar3d_restatement_reproduces_the_reference_paths():
    """3D twin (ThreeD_generate_paths.py:14-97), including its `== 255` offset comparisons (oracle/astar_oracle.py)."""
    g = H.load("astar3d")
    for i in range(int(g["n_cases"])):
        path = ao.astar3d(g[f"occ_{i}"], g[f"start_{i}"], g[f"goal_{i}"])
        if not bool(g[f"reached_{i}"]):
            assert path is None
            continue
        assert np.array_equal(np.array(path, np.int32).reshape(-1, 3), g[f"path_{i}"]), f"case {i}"
        assert abs(ao.path_length(path) - float(g[f"length_{i}"])) <= 1e-12 * max(1.0, float(g[f"length_{i}"]))
