"""GPU: the drop-in modules (same names/signatures as the reference's rrt_star / neural_rrt / ThreeD_*) driven exactly
like the reference's own callers, with the shared Philox stream injected through the module's `random` attribute."""
import numpy as np
import pytest

from oracle.philox_stream import StreamRandom
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _module(name):
    import importlib
    return importlib.import_module("mgr_sampling_cnn_b200." + name)


CASES = [("plan2d_uniform_stream", "rrt_star"), ("plan2d_neural_stream", "neural_rrt"),
         ("plan3d_uniform_stream", "ThreeD_rrt_star"), ("plan3d_neural_stream", "ThreeD_neural_rrt")]


@pytest.mark.parametrize("fixture,modname", CASES)
def test_dropin_rrt_star_matches_reference(fixture, modname, cuda_device):
    g = H.load(fixture)
    mod = _module(modname)
    dim = H.fixture_dim(fixture)
    mod.random = StreamRandom(int(g["stream_seed"]), int(g["stream_task"]))
    start, goal = tuple(int(v) for v in g["start"]), tuple(int(v) for v in g["goal"])
    mi, thr = int(g["max_iterations"]), float(g["goal_threshold"])
    if "neural" in modname:
        if dim == 2:
            occ = np.repeat((g["occ_map"].astype(np.float32) / 255.0)[:, :, None], 3, axis=2)
            occ_arg = occ
        else:
            occ_arg = (g["occ_map"] != 0).astype(np.float32)
        planner = mod.RRTStar(occ_arg, g["heat_map"], start, goal, mi, thr, float(g["neural_bias"]))
    else:
        occ_arg = g["occ_map"] if dim == 2 else g["occ_map"].astype(np.float64)
        planner = mod.RRTStar(occ_arg, start, goal, mi, thr)
    path, iters = planner.rrt_star()
    assert iters == int(g["ref_iterations"]) == planner.iteration_no
    assert len(planner.nodes) == int(g["ref_n_nodes"])
    assert np.array_equal(np.array(path, dtype=np.float64).reshape(-1, dim), g["ref_path"])
    assert planner.best_distance == float(g["ref_best_distance"])
    pos = np.array([n.position for n in planner.nodes], dtype=np.float64)
    assert np.array_equal(pos, g["ref_pos"][:len(planner.nodes)])
    index = {id(n): i for i, n in enumerate(planner.nodes)}
    par = np.array([-1 if n.parent is None else index.get(id(n.parent), len(planner.nodes)) for n in planner.nodes])
    assert np.array_equal(par, g["ref_parent"][:len(planner.nodes)])
    assert mod.calculate_length(path) == sum(
        float(np.sqrt(np.add.reduce((a - b) * (a - b)))) for a, b in zip(g["ref_path"][:-1], g["ref_path"][1:]))
    assert mod.MAX_ITERATIONS == 5000 and mod.GOAL_THRESHOLD == (5.0 if dim == 2 else 3.0)


@pytest.mark.parametrize("fixture,modname", [("plan2d_uniform_mt_0", "rrt_star"), ("plan3d_uniform_mt_0", "ThreeD_rrt_star")])
def test_dropin_methods_match_reference_calls(fixture, modname, cuda_device):
    """is_collision_free on the exact (p1, p2) pairs the reference evaluated; nearest / near against scipy's formula."""
    g = H.load(fixture)
    mod = _module(modname)
    dim = H.fixture_dim(fixture)
    occ = g["occ_map"] if dim == 2 else g["occ_map"].astype(np.float64)
    planner = mod.RRTStar(occ, tuple(int(v) for v in g["start"]), tuple(int(v) for v in g["goal"]), 100, 5.0)
    v = g["ref_verdicts"]
    rows = list(range(0, len(v), max(len(v) // 60, 1)))
    for r in rows:
        p1, p2 = tuple(v[r, 2:2 + dim]), tuple(v[r, 2 + dim:2 + 2 * dim])
        assert planner.is_collision_free(p1, p2) == bool(v[r, -1]), r
    assert planner.is_collision_free((5.0,) * dim, (5.0,) * dim) is False  # point1 == point2
    # tree queries on the reference's final tree
    n = int(g["ref_n_nodes"])
    planner.nodes = [mod.Node(tuple(p)) for p in g["ref_pos"][:n]]
    rng = np.random.default_rng(0)
    for _ in range(10):
        q = tuple(float(x) for x in rng.integers(0, 80 if dim == 3 else 512, size=dim))
        d = np.sqrt(np.add.reduce((g["ref_pos"][:n] - np.array(q)) ** 2, axis=1))
        assert planner.find_nearest_neighbor(q) is planner.nodes[int(np.argmin(d))]
        planner.search_radius = 30.0 if dim == 2 else 12.0
        near = planner.find_nearby_nodes(mod.Node(q))
        assert [id(m) for m in near] == [id(planner.nodes[j]) for j in np.nonzero(d <= planner.search_radius)[0]]
    planner.iteration_no = 100
    assert planner.compute_search_radius(dim) == H.load("radius")["r2d_512" if dim == 2 else "r3d_80"][99]
    s = planner.generate_random_sample()
    assert len(s) == dim and (occ[s] != 0 if dim == 2 else occ[s] == 0)


def test_evaluation_harness_table(cuda_device):
    """The batched counterpart of test_network_and_algorithms.main(): same columns, lengths = calculate_length(path)."""
    from mgr_sampling_cnn_b200 import test_network_and_algorithms as harness, engine, workload
    harness.MAX_ITERATIONS = 600
    try:
        df, summary = harness.main(n_tasks=20, device=str(cuda_device))
    finally:
        harness.MAX_ITERATIONS = 5000
    assert list(df.columns.levels[0]) == ["A*", "Neural+RRT*", "RRT*"] and len(df) == 20
    assert list(summary.index) == ["mean", "median"]
    assert (df[("RRT*", "Iterations")].astype(int) <= 600).all() and (df[("Neural+RRT*", "Length")].astype(float) >= 0).all()
    # A* column: the shortest grid path bounds every solved RRT* path from below up to the goal threshold (5 px: the RRT* path
    # ends at the goal, but its vertices are free-space points, not grid moves: allow the two end snaps)
    astar = df[("A*", "Length")].astype(float).to_numpy()
    assert np.isfinite(astar).all() and (df[("A*", "Iterations")] == "-").all() and (df[("A*", "Time")].astype(float) > 0).all()
    assert ("A*", "Length") in summary.columns
    # Length column == calculate_length of the returned path points
    wl = workload.make_2d(20)
    import torch
    bits = engine.pack_occupancy(torch.from_numpy(wl.occ_maps).to(cuda_device), 2, cuda_device)
    res = engine.plan_batch(bits, wl.task_to_map, wl.starts, wl.goals, dim=2, shape=wl.shape, max_iterations=600, seed=0,
                            keep_tree=True)
    for b in range(20):
        assert harness.calculate_length(res.path_points(b)) == float(res.path_len[b]) == float(df[("RRT*", "Length")][b])
