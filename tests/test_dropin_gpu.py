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


@pytest.mark.parametrize("dim", [2, 3])
def test_generate_neural_sample_continues_the_stream_like_the_reference(dim, cuda_device):
    """300 direct generate_neural_sample() calls: the reference's recorded points AND its draw count, call by call, on
    the injected stream (no bias draw is consumed by the method itself; 3D redraws until the voxel is free)."""
    g = H.load("neural_draws")
    mod = _module("neural_rrt" if dim == 2 else "ThreeD_neural_rrt")
    if dim == 2:
        occ = np.repeat((g["occ2d"].astype(np.float32) / 255.0)[:, :, None], 3, axis=2)
        heat, seed, task, ref, nd = g["heat2d"], int(g["seed2d"]), int(g["task2d"]), g["draws2d"], int(g["ndraws2d"])
    else:
        occ = (g["occ3d"] != 0).astype(np.float32)
        heat, seed, task, ref, nd = g["heat3d"], int(g["seed3d"]), int(g["task3d"]), g["draws3d"], int(g["ndraws3d"])
    rnd = mod.random = StreamRandom(seed, task)
    planner = mod.RRTStar(occ, heat, (1,) * dim, (1,) * dim, 10, 1.0, 1.0)
    out = [planner.generate_neural_sample() for _ in range(60)]
    assert np.array_equal(np.array(out, np.int32), ref[:60])
    # host-side draws in between (what the reference's own loop does with `random.random() < bias`) keep the lock-step
    u_ref = StreamRandom(seed, task)
    u_ref.draws = rnd.draws
    assert rnd.random() == u_ref.random()
    rnd.draws -= 1
    out += [planner.generate_neural_sample() for _ in range(240)]
    assert np.array_equal(np.array(out, np.int32), ref) and rnd.draws == nd
    s = planner.generate_random_sample()
    assert rnd.draws >= nd + dim and (g["occ2d"][s] != 0 if dim == 2 else g["occ3d"][s] == 0)


def test_rrt_star_leaves_the_stream_where_the_reference_left_it(cuda_device):
    g = H.load("plan2d_neural_stream")
    mod = _module("neural_rrt")
    rnd = mod.random = StreamRandom(int(g["stream_seed"]), int(g["stream_task"]))
    occ = np.repeat((g["occ_map"].astype(np.float32) / 255.0)[:, :, None], 3, axis=2)
    planner = mod.RRTStar(occ, g["heat_map"], tuple(int(v) for v in g["start"]), tuple(int(v) for v in g["goal"]),
                          int(g["max_iterations"]), float(g["goal_threshold"]), float(g["neural_bias"]))
    _, iters = planner.rrt_star()
    assert iters == int(g["ref_iterations"]) and rnd.draws == int(g["ref_draws"])


@pytest.mark.parametrize("fixture,modname", [("plan2d_uniform_mt_0", "rrt_star"), ("plan3d_uniform_mt_0", "ThreeD_rrt_star")])
def test_main_loop_from_the_callable_methods(fixture, modname, cuda_device):
    """The reference's loop body (rrt_star.py:184-215) written out with the drop-in's individually callable methods --
    find_nearest_neighbor, steer, can_connect_nodes, rewire_tree, goal_reached -- on the reference's recorded samples:
    node positions, parents and costs after 120 iterations equal the reference's tree at that point."""
    g = H.load(fixture)
    mod = _module(modname)
    dim = H.fixture_dim(fixture)
    occ = g["occ_map"] if dim == 2 else g["occ_map"].astype(np.float64)
    start, goal = tuple(int(v) for v in g["start"]), tuple(int(v) for v in g["goal"])
    planner = mod.RRTStar(occ, start, goal, 120, float(g["goal_threshold"]))
    ref = mod.RRTStar(occ, start, goal, 120, float(g["goal_threshold"]))
    smp = np.zeros((120, dim))
    smp[:] = g["ref_samples"][:120]
    res = __import__("mgr_sampling_cnn_b200.engine", fromlist=["x"]).plan_batch(
        ref._bits, np.zeros(1, np.int32), np.array([start], np.int32), np.array([goal], np.int32), dim=dim, shape=occ.shape,
        max_iterations=120, goal_threshold=float(g["goal_threshold"]), samples=smp.astype(np.int16)[None], keep_tree=True)
    for i in range(120):
        planner.iteration_no = i + 1
        planner.search_radius = planner.compute_search_radius(dim)
        sample = tuple(int(v) for v in g["ref_samples"][i])
        nearest = planner.find_nearest_neighbor(sample)
        new = planner.steer(nearest, sample)
        if planner.can_connect_nodes(nearest, new):
            planner.rewire_tree(new)
            if planner.goal_reached(new, planner.goal):
                break
            planner.nodes.append(new)
    n = len(planner.nodes)
    assert n == int(res.n_nodes[0]) and n > 40
    index = {id(m): j for j, m in enumerate(planner.nodes)}
    assert np.array_equal(np.array([m.position for m in planner.nodes], np.float64), res.pos[0, :n].cpu().numpy())
    assert np.array_equal(np.array([m.cost for m in planner.nodes]), res.cost[0, :n].cpu().numpy())
    assert np.array_equal(np.array([-1 if m.parent is None else index[id(m.parent)] for m in planner.nodes]),
                          res.parent[0, :n].cpu().numpy())


def _write_task_files(tmp_path, three_d, n=3):
    """A tiny evaluation set in the reference's on-disk format (generate_tasks.py:76-82 names; train.py:38-60 reader)."""
    from mgr_sampling_cnn_b200 import workload
    wl = workload.make_3d(n, size=32) if three_d else workload.make_2d(n, size=128)
    (tmp_path / "images").mkdir()
    (tmp_path / "masks").mkdir()
    names = []
    for b in range(n):
        occ = wl.occ_maps[wl.task_to_map[b]]
        if three_d:
            s, f = wl.starts[b], wl.goals[b]
            name = f"map_0_path{b}_sx{s[0]}_sy{s[1]}_sz{s[2]}_fx{f[0]}_fy{f[1]}_fz{f[2]}.npy"
            np.save(tmp_path / "images" / name, occ)
            np.save(tmp_path / "masks" / name, np.random.default_rng(b).random(occ.shape))
        else:
            import cv2
            (sy, sx), (fy, fx) = wl.starts[b], wl.goals[b]
            name = f"map_0_path{b}_sx{sx}_sy{sy}_fx{fx}_fy{fy}.png"
            cv2.imwrite(str(tmp_path / "images" / name), occ)
            cv2.imwrite(str(tmp_path / "masks" / name), (np.random.default_rng(b).random(occ.shape) * 255).astype(np.uint8))
        names.append(name)
    return wl, names


@pytest.mark.parametrize("three_d", [False, True])
def test_generate_paths_demos(three_d, tmp_path, cuda_device):
    """generate_paths() of the four modules on task files in the reference's directory layout (weights: a saved random-init
    state_dict standing in for the trained .pth).  Start / finish as parsed from the file names; a path ends at the goal."""
    import torch
    wl, names = _write_task_files(tmp_path, three_d)
    uni = _module("ThreeD_rrt_star" if three_d else "rrt_star")
    neu = _module("ThreeD_neural_rrt" if three_d else "neural_rrt")
    uni.MAPS_DIRECTORY = str(tmp_path / "images" / ("*.npy" if three_d else "*.png"))
    assert [p.split("/")[-1] for p in uni.get_blank_maps_list()] == sorted(names)
    uni.MAX_ITERATIONS = neu.MAX_ITERATIONS = 1500
    try:
        map_path, path, iters = uni.generate_paths()
        s, f = uni.get_start_finish_coordinates(map_path)
        assert 1 <= iters <= 1500 and (not path or tuple(path[0]) == s)
        from mgr_sampling_cnn_b200 import ThreeD_train, train
        torch.manual_seed(0)
        model = (ThreeD_train.ThreeD_UNet_cooler() if three_d else train.UNet_cooler())
        torch.save(model.state_dict(), tmp_path / "weights.pth")
        neu.BASE_PATH, neu.MODEL_PATH = tmp_path, str(tmp_path / "weights.pth")
        assert len(neu.get_blank_maps_list()) == 3
        out = neu.generate_paths(index=0)
        assert 1 <= out["neural"][1] <= 1500 and out["neural"][2] > 0
        if not three_d:
            assert 1 <= out["plain"][1] <= 1500
    finally:
        uni.MAX_ITERATIONS = neu.MAX_ITERATIONS = 5000
