"""GPU: the CUDA planner (through the C ABI) against the reference's golden traces and the CPU oracle.  Bit-exact:
node positions, costs, parent indices, iteration counts, best/goal nodes, path length, and the ordered collision-verdict
list the reference would produce."""
import numpy as np
import pytest
import torch

from mgr_sampling_cnn_b200 import engine
from oracle import planner_oracle as po
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _run_gpu(name, g, dev, samples=None, trace_cap=0, cdf=None, seed=0, task_id=0):
    dim = H.fixture_dim(name)
    mi, thr = int(g["max_iterations"]), float(g["goal_threshold"])
    bits = engine.pack_occupancy(g["occ_map"][None], dim, dev)
    kw = {}
    if samples is not None:
        kw["samples"] = samples.astype(np.int16)[None]
    res = engine.plan_batch(bits, np.zeros(1, np.int32), g["start"][None], g["goal"][None], dim=dim,
                            shape=g["occ_map"].shape, max_iterations=mi, goal_threshold=thr, cdf=cdf,
                            neural_bias=float(g["neural_bias"]) if "neural_bias" in g else 0.75, seed=seed,
                            task_ids=np.array([task_id], np.int32), keep_tree=True, trace_cap=trace_cap, **kw)
    torch.cuda.synchronize()
    return res


def _check_against_reference(res, g, dim):
    n_tot = len(g["ref_parent"])
    assert int(res.iterations[0]) == int(g["ref_iterations"])
    assert int(res.n_nodes[0]) == int(g["ref_n_nodes"])
    assert (int(res.status[0]) == 0) == bool(g["ref_solved"])
    assert int(res.best_node[0]) == int(g["ref_best_node"])
    assert float(res.best_dist[0]) == float(g["ref_best_distance"])
    assert np.array_equal(res.parent[0, :n_tot].cpu().numpy(), g["ref_parent"])
    assert np.array_equal(res.pos[0, :n_tot].cpu().numpy(), g["ref_pos"])
    assert np.array_equal(res.cost[0, :n_tot].cpu().numpy(), g["ref_cost"])
    assert np.array_equal(res.path_points(0), g["ref_path"])
    ref_len = 0.0
    for a, b in zip(g["ref_path"][:-1], g["ref_path"][1:]):
        ref_len += float(np.sqrt(np.add.reduce((a - b) * (a - b))))  # scipy euclidean
    assert float(res.path_len[0]) == ref_len


@pytest.mark.parametrize("name", H.PLAN_FIXTURES)
@pytest.mark.parametrize("trace", [False, True])
def test_replay_reference_samples_bit_exact(name, trace, cuda_device):
    """Parity level A: the reference's recorded sample points drive the CUDA planner."""
    g = H.load(name)
    dim = H.fixture_dim(name)
    mi = int(g["max_iterations"])
    res = _run_gpu(name, g, cuda_device, samples=H.padded_samples(g, mi, dim), trace_cap=65536 if trace else 0)
    _check_against_reference(res, g, dim)
    if trace:
        ref_it, ref_v, _ = H.ref_trace_int(g, dim)
        o = po.plan(dim, g["occ_map"].shape, H.occ_free(name, g), po.radius_table(dim, g["occ_map"].shape, mi),
                    H.padded_samples(g, mi, dim), g["start"], g["goal"], mi, float(g["goal_threshold"]), pow_mode=1,
                    trace_cap=80000)
        tn = int(res.trace_n[0])
        assert tn == len(ref_it) == o["trace_len"]
        tr = res.trace[0, :tn].cpu().numpy()
        assert np.array_equal(tr[:, 0], ref_it) and np.array_equal(tr[:, 3], ref_v)  # vs the reference itself
        assert np.array_equal(tr, o["trace"])  # vs the oracle, including (kind, node index)


@pytest.mark.parametrize("name", H.STREAM_FIXTURES)
def test_shared_stream_end_to_end_bit_exact(name, cuda_device):
    """Parity level B: only the Philox uniforms are shared; CUDA CDF + CUDA sampler + CUDA planner vs the reference."""
    g = H.load(name)
    dim = H.fixture_dim(name)
    heat = H.planner_heat(name, g)
    cdf = None
    if heat is not None:
        cdf = engine.build_cdf(heat[None], engine.HEAT_MODE_RAW, cuda_device)
        assert np.array_equal(cdf[0].cpu().numpy(), po.build_cdf(heat)[0])
    res = _run_gpu(name, g, cuda_device, cdf=cdf, seed=int(g["stream_seed"]), task_id=int(g["stream_task"]))
    it = int(g["ref_iterations"])
    assert np.array_equal(res.samples[0, :it].cpu().numpy(), g["ref_samples"].astype(np.int16))
    if heat is not None:
        assert np.array_equal(res.neural_flag[0, :it].cpu().numpy(), g["ref_neural_flags"])
    _check_against_reference(res, g, dim)


def _oracle_batch(dim, shape, free_maps, t2m, starts, goals, seed, mi, thr, cdfs=None, bias=0.75):
    rad = po.radius_table(dim, shape, mi)
    outs = []
    for b in range(len(starts)):
        smp, _, nd, st = po.draw_samples(dim, shape, free_maps[t2m[b]], None if cdfs is None else cdfs[b], bias, seed, b, mi)
        assert st == 0
        o = po.plan(dim, shape, free_maps[t2m[b]], rad, smp.astype(np.float64), starts[b], goals[b], mi, thr, pow_mode=1)
        o["samples"] = smp
        outs.append(o)
    return outs


def _compare_batch(res, outs):
    for b, o in enumerate(outs):
        n = o["n_nodes"] + (1 if o["goal_node"] >= 0 else 0)
        assert int(res.n_nodes[b]) == o["n_nodes"], b
        assert int(res.iterations[b]) == o["iterations"], b
        assert int(res.status[b]) == o["status"], b
        assert int(res.goal_node[b]) == o["goal_node"] and int(res.best_node[b]) == o["best_node"], b
        assert np.array_equal(res.parent[b, :n].cpu().numpy(), o["parent"]), b
        assert np.array_equal(res.pos[b, :n].cpu().numpy(), o["pos"]), b
        assert np.array_equal(res.cost[b, :n].cpu().numpy(), o["cost"]), b
        assert int(res.path_n[b]) == len(o["path_idx"]), b
        assert np.array_equal(res.path_idx[b, :len(o["path_idx"])].cpu().numpy(), o["path_idx"]), b
        assert float(res.path_len[b]) == o["path_length"], b


def test_batch_2d_uniform_full_iterations_vs_oracle(cuda_device):
    """Config-3 shape: 2 maps x 12 tasks, MAX_ITERATIONS = 5000 (full size), uniform sampling."""
    mt = H.load("maps_tasks")
    maps = np.stack([mt["occ2d_0"], mt["occ2d_1"]])
    rng = np.random.default_rng(5)
    starts, goals, t2m = [], [], []
    for m in range(2):
        s, g = H.random_tasks_2d(rng, maps[m], 12)
        starts.append(s), goals.append(g), t2m.extend([m] * 12)
    starts, goals, t2m = np.concatenate(starts), np.concatenate(goals), np.array(t2m, np.int32)
    bits = engine.pack_occupancy(maps, 2, cuda_device)
    res = engine.plan_batch(bits, t2m, starts, goals, dim=2, shape=(512, 512), max_iterations=5000, seed=99, keep_tree=True)
    torch.cuda.synchronize()
    free = [po.occ_free_2d(m) for m in maps]
    _compare_batch(res, _oracle_batch(2, (512, 512), free, t2m, starts, goals, 99, 5000, 5.0))
    # more tasks than CTAs is exercised by the queue: same call, tree not kept, results identical
    res2 = engine.plan_batch(bits, t2m, starts, goals, dim=2, shape=(512, 512), max_iterations=5000, seed=99)
    torch.cuda.synchronize()
    assert torch.equal(res2.n_nodes, res.n_nodes) and torch.equal(res2.path_len, res.path_len)
    for b in range(len(starts)):
        n = int(res.n_nodes[b])
        assert torch.equal(res2.parent[b, :n], res.parent[b, :n]) and torch.equal(res2.cost[b, :n], res.cost[b, :n])


def test_batch_3d_neural_vs_oracle(cuda_device):
    mt = H.load("maps_tasks")
    occ = mt["occ3d_0"]
    rng = np.random.default_rng(6)
    starts, goals = H.random_tasks_3d(rng, occ, 8)
    heat = np.zeros((8, 80, 80, 80), dtype=np.float32)
    for b in range(8):  # a fuzzy tube of weights in (0.96, 1] around the start-goal segment, zeros elsewhere
        t = np.linspace(0, 1, 60)[:, None]
        pts = np.rint(starts[b] + t * (goals[b] - starts[b])).astype(int)
        for d in rng.integers(-2, 3, size=(6, 3)):
            q = np.clip(pts + d, 0, 79)
            heat[b, q[:, 0], q[:, 1], q[:, 2]] = (0.961 + 0.039 * rng.random(len(q))).astype(np.float32)
    h = (heat - 250).astype(np.float32)
    cdf = engine.build_cdf(h, engine.HEAT_MODE_RAW, cuda_device)
    cdfs = [po.build_cdf(h[b])[0] for b in range(8)]
    for b in range(8):
        assert np.array_equal(cdf[b].cpu().numpy(), cdfs[b])
    bits = engine.pack_occupancy(occ[None], 3, cuda_device)
    t2m = np.zeros(8, np.int32)
    res = engine.plan_batch(bits, t2m, starts, goals, dim=3, shape=(80, 80, 80), max_iterations=1500, cdf=cdf, seed=7,
                            keep_tree=True)
    torch.cuda.synchronize()
    _compare_batch(res, _oracle_batch(3, (80, 80, 80), [po.occ_free_3d(occ)], t2m, starts, goals, 7, 1500, 3.0, cdfs))


def test_edge_cases(cuda_device):
    """Empty batch; start == goal cell neighbourhood; a task with no free sample cell reports SAMPLER_STUCK."""
    occ = np.full((64, 64), 255, np.uint8)
    bits = engine.pack_occupancy(occ[None], 2, cuda_device)
    res = engine.plan_batch(bits, np.zeros(0, np.int32), np.zeros((0, 2), np.int32), np.zeros((0, 2), np.int32), dim=2,
                            shape=(64, 64), max_iterations=50)
    assert res.n_nodes.numel() == 0
    # goal within threshold of the first connectable node
    res = engine.plan_batch(bits, np.zeros(1, np.int32), np.array([[10, 10]]), np.array([[12, 12]]), dim=2,
                            shape=(64, 64), max_iterations=200, seed=3, keep_tree=True)
    torch.cuda.synchronize()
    o = _oracle_batch(2, (64, 64), [po.occ_free_2d(occ)], np.zeros(1, np.int32), np.array([[10, 10]]),
                      np.array([[12, 12]]), 3, 200, 5.0)
    _compare_batch(res, o)
    assert int(res.status[0]) == 0
    # fully blocked map: generate_random_sample would spin forever in the reference
    blocked = np.zeros((64, 64), np.uint8)
    bits = engine.pack_occupancy(blocked[None], 2, cuda_device)
    res = engine.plan_batch(bits, np.zeros(1, np.int32), np.array([[10, 10]]), np.array([[40, 40]]), dim=2,
                            shape=(64, 64), max_iterations=50, max_reject=1000)
    torch.cuda.synchronize()
    assert int(res.status[0]) == 3 and int(res.path_n[0]) == 0
