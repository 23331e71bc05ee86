"""TEST INFRASTRUCTURE — ctypes wrapper around oracle/_build/libnrrt_oracle.so (oracle/planner_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product package `mgr_sampling_cnn_b200` never does.
"""
import ctypes as C
import math
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnrrt_oracle.so")

ST_SOLVED, ST_MAX_ITER_BEST_EFFORT, ST_NO_NODE_CONNECTED, ST_SAMPLER_STUCK = 0, 1, 2, 3


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "planner_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.nrrto_uniform.restype = C.c_double
        L.nrrto_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64]
        L.nrrto_pairwise_sum_f32.restype = C.c_float
        L.nrrto_pairwise_sum_f32.argtypes = [C.c_void_p, C.c_long]
        L.nrrto_build_cdf.restype = None
        L.nrrto_build_cdf.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nrrto_cdf_search.restype = C.c_long
        L.nrrto_cdf_search.argtypes = [C.c_void_p, C.c_long, C.c_double]
        L.nrrto_draw_samples.restype = C.c_int
        L.nrrto_draw_samples.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_uint64, C.c_uint32,
                                         C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nrrto_plan.restype = C.c_int
        L.nrrto_plan.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int, C.c_double, C.c_int] + [C.c_void_p] * 10 + [C.c_void_p, C.c_int, C.c_void_p,
                                                                                      C.c_void_p, C.c_void_p]
        L.nrrto_path.restype = C.c_int
        L.nrrto_path.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.nrrto_plan_batch.restype = C.c_int
        L.nrrto_plan_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int, C.c_double,
                                       C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def radius_table(dim: int, shape, max_iter: int) -> np.ndarray:
    """A.1 compute_search_radius for i = 1..max_iter (rrt_star.py:174-182, ThreeD_rrt_star.py:183-191), Python `math`.

    2D: map_height, map_width = shape; V = W*H.   3D: map_height, map_width, map_depth = shape; V = W*H*W (sic, :187).
    """
    if dim == 2:
        vol = shape[1] * shape[0]
    else:
        vol = shape[1] * shape[0] * shape[1]
    leb = math.pow(math.pi, dim / 2.0) / math.gamma((dim / 2.0) + 1)
    out = np.empty(max_iter, dtype=np.float64)
    for i in range(1, max_iter + 1):
        out[i - 1] = math.pow(2 * (1 + 1.0 / dim) * (vol / leb) * (math.log(i) / i), 1.0 / dim)
    return out


def occ_free_2d(occ_map: np.ndarray) -> np.ndarray:
    """2D: free iff value != 0 (rrt_star.py:67,120)."""
    return (np.asarray(occ_map) != 0).astype(np.uint8)


def occ_free_3d(occ_map: np.ndarray) -> np.ndarray:
    """3D: free iff value == 0 (ThreeD_rrt_star.py:69,128)."""
    return (np.asarray(occ_map) == 0).astype(np.uint8)


def uniform(seed: int, task: int, k: int) -> float:
    return lib().nrrto_uniform(seed, task, k)


def pairwise_sum_f32(a: np.ndarray) -> np.float32:
    a = np.ascontiguousarray(a, dtype=np.float32)
    return np.float32(lib().nrrto_pairwise_sum_f32(_p(a), a.size))


def build_cdf(h: np.ndarray):
    """h = planner-side heat map (float32, any shape, flattened row-major). Returns (cdf float32[N], min, sum)."""
    h = np.ascontiguousarray(h, dtype=np.float32).ravel()
    cdf = np.empty_like(h)
    m, s = C.c_float(), C.c_float()
    lib().nrrto_build_cdf(_p(h), h.size, _p(cdf), C.byref(m), C.byref(s))
    return cdf, np.float32(m.value), np.float32(s.value)


def cdf_search(cdf: np.ndarray, u: float) -> int:
    return lib().nrrto_cdf_search(_p(cdf), cdf.size, u)


def draw_samples(dim, shape, occ_free, cdf, neural_bias, seed, task, n_iter, max_reject=100000):
    shape_a = np.asarray(list(shape) + [1] * (3 - len(shape)), dtype=np.int32)
    occ_free = np.ascontiguousarray(occ_free, dtype=np.uint8)
    samples = np.zeros((n_iter, dim), dtype=np.int32)
    flags = np.zeros(n_iter, dtype=np.int8)
    nd = C.c_int64()
    st = lib().nrrto_draw_samples(dim, _p(shape_a), _p(occ_free), _p(cdf), float(neural_bias), int(seed), int(task),
                                  n_iter, max_reject, _p(samples), _p(flags), C.byref(nd))
    return samples, flags, nd.value, st


def plan(dim, shape, occ_free, radius, samples, start, goal, max_iter, goal_thr, pow_mode=0, trace_cap=0):
    """One plan driven by pre-drawn samples. Returns a dict mirroring NrrtPlanOut (+ trace)."""
    shape_a = np.asarray(list(shape) + [1] * (3 - len(shape)), dtype=np.int32)
    occ_free = np.ascontiguousarray(occ_free, dtype=np.uint8)
    radius = np.ascontiguousarray(radius, dtype=np.float64)
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    assert samples.shape[0] >= max_iter and radius.shape[0] >= max_iter
    start = np.asarray(start, dtype=np.float64)
    goal = np.asarray(goal, dtype=np.float64)
    pos = np.zeros((max_iter + 2, dim), dtype=np.float64)
    cost = np.zeros(max_iter + 2, dtype=np.float64)
    parent = np.full(max_iter + 2, -1, dtype=np.int32)
    scratch = np.zeros(max_iter + 2, dtype=np.int32)
    trace = np.zeros((max(trace_cap, 1), 4), dtype=np.int32)
    seg = np.zeros((max(trace_cap, 1), 2 * dim), dtype=np.float64)
    n_nodes, iters, status, goal_node, best_node, tl, cl = (C.c_int32() for _ in range(7))
    best = C.c_double()
    lib().nrrto_plan(dim, _p(shape_a), _p(occ_free), _p(radius), _p(samples), _p(start), _p(goal), max_iter,
                     float(goal_thr), pow_mode, _p(pos), _p(cost), _p(parent), _p(scratch),
                     C.byref(n_nodes), C.byref(iters), C.byref(status), C.byref(goal_node), C.byref(best_node),
                     C.byref(best), _p(trace) if trace_cap else None, trace_cap, C.byref(tl),
                     _p(seg) if trace_cap else None, C.byref(cl))
    n_tot = n_nodes.value + (1 if goal_node.value >= 0 else 0)
    end = goal_node.value if goal_node.value >= 0 else best_node.value
    path_idx = np.zeros(max_iter + 2, dtype=np.int32)
    length = C.c_double(0.0)
    m = 0
    if end >= 0:
        m = lib().nrrto_path(dim, _p(pos), _p(parent), end, max_iter + 2, _p(path_idx), C.byref(length))
    return {
        "pos": pos[:n_tot].copy(), "cost": cost[:n_tot].copy(), "parent": parent[:n_tot].copy(),
        "n_nodes": n_nodes.value, "iterations": iters.value, "status": status.value, "goal_node": goal_node.value,
        "best_node": best_node.value, "best_distance": best.value, "path_idx": path_idx[:m].copy(),
        "path": pos[path_idx[:m]].copy(), "path_length": length.value,
        "trace": trace[:min(tl.value, trace_cap)].copy(), "trace_len": tl.value,
        "segments": seg[:min(tl.value, trace_cap)].copy(), "clamped": cl.value,
    }


def plan_batch(dim, shape, occ_free_maps, task_to_map, cdf, neural_bias, radius, starts, goals, seed, task_ids,
               max_iter, goal_thr, max_reject=100000, pow_mode=1, n_threads=None):
    """CPU baseline: B independent plans over host threads (one C call per task; ctypes drops the GIL)."""
    B = len(task_ids)
    shape_a = np.asarray(list(shape) + [1] * (3 - len(shape)), dtype=np.int32)
    occ_free_maps = np.ascontiguousarray(occ_free_maps, dtype=np.uint8)
    task_to_map = np.ascontiguousarray(task_to_map, dtype=np.int32)
    starts = np.ascontiguousarray(starts, dtype=np.int32)
    goals = np.ascontiguousarray(goals, dtype=np.int32)
    task_ids = np.ascontiguousarray(task_ids, dtype=np.int32)
    radius = np.ascontiguousarray(radius, dtype=np.float64)
    if cdf is not None:
        cdf = np.ascontiguousarray(cdf, dtype=np.float32)
    n_nodes = np.zeros(B, dtype=np.int32)
    iters = np.zeros(B, dtype=np.int32)
    status = np.zeros(B, dtype=np.int32)
    plen = np.zeros(B, dtype=np.float64)
    L = lib()
    ncell = int(np.prod(shape))

    def one(b):
        L.nrrto_plan_batch(dim, _p(shape_a), _p(occ_free_maps), _p(task_to_map[b:]),
                           None if cdf is None else C.c_void_p(cdf.ctypes.data + 4 * ncell * b), float(neural_bias),
                           _p(radius), _p(starts[b:]), _p(goals[b:]), int(seed), _p(task_ids[b:]), 1, max_iter,
                           float(goal_thr), max_reject, pow_mode, _p(n_nodes[b:]), _p(iters[b:]), _p(status[b:]),
                           _p(plen[b:]))

    n_threads = n_threads or os.cpu_count() or 1
    with ThreadPoolExecutor(max_workers=n_threads) as ex:
        list(ex.map(one, range(B)))
    return {"n_nodes": n_nodes, "iterations": iters, "status": status, "path_length": plen, "threads": n_threads}
