"""Batched planning engine: thousands of independent RRT* tasks per call on one B200 through libnrrt.so.

PyTorch is used for device memory and streams only; every computation on the planning path is a hand-written CUDA
kernel behind the C ABI in include/nrrt.h.  The reference's Python classes (`rrt_star.RRTStar` etc.) are thin B=1 views
over this engine (see the sibling drop-in modules).
"""
import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import NrrtPlanOut, NrrtPlanParams, check

HEAT_MODE_2D, HEAT_MODE_3D, HEAT_MODE_RAW = 0, 1, 2
_DTYPE_CODE = {torch.uint8: 0, torch.bool: 0, torch.float32: 1, torch.float64: 2}
_radius_cache = {}


def radius_table(dim: int, shape: Sequence[int], max_iterations: int) -> np.ndarray:
    """compute_search_radius for i = 1..max_iterations (rrt_star.py:174-182, ThreeD_rrt_star.py:183-191).

    Task-independent, so it is evaluated once on the host with Python `math` — exactly the reference's expression,
    including the 3D volume `map_width * map_height * map_width` (ThreeD_rrt_star.py:187) and r_1 = 0.
    """
    key = (dim, tuple(int(s) for s in shape), int(max_iterations))
    tab = _radius_cache.get(key)
    if tab is None:
        if dim == 2:
            map_height, map_width = key[1][:2]
            volume = map_width * map_height
        else:
            map_height, map_width, _ = key[1][:3]
            volume = map_width * map_height * map_width
        lebesgue = math.pow(math.pi, dim / 2.0) / math.gamma((dim / 2.0) + 1)
        tab = np.empty(max_iterations, dtype=np.float64)
        for i in range(1, max_iterations + 1):
            tab[i - 1] = math.pow(2 * (1 + 1.0 / dim) * (volume / lebesgue) * (math.log(i) / i), 1.0 / dim)
        _radius_cache[key] = tab
    return tab


def _dev(t, dtype=None, device=None):
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t))
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if device is not None and t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(device) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.NrrtError("mgr_sampling_cnn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device(device if device is not None else "cuda")


def pack_occupancy(occ_maps, dim: int, device=None) -> torch.Tensor:
    """[n_maps, *shape] occupancy (uint8 / float32 / float64) -> bit-packed [n_maps, words] uint32 (1 = free).

    2D: free iff value != 0 (rrt_star.py:67,120).  3D: free iff value == 0 (ThreeD_rrt_star.py:69,128).
    """
    device = _require_cuda(device)
    occ = _dev(occ_maps, device=device)
    if occ.dtype == torch.bool:
        occ = occ.view(torch.uint8)
    if occ.dtype not in _DTYPE_CODE:
        occ = occ.to(torch.float32)
    n_maps = occ.shape[0]
    cells = int(np.prod(occ.shape[1:]))
    lib = _lib.load()
    words = lib.nrrt_words_per_map(cells)
    bits = torch.empty((n_maps, words), dtype=torch.int32, device=device)
    _lib.count_launch()
    check(lib.nrrt_pack_occupancy(_ptr(occ), _DTYPE_CODE[occ.dtype], 1 if dim == 2 else 0, n_maps, cells, _ptr(bits),
                                  _stream_ptr(device)))
    return bits


def build_cdf(logits, mode: int, device=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Network output (or a ready heat map, mode RAW) [B, cells...] float32 -> sampler CDF [B, cells] float32.

    Post-processing + `generate_neural_sample`'s normalisation/cumsum with NumPy's fp32 rounding,
    hoisted out of the per-draw path.
    """
    device = _require_cuda(device)
    x = _dev(logits, torch.float32, device)
    B = x.shape[0]
    x = x.reshape(B, -1)
    cells = x.shape[1]
    lib = _lib.load()
    ws_bytes = lib.nrrt_cdf_workspace_bytes(B, cells)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    cdf = torch.empty((B, cells), dtype=torch.float32, device=device) if out is None else out
    _lib.count_launch()
    check(lib.nrrt_cdf_build(_ptr(x), mode, B, cells, _ptr(cdf), _ptr(ws), ws_bytes, _stream_ptr(device)))
    return cdf


@dataclass
class PlanResult:
    """Per-task results (device tensors).  Mirrors NrrtPlanOut in include/nrrt.h."""
    n_nodes: torch.Tensor
    iterations: torch.Tensor
    status: torch.Tensor
    goal_node: torch.Tensor
    best_node: torch.Tensor
    best_dist: torch.Tensor
    path_n: torch.Tensor
    path_idx: torch.Tensor
    path_len: torch.Tensor
    cost: torch.Tensor
    parent: torch.Tensor
    pos: Optional[torch.Tensor] = None
    samples: Optional[torch.Tensor] = None
    neural_flag: Optional[torch.Tensor] = None
    n_draws: Optional[torch.Tensor] = None
    trace: Optional[torch.Tensor] = None
    trace_n: Optional[torch.Tensor] = None

    def solved_fraction(self) -> float:
        return float((self.status == _lib.NRRT_SOLVED).float().mean().item())

    def path_points(self, b: int) -> np.ndarray:
        """find_path() of task b as an (m, dim) float64 array (needs keep_tree=True)."""
        if self.pos is None:
            raise ValueError("plan_batch(keep_tree=True) is required for path_points")
        m = int(self.path_n[b].item())
        idx = self.path_idx[b, :m].long()
        return self.pos[b, idx].cpu().numpy()


class PlannerBuffers:
    """Caller-owned device buffers for one (dim, B, max_iterations) configuration, reusable across calls."""

    def __init__(self, dim, B, max_iterations, path_cap, keep_tree, trace_cap, device):
        stride = max_iterations + 2
        self.node_stride = stride
        self.key = (int(dim), int(B), int(max_iterations), max(int(path_cap), 1), int(trace_cap))
        i32, f64 = torch.int32, torch.float64
        z = lambda *s, dtype=i32: torch.empty(s, dtype=dtype, device=device)  # noqa: E731
        self.cost = z(B, stride, dtype=f64)
        self.parent = z(B, stride)
        self.pos = z(B, stride, dim, dtype=f64)  # also the kernel's store for nodes with non-integer coordinates
        self.n_nodes, self.iterations, self.status = z(B), z(B), z(B)
        self.goal_node, self.best_node = z(B), z(B)
        self.best_dist = z(B, dtype=f64)
        self.path_n = z(B)
        self.path_idx = z(B, max(path_cap, 1))
        self.path_len = z(B, dtype=f64)
        self.trace = z(B, trace_cap, 4) if trace_cap > 0 else None
        self.trace_n = z(B) if trace_cap > 0 else None
        self.samples = z(B, max_iterations, dim, dtype=torch.int16)
        self.neural_flag = z(B, max_iterations, dtype=torch.int8)
        self.n_draws = z(B, dtype=torch.int64)
        self.draw_status = z(B)
        self.queue = z(1)


def plan_batch(occ_bits: torch.Tensor, task_to_map, starts, goals, *, dim: int, shape: Sequence[int],
               max_iterations: int = 5000, goal_threshold: Optional[float] = None, cdf: Optional[torch.Tensor] = None,
               neural_bias: float = 0.75, seed: int = 0, task_ids=None, samples=None, keep_tree: bool = False,
               trace_cap: int = 0, path_cap: int = 512, max_reject: int = 100000,
               buffers: Optional[PlannerBuffers] = None, draw_start: int = 0, events: Optional[list] = None) -> PlanResult:
    """Run B independent RRT* plans (`RRTStar(...).rrt_star()` of the four reference modules) on the current stream.

    occ_bits   [n_maps, words] from pack_occupancy;  task_to_map [B];  starts/goals [B, dim] int (reference order:
               (y, x) in 2D, (x, y, z) in 3D);  cdf [B, cells] from build_cdf selects the neural modules, None the uniform ones.
    samples    optional pre-drawn sample points [B, max_iterations, dim] int16 (parity level A); otherwise they are
               drawn on the device from the Philox stream (seed, task_ids[b]), starting at draw `draw_start`
               (PlanResult.n_draws = where every stream stopped: pass it on to continue a stream).
    """
    device = occ_bits.device
    if device.type != "cuda":
        raise _lib.NrrtError("occ_bits must live on a CUDA device")
    lib = _lib.load()
    shape = [int(s) for s in shape]
    if goal_threshold is None:
        goal_threshold = 5.0 if dim == 2 else 3.0
    starts = _dev(starts, torch.int32, device).reshape(-1, dim)
    goals = _dev(goals, torch.int32, device).reshape(-1, dim)
    B = starts.shape[0]
    task_to_map = _dev(task_to_map, torch.int32, device).reshape(-1)
    task_ids = _dev(np.arange(B, dtype=np.int32) if task_ids is None else task_ids, torch.int32, device).reshape(-1)
    if not (goals.shape[0] == B and task_to_map.shape[0] == B and task_ids.shape[0] == B):
        raise ValueError("starts, goals, task_to_map and task_ids must agree on the batch size")
    radius = _dev(radius_table(dim, shape, max_iterations), torch.float64, device)
    buf = buffers or PlannerBuffers(dim, B, max_iterations, path_cap, keep_tree, trace_cap, device)
    kd, kB, kit, kpath, ktrace = buf.key
    if kd != dim or kB < B or kit != max_iterations or kpath < path_cap or ktrace < trace_cap or buf.cost.device != device:
        raise ValueError(f"PlannerBuffers were sized for (dim, B, max_iterations, path_cap, trace_cap) = {buf.key} on "
                         f"{buf.cost.device}; this call needs ({dim}, {B}, {max_iterations}, {path_cap}, {trace_cap}) on {device}")

    p = NrrtPlanParams()
    p.dim = dim
    for k in range(3):
        p.shape[k] = shape[k] if k < len(shape) else 1
    p.max_iterations, p.node_stride, p.max_reject, p.path_cap = max_iterations, buf.node_stride, max_reject, path_cap
    p.goal_threshold, p.neural_bias, p.seed = float(goal_threshold), float(neural_bias), int(seed) & (2 ** 64 - 1)
    p.draw_start, p.sampler_mode = int(draw_start), 0
    st = _stream_ptr(device)

    draw_status = None
    if samples is None:
        if cdf is not None:
            cdf = _dev(cdf, torch.float32, device)
        _lib.count_launch()
        check(lib.nrrt_draw_samples(C.byref(p), _ptr(occ_bits), _ptr(task_to_map), _ptr(cdf), _ptr(task_ids), B,
                                    _ptr(buf.samples), _ptr(buf.neural_flag), _ptr(buf.n_draws), _ptr(buf.draw_status), st))
        samples_t, draw_status = buf.samples, buf.draw_status
        if events is not None:  # measurement only: a CUDA event between the draw pass and the planner
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            events.append(e)
    else:
        samples_t = _dev(samples, torch.int16, device).reshape(B, max_iterations, dim)

    out = NrrtPlanOut()
    out.pos, out.cost, out.parent = _ptr(buf.pos), _ptr(buf.cost), _ptr(buf.parent)
    out.n_nodes, out.iterations, out.status = _ptr(buf.n_nodes), _ptr(buf.iterations), _ptr(buf.status)
    out.goal_node, out.best_node, out.best_dist = _ptr(buf.goal_node), _ptr(buf.best_node), _ptr(buf.best_dist)
    out.path_n, out.path_idx, out.path_len = _ptr(buf.path_n), _ptr(buf.path_idx), _ptr(buf.path_len)
    out.trace, out.trace_n, out.trace_cap = _ptr(buf.trace), _ptr(buf.trace_n), trace_cap
    _lib.count_launch()
    check(lib.nrrt_plan_batch(C.byref(p), _ptr(occ_bits), _ptr(task_to_map), _ptr(radius), _ptr(samples_t), _ptr(starts),
                              _ptr(goals), _ptr(draw_status), C.byref(out), B, _ptr(buf.queue), st))
    v = (lambda t: t) if kB == B else (lambda t: None if t is None else t[:B])  # a larger buffer set: views of its first B rows
    return PlanResult(n_nodes=v(buf.n_nodes), iterations=v(buf.iterations), status=v(buf.status), goal_node=v(buf.goal_node),
                      best_node=v(buf.best_node), best_dist=v(buf.best_dist), path_n=v(buf.path_n), path_idx=v(buf.path_idx),
                      path_len=v(buf.path_len), cost=v(buf.cost), parent=v(buf.parent), pos=v(buf.pos), samples=v(samples_t),
                      neural_flag=v(buf.neural_flag) if samples is None else None,
                      n_draws=v(buf.n_draws) if samples is None else None, trace=v(buf.trace), trace_n=v(buf.trace_n))


def grid_paths(occ_bits: torch.Tensor, task_to_map, starts, goals, shape: Sequence[int], path_cap: int = 2048,
               n_slots: Optional[int] = None):
    """generate_paths.astar / ThreeD_generate_paths.astar for a batch: (length [B], n_steps [B, 3], paths int32 [B, path_cap, dim]
    start -> goal in array-axis order, path_n [B]) on the device -- one shortest path per task (the mask makers' input; see
    nrrt_grid_paths_ex in include/nrrt.h for how ties between equal-cost paths are broken)."""
    return grid_path_lengths(occ_bits, task_to_map, starts, goals, shape, n_slots=n_slots, path_cap=path_cap)


def grid_path_lengths(occ_bits: torch.Tensor, task_to_map, starts, goals, shape: Sequence[int], n_slots: Optional[int] = None,
                      axis_only: bool = False, path_cap: int = 0):
    """Shortest grid path of every task = the reference's A* (generate_paths.py:11-74, ThreeD_generate_paths.py:14-97): its
    path length column (test_network_and_algorithms.py:58-70) and its task filter (no path -> the task is deleted,
    generate_paths.py:111-113).  occ_bits from pack_occupancy; starts / goals [B, dim] in array-axis order ((y, x) / (x, y, z)).
    Returns (length fp64 [B] (NaN = no path), n_steps int32 [B, 3] = steps of cost 1, sqrt(2), sqrt(3), -1 = no path) on
    the device; length = n_steps . (1, sqrt(2), sqrt(3)).  axis_only: axis moves only, i.e. reachability through 4- / 6-connected
    free cells (the task filter of workload.py; in 2D it accepts exactly the tasks the reference's A* solves)."""
    lib = _lib.load()
    dev = occ_bits.device
    if dev.type != "cuda":
        raise _lib.NrrtError("grid_path_lengths needs CUDA tensors: there is no CPU fallback")
    dim = len(shape)
    if dim not in (2, 3):
        raise ValueError("shape must be (H, W) or (X, Y, Z)")
    t2m = _dev(task_to_map, torch.int32, dev).reshape(-1)
    st = _dev(starts, torch.int32, dev).reshape(-1, dim).contiguous()
    gl = _dev(goals, torch.int32, dev).reshape(-1, dim).contiguous()
    B = t2m.shape[0]
    cells = int(np.prod(shape))
    n_slots = n_slots or torch.cuda.get_device_properties(dev).multi_processor_count * (2 if dim == 2 else 1)  # resident CTAs
    work = torch.empty((max(min(n_slots, B), 1), cells), dtype=torch.int64, device=dev)
    queue = torch.zeros(1, dtype=torch.int32, device=dev)
    steps = torch.empty((B, 3), dtype=torch.int32, device=dev)
    length = torch.empty(B, dtype=torch.float64, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.count_launch()
    s = [int(v) for v in shape] + [1]
    if path_cap > 0:
        paths = torch.zeros((B, path_cap, dim), dtype=torch.int32, device=dev)
        path_n = torch.zeros(B, dtype=torch.int32, device=dev)
        check(lib.nrrt_grid_paths_ex(dim, _ptr(occ_bits), occ_bits.shape[1], _ptr(t2m), _ptr(st), _ptr(gl), B, s[0], s[1], s[2],
                                     1 if axis_only else 0, _ptr(work), work.shape[0], _ptr(queue), _ptr(steps), _ptr(length),
                                     _ptr(paths), path_cap, _ptr(path_n), stream))
        return length, steps, paths, path_n
    check(lib.nrrt_grid_paths(dim, _ptr(occ_bits), occ_bits.shape[1], _ptr(t2m), _ptr(st), _ptr(gl), B, s[0], s[1], s[2],
                              1 if axis_only else 0, _ptr(work), work.shape[0], _ptr(queue), _ptr(steps), _ptr(length), stream))
    return length, steps


class GuidedPlanner(object):
    """The whole hot path for a batch of tasks: CNN probability maps -> sampler CDFs -> sample draws -> guided RRT*.

    Equivalent to running, for every task, `model(image, coords)` + post-processing + `neural_rrt.RRTStar(...).rrt_star()`
    (neural_rrt.py:319-369, test_network_and_algorithms.py:107-134; 3D twins), but batched on one GPU.  `model` is a
    `train.UNet_cooler` / `ThreeD_train.ThreeD_UNet_cooler` from this package (eval mode, on the CUDA device).
    """

    def __init__(self, model, dim: int, shape: Sequence[int], max_iterations: int = 5000,
                 goal_threshold: Optional[float] = None, neural_bias: float = 0.75, micro_batch: Optional[int] = None,
                 pipeline_batch: int = 4096, path_cap: int = 512, max_reject: int = 100000, keep_tree: bool = False):
        self.model, self.dim, self.shape = model, dim, tuple(int(s) for s in shape)
        self.max_iterations, self.neural_bias = max_iterations, neural_bias
        self.goal_threshold = goal_threshold if goal_threshold is not None else (5.0 if dim == 2 else 3.0)
        # tasks per decoder launch: four maps' worth in 2D (fewer, longer launches: +5 % over one map's worth), half a map in 3D
        self.micro_batch = micro_batch or (40 if dim == 2 else 10)
        self.pipeline_batch, self.path_cap, self.max_reject = pipeline_batch, path_cap, max_reject
        self.keep_tree = keep_tree  # plan() returns the trees, samples and flags too (single pipeline batch only)
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise _lib.NrrtError("GuidedPlanner needs the model on a CUDA device")
        self.cells = int(np.prod(self.shape))
        self._bufs = None
        self._logits = None
        self._cdf = None
        self.share_encoder = True
        self.single_call = micro_batch is None  # nrrt_conv_forward_* (fixed micro-batch); False: the Python layer graph
        self.encoder_batch = 8  # distinct maps per encoder pass

    def _ensure(self, B):
        if self._bufs is None or self._logits.shape[0] != B:
            self._logits = torch.empty((B, self.cells), dtype=torch.float32, device=self.device)
            self._cdf = torch.empty((B, self.cells), dtype=torch.float32, device=self.device)
            self._bufs = PlannerBuffers(self.dim, B, self.max_iterations, self.path_cap, False, 0, self.device)

    def probability_maps(self, occ_maps: torch.Tensor, task_to_map: torch.Tensor, coords: torch.Tensor,
                         out: Optional[torch.Tensor] = None, task_to_map_host: Optional[np.ndarray] = None) -> torch.Tensor:
        """CNN logits [B, cells] fp32 for tasks (occ_maps[task_to_map[b]], coords[b]); occ_maps uint8 on the device.

        The encoder runs once per distinct map of a micro-batch (tasks of one map are adjacent in the reference's
        task files, generate_tasks.py:18-24), the decoder once per task.

        3D: occ_maps are the task files' arrays [x][y][z]; the network sees them last axis first ([z][x][y], the
        dataset's ToTensorV2, ThreeD_train.py:59-66) and its output is consumed un-transposed as heat[x][y][z]
        (ThreeD_test_network_and_algorithms.py:98-105, transpose commented out) -- reproduced as is, so only cubic
        volumes are meaningful (as in the reference)."""
        if self.dim == 3 and len(set(self.shape)) != 1:
            raise ValueError("3D maps must be cubic: the reference feeds the network [z][x][y] and reads it back as [x][y][z]")
        plan = self.model._plan(self.device)
        B = task_to_map.shape[0]
        out = out if out is not None else torch.empty((B, self.cells), dtype=torch.float32, device=self.device)
        mb = self.micro_batch
        t2m_h = task_to_map_host if task_to_map_host is not None else task_to_map.cpu().numpy()
        grid = (1,) + self.shape if self.dim == 2 else self.shape
        grouped = not np.any(np.diff(t2m_h) < 0)
        if self.share_encoder and grouped and self.single_call:
            prep = None
        else:
            prep = plan.coords_prepare(coords, grid)  # coordinate branch (+ 2D epilogue polynomials) for every task at once
        if not self.share_encoder or not grouped:
            # tasks not grouped by map (or sharing disabled): one encoder pass per task
            for i in range(0, B, mb):
                plan.forward_maps(occ_maps, task_to_map[i:i + mb], coords[i:i + mb], logits_out=out[i:i + mb],
                                  prep=plan.slice_prep(prep, i, i + mb))
            return out
        uniq, first = np.unique(t2m_h, return_index=True)
        if self.single_call:
            # tasks are grouped by map: the whole forward is ONE library call (nrrt_conv_forward_*, csrc/forward.cu), which
            # encodes 8 distinct maps at a time and decodes their tasks in micro-batches
            row = np.searchsorted(uniq, t2m_h).astype(np.int32)
            runs = np.diff(np.append(first, B))
            # scheduling hint: the usual number of tasks per map (the library applies it per micro-batch, only where every
            # map of the micro-batch has exactly that many tasks; a truncated last map just falls back)
            per = int(np.bincount(runs).argmax()) if len(runs) else 0
            plan.forward_tasks(occ_maps, uniq.astype(np.int32), row, coords, out, tasks_per_map=per)
            return out
        # the same graph driven layer by layer from Python (cnn.py; kept for the tests that compare the two)
        G = self.encoder_batch
        rel = np.searchsorted(uniq, t2m_h).astype(np.int32) % G  # row of the task's map inside its encoder pass
        uniq_d = torch.from_numpy(uniq.astype(np.int32)).to(self.device, non_blocking=True)
        rel_d = torch.from_numpy(rel).to(self.device, non_blocking=True)
        bounds = list(first) + [B]
        for g0 in range(0, len(uniq), G):
            g1 = min(len(uniq), g0 + G)
            enc = plan.encode(grid, g1 - g0, occ=occ_maps, maps=uniq_d[g0:g1])
            lo, hi = bounds[g0], bounds[g1]
            for i in range(lo, hi, mb):
                j = min(hi, i + mb)
                # scheduling hint: tasks per map when every map of the micro-batch has the same number of tasks
                runs = np.diff(np.flatnonzero(np.diff(rel[i:j], prepend=-1, append=-1)))
                rg = int(runs[0]) if len(runs) and np.all(runs == runs[0]) else 0
                plan.decode(enc, j - i, grid, coords[i:j], plan.slice_prep(prep, i, j), rel_d[i:j], out[i:j], res_group=rg)
        return out

    def plan(self, occ_maps, task_to_map, starts, goals, coords, seed: int = 0, task_ids=None, record_events=False):
        """occ_maps [n_maps, *shape] uint8 (2D: 0 = obstacle; 3D: 0 = free), host or device; the rest per task.
        Returns one PlanResult of device tensors holding the per-task result records (copies: a later plan() call does
        not disturb them).  With keep_tree=True the result also carries the trees / samples / flags as VIEWS of the
        planner's reusable buffers, valid until the next plan() call; that mode takes B <= pipeline_batch."""
        dev = self.device
        occ = _dev(occ_maps, torch.uint8, dev)
        t2m_host = (task_to_map.cpu().numpy() if isinstance(task_to_map, torch.Tensor) else np.asarray(task_to_map)).reshape(-1)
        t2m = _dev(task_to_map, torch.int32, dev).reshape(-1)
        starts = _dev(starts, torch.int32, dev).reshape(-1, self.dim)
        goals = _dev(goals, torch.int32, dev).reshape(-1, self.dim)
        coords = _dev(coords, torch.float32, dev)
        B = t2m.shape[0]
        if self.keep_tree and B > self.pipeline_batch:
            raise ValueError("keep_tree=True needs B <= pipeline_batch")
        ids = _dev(np.arange(B, dtype=np.int32) if task_ids is None else task_ids, torch.int32, dev).reshape(-1)
        bits = pack_occupancy(occ, self.dim, dev)
        mode = HEAT_MODE_2D if self.dim == 2 else HEAT_MODE_3D
        outs = []
        ev = []
        for b0 in range(0, B, self.pipeline_batch):
            b1 = min(B, b0 + self.pipeline_batch)
            n = b1 - b0
            self._ensure(n)
            if record_events:
                ev.append([torch.cuda.Event(enable_timing=True) for _ in range(4)])
                ev[-1][0].record()
            prof = self.model._plan(dev).profiler
            if prof is not None:
                prof.active = bool(record_events)
            self.probability_maps(occ, t2m[b0:b1], coords[b0:b1], out=self._logits, task_to_map_host=t2m_host[b0:b1])
            if prof is not None:
                prof.active = False
            if record_events:
                ev[-1][1].record()
            build_cdf(self._logits, mode, dev, out=self._cdf)
            if record_events:
                ev[-1][2].record()
            res = plan_batch(bits, t2m[b0:b1], starts[b0:b1], goals[b0:b1], dim=self.dim, shape=self.shape,
                             max_iterations=self.max_iterations, goal_threshold=self.goal_threshold, cdf=self._cdf,
                             neural_bias=self.neural_bias, seed=seed, task_ids=ids[b0:b1], path_cap=self.path_cap,
                             max_reject=self.max_reject, buffers=self._bufs, events=ev[-1] if record_events else None)
            if record_events:
                ev[-1][3].record()
            # the buffers are reused by the next pipeline batch / the next call: the result records are copies
            res = PlanResult(**{k: (v.clone() if isinstance(v, torch.Tensor) and k in _COMPACT else (v if self.keep_tree else None))
                                for k, v in res.__dict__.items()})
            outs.append(res)
        self._events = ev
        if len(outs) == 1:
            return outs[0]
        return PlanResult(**{k: (torch.cat([getattr(o, k) for o in outs]) if k in _COMPACT else None)
                             for k in outs[0].__dict__})

    def last_logits(self) -> torch.Tensor:
        """Raw network output [B, cells] fp32 of the last pipeline batch of the last plan() call (a view of the reusable buffer)."""
        return self._logits

    def stage_ms(self):
        """(cnn, cdf, draw, plan) milliseconds of the last plan(record_events=True) call, after a synchronize."""
        c = d = w = p = 0.0
        for e in self._events:  # [start, after cnn, after cdf, end, between draw and plan]
            c += e[0].elapsed_time(e[1])
            d += e[1].elapsed_time(e[2])
            w += e[2].elapsed_time(e[4])
            p += e[4].elapsed_time(e[3])
        return c, d, w, p


_COMPACT = ("n_nodes", "iterations", "status", "goal_node", "best_node", "best_dist", "path_n", "path_idx", "path_len", "n_draws")
