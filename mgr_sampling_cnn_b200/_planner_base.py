"""Shared implementation of the four drop-in planner modules (rrt_star, neural_rrt, ThreeD_rrt_star, ThreeD_neural_rrt).

Each module exposes the reference's names (`Node`, `RRTStar`, `MAX_ITERATIONS`, `GOAL_THRESHOLD`, the file-name parsers);
`RRTStar.rrt_star()` runs the batched CUDA engine with B = 1 and then materialises the reference's Python-side state
(`nodes`, `best_node`, `best_distance`, `iteration_no`, `search_radius`).  The individually callable methods run the
same CUDA primitives one call at a time (nrrt_collision_free / nrrt_tree_search / nrrt_draw_samples); only scalar
bookkeeping (steer, goal test, parent walk, radius lookup) is host Python.

Randomness: the reference draws from the module attribute `random` (CPython's module).  Here every instance consumes ONE
Philox stream (seed, task) through a draw counter, whichever method draws -- `rrt_star()`, `generate_random_sample()`,
`generate_neural_sample()` -- the way the reference's methods all consume the one module-level generator.  If the
module's `random` attribute carries `.seed` / `.task` / `.draws` (a StreamRandom, the injection point of SURVEY.md §8b)
that stream is used and its `.draws` counter is read and advanced, so reference and drop-in stay in lock-step call by
call; otherwise the instance takes a 64-bit seed from `random.getrandbits(64)` on first use (`random.seed(s)` keeps
runs reproducible) and counts its own draws.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib, engine


class Node(object):
    def __init__(self, position, cost: float = 0.0):
        self.position = position
        self.parent = None
        self.children = []
        self.cost = cost


def get_from_string(path: str, start: str, finish: str) -> str:
    return path[path.find(start) + 3:path.find(finish)]


class RRTStarBase(object):
    """dim, neural and the module (for its `random` attribute) are bound by the concrete subclasses."""
    _dim = 2
    _neural = False
    _module = None

    def _init_common(self, occ_map, heat_map, start, goal, max_iterations, goal_threshold, neural_bias):
        self.start_node = Node(start)
        self.goal = goal
        self.max_iterations = max_iterations
        self.iteration_no = None
        self.search_radius = None
        self.goal_threshold = goal_threshold
        self.nodes = [self.start_node]
        self.best_distance = float('inf')
        self.best_node = None
        self.neural_bias = neural_bias
        self.occ_map = occ_map
        self.heat_map = heat_map
        if not torch.cuda.is_available():
            raise _lib.NrrtError("RRTStar needs a CUDA device (sm_100a): this package has no CPU fallback")
        self._device = torch.device("cuda")
        occ = np.ascontiguousarray(self.occ_map)
        if occ.dtype not in (np.uint8, np.float32, np.float64):
            occ = occ.astype(np.float64)
        self._bits = engine.pack_occupancy(occ[None], self._dim, self._device)
        self._cdf = None
        if self._neural:
            # the planner-side heat map is consumed as float32, like the reference's callers provide it
            h = np.ascontiguousarray(np.asarray(self.heat_map, dtype=np.float32)).reshape(1, -1)
            self._cdf = engine.build_cdf(h, engine.HEAT_MODE_RAW, self._device)
        self._draw_k = 0
        self._own_seed = None
        self.status = None

    # ---- stream identity -------------------------------------------------------------------------------------------
    def _stream(self):
        """(seed, task, first draw) of the next draw of this instance."""
        rnd = self._module.random
        if hasattr(rnd, "seed") and hasattr(rnd, "task") and not callable(getattr(rnd, "seed")):
            return int(rnd.seed), int(rnd.task), int(getattr(rnd, "draws", 0))
        if self._own_seed is None:
            self._own_seed = int(rnd.getrandbits(64))
        return self._own_seed, 0, self._draw_k

    def _advance(self, draws: int):
        rnd = self._module.random
        if hasattr(rnd, "seed") and hasattr(rnd, "task") and hasattr(rnd, "draws") and not callable(getattr(rnd, "seed")):
            rnd.draws = int(draws)
            if hasattr(rnd, "_buf"):
                rnd._buf = None  # StreamRandom's look-ahead buffer is positioned by .draws
        else:
            self._draw_k = int(draws)

    @property
    def _shape(self):
        return tuple(int(s) for s in np.shape(self.occ_map))[:self._dim]

    # ---- the hot path ----------------------------------------------------------------------------------------------
    def rrt_star(self):
        for pt in (self.start_node.position, self.goal):
            if any(float(v) != int(v) for v in pt):
                raise _lib.NrrtError("start and goal must be integer-valued cells (as the reference's callers pass them)")
        seed, task, first = self._stream()
        res = engine.plan_batch(self._bits, np.zeros(1, np.int32), np.asarray(self.start_node.position, np.float64)[None],
                                np.asarray(self.goal, np.float64)[None], dim=self._dim, shape=self._shape,
                                max_iterations=self.max_iterations, goal_threshold=self.goal_threshold, cdf=self._cdf,
                                neural_bias=self.neural_bias if self._neural else 0.0, seed=seed,
                                task_ids=np.array([task], np.int32), keep_tree=True,
                                path_cap=self.max_iterations + 2, draw_start=first)
        torch.cuda.synchronize()
        # the stream stops where the reference's loop stopped drawing: after the sample of the last executed iteration
        iters = int(res.iterations[0])
        if 0 < iters < self.max_iterations and int(res.status[0]) != _lib.NRRT_SAMPLER_STUCK:
            self._advance(self._draw(iters, 0, seed, task, first)[1])
        else:
            self._advance(int(res.n_draws[0]))
        n = int(res.n_nodes[0])
        goal_node = int(res.goal_node[0])
        best = int(res.best_node[0])
        n_tot = n + (1 if goal_node >= 0 else 0)
        pos = res.pos[0, :n_tot].cpu().numpy()
        cost = res.cost[0, :n_tot].cpu().numpy()
        parent = res.parent[0, :n_tot].cpu().numpy()
        objs = [Node(tuple(pos[j].tolist()), float(cost[j])) for j in range(n_tot)]
        objs[0] = self.start_node
        self.start_node.cost = float(cost[0])
        for j in range(1, n_tot):
            objs[j].parent = objs[parent[j]]
            objs[parent[j]].children.append(objs[j])
        self.nodes = objs[:n]
        self.status = int(res.status[0])
        self.iteration_no = int(res.iterations[0])
        self.search_radius = self.compute_search_radius(self._dim) if self.iteration_no else None
        self.best_distance = float(res.best_dist[0])
        self.best_node = objs[best] if best >= 0 else None
        if goal_node >= 0:
            objs[goal_node].position = self.goal  # goal_node.position = self.goal
        if self.status == _lib.NRRT_SAMPLER_STUCK:
            raise _lib.NrrtError("sampler found no admissible cell after max_reject draws "
                                 "(the reference would loop forever here)")
        end = objs[goal_node] if goal_node >= 0 else self.best_node
        return self.find_path(end), self.iteration_no

    # ---- individually callable methods ----------------------------------------------------------------------------
    def find_nearest_neighbor(self, sample):
        pos = np.array([n.position for n in self.nodes], dtype=np.float64).reshape(len(self.nodes), self._dim)
        nodes = torch.from_numpy(pos).to(self._device)
        q = torch.tensor([list(map(float, sample))], dtype=torch.float64, device=self._device)
        out = torch.empty(1, dtype=torch.int32, device=self._device)
        _lib.check(_lib.load().nrrt_tree_search(self._dim, C.c_void_p(nodes.data_ptr()), len(self.nodes),
                                                C.c_void_p(q.data_ptr()), None, 1, C.c_void_p(out.data_ptr()), None, None))
        torch.cuda.synchronize()
        return self.nodes[int(out[0])]

    def find_nearby_nodes(self, node):
        n = len(self.nodes)
        pos = np.array([m.position for m in self.nodes], dtype=np.float64).reshape(n, self._dim)
        nodes = torch.from_numpy(pos).to(self._device)
        q = torch.tensor([list(map(float, node.position))], dtype=torch.float64, device=self._device)
        r = torch.tensor([float(self.search_radius)], dtype=torch.float64, device=self._device)
        nearest = torch.empty(1, dtype=torch.int32, device=self._device)
        mask = torch.zeros((n + 31) // 32, dtype=torch.int32, device=self._device)
        _lib.check(_lib.load().nrrt_tree_search(self._dim, C.c_void_p(nodes.data_ptr()), n, C.c_void_p(q.data_ptr()),
                                                C.c_void_p(r.data_ptr()), 1, C.c_void_p(nearest.data_ptr()),
                                                C.c_void_p(mask.data_ptr()), None))
        torch.cuda.synchronize()
        bits = mask.cpu().numpy().view(np.uint32)
        return [self.nodes[j] for j in range(n) if (bits[j >> 5] >> (j & 31)) & 1]

    def is_collision_free(self, point1, point2) -> bool:
        p1 = torch.tensor([list(map(float, point1))], dtype=torch.float64, device=self._device)
        p2 = torch.tensor([list(map(float, point2))], dtype=torch.float64, device=self._device)
        seg = torch.zeros(1, dtype=torch.int32, device=self._device)
        out = torch.empty(1, dtype=torch.int32, device=self._device)
        shape = (C.c_int32 * 3)(*(list(self._shape) + [1] * (3 - self._dim)))
        _lib.check(_lib.load().nrrt_collision_free(self._dim, shape, C.c_void_p(self._bits.data_ptr()),
                                                   C.c_void_p(seg.data_ptr()), C.c_void_p(p1.data_ptr()),
                                                   C.c_void_p(p2.data_ptr()), 1, C.c_void_p(out.data_ptr()), None))
        torch.cuda.synchronize()
        return bool(int(out[0]))

    def can_connect_nodes(self, from_node, to_node) -> bool:
        if self.is_collision_free(from_node.position, to_node.position):
            from_node.children.append(to_node)
            to_node.parent = from_node
            return True
        return False

    def steer(self, from_node, to_point):
        d = [to_point[k] - from_node.position[k] for k in range(self._dim)]
        dist = math.sqrt(sum(v ** 2 for v in d))
        if dist > self.search_radius:
            d = [v * self.search_radius / dist for v in d]
        dist = math.sqrt(sum(v ** 2 for v in d))
        node = Node(tuple(from_node.position[k] + d[k] for k in range(self._dim)), from_node.cost + dist)
        node.parent = from_node
        return node

    def goal_reached(self, node, goal) -> bool:
        dist = _euclid(node.position, goal)
        if dist < self.best_distance:
            self.best_distance = dist
            self.best_node = node
        return dist <= self.goal_threshold

    def find_path(self, goal):
        path = []
        cur = goal
        while cur is not None:
            path.append(cur.position)
            cur = cur.parent
        path.reverse()
        return path

    def lebesgue_measure(self, dim: int) -> float:
        return math.pow(math.pi, dim / 2.0) / math.gamma((dim / 2.0) + 1)

    def compute_search_radius(self, dim: int) -> float:
        return float(engine.radius_table(dim, self._shape, max(self.iteration_no, 1))[self.iteration_no - 1])

    def _draw(self, n, mode, seed, task, first):
        """n samples through nrrt_draw_samples from draw `first` of stream (seed, task); mode = NrrtPlanParams.sampler_mode.
        Returns (samples int16 [n, dim] on the host, index of the draw after the last one consumed)."""
        p = _lib.NrrtPlanParams()
        p.dim = self._dim
        shp = list(self._shape) + [1] * (3 - self._dim)
        for k in range(3):
            p.shape[k] = shp[k]
        p.max_iterations, p.node_stride, p.max_reject, p.path_cap = n, n + 2, 100000, 0
        p.goal_threshold, p.seed = float(self.goal_threshold), seed & (2 ** 64 - 1)
        p.neural_bias = float(self.neural_bias) if self._neural else 0.0
        p.draw_start, p.sampler_mode = int(first), int(mode)
        dev = self._device
        t2m = torch.zeros(1, dtype=torch.int32, device=dev)
        ids = torch.tensor([task], dtype=torch.int32, device=dev)
        smp = torch.zeros((1, n, self._dim), dtype=torch.int16, device=dev)
        st = torch.zeros(1, dtype=torch.int32, device=dev)
        nd = torch.zeros(1, dtype=torch.int64, device=dev)
        cdf = self._cdf if self._cdf is not None else None
        _lib.count_launch()
        _lib.check(_lib.load().nrrt_draw_samples(C.byref(p), C.c_void_p(self._bits.data_ptr()), C.c_void_p(t2m.data_ptr()),
                                                 None if cdf is None else C.c_void_p(cdf.data_ptr()), C.c_void_p(ids.data_ptr()),
                                                 1, C.c_void_p(smp.data_ptr()), None, C.c_void_p(nd.data_ptr()),
                                                 C.c_void_p(st.data_ptr()), None))
        torch.cuda.synchronize()
        if int(st[0]) != 0:
            raise _lib.NrrtError("sampler found no admissible cell (the reference would loop forever)")
        return smp[0].cpu().numpy(), int(nd[0])

    def _sample_via_kernel(self, force_neural):
        """generate_*_sample called directly: one sample from the instance's stream, continuing its draw counter and
        consuming exactly the draws the reference's method consumes (no `random() < neural_bias` draw)."""
        seed, task, first = self._stream()
        smp, nxt = self._draw(1, 1 if force_neural else 2, seed, task, first)
        self._advance(nxt)
        return tuple(int(v) for v in smp[0].tolist())

    def generate_random_sample(self):
        return self._sample_via_kernel(False)

    def rewire_tree(self, new_node):
        """rrt_star.py:125-147 on host objects (the batched kernel fuses this into the main loop; this is the callable
        method): near set and collision tests through the CUDA primitives, both loops in the reference's order."""
        nearby_nodes = self.find_nearby_nodes(new_node)
        for nearby_node in nearby_nodes:
            new_cost = nearby_node.cost + _euclid(nearby_node.position, new_node.position)
            if new_cost < new_node.cost:
                if self.is_collision_free(nearby_node.position, new_node.position):
                    new_node.parent.children.remove(new_node)
                    new_node.parent = nearby_node
                    nearby_node.children.append(new_node)
                    new_node.cost = new_cost
        for node in nearby_nodes:
            redone_cost = new_node.cost + _euclid(new_node.position, node.position)
            if redone_cost < node.cost:
                if self.is_collision_free(new_node.position, node.position):
                    node.parent.children.remove(node)
                    node.parent = new_node
                    new_node.children.append(node)
                    node.cost = redone_cost

    # plotting is out of scope: the reference's matplotlib methods exist as no-ops so that its demo code runs
    def visualize_tree(self, *a, **k):
        return None

    def visualize_path(self, *a, **k):
        return None


def _euclid(u, v) -> float:
    """scipy.spatial.distance.euclidean: sqrt of the left-to-right sum of squared differences."""
    s = 0.0
    for a, b in zip(u, v):
        d = float(a) - float(b)
        s = s + d * d
    return math.sqrt(s)


def calculate_length(path) -> float:
    """test_network_and_algorithms.py:52-56."""
    total = 0.0
    for i in range(len(path) - 1):
        total += _euclid(path[i], path[i + 1])
    return total
