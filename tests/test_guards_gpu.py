"""GPU: bounds and race checks of our own (compute-sanitizer is closed on this GPU pool, profiles/sanitizer_r2.md).
 * guard bands: every caller-owned buffer a kernel family writes (planner outputs, sample buffers, the forward's workspace and
   logits) is carved out of an arena filled with a byte pattern, with a band before and after it; after the kernels have run the
   bands must be untouched and the results equal to a run on ordinary buffers;
 * repeatability: kernels that hand data between warps without a block barrier (plan_kernel's register hand-off of the newest
   node, the conv kernels' mbarrier pipelines across a CTA pair) give bit-identical results run after run -- a missing
   barrier / phase bug shows up as run-to-run differences."""
import ctypes as C

import numpy as np
import pytest
import torch

from mgr_sampling_cnn_b200 import ThreeD_train, _lib, engine, train, workload

pytestmark = pytest.mark.gpu
GUARD = 4096
PATTERN = 0xA5


class Arena(object):
    def __init__(self, nbytes, dev):
        self.buf = torch.full((nbytes,), PATTERN, dtype=torch.uint8, device=dev)
        self.off = 0
        self.bands = []

    def take(self, like: torch.Tensor) -> torch.Tensor:
        n = like.numel() * like.element_size()
        start = (self.off + GUARD + 255) & ~255
        self.bands.append((self.off, start))
        view = self.buf[start:start + n].view(like.dtype).view(like.shape)
        self.off = start + n
        return view

    def check(self):
        self.bands.append((self.off, min(self.off + GUARD, self.buf.numel())))
        for a, b in self.bands:
            assert bool((self.buf[a:b] == PATTERN).all()), f"guard band [{a}, {b}) was written"


@pytest.mark.parametrize("dim", [2, 3])
def test_planner_and_sampler_stay_inside_their_buffers(dim, cuda_device):
    wl = workload.make_2d(12, size=128) if dim == 2 else workload.make_3d(12, size=32)
    bits = engine.pack_occupancy(wl.occ_maps, dim, cuda_device)
    heat = np.random.default_rng(1).normal(0, 1, size=(12,) + wl.shape).astype(np.float32)
    cdf = engine.build_cdf(heat, engine.HEAT_MODE_2D if dim == 2 else engine.HEAT_MODE_3D, cuda_device)
    kw = dict(dim=dim, shape=wl.shape, max_iterations=600, cdf=cdf, seed=5, trace_cap=3000, path_cap=64)
    ref = engine.plan_batch(bits, wl.task_to_map, wl.starts, wl.goals, **kw)
    bufs = engine.PlannerBuffers(dim, 12, 600, 64, True, 3000, cuda_device)
    names = [k for k, v in vars(bufs).items() if isinstance(v, torch.Tensor)]
    arena = Arena(sum(getattr(bufs, k).numel() * getattr(bufs, k).element_size() + 2 * GUARD + 512 for k in names) + GUARD, cuda_device)
    for k in names:
        setattr(bufs, k, arena.take(getattr(bufs, k)))
    res = engine.plan_batch(bits, wl.task_to_map, wl.starts, wl.goals, buffers=bufs, **kw)
    torch.cuda.synchronize()
    arena.check()
    for k in ("n_nodes", "iterations", "status", "goal_node", "best_node", "path_n", "trace_n", "n_draws"):
        assert torch.equal(getattr(res, k), getattr(ref, k)), k
    drawn = ref.status != _lib.NRRT_SAMPLER_STUCK  # a stuck sampler stops writing sample points where it gives up
    assert torch.equal(res.samples[drawn], ref.samples[drawn]) and torch.equal(res.neural_flag[drawn], ref.neural_flag[drawn])
    assert torch.equal(res.path_len, ref.path_len)
    for b in range(12):  # rows beyond a task's own nodes / trace records are never written: compare what is defined
        if not bool(drawn[b]):
            continue  # the planner skips a task whose sampler gave up: no tree rows are written
        n = int(ref.n_nodes[b]) + (1 if int(ref.goal_node[b]) >= 0 else 0)
        assert torch.equal(res.parent[b, :n], ref.parent[b, :n]) and torch.equal(res.cost[b, :n], ref.cost[b, :n])
        assert torch.equal(res.pos[b, :n], ref.pos[b, :n])
        t = min(int(ref.trace_n[b]), 3000)
        assert torch.equal(res.trace[b, :t], ref.trace[b, :t])


@pytest.mark.parametrize("dim", [2, 3])
def test_forward_stays_inside_workspace_and_logits(dim, cuda_device):
    wl = workload.make_2d(13, size=128) if dim == 2 else workload.make_3d(7, size=32)
    n = len(wl.task_to_map)
    torch.manual_seed(0)
    model = (train.UNet_cooler() if dim == 2 else ThreeD_train.ThreeD_UNet_cooler()).eval().to(cuda_device)
    plan = model._plan(cuda_device)
    occ = torch.from_numpy(wl.occ_maps).to(cuda_device)
    coords = torch.from_numpy(wl.coords).to(cuda_device)
    cells = int(np.prod(wl.shape))
    ref = plan.forward_tasks(occ, np.unique(wl.task_to_map).astype(np.int32), wl.task_to_map, coords,
                             torch.empty((n, cells), dtype=torch.float32, device=cuda_device)).clone()
    W = plan.c_weights()
    lib = _lib.load()
    d, h, w = (1,) + wl.shape if dim == 2 else (wl.shape[2], wl.shape[0], wl.shape[1])
    need = lib.nrrt_conv_forward_workspace_bytes(C.byref(W), d, h, w, n)
    arena = Arena(need + 4 * n * cells + 4 * GUARD + 2048, cuda_device)
    ws = arena.take(torch.empty(need, dtype=torch.uint8, device="meta"))
    out = arena.take(torch.empty((n, cells), dtype=torch.float32, device="meta"))
    ids, rows = np.unique(wl.task_to_map).astype(np.int32), np.ascontiguousarray(wl.task_to_map, dtype=np.int32)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    if dim == 2:
        rc = lib.nrrt_conv_forward_2d(C.byref(W), p(occ), ids.ctypes.data_as(C.c_void_p), len(ids), rows.ctypes.data_as(C.c_void_p),
                                      p(coords), n, h, w, 10, p(out), p(ws), need, None)
    else:
        rc = lib.nrrt_conv_forward_3d(C.byref(W), p(occ), ids.ctypes.data_as(C.c_void_p), len(ids), rows.ctypes.data_as(C.c_void_p),
                                      p(coords), n, d, h, w, 10, p(out), p(ws), need, None)
    assert rc == 0, lib.nrrt_last_error()
    torch.cuda.synchronize()
    arena.check()
    assert torch.equal(out, ref)


@pytest.mark.parametrize("dim", [2, 3])
def test_results_are_repeatable_run_after_run(dim, cuda_device):
    wl = workload.make_2d(13, size=128) if dim == 2 else workload.make_3d(20, size=32)
    torch.manual_seed(0)
    model = (train.UNet_cooler() if dim == 2 else ThreeD_train.ThreeD_UNet_cooler()).eval().to(cuda_device)
    gp = engine.GuidedPlanner(model, dim, wl.shape, max_iterations=1500, keep_tree=True)
    first = None
    for _ in range(6):
        res = gp.plan(wl.occ_maps, wl.task_to_map, wl.starts, wl.goals, wl.coords, seed=9)
        torch.cuda.synchronize()
        live = torch.arange(res.parent.shape[1], device=cuda_device)[None, :] < res.n_nodes[:, None]   # defined tree rows
        cur = (gp.last_logits().clone(), res.samples.clone(), res.iterations.clone(), torch.where(live, res.parent, -7),
               torch.where(live, res.cost, 0.0), res.path_len.clone())
        if first is None:
            first = cur
        else:
            for a, b in zip(first, cur):
                assert torch.equal(a, b)
