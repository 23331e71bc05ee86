"""compute-sanitizer target: one small invocation of every kernel family of libnrrt.so (planner incl. TRACE, draw, CDF,
pack, grid paths, tree search / collision primitives, and the CNN forward: stem, conv pair kernels with staged residual /
TMA store / polynomial MMA, folded layers, coordinate branch).  Small sizes: the sanitizer serialises everything.

    compute-sanitizer --tool memcheck  python scripts/sanitize_probe.py [planner|cnn|all]
    compute-sanitizer --tool racecheck python scripts/sanitize_probe.py ...
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mgr_sampling_cnn_b200 import ThreeD_train, engine, train, workload  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0")
if what in ("planner", "all"):
    for dim in (2, 3):
        wl = workload.make_2d(6, size=128) if dim == 2 else workload.make_3d(6, size=32)
        bits = engine.pack_occupancy(wl.occ_maps, dim, dev)
        heat = np.random.default_rng(0).normal(0, 1, size=(6,) + wl.shape).astype(np.float32)
        cdf = engine.build_cdf(heat, engine.HEAT_MODE_2D if dim == 2 else engine.HEAT_MODE_3D, dev)
        for c, tc in ((None, 0), (cdf, 0), (None, 4096)):
            res = engine.plan_batch(bits, wl.task_to_map, wl.starts, wl.goals, dim=dim, shape=wl.shape, max_iterations=300, cdf=c,
                                    seed=1, keep_tree=True, trace_cap=tc)
        engine.grid_path_lengths(bits, wl.task_to_map, wl.starts, wl.goals, wl.shape)
        torch.cuda.synchronize()
        print("planner", dim, "ok: statuses", res.status.cpu().tolist())
if what in ("cnn", "all"):
    for dim in (2, 3):
        wl = workload.make_2d(4, size=128) if dim == 2 else workload.make_3d(3, size=32)
        torch.manual_seed(0)
        model = (train.UNet_cooler() if dim == 2 else ThreeD_train.ThreeD_UNet_cooler()).eval().to(dev)
        gp = engine.GuidedPlanner(model, dim, wl.shape, max_iterations=200)
        res = gp.plan(wl.occ_maps, wl.task_to_map, wl.starts, wl.goals, wl.coords, seed=2)
        torch.cuda.synchronize()
        print("cnn", dim, "ok: logits", float(gp.last_logits().abs().max()), "statuses", res.status.cpu().tolist())
