"""Scratch timing of nrrt_grid_paths_2d on the 2D BASELINE workload (CUDA events; not the bench contract).
    python scripts/perf_gridpath.py [tasks] [dim]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mgr_sampling_cnn_b200 import engine, workload  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = workload.make_2d(B, filter_unreachable=False) if dim == 2 else workload.make_3d(B, filter_unreachable=False)
bits = engine.pack_occupancy(torch.from_numpy(wl.occ_maps).to(dev), dim, dev)
t2m, st, gl = (torch.from_numpy(x).to(dev) for x in (wl.task_to_map, wl.starts, wl.goals))
engine.grid_path_lengths(bits, t2m[:8], st[:8], gl[:8], wl.shape)
ts = []
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    length, steps = engine.grid_path_lengths(bits, t2m, st, gl, wl.shape)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
L = length.cpu().numpy()
ok = np.isfinite(L)
cells = float(steps.cpu().numpy()[ok].sum()) if ok.any() else 0.0
print(f"grid paths {dim}D B={B}: {min(ts):.1f} ms -> {B / min(ts) * 1e3:.0f} tasks/s; reachable {ok.mean():.3f}; mean length {L[ok].mean():.1f}; "
      f"mean steps {cells / max(ok.sum(), 1):.0f}")
