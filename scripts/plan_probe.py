"""Scratch: one 2D uniform plan_batch call (ncu target for plan_kernel).  python scripts/plan_probe.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgr_sampling_cnn_b200 import engine, workload
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
wl = workload.make_2d(B)
occ = torch.from_numpy(wl.occ_maps).to(dev)
bits = engine.pack_occupancy(occ, 2, dev)
t2m, st, gl = (torch.from_numpy(x).to(dev) for x in (wl.task_to_map, wl.starts, wl.goals))
buf = engine.PlannerBuffers(2, B, 5000, 512, False, 0, dev)
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); res = engine.plan_batch(bits, t2m, st, gl, dim=2, shape=(512, 512), seed=1, buffers=buf); b.record()
    torch.cuda.synchronize()
    it = res.iterations.cpu().numpy()
    print(f"B={B}: {a.elapsed_time(b):.1f} ms, iterations total {it.sum()}, mean {it.mean():.0f}")
