"""Scratch timing of the planner-side kernels on the BASELINE workloads (CUDA events; not the bench contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgr_sampling_cnn_b200 import engine, workload

def ev_time(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), r

dev = torch.device("cuda:0")
B2 = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
t0 = time.time(); wl = workload.make_2d(B2); print("gen 2D", time.time() - t0, "redrawn", wl.redrawn)
occ = torch.from_numpy(wl.occ_maps).to(dev)
t, bits = ev_time(lambda: engine.pack_occupancy(occ, 2, dev)); print("pack 2D ms", t)
t2m, st, gl = (torch.from_numpy(x).to(dev) for x in (wl.task_to_map, wl.starts, wl.goals))
buf = engine.PlannerBuffers(2, B2, 5000, 512, False, 0, dev)
t, res = ev_time(lambda: engine.plan_batch(bits, t2m, st, gl, dim=2, shape=(512, 512), seed=1, buffers=buf))
it = res.iterations.cpu().numpy(); stt = res.status.cpu().numpy()
print(f"2D uniform B={B2}: {t:.1f} ms -> {B2 / t * 1e3:.0f} plans/s; solved {np.mean(stt == 0):.3f}; iters mean {it.mean():.0f} median {np.median(it):.0f}; nodes mean {res.n_nodes.float().mean().item():.0f}")
# neural with random logits
Bn = min(B2, 1024)
logits = torch.randn(Bn, 512 * 512, device=dev)
t, cdf = ev_time(lambda: engine.build_cdf(logits, engine.HEAT_MODE_2D, dev)); print(f"cdf 2D B={Bn} ms {t:.2f} -> {Bn*262144*8/t/1e6:.1f} GB/s")
bufn = engine.PlannerBuffers(2, Bn, 5000, 512, False, 0, dev)
t, res = ev_time(lambda: engine.plan_batch(bits, t2m[:Bn], st[:Bn], gl[:Bn], dim=2, shape=(512, 512), seed=1, cdf=cdf, buffers=bufn))
it = res.iterations.cpu().numpy(); stt = res.status.cpu().numpy()
print(f"2D neural(random heat) B={Bn}: {t:.1f} ms -> {Bn / t * 1e3:.0f} plans/s; solved {np.mean(stt == 0):.3f}; iters mean {it.mean():.0f}")
# 3D
B3 = max(B2 // 4, 8)
t0 = time.time(); wl3 = workload.make_3d(B3); print("gen 3D", time.time() - t0, "redrawn", wl3.redrawn)
occ3 = torch.from_numpy(wl3.occ_maps).to(dev)
t, bits3 = ev_time(lambda: engine.pack_occupancy(occ3, 3, dev)); print("pack 3D ms", t)
t2m3, st3, gl3 = (torch.from_numpy(x).to(dev) for x in (wl3.task_to_map, wl3.starts, wl3.goals))
buf3 = engine.PlannerBuffers(3, B3, 5000, 512, False, 0, dev)
t, res = ev_time(lambda: engine.plan_batch(bits3, t2m3, st3, gl3, dim=3, shape=(80, 80, 80), seed=1, buffers=buf3))
it = res.iterations.cpu().numpy(); stt = res.status.cpu().numpy()
print(f"3D uniform B={B3}: {t:.1f} ms -> {B3 / t * 1e3:.0f} plans/s; solved {np.mean(stt == 0):.3f}; iters mean {it.mean():.0f} median {np.median(it):.0f}")
