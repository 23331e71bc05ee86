"""Scratch timing of UNetPlan.coords_prepare (coordinate branch + epilogue polynomials) for a full batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgr_sampling_cnn_b200 import train, workload, _lib
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wl = workload.make_2d(min(B, 200))
coords = torch.from_numpy(np.resize(wl.coords, (B,) + wl.coords.shape[1:])).to(dev)
torch.manual_seed(0)
plan = train.UNet_cooler().eval().to(dev)._plan(dev)
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); prep = plan.coords_prepare(coords, (1, 512, 512)); b.record(); torch.cuda.synchronize()
    print(f"coords_prepare B={B}: {a.elapsed_time(b):.2f} ms")

# component timing through a monkey-patched library handle
import ctypes as C
lib = plan.lib
times = {}
class Timed(object):
    def __init__(self, lib): self._lib = lib
    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        def call(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = fn(*a); e1.record(); torch.cuda.synchronize()
            times.setdefault(name, []).append(e0.elapsed_time(e1)); return r
        return call
plan.lib = Timed(lib)
plan.coords_prepare(coords, (1, 512, 512))
for k, v in times.items():
    print(k, [round(x, 2) for x in v], "sum", round(sum(v), 2))
