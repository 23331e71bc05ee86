"""Scratch timing of the mask makers (CUDA events; not the bench contract): 2D 4096 tasks at 512^2, 3D 1024 tasks at 80^3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgr_sampling_cnn_b200 import generate_paths as gp, ThreeD_enlarge_path as ep
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


B, H = 4096, 512
occ = torch.full((410, H, H), 255, dtype=torch.uint8, device=dev)
occ[:, 100:200, 150:300] = 0
n = rng.integers(200, 800, B).astype(np.int32)
paths = np.zeros((B, 800, 2), np.int32)
for b in range(B):
    s, g = rng.integers(0, H, 2), rng.integers(0, H, 2)
    t = np.linspace(0, 1, n[b])[:, None]
    paths[b, :n[b]] = np.rint(s + (g - s) * t)
t2m = torch.arange(B, device=dev, dtype=torch.int32) // 10
pts, pn = torch.from_numpy(paths).to(dev), torch.from_numpy(n).to(dev)
t = timed(lambda: gp.path_masks(occ, t2m, pts, pn))
print(f"masks 2D B={B}: {t:.3f} ms -> {B * H * H * 2 / t / 1e6:.1f} GB/s algorithmic (2 B/cell), {B / t * 1e3:.0f} masks/s")

B, S = 1024, 80
occ3 = (torch.rand((103, S, S, S), device=dev) < 0.1).to(torch.uint8) * 255
vols = torch.zeros((B, S, S, S), dtype=torch.uint8, device=dev)
for b in range(0, B):
    s, g = rng.integers(0, S, 3), rng.integers(0, S, 3)
    p = np.rint(s + (g - s) * np.linspace(0, 1, 100)[:, None]).astype(np.int64)
    vols[b][p[:, 0], p[:, 1], p[:, 2]] = 255
t2m3 = torch.arange(B, device=dev, dtype=torch.int32) // 10
t = timed(lambda: ep.enlarge_paths(occ3, t2m3, vols))
print(f"enlarge 3D B={B}: {t:.3f} ms -> {B * S ** 3 * 3 / t / 1e6:.1f} GB/s algorithmic (3 B/cell), {B / t * 1e3:.0f} masks/s")
