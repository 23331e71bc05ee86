"""Scratch timing of nrrt_cdf_build (CUDA events; not the bench contract): 2D 4096 x 512^2 logits, 3D 1024 x 80^3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgr_sampling_cnn_b200 import engine
dev = torch.device("cuda:0")
for name, B, cells, mode in (("2D", 4096, 512 * 512, engine.HEAT_MODE_2D), ("3D", 1024, 80 ** 3, engine.HEAT_MODE_3D)):
    logits = torch.randn(B, cells, device=dev)
    out = torch.empty_like(logits)
    engine.build_cdf(logits, mode, dev, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); engine.build_cdf(logits, mode, dev, out=out); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = min(ts)
    print(f"cdf {name} B={B}: {t:.3f} ms -> {B * cells * 8 / t / 1e6:.1f} GB/s algorithmic (8 B/cell)")
    ref = np.cumsum(((np.clip(logits[0].cpu().numpy(), -3, 1) + 10).astype(np.float32) if mode == 0 else None)) if False else None
