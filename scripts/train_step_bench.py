"""Scratch / profiling driver: UNetTrainer steps at B x 512 x 512 on synthetic data (CUDA-event timing; not the bench contract).
usage: python scripts/train_step_bench.py [steps] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgr_sampling_cnn_b200 import _lib, train, trainer
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = train.UNet_cooler()
tr = trainer.UNetTrainer(model, B, 512, 512, dev)
rng = np.random.default_rng(0)
x = torch.from_numpy((rng.random((B, 1, 512, 512)) > 0.2).astype(np.float32)).repeat(1, 3, 1, 1).to(dev)
y = torch.from_numpy(rng.random((B, 1, 512, 512)).astype(np.float32)).to(dev)
coords = torch.from_numpy(rng.integers(0, 512, (B, 1, 2, 2)).astype(np.float32)).to(dev)
for it in range(steps):
    n0 = _lib.launches
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); loss = tr.training_step((x, y, coords)); b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print("step %d: %.2f ms, %.1f samples/s, %.1f TFLOP/s algorithmic, loss %.5f, %d launches" % (
        it, ms, B / ms * 1e3, tr.flops_per_step() / ms / 1e9, float(loss.item()), _lib.launches - n0))
