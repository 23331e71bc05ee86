"""Scratch: where the CNN time of one 4096-task batch goes besides the tensor-core layers (CUDA events around
coords_prepare, the stem kernel, poly_relu; not the bench contract).   python scripts/perf_cnn_split.py [tasks] [dim]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mgr_sampling_cnn_b200 import _lib, engine, train, workload  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = workload.make_2d(B) if dim == 2 else workload.make_3d(B)
torch.manual_seed(0)
if dim == 2:
    model = train.UNet_cooler().eval().to(dev)
else:
    from mgr_sampling_cnn_b200 import ThreeD_train
    model = ThreeD_train.ThreeD_UNet_cooler().eval().to(dev)
gp = engine.GuidedPlanner(model, dim, wl.shape)
lib = _lib.load()
names = ["nrrt_conv_layer", "nrrt_stem_conv_maps", "nrrt_coords_branch", "nrrt_coords_poly_gemm", "nrrt_poly16", "nrrt_poly_fold",
         "nrrt_poly_relu", "nrrt_poly_relu_3d", "nrrt_coords_fill", "nrrt_coords_pieces_3d", "nrrt_cdf_build", "nrrt_draw_samples", "nrrt_plan_batch", "nrrt_pack_occupancy"]
events = {n: [] for n in names}
for name in names:
    fn = getattr(lib, name)

    def wrapped(*a, _fn=fn, _name=name):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = _fn(*a)
        e1.record()
        events[_name].append((e0, e1))
        return rc
    setattr(lib, name, wrapped)
for it in range(2):
    for v in events.values():
        v.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    res = gp.plan(wl.occ_maps, wl.task_to_map, wl.starts, wl.goals, wl.coords, seed=0)
    t1.record()
    torch.cuda.synchronize()
tot = 0.0
for name, v in events.items():
    ms = sum(a.elapsed_time(b) for a, b in v)
    tot += ms
    print(f"{name:24s} {len(v):6d} calls {ms:9.2f} ms")
print(f"sum {tot:.1f} ms of {t0.elapsed_time(t1):.1f} ms (events between every call serialise nothing; gaps = host / torch ops)")
if res.n_draws is not None:
    nd = res.n_draws.double().cpu().numpy()
    import numpy as np  # noqa: E402
    print(f"draws per task: mean {nd.mean():.0f}, median {np.median(nd):.0f}, max {nd.max():.0f}; per iteration {nd.sum() / res.iterations.double().sum().item():.2f}"
          f" (draw kernel always draws max_iterations samples: {nd.mean() / 5000:.2f} per sample)")
