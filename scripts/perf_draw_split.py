"""Scratch: split of the draw / plan kernels inside GuidedPlanner.plan on the 2D neural bench workload (CUDA events
around the two C-ABI calls; not the bench contract).   python scripts/perf_draw_split.py [tasks]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mgr_sampling_cnn_b200 import _lib, engine, train, workload  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wl = workload.make_2d(B)
torch.manual_seed(0)
model = train.UNet_cooler().eval().to(dev)
gp = engine.GuidedPlanner(model, 2, wl.shape)
lib = _lib.load()
events = {"nrrt_draw_samples": [], "nrrt_plan_batch": []}
for name in events:
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
    res = gp.plan(wl.occ_maps, wl.task_to_map, wl.starts, wl.goals, wl.coords, seed=0)
    torch.cuda.synchronize()
for name, v in events.items():
    print(name, " ".join(f"{a.elapsed_time(b):.2f}" for a, b in v), "ms")
