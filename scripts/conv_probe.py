"""Scratch: one encoder pass (8 maps) + decoder micro-batches (40 tasks = 4 maps, the bench's micro-batch) of the 2D CNN at 512x512, for ncu captures
and CUDA-event timing of individual layers.  Not the bench contract.

    python scripts/conv_probe.py [n_decode]     prints the per-layer event timings of the last decode
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mgr_sampling_cnn_b200 import cnn, train, workload  # noqa: E402

dev = torch.device("cuda:0")
n_dec = int(sys.argv[1]) if len(sys.argv) > 1 else 2
variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["base"]
wl = workload.make_2d(80)
MB = 40  # tasks per decoder launch, as GuidedPlanner runs them (engine.py)
torch.manual_seed(0)
model = train.UNet_cooler().eval().to(dev)
occ = torch.from_numpy(wl.occ_maps).to(dev)
coords = torch.from_numpy(wl.coords).to(dev)
grid = (1, 512, 512)
maps = torch.arange(8, dtype=torch.int32, device=dev)
ref_logits = None
for variant in variants:
    cnn.UNetPlan.res_mma = "res_mma" in variant
    model.invalidate_plan()
    plan = model._plan(dev)
    prof = cnn.ConvProfiler(sample_every=1)
    prep = plan.coords_prepare(coords, grid)
    prof.active = True
    enc = plan.encode(grid, 8, occ=occ, maps=maps)
    prof.active = False
    torch.cuda.synchronize()
    prof.collect()
    enc_bytes, enc_launches = prof.bytes_seen, prof.seen
    prof.reset()
    logits = torch.empty((MB, 512 * 512), dtype=torch.float32, device=dev)
    for it in range(n_dec):
        prof.active = it >= max(n_dec - 4, 1)
        m = (it % 2) * MB
        rel = (torch.arange(m, m + MB, device=dev) // 10).to(torch.int32)
        plan.decode(enc, MB, grid, coords[m:m + MB], plan.slice_prep(prep, m, m + MB), rel, logits, res_group=10)
    prof.active = False
    torch.cuda.synchronize()
    prof.collect()
    print(f"== variant {variant}")
    if ref_logits is None:
        ref_logits = logits.clone()
    else:
        print(f"max |logits - base| = {(logits - ref_logits).abs().max().item():.3e}")
    for name, (n, fl, ms) in sorted(prof.by_layer().items(), key=lambda kv: -kv[1][2]):
        print(f"{name:10s} {ms / n:8.4f} ms  {fl / max(ms, 1e-9) / 1e9:8.1f} TFLOP/s")
    n_prof = max(n_dec - max(n_dec - 4, 1), 1)  # decode passes recorded
    dec_bytes, dec_launches = prof.bytes_seen / n_prof, prof.seen // n_prof
    print(f"algorithmic DRAM bytes of one encoder pass (8 maps) + one decoder micro-batch ({MB} tasks): "
          f"{(enc_bytes + dec_bytes) / (enc_launches + dec_launches):.0f} per launch over {enc_launches + dec_launches} launches")
    prof.reset()
cnn.UNetPlan.res_mma = False

if plan.use_fold(grid):  # the decomposition below times the unfolded conv7.0 / conv8.0
    sys.exit(0)
# epilogue cost decomposition of conv8.0 / conv7.0 (timing only: the variants without poly / residual are not the layer)
import ctypes as C  # noqa: E402
lib = plan.lib
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
bf = dict(plan._dec_buffers(10, grid))
bf.update({k: enc[k] for k in ("r7", "r8")})
g = plan._levels(grid)
pr = plan.slice_prep(prep, 0, 10)
rel = torch.zeros(10, dtype=torch.int32, device=dev)
plan.set_profiler(None)
for name, layer, x, lv, out, res, poly in (("c8.0", plan.c8[0], bf["t8"], 0, bf["j"], bf["r8"], pr["e_c8"]),
                                           ("c7.0", plan.c7[0], bf["t7"], 1, bf["g7"], bf["r7"], pr["e_c7"])):
    for tag, kw in (("poly+res", dict(res=res, poly=poly, res_idx=rel)), ("res", dict(res=res, res_idx=rel)),
                    ("poly", dict(poly=poly)), ("plain", dict())):
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            layer.run(lib, st, x, 0, g[lv], out, 0, kw.get("res"), 0, poly=kw.get("poly"), res_idx=kw.get("res_idx"))
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(f"{name} {tag:9s} {min(ts):.4f} ms")
