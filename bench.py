#!/usr/bin/env python
"""bench.py — plans solved per second of the neural-guided RRT* hot path (BASELINE.json metric), B200.

    python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N ...            the reference's CPU implementation of the same path

A step = one pass of the hot path over one batch of synthetic tasks: CNN probability maps (tcgen05 implicit-GEMM convs)
-> sampler CDFs -> sample draws -> guided RRT* for every task.  Workload at N = 1: BASELINE.json configs[1], "2D batch of
4096 tasks" (410 maps x 10 tasks from this repo's generate_maps / generate_tasks drop-ins, random-init 2D CNN seed 0,
neural_bias 0.75, MAX_ITERATIONS 5000, GOAL_THRESHOLD 5.0).  Weak scaling: every rank runs its own 4096 tasks.
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "neural-RRT* plans solved/sec"
UNIT = "plans/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--workload", default="2d_neural", choices=["2d_neural", "2d_uniform", "3d_neural", "3d_uniform"])
    ap.add_argument("--tasks", type=int, default=None, help="tasks per GPU (default: 4096 for 2D, 1024 for 3D)")
    ap.add_argument("--ref-sample", type=int, default=None, help="tasks per step of the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-every", type=int, default=16, help="time the conv launches of every n-th micro-batch")
    ap.add_argument("--micro-batch", type=int, default=None, help="tasks per decoder launch (default: GuidedPlanner's)")
    ap.add_argument("--encoder-batch", type=int, default=None, help="distinct maps per encoder pass (default: GuidedPlanner's)")
    ap.add_argument("--layer-table", default=None, help="write the per-layer CUDA-event timings of the sampled launches (csv)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tensor": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops"))), "hbm": float(p["hbm_gbs"]),
                "which": "measured (MEASURED_PEAKS.json, sustained bf16 cuBLAS)"}
    return {"tensor": 1400.0, "hbm": 6650.0, "which": "fallback (B200_PROFILING.md)"}


def conv_traffic():
    """DRAM bytes per conv launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu capture of one
    encoder pass + one decoder micro-batch (profiles/conv_traffic_r1.json), averaged over launches like `achieved`."""
    path = os.path.join(ROOT, "profiles", "conv_traffic_r1.json")
    if not os.path.exists(path):
        return None, "no ncu capture committed"
    with open(path) as f:
        t = json.load(f)
    note = t["note"]
    if "algorithmic_bytes_per_launch" in t:
        note += "; algorithmic bytes of the same launches: %.0f per launch (%s)" % (t["algorithmic_bytes_per_launch"], t["algorithmic_note"])
    return t["avg_dram_bytes_per_launch"], note


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])), smax.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}
        os.unlink(self.f.name)
        return out


def make_workload(name, tasks, rank):
    from mgr_sampling_cnn_b200 import workload
    per = 10
    first_map = rank * ((tasks + per - 1) // per)
    if name.startswith("2d"):
        return workload.make_2d(tasks, first_map=first_map)
    return workload.make_3d(tasks, first_map=first_map)


def make_model(dim, device=None):
    from mgr_sampling_cnn_b200 import ThreeD_train, train
    torch.manual_seed(0)
    model = (train.UNet_cooler() if dim == 2 else ThreeD_train.ThreeD_UNet_cooler()).eval()
    return model.to(device) if device is not None else model


def cpu_arm(args, wl, n_sample, model=None):
    """The reference's path on host cores (oracle port): returns (plans/s, dict describing the run)."""
    from oracle import cpu_path
    dim = wl.dim
    uniform = args.workload.endswith("uniform")
    sd = None
    if not uniform:
        model = model or make_model(dim)
        sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    idx = np.arange(n_sample)
    maps_used, t2m = np.unique(wl.task_to_map[idx], return_inverse=True)
    threads = os.cpu_count() or 1
    secs, s_cnn, s_plan, res = cpu_path.run(sd, dim, wl.shape, wl.occ_maps[maps_used], t2m.astype(np.int32), wl.starts[idx],
                                            wl.goals[idx], wl.coords[idx], 0, idx.astype(np.int32), threads=threads,
                                            uniform_only=uniform)
    info = {"value": n_sample / secs, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {n_sample} tasks of the same workload, same Philox streams: torch-CPU fp32 forward "
                      f"({s_cnn:.1f} s) + C port of sampler/planner over {threads} threads ({s_plan:.1f} s)",
            "solved_fraction": float(np.mean(res["status"] == 0)), "mean_iterations": float(np.mean(res["iterations"]))}
    return n_sample / secs, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dim = 2 if args.workload.startswith("2d") else 3
    tasks = args.tasks or (4096 if dim == 2 else 1024)
    n_sample = args.ref_sample or (16 if not args.workload.endswith("uniform") else 32)
    wl = make_workload(args.workload, max(n_sample, 10), 0)
    model = None if args.workload.endswith("uniform") else make_model(dim)
    for _ in range(args.warmup):
        cpu_arm(args, wl, n_sample, model)
    t0 = time.perf_counter()
    info = None
    for _ in range(args.steps):
        _, info = cpu_arm(args, wl, n_sample, model)
    secs = time.perf_counter() - t0
    value = n_sample * args.steps / secs
    info["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 planner / f32 CNN", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, tasks), "tasks_per_step": n_sample},
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(w, tasks):
    return {"2d_neural": f"2D batch of {tasks} tasks: CNN probability maps + guided RRT* (BASELINE configs[1])",
            "2d_uniform": f"2D uniform-sampling RRT* batch of {tasks} tasks (BASELINE configs[2])",
            "3d_neural": f"3D voxel batch of {tasks} tasks: 3D CNN + guided RRT* (BASELINE configs[3])",
            "3d_uniform": f"3D uniform-sampling RRT* batch of {tasks} tasks"}[w]


def run_graft(args):
    import torch.distributed as dist
    from mgr_sampling_cnn_b200 import _lib, cnn, engine

    # rank 0 prints ONE JSON line: libraries that write to stdout (NCCL's version banner) go to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dim = 2 if args.workload.startswith("2d") else 3
    uniform = args.workload.endswith("uniform")
    tasks = args.tasks or (4096 if dim == 2 else 1024)
    wl = make_workload(args.workload, tasks, rank)
    B = len(wl.task_to_map)

    # host (pinned) and device copies of the inputs
    host = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in
            (("occ", wl.occ_maps), ("t2m", wl.task_to_map), ("starts", wl.starts), ("goals", wl.goals), ("coords", wl.coords))}
    d = {k: v.to(dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    model = gp = prof = None
    if not uniform:
        model = make_model(dim, dev)
        gp = engine.GuidedPlanner(model, dim, wl.shape, micro_batch=args.micro_batch)
        if args.encoder_batch:
            gp.encoder_batch = args.encoder_batch
        prof = cnn.ConvProfiler()
        model._plan(dev).set_profiler(prof)
    bits = engine.pack_occupancy(d["occ"], dim, dev)
    bufs = engine.PlannerBuffers(dim, B, 5000, 512, False, 0, dev) if uniform else None
    result_keys = ("status", "iterations", "n_nodes", "path_n", "path_len")
    gather_buf = None

    def step(inp, from_host, profile):
        if from_host:
            inp = {k: v.to(dev, non_blocking=True) for k, v in inp.items()}
        if uniform:
            b = bits if not from_host else engine.pack_occupancy(inp["occ"], dim, dev)
            res = engine.plan_batch(b, inp["t2m"], inp["starts"], inp["goals"], dim=dim, shape=wl.shape, seed=0, buffers=bufs)
        else:
            gp.profile_every = args.profile_every if profile else 0
            res = gp.plan(inp["occ"], inp["t2m"], inp["starts"], inp["goals"], inp["coords"], seed=0, record_events=profile)
        # fixed-size result records; NCCL gather to rank 0 is the only collective of the path
        rec = torch.stack([res.status.double(), res.iterations.double(), res.n_nodes.double(), res.path_n.double(),
                           res.path_len], 1)
        if world > 1:
            nonlocal gather_buf
            if gather_buf is None and rank == 0:
                gather_buf = [torch.empty_like(rec) for _ in range(world)]
            dist.gather(rec, gather_buf if rank == 0 else None, dst=0)
        if from_host:
            rec = rec.cpu()  # device -> host read of the step's results
        return rec, res

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 1)):
        rec, res = step(d, False, False)
    sync_all()
    launches0 = _lib.launches
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    stage = [0.0, 0.0, 0.0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        rec, res = step(d, False, True)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = _lib.launches - launches0
    clk = clocks.stop() if rank == 0 else None
    if gp is not None:
        stage = list(gp.stage_ms())
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max / 1e3)

    # end to end through the public API with HOST buffers (pinned): H2D of the inputs + D2H of the results every step
    step(host, True, False)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        rec_h, _ = step(host, True, False)
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(t.item()) / 1e3)
    d2h_bytes = rec_h.numel() * rec_h.element_size()

    status = res.status.cpu().numpy()
    iters = res.iterations.cpu().numpy()
    if rank == 0:
        pk = peaks()
        roofline = None
        if prof is not None:
            n_l, fl, t_ms = prof.totals()
            ach = fl / (t_ms / 1e3) / 1e12 if t_ms > 0 else 0.0
            traffic, traffic_note = conv_traffic()
            roofline = {"bound": "tensor", "kernel": "conv_gemm2_kernel / conv_halo2_kernel (tcgen05 implicit-GEMM convolution, every "
                                                     "layer shape of the forward)",
                        "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ach / pk["tensor"],
                        "traffic": traffic, "traffic_note": traffic_note,
                        "algorithmic_bytes_per_launch_avg": (sum(prof.bytes) / max(len(prof.bytes), 1)), "peak_source": pk["which"], "launches_timed": n_l,
                        "avg_launch_ms": t_ms / max(n_l, 1), "flops_per_launch_avg": fl / max(n_l, 1),
                        "step_share": {"cnn_ms": stage[0], "cdf_ms": stage[1], "draw_plan_ms": stage[2]}}  # last timed step
            if args.layer_table:
                with open(args.layer_table, "w") as f:
                    f.write("layer,launches,gflop_per_launch,ms_per_launch,tflops\n")
                    for name, (n, fl_l, ms_l) in sorted(prof.by_layer().items(), key=lambda kv: -kv[1][2]):
                        f.write(f"{name},{n},{fl_l / n / 1e9:.2f},{ms_l / n:.4f},{fl_l / max(ms_l, 1e-9) / 1e9:.1f}\n")
        else:
            roofline = {"bound": "hbm", "kernel": "plan_kernel", "achieved": None, "peak": pk["hbm"], "unit": "GB/s",
                        "frac": None, "traffic": None, "peak_source": pk["which"],
                        "note": "planner state lives in shared memory; see DESIGN.md"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                _, cpu = cpu_arm(args, wl, args.ref_sample or (16 if not uniform else 32), model=None)
            except Exception as exc:  # the baseline is informational; never lose the GPU line over it
                cpu = {"error": repr(exc)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64 planner / f16 tensor-core CNN (f32 accumulate)", "data": "synthetic",
                "config": {"workload": workload_name(args.workload, B), "tasks_per_gpu": B, "maps_per_gpu": int(wl.occ_maps.shape[0]),
                           "max_iterations": 5000, "neural_bias": 0.75, "weights": "random-init seed 0",
                           "task_filter": "start/goal in one free component (redrawn %d)" % wl.redrawn,
                           "l2": "inputs larger than L2 (activations and CDFs are GBs per step)",
                           "solved_fraction": float(np.mean(status == 0)), "mean_iterations": float(iters.mean()),
                           "median_iterations": float(np.median(iters))},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
                "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_graft(args)


if __name__ == "__main__":
    main()
