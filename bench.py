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
    ap.add_argument("--profile-every", type=int, default=5, help="bracket every n-th conv launch of the timed steps with CUDA events")
    ap.add_argument("--no-sections", action="store_true", help="skip the secondary sections (2D uniform, 3D 1024, 3D sweep)")
    ap.add_argument("--section-steps", type=int, default=2)
    ap.add_argument("--sweep-tasks", type=int, default=4096, help="total tasks of the strong-scaling 3D section")
    ap.add_argument("--micro-batch", type=int, default=None, help="tasks per decoder launch (default: GuidedPlanner's)")
    ap.add_argument("--encoder-batch", type=int, default=None, help="distinct maps per encoder pass (default: GuidedPlanner's)")
    ap.add_argument("--train-steps", type=int, default=6000, help="optimisation steps of the trained_guidance section (0 = skip it)")
    ap.add_argument("--train-maps", type=int, default=800, help="generated maps (x 10 tasks) of its training set")
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
    encoder pass + one decoder micro-batch of 40 tasks, the bench's own micro-batch (profiles/conv_traffic_r2.json), averaged
    over launches like `achieved`.  The like-for-like algorithmic figure of the SAME 29 launches is in the note (the step's
    own launch mix, `algorithmic_bytes_per_launch_avg`, has more decoder launches per encoder pass)."""
    path = os.path.join(ROOT, "profiles", "conv_traffic_r2.json")
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


FLOPS_NOMINAL = {2: 821.91e9, 3: 2522.27e9}  # 2 x MACs of the 35 conv layers of one reference forward


def _rank_workload(name, t0, t1):
    """Tasks [t0, t1) of the global task list of a workload (task t = task t % 10 of map t // 10; map m is drawn with
    random.seed(1000 + m), its tasks with random.seed(2000 + m): SURVEY 8(d)).  Returns (workload, global task ids)."""
    from mgr_sampling_cnn_b200 import workload
    per = 10
    m0, m1 = t0 // per, (t1 - 1) // per
    wl = (workload.make_2d if name.startswith("2d") else workload.make_3d)((m1 - m0 + 1) * per, first_map=m0)
    lo, hi = t0 - m0 * per, t1 - m0 * per
    maps_used, t2m = np.unique(wl.task_to_map[lo:hi], return_inverse=True)
    wl2 = type(wl)(wl.dim, wl.shape, wl.occ_maps[maps_used], t2m.astype(np.int32), wl.starts[lo:hi], wl.goals[lo:hi],
                   wl.coords[lo:hi], wl.filtered, wl.redrawn)
    return wl2, np.arange(t0, t1, dtype=np.int32)


class Arm(object):
    """One workload resident on this rank's GPU + the callable that runs one step of the hot path over it."""

    def __init__(self, name, wl, ids, dev, args, profile=False):
        from mgr_sampling_cnn_b200 import cnn, engine
        self.name, self.wl, self.dev, self.engine = name, wl, dev, engine
        self.dim = wl.dim
        self.uniform = name.endswith("uniform")
        self.B = len(wl.task_to_map)
        self.host = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in
                     (("occ", wl.occ_maps), ("t2m", wl.task_to_map), ("starts", wl.starts), ("goals", wl.goals),
                      ("coords", wl.coords), ("ids", ids))}
        self.d = {k: v.to(dev) for k, v in self.host.items()}
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.host.values())
        self.gp = self.prof = self.model = None
        if not self.uniform:
            self.model = make_model(self.dim, dev)
            self.gp = engine.GuidedPlanner(self.model, self.dim, wl.shape, micro_batch=args.micro_batch)
            if args.encoder_batch:
                self.gp.encoder_batch = args.encoder_batch
            if profile:
                self.prof = cnn.ConvProfiler(args.profile_every)
                self.model._plan(dev).set_profiler(self.prof)
        else:
            self.bits = engine.pack_occupancy(self.d["occ"], self.dim, dev)
            self.bufs = engine.PlannerBuffers(self.dim, self.B, 5000, 512, False, 0, dev)
        self.uni_events = []

    def step(self, from_host=False, profile=False):
        """One pass of the hot path; returns (records [B, 5] fp64 on the device: status, iterations, n_nodes, path points,
        path length; PlanResult)."""
        eng, dev = self.engine, self.dev
        inp = {k: v.to(dev, non_blocking=True) for k, v in self.host.items()} if from_host else self.d
        if self.uniform:
            b = self.bits if not from_host else eng.pack_occupancy(inp["occ"], self.dim, dev)
            ev = None
            if profile:
                ev = [torch.cuda.Event(enable_timing=True)]
                ev[0].record()
            res = eng.plan_batch(b, inp["t2m"], inp["starts"], inp["goals"], dim=self.dim, shape=self.wl.shape, seed=0,
                                 task_ids=inp["ids"], buffers=self.bufs, events=ev)
            if profile:
                ev.append(torch.cuda.Event(enable_timing=True))
                ev[-1].record()
                self.uni_events = ev  # [start, between draw and plan, end]
        else:
            res = self.gp.plan(inp["occ"], inp["t2m"], inp["starts"], inp["goals"], inp["coords"], seed=0, task_ids=inp["ids"],
                               record_events=profile)
        rec = torch.stack([res.status.double(), res.iterations.double(), res.n_nodes.double(), res.path_n.double(),
                           res.path_len], 1)
        return rec, res

    def stage_ms(self):
        if self.uniform:
            e = self.uni_events
            return (0.0, 0.0, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])) if len(e) == 3 else (0.0,) * 4
        return self.gp.stage_ms()


def train_section(args, dev, rank, world, pk, sync_all, max_over_ranks):
    """UNet_cooler.training_step + Adam (train.py:205-234) on the B200 kernels, batch_size 3 per GPU (the
    reference's MapsDataModule default, train.py:73) at 512 x 512, synthetic maps / masks.  N > 1: data parallel, one
    all-reduce of the flat gradient buffer per step (NCCL), weak scaling."""
    from mgr_sampling_cnn_b200 import _lib, train, trainer
    torch.manual_seed(0)
    model = train.UNet_cooler()
    B, S = 3, 512
    tr = trainer.UNetTrainer(model, B, S, S, dev)
    rng = np.random.default_rng(100 + rank)
    x = torch.from_numpy((rng.random((B, 1, S, S)) > 0.2).astype(np.float32)).repeat(1, 3, 1, 1).to(dev)
    y = torch.from_numpy(rng.random((B, 1, S, S)).astype(np.float32)).to(dev)
    coords = torch.from_numpy(rng.integers(0, S, (B, 1, 2, 2)).astype(np.float32)).to(dev)
    for _ in range(2):
        tr.training_step((x, y, coords))
    sync_all()
    steps = max(args.section_steps, 2) * 5
    n0 = _lib.launches
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        loss = tr.training_step((x, y, coords))
    b.record()
    sync_all()
    ms = max_over_ranks(a.elapsed_time(b)) / steps
    # the weight-gradient kernel alone on the largest layer (conv8[0]: 192 -> 128 channels at 512 x 512, train.py:157)
    cv = tr.convs["c8.0"]
    g = tr._buf("dx_c8.2", (B, S, S, 128))
    tr._wgrad(cv, g, 0, tr.cat8, 0)
    torch.cuda.synchronize(dev)
    a.record()
    for _ in range(5):
        tr._wgrad(cv, g, 0, tr.cat8, 0)
    b.record()
    torch.cuda.synchronize(dev)
    wg_ms = a.elapsed_time(b) / 5
    wg_flops = 2.0 * B * S * S * 128 * 192 * 9
    sec = {"workload": "UNet_cooler training step (forward in train mode, BCEWithLogits, backward, Adam), batch %d per GPU at %dx%d, synthetic" % (B, S, S),
           "scaling": "weak (data parallel: one NCCL all-reduce of the %d-float gradient buffer per step)" % tr.n_params if world > 1 else "weak",
           "steps": steps, "warmup": 2, "ms_per_step": ms, "value": world * B / (ms / 1e3), "unit": "samples/s",
           "flops_per_step_algorithmic": tr.flops_per_step(), "tflops_algorithmic": world * tr.flops_per_step() / ms / 1e9,
           "launches_per_step": (_lib.launches - n0) // steps, "loss_last": float(loss.item()),
           "dtype": "f16 tensor-core operands, f32 accumulate / parameters / optimizer state",
           "kernels": {"wgrad_kernel": {"bound": "tensor", "layer": "conv8[0] weight gradient, 192 -> 128 channels, K = %d pixels" % (B * S * S),
                                        "ms": wg_ms, "achieved": wg_flops / wg_ms / 1e9, "peak": pk["tensor"], "unit": "TFLOP/s",
                                        "frac": wg_flops / wg_ms / 1e9 / pk["tensor"]}}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:   # the oracle's torch fp32 step on the host cores, once (kind "port": the reference's Lightning loop cannot travel)
            from oracle import train_oracle
            sd = {k: v.clone() for k, v in model.state_dict().items()}
            t0 = time.time()
            train_oracle.forward_backward(sd, x[:2].cpu(), y[:2].cpu(), coords[:2].cpu())
            dt = time.time() - t0
            sec["cpu_baseline"] = {"value": 2 / dt, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": "one forward + backward of 2 samples (oracle/train_oracle.py, torch fp32 CPU)"}
        except Exception as exc:
            sec["cpu_baseline"] = {"error": repr(exc)}
    del tr
    torch.cuda.empty_cache()
    return sec


def trained_guidance_section(args, dev):
    """What SURVEY 8 f3 is for: train UNet_cooler on the device (generated maps, tasks, A* paths, masks: maps 5000.. of the
    generators, disjoint from the benchmark's maps 0..) and plan 1000 benchmark tasks with uniform sampling, random-init
    guidance and trained guidance (same tasks, same sample streams).  N = 1 only."""
    from mgr_sampling_cnn_b200 import engine, train, trainer, workload
    S, n_eval = 512, 1000
    wl = workload.make_on_device(2, n_eval, first_map=0, device=dev)
    bits = engine.pack_occupancy(wl.occ_maps, 2, dev)

    def stats(res, label):
        st, it = res.status.cpu().numpy(), res.iterations.cpu().numpy()
        pl = res.path_len.cpu().numpy()
        return {"label": label, "solved_fraction": float((st == 0).mean()), "mean_iterations": float(it.mean()),
                "median_iterations": float(np.median(it)), "mean_path_length_solved": float(pl[st == 0].mean()) if (st == 0).any() else None,
                "sampler_stuck": int((st == 3).sum()), "no_node_connected": int((st == 2).sum())}

    def guided(model, label):
        planner = engine.GuidedPlanner(model.to(dev).eval(), 2, (S, S), max_iterations=5000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        planner.plan(wl.occ_maps, wl.task_to_map, wl.starts, wl.goals, wl.coords, seed=7)
        a.record()
        res = planner.plan(wl.occ_maps, wl.task_to_map, wl.starts, wl.goals, wl.coords, seed=7)
        b.record()
        torch.cuda.synchronize(dev)
        out = stats(res, label)
        out["plans_per_s"] = n_eval / (a.elapsed_time(b) / 1e3)
        return out

    rows = [stats(engine.plan_batch(bits, wl.task_to_map, wl.starts, wl.goals, dim=2, shape=(S, S), max_iterations=5000, seed=7),
                  "uniform sampling (rrt_star.py)")]
    torch.manual_seed(0)
    model = train.UNet_cooler()
    rows.append(guided(model, "neural, random-init weights (BASELINE configs[1])"))
    fit = trainer.fit_on_generated(model.cpu(), args.train_maps, args.train_steps, first_map=5000, device=dev)
    rows.append(guided(model, "neural, trained %d steps on %d generated tasks" % (fit["steps"], fit["train_tasks"])))
    del model
    torch.cuda.empty_cache()
    return {"workload": "2D, %d benchmark tasks (maps 0..%d), 5000 iterations, neural_bias 0.75" % (n_eval, n_eval // 10 - 1),
            "training": fit, "plans": rows,
            "note": "training data and evaluation maps come from the reference's own generators (disjoint seeds); everything runs on the GPU"}


def run_graft(args):
    import torch.distributed as dist
    from mgr_sampling_cnn_b200 import _lib, sharding

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

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def run_arm(arm, gidx, n_global, steps, warmup, from_host=False, profile=False):
        """warmup untimed + `steps` timed steps of `arm`, each followed by the path's only collective: the gather of this
        rank's result records into rows `gidx` of the [n_global, 5] table on rank 0.  Returns (ms per step: max over
        ranks, last table on rank 0, last PlanResult)."""
        out = res = None

        def one(prof):
            nonlocal out, res
            rec, res = arm.step(from_host, prof)
            out = sharding.gather_records(rec, gidx, n_global, dst=0)
            if from_host and out is not None:
                out = out.cpu()  # device -> host read of the step's results
        for _ in range(warmup):
            one(False)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one(profile)
        e1.record()
        sync_all()
        return max_over_ranks(e0.elapsed_time(e1)) / max(steps, 1), out, res

    # ---- headline: args.workload, `tasks` per GPU, weak scaling (rank r owns global tasks [r B, (r + 1) B)) ----
    dim = 2 if args.workload.startswith("2d") else 3
    tasks = args.tasks or (4096 if dim == 2 else 1024)
    per_rank = ((tasks + 9) // 10) * 10  # whole maps per rank; the last map of a rank is truncated to `tasks`
    wl, ids = _rank_workload(args.workload, rank * per_rank, rank * per_rank + tasks)
    arm = Arm(args.workload, wl, ids, dev, args, profile=True)
    B = arm.B
    gidx = torch.arange(rank * B, (rank + 1) * B, device=dev)
    clocks = ClockSampler(local)
    run_arm(arm, gidx, world * B, 0, max(args.warmup, 1))
    launches0 = _lib.launches
    if rank == 0:
        clocks.start()
    ms_step, _, res = run_arm(arm, gidx, world * B, args.steps, 0, profile=True)
    launches = _lib.launches - launches0
    clk = clocks.stop() if rank == 0 else None
    stage = arm.stage_ms()
    value = world * B / (ms_step / 1e3)
    # end to end through the public API with HOST buffers (pinned): H2D of the inputs + D2H of the results every step
    ms_e2e, rec_h, _ = run_arm(arm, gidx, world * B, args.steps, 1, from_host=True)
    e2e_value = world * B / (ms_e2e / 1e3)
    arm_h2d, arm_uniform = arm.h2d_bytes, arm.uniform
    d2h_bytes = (rec_h.numel() * rec_h.element_size()) if rec_h is not None else 0
    status = res.status.cpu().numpy()
    iters = res.iterations.cpu().numpy()
    n_draws = float(res.n_draws.sum().item()) if res.n_draws is not None else 0.0
    pk = peaks()

    roofline, kernels = None, {}
    cells = int(np.prod(wl.shape))
    if arm.prof is not None:
        prof = arm.prof.collect()
        n_l, fl, t_ms = prof.totals()
        ach = fl / (t_ms / 1e3) / 1e12 if t_ms > 0 else 0.0
        traffic, traffic_note = conv_traffic()
        fl_step = prof.flops_seen / max(args.steps, 1)   # FLOPs the conv kernels executed in one step (every launch counted)
        roofline = {"bound": "tensor", "kernel": "conv_gemm2_kernel / conv_halo2_kernel (tcgen05 implicit-GEMM convolution, every "
                                                 "layer shape of the forward)",
                    "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ach / pk["tensor"],
                    "traffic": traffic, "traffic_note": traffic_note,
                    "how": f"every {prof.sample_every}th conv launch of the timed steps bracketed by CUDA events (uniform over the launch "
                           "sequence: flops-weighted); achieved = sum(flops) / sum(time) of that sample",
                    "algorithmic_bytes_per_launch_avg": prof.bytes_seen / max(prof.seen, 1), "peak_source": pk["which"],
                    "launches_timed": n_l, "launches_per_step": prof.seen / max(args.steps, 1),
                    "avg_launch_ms": t_ms / max(n_l, 1), "flops_per_launch_avg": fl / max(n_l, 1),
                    "flops_executed_per_task": fl_step / B, "flops_nominal_per_task": FLOPS_NOMINAL[dim],
                    "flops_note": "nominal = the reference's forward per task (SURVEY 8d); executed = after the exact rewrites of "
                                  "DESIGN 4.0 (encoder once per map, linear first convs hoisted per map, coordinate channels as "
                                  "epilogue polynomial, ConvTranspose folded): work removed, not skipped -- each rewrite has an "
                                  "equality test",
                    "step_tflops_executed": fl_step / (ms_step / 1e3) / 1e12,
                    "step_frac": fl_step / (ms_step / 1e3) / 1e12 / pk["tensor"],
                    "cnn_stage_frac": fl_step / (max(stage[0], 1e-9) / 1e3) / 1e12 / pk["tensor"],
                    "step_share": {"cnn_ms": stage[0], "cdf_ms": stage[1], "draw_ms": stage[2], "plan_ms": stage[3]}}  # last timed step
        cdf_bytes = 8.0 * B * cells
        kernels["cdf_stream_kernel"] = {"bound": "hbm", "ms": stage[1], "algorithmic_bytes": cdf_bytes,
                                        "achieved": cdf_bytes / max(stage[1], 1e-9) / 1e6, "peak": pk["hbm"], "unit": "GB/s",
                                        "frac": cdf_bytes / max(stage[1], 1e-9) / 1e6 / pk["hbm"]}
        if args.layer_table:
            with open(args.layer_table, "w") as f:
                f.write("layer,launches,gflop_per_launch,ms_per_launch,tflops\n")
                for name, (n, fl_l, ms_l) in sorted(prof.by_layer().items(), key=lambda kv: -kv[1][2]):
                    f.write(f"{name},{n},{fl_l / n / 1e9:.2f},{ms_l / n:.4f},{fl_l / max(ms_l, 1e-9) / 1e9:.1f}\n")
    else:
        roofline = {"bound": "hbm", "kernel": "plan_kernel", "achieved": None, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": None, "traffic": None, "peak_source": pk["which"],
                    "note": "planner state lives in shared memory; see DESIGN.md",
                    "step_share": {"cnn_ms": 0.0, "cdf_ms": 0.0, "draw_ms": stage[2], "plan_ms": stage[3]}}
    it_sum = float(iters.sum())
    kernels["draw_kernel"] = {"bound": "latency (dependent L2/DRAM probes, one warp per task)", "ms": stage[2], "draws": n_draws,
                              "draws_per_s": n_draws / max(stage[2], 1e-9) * 1e3,
                              "probe_sector_bytes": 32.0 * 18 * 0.75 * it_sum if not arm_uniform else 0.0,
                              "achieved": (32.0 * 18 * 0.75 * it_sum / max(stage[2], 1e-9) / 1e6) if not arm_uniform else None,
                              "peak": pk["hbm"], "unit": "GB/s"}
    kernels["plan_kernel"] = {"bound": "latency (chain of dependent iterations, tree in shared memory)", "ms": stage[3],
                              "iterations": it_sum, "iterations_per_s": it_sum / max(stage[3], 1e-9) * 1e3,
                              "hbm_algorithmic_bytes": float(B) * (cells / 8 + 5000 * 2 * dim) + float(res.n_nodes.sum().item()) * 20,
                              "unit": "GB/s", "peak": pk["hbm"],
                              "achieved": (float(B) * (cells / 8 + 5000 * 2 * dim) + float(res.n_nodes.sum().item()) * 20) / max(stage[3], 1e-9) / 1e6}

    # ---- secondary sections (the other BASELINE configs), a few steps each; the headline above stays on args.workload ----
    sections = {}
    if not args.no_sections:
        del arm
        torch.cuda.empty_cache()
        plan = []
        if args.workload != "2d_uniform":
            plan.append(("2d_uniform", "2d_uniform", 4096, "weak"))          # BASELINE configs[2]
        if args.workload != "3d_neural":
            plan.append(("3d_neural_1024", "3d_neural", 1024, "weak"))       # BASELINE configs[3]
        plan.append(("3d_sweep", "3d_neural", args.sweep_tasks, "strong"))   # BASELINE configs[4], scaled (see note)
        for key, wname, n_tasks, kind in plan:
            if kind == "weak":
                pr = ((n_tasks + 9) // 10) * 10
                swl, sids = _rank_workload(wname, rank * pr, rank * pr + n_tasks)
                sg, n_global = torch.arange(rank * n_tasks, (rank + 1) * n_tasks, device=dev), world * n_tasks
            else:
                idx = sharding.shard_indices(n_tasks, rank, world, interleaved=False)
                swl, sids = _rank_workload(wname, int(idx[0]), int(idx[-1]) + 1)
                sg, n_global = torch.from_numpy(idx).to(dev), n_tasks
            sarm = Arm(wname, swl, sids, dev, args)
            ms_s, out, sres = run_arm(sarm, sg, n_global, args.section_steps, 1)
            done = n_global
            sec = {"workload": workload_name(wname, done if kind == "strong" else sarm.B), "scaling": kind,
                   "tasks_total": done, "tasks_per_gpu": sarm.B, "steps": args.section_steps, "warmup": 1,
                   "ms_per_step": ms_s, "value": done / (ms_s / 1e3), "unit": UNIT}
            if rank == 0 and out is not None:
                o = out.cpu().numpy()
                sec.update({"solved_fraction": float(np.mean(o[:, 0] == 0)), "mean_iterations": float(o[:, 1].mean()),
                            "sampler_stuck": int(np.sum(o[:, 0] == 3)),
                            # identical at every N when the sharded result equals the single-GPU result (strong sections)
                            "checksum": {"iterations_sum": int(o[:, 1].sum()), "nodes_sum": int(o[:, 2].sum()),
                                         "path_len_sum": float(o[:, 4].sum())}})
            sections[key] = sec
            del sarm
            torch.cuda.empty_cache()
        sections["train_step"] = train_section(args, dev, rank, world, pk, sync_all, max_over_ranks)
        if world == 1 and args.train_steps > 0:
            sections["trained_guidance"] = trained_guidance_section(args, dev)
        sections["3d_sweep"]["note"] = (f"BASELINE configs[4] is 65536 tasks; the sweep here is {args.sweep_tasks} tasks of the same generator "
                                        "(fixed total, contiguous shards via sharding.shard_indices, records gathered with "
                                        "sharding.gather_records) so that every N fits the bench budget; --sweep-tasks 65536 runs it whole")

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                _, cpu = cpu_arm(args, wl, args.ref_sample or (16 if not args.workload.endswith("uniform") else 32), model=None)
            except Exception as exc:  # the baseline is informational; never lose the GPU line over it
                cpu = {"error": repr(exc)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64 planner / f16 tensor-core CNN (f32 accumulate)", "data": "synthetic",
                "config": {"workload": workload_name(args.workload, B), "tasks_per_gpu": B, "maps_per_gpu": int(wl.occ_maps.shape[0]),
                           "max_iterations": 5000, "neural_bias": 0.75, "weights": "random-init seed 0",
                           "task_filter": "start/goal in one free component (redrawn %d)" % wl.redrawn,
                           "l2": "inputs larger than L2 (activations and CDFs are GBs per step)",
                           "sharding": "rank r owns global tasks [r B, (r+1) B); records gathered to rank 0 (sharding.gather_records)",
                           "solved_fraction": float(np.mean(status == 0)), "mean_iterations": float(iters.mean()),
                           "median_iterations": float(np.median(iters))},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": arm_h2d, "d2h_bytes_per_step": d2h_bytes},
                "gpu_launches": launches, "clocks": clk, "roofline": roofline, "kernels": kernels, "sections": sections,
                "cpu_baseline": cpu}
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
