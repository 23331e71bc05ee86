"""TEST INFRASTRUCTURE. Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/run_reference.py).

Run in the authoring container only (needs /root/reference):   python oracle/gen_golden.py [--only NAME]
The fixtures are committed; the GPU box and the CPU test-suite read them and never touch /root/reference.
Each fixture stores the exact inputs, the reference's outputs, and the library versions that defined the arithmetic
(the reference pins nothing itself, so "the reference executed in this container" is the pin).
"""
import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(_HERE), "tests", "golden")


def run_job(**job):
    with tempfile.TemporaryDirectory() as td:
        jp, op = os.path.join(td, "job.npz"), os.path.join(td, "out.npz")
        np.savez(jp, **job)
        env = dict(os.environ, OMP_NUM_THREADS=os.environ.get("OMP_NUM_THREADS", "8"))
        subprocess.check_call([sys.executable, os.path.join(_HERE, "run_reference.py"), jp, op], env=env)
        with np.load(op, allow_pickle=False) as z:
            return {k: z[k] for k in z.files}


def versions():
    import scipy
    import torch
    return np.array([f"python {sys.version.split()[0]}", f"numpy {np.__version__}", f"scipy {scipy.__version__}",
                     f"torch {torch.__version__}"])


def save(name, **arrays):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, versions=versions(), **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def parse_name_2d(name):
    import re
    m = re.search(r"_sx(\d+)_sy(\d+)_fx(\d+)_fy(\d+)\.png", name)
    sx, sy, fx, fy = (int(v) for v in m.groups())
    return (sy, sx), (fy, fx)  # rrt_star.py:22-28 returns (y, x)


def parse_name_3d(name):
    import re
    m = re.search(r"_sx(\d+)_sy(\d+)_sz(\d+)_fx(\d+)_fy(\d+)_fz(\d+)\.npy", name)
    v = [int(t) for t in m.groups()]
    return tuple(v[:3]), tuple(v[3:])


def synth_heat_2d(occ, start, goal, seed):
    """A clamp(-3,1)-range logit map: corridor around the start-goal segment + noise, quantised to 1/64 (exact fp32)."""
    rng = np.random.default_rng(seed)
    H, W = occ.shape
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    p0, p1 = np.array(start, float), np.array(goal, float)
    d = p1 - p0
    t = np.clip(((yy - p0[0]) * d[0] + (xx - p0[1]) * d[1]) / (d @ d), 0, 1)
    dist = np.hypot(yy - (p0[0] + t * d[0]), xx - (p0[1] + t * d[1]))
    logit = 1.5 - dist / 25.0 + rng.normal(0, 0.3, size=(H, W))
    logit = np.clip(np.round(logit * 64) / 64, -3, 1).astype(np.float32)
    return logit.reshape(1, H, W, 1)


def synth_heat_3d(occ, start, goal, seed):
    """A normalised+thresholded 3D map like ThreeD_neural_rrt.py:339-346 produces: zeros except a tube in (0.96, 1]."""
    rng = np.random.default_rng(seed)
    S = occ.shape
    g = np.mgrid[0:S[0], 0:S[1], 0:S[2]].astype(np.float64)
    p0, p1 = np.array(start, float), np.array(goal, float)
    d = p1 - p0
    t = np.clip(sum((g[k] - p0[k]) * d[k] for k in range(3)) / (d @ d), 0, 1)
    dist = np.sqrt(sum((g[k] - (p0[k] + t * d[k])) ** 2 for k in range(3)))
    v = np.exp(-dist / 40.0) * (1 - 0.01 * rng.random(S))
    v = ((v - v.min()) / (v.max() - v.min())).astype(np.float32)
    v[~(v > 0.96)] = 0
    return v


def gen_radius():
    r2 = run_job(job="radius", module="rrt_star", shape=np.array([512, 512]), n=5000)["radius"]
    r3 = run_job(job="radius", module="ThreeD_rrt_star", shape=np.array([80, 80, 80]), n=5000)["radius"]
    r2b = run_job(job="radius", module="rrt_star", shape=np.array([96, 160]), n=300)["radius"]
    r3b = run_job(job="radius", module="ThreeD_rrt_star", shape=np.array([24, 40, 32]), n=300)["radius"]
    save("radius", r2d_512=r2, r3d_80=r3, r2d_96x160=r2b, r3d_24x40x32=r3b)


def gen_maps():
    out = {}
    for i, (ms, ts) in enumerate([(1000, 2000), (1001, 2001)]):
        g = run_job(job="gen", three_d=False, n_tasks=10, map_seed=ms, task_seed=ts, w=512, h=512, d=0)
        out[f"occ2d_{i}"] = g["occ_map"]
        out[f"names2d_{i}"] = g["names"]
        out[f"seeds2d_{i}"] = np.array([ms, ts])
        g = run_job(job="gen", three_d=True, n_tasks=10, map_seed=ms, task_seed=ts, w=80, h=80, d=80)
        out[f"occ3d_{i}"] = (g["occ_map"] != 0).astype(np.uint8) * 255
        out[f"names3d_{i}"] = g["names"]
        out[f"seeds3d_{i}"] = np.array([ms, ts])
    save("maps_tasks", **out)
    return out


def gen_plans():
    mt = np.load(os.path.join(GOLDEN, "maps_tasks.npz"))
    # ---- 2D uniform, CPython MT stream (level-A: samples are replayed) ----
    occ = mt["occ2d_0"]
    for i, (task_no, seed, iters) in enumerate([(0, 11, 700), (3, 12, 500)]):
        start, goal = parse_name_2d(str(mt["names2d_0"][task_no]))
        tr = run_job(job="plan", module="rrt_star", occ_map=occ, start=np.array(start), goal=np.array(goal),
                     max_iterations=iters, goal_threshold=5.0, mt_seed=seed)
        save(f"plan2d_uniform_mt_{i}", occ_map=occ, start=np.array(start), goal=np.array(goal),
             max_iterations=np.int32(iters), goal_threshold=np.float64(5.0), **{"ref_" + k: v for k, v in tr.items()})
    # ---- 2D uniform, Philox stream (level-B: both sides run their own sampler) ----
    occ = mt["occ2d_1"]
    start, goal = parse_name_2d(str(mt["names2d_1"][1]))
    tr = run_job(job="plan", module="rrt_star", occ_map=occ, start=np.array(start), goal=np.array(goal),
                 max_iterations=600, goal_threshold=5.0, stream_seed=77, stream_task=5)
    save("plan2d_uniform_stream", occ_map=occ, start=np.array(start), goal=np.array(goal), max_iterations=np.int32(600),
         goal_threshold=np.float64(5.0), stream_seed=np.int64(77), stream_task=np.int32(5),
         **{"ref_" + k: v for k, v in tr.items()})
    # ---- 2D neural, Philox stream ----
    heat = synth_heat_2d(occ, start, goal, 5)
    occ_rgb = np.repeat((occ.astype(np.float32) / 255.0)[:, :, None], 3, axis=2)  # train.py:41-44 -> HWC float [0,1]
    tr = run_job(job="plan", module="neural_rrt", occ_map=occ_rgb, heat_map=heat, start=np.array(start),
                 goal=np.array(goal), max_iterations=600, goal_threshold=5.0, neural_bias=0.75, stream_seed=78,
                 stream_task=9)
    save("plan2d_neural_stream", occ_map=occ, heat_map=heat, start=np.array(start), goal=np.array(goal),
         max_iterations=np.int32(600), goal_threshold=np.float64(5.0), neural_bias=np.float64(0.75),
         stream_seed=np.int64(78), stream_task=np.int32(9), **{"ref_" + k: v for k, v in tr.items()})
    # ---- 3D uniform (MT + stream) ----
    occ3 = mt["occ3d_0"].astype(np.float64)
    start, goal = parse_name_3d(str(mt["names3d_0"][2]))
    tr = run_job(job="plan", module="ThreeD_rrt_star", occ_map=occ3, start=np.array(start), goal=np.array(goal),
                 max_iterations=700, goal_threshold=3.0, mt_seed=21)
    save("plan3d_uniform_mt_0", occ_map=mt["occ3d_0"], start=np.array(start), goal=np.array(goal),
         max_iterations=np.int32(700), goal_threshold=np.float64(3.0), **{"ref_" + k: v for k, v in tr.items()})
    occ3 = mt["occ3d_1"].astype(np.float64)
    start, goal = parse_name_3d(str(mt["names3d_1"][4]))
    tr = run_job(job="plan", module="ThreeD_rrt_star", occ_map=occ3, start=np.array(start), goal=np.array(goal),
                 max_iterations=600, goal_threshold=3.0, stream_seed=79, stream_task=3)
    save("plan3d_uniform_stream", occ_map=mt["occ3d_1"], start=np.array(start), goal=np.array(goal),
         max_iterations=np.int32(600), goal_threshold=np.float64(3.0), stream_seed=np.int64(79),
         stream_task=np.int32(3), **{"ref_" + k: v for k, v in tr.items()})
    # ---- 3D neural, Philox stream (dataset normalisation maps 255 -> 1, ThreeD_train.py:48-50) ----
    occ3n = (mt["occ3d_1"] != 0).astype(np.float32)
    heat3 = synth_heat_3d(occ3n, start, goal, 6)
    tr = run_job(job="plan", module="ThreeD_neural_rrt", occ_map=occ3n, heat_map=heat3, start=np.array(start),
                 goal=np.array(goal), max_iterations=600, goal_threshold=3.0, neural_bias=0.75, stream_seed=80,
                 stream_task=4)
    save("plan3d_neural_stream", occ_map=mt["occ3d_1"], heat_map=heat3, start=np.array(start), goal=np.array(goal),
         max_iterations=np.int32(600), goal_threshold=np.float64(3.0), neural_bias=np.float64(0.75),
         stream_seed=np.int64(80), stream_task=np.int32(4), **{"ref_" + k: v for k, v in tr.items()})


def gen_draws():
    mt = np.load(os.path.join(GOLDEN, "maps_tasks.npz"))
    occ = mt["occ2d_0"]
    start, goal = parse_name_2d(str(mt["names2d_0"][5]))
    heat = synth_heat_2d(occ, start, goal, 15)
    occ_rgb = np.repeat((occ.astype(np.float32) / 255.0)[:, :, None], 3, axis=2)
    d2 = run_job(job="neural_draws", module="neural_rrt", occ_map=occ_rgb, heat_map=heat, stream_seed=5, stream_task=1,
                 count=300)
    occ3n = (mt["occ3d_0"] != 0).astype(np.float32)
    s3, g3 = parse_name_3d(str(mt["names3d_0"][5]))
    heat3 = synth_heat_3d(occ3n, s3, g3, 16)
    d3 = run_job(job="neural_draws", module="ThreeD_neural_rrt", occ_map=occ3n, heat_map=heat3, stream_seed=6,
                 stream_task=2, count=300)
    save("neural_draws", occ2d=occ, heat2d=heat, draws2d=d2["draws_xyz"], ndraws2d=d2["draws"], seed2d=np.int64(5),
         task2d=np.int32(1), occ3d=mt["occ3d_0"], heat3d=heat3, draws3d=d3["draws_xyz"], ndraws3d=d3["draws"],
         seed3d=np.int64(6), task3d=np.int32(2))


def gen_cnn():
    rng = np.random.default_rng(3)
    x = (rng.random((2, 1, 64, 64)) > 0.3).astype(np.float32).repeat(3, axis=1)
    coords = np.array([[[[5., 9.], [50., 40.]]], [[[60., 3.], [12., 33.]]]], dtype=np.float32)
    for bn in (False, True):
        o = run_job(job="cnn", three_d=False, seed=0, x=x, coords=coords, randomize_bn=bn, dump_keys=not bn)
        extra = {} if bn else {"keys": o["keys"], "shapes": o["shapes"]}
        save("cnn2d_seed0" + ("_bn" if bn else ""), x=x, coords=coords, logits=o["logits"], seed=np.int64(0),
             randomize_bn=np.int8(bn), **extra)
    x3 = (rng.random((2, 16, 16, 16)) > 0.7).astype(np.float32)
    c3 = np.array([[[[1., 2., 3.], [14., 9., 11.]]], [[[8., 0., 15.], [2., 13., 6.]]]], dtype=np.float32)
    for bn in (False, True):
        o = run_job(job="cnn", three_d=True, seed=0, x=x3, coords=c3, randomize_bn=bn, dump_keys=not bn)
        extra = {} if bn else {"keys": o["keys"], "shapes": o["shapes"]}
        save("cnn3d_seed0" + ("_bn" if bn else ""), x=x3, coords=c3, logits=o["logits"], seed=np.int64(0),
             randomize_bn=np.int8(bn), **extra)


def gen_cnn_full():
    """One sample at each BASELINE map size (512^2, 80^3) pushed through the reference's own dataset class, model and
    harness glue (run_reference.run_dataset_cnn).  Stores the map (bit-packed), the tensors the reference fed to the model,
    the fp32 CPU logits and what the harness hands to the planner."""
    mt = np.load(os.path.join(GOLDEN, "maps_tasks.npz"))
    for three_d, key, task_no in ((False, "2d", 0), (True, "3d", 3)):
        occ = mt[f"occ{key}_0"]
        name = os.path.basename(str(mt[f"names{key}_0"][task_no]))
        o = run_job(job="dataset_cnn", three_d=three_d, seed=0, occ_map=occ, name=name)
        img = o["image"]
        assert set(np.unique(img)) <= {0.0, 1.0}
        # the planner-side map must be the file's map again (3D: ToTensorV2 moved the last axis first, the harness undoes it)
        pm = o["planner_occ_map"]
        assert np.array_equal((pm[..., 0] if not three_d else pm) != 0, occ != 0)
        heat = o["heat_map"]
        save("cnn3d_80_seed0" if three_d else "cnn2d_512_seed0", occ_bits=np.packbits(occ != 0), occ_shape=np.array(occ.shape),
             file_name=np.array(name), image_bits=np.packbits(img != 0), image_shape=np.array(img.shape), coords=o["coords"],
             logits=o["logits"], start=o["start"], finish=o["finish"], seed=np.int64(0),
             heat_nonzero=np.int64(np.count_nonzero(heat)) if three_d else np.int64(-1),
             heat_argmax=np.array(np.unravel_index(np.argmax(o["logits"][0, 0]), o["logits"][0, 0].shape)))


def gen_long_plans():
    """Plans that run (close to) all MAX_ITERATIONS = 5000 iterations in the reference itself: one best-effort and one
    late-solved task per dimension, uniform sampler on the Philox stream (level B).  Minutes of CPU each."""
    mt = np.load(os.path.join(GOLDEN, "maps_tasks.npz"))
    for name, occ_key, names_key, task_no, seed in (("plan2d_long_0", "occ2d_0", "names2d_0", 0, 91),
                                                    ("plan2d_long_1", "occ2d_1", "names2d_1", 9, 91),
                                                    ("plan3d_long_0", "occ3d_0", "names3d_0", 0, 92),
                                                    ("plan3d_long_1", "occ3d_0", "names3d_0", 5, 92)):
        three_d = "3d" in name
        occ = mt[occ_key]
        start, goal = (parse_name_3d if three_d else parse_name_2d)(str(mt[names_key][task_no]))
        thr = 3.0 if three_d else 5.0
        tr = run_job(job="plan", module="ThreeD_rrt_star" if three_d else "rrt_star",
                     occ_map=occ.astype(np.float64) if three_d else occ, start=np.array(start), goal=np.array(goal),
                     max_iterations=5000, goal_threshold=thr, stream_seed=seed, stream_task=task_no)
        save(name, occ_map=np.packbits(occ != 0), occ_shape=np.array(occ.shape), start=np.array(start), goal=np.array(goal),
             max_iterations=np.int32(5000), goal_threshold=np.float64(thr), stream_seed=np.int64(seed),
             stream_task=np.int32(task_no), **{"ref_" + k: v for k, v in tr.items()})


STEPS = {"radius": gen_radius, "maps": gen_maps, "plans": gen_plans, "draws": gen_draws, "cnn": gen_cnn, "long_plans": gen_long_plans, "cnn_full": gen_cnn_full}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    for name, fn in STEPS.items():
        if a.only in (None, name):
            fn()
