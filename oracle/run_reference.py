"""TEST INFRASTRUCTURE. Runs the UNMODIFIED reference (read-only /root/reference) in a private process and records traces.

Only used in the authoring container to (1) pin `oracle/planner_oracle.c` / `oracle/sampler_oracle.py` /
`oracle/cnn_oracle.py` against the real reference and (2) generate the golden fixtures under `tests/golden/`
(`oracle/gen_golden.py`).  Nothing on the GPU box imports this file: /root/reference does not exist there.

Invocation:  python oracle/run_reference.py <job.npz> <out.npz>
The process sees only [oracle/ref_shim, oracle/, REFERENCE_DIR] on sys.path (SURVEY.md §8c "Name-collision pitfall").
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_DIR = os.environ.get("NRRT_REFERENCE_DIR", "/root/reference")


def _private_path():
    repo_root = os.path.dirname(_HERE)
    keep = [p for p in sys.path if p and os.path.abspath(p) not in (repo_root, _HERE)]
    sys.path[:] = [os.path.join(_HERE, "ref_shim"), _HERE, REFERENCE_DIR] + keep


_private_path()

import math  # noqa: E402
import random as _py_random  # noqa: E402

import numpy as np  # noqa: E402


def _patch_torchvision():
    """`resnet18(pretrained=True)` would download (train.py:115, ThreeD_train.py:112); BASELINE uses random init."""
    import torchvision.models as tvm
    import torchvision.models.video as tvv
    _r18, _r3d = tvm.resnet18, tvv.r3d_18
    tvm.resnet18 = lambda *a, **k: _r18(weights=None)
    tvv.r3d_18 = lambda *a, **k: _r3d(weights=None)
    tvm.video.r3d_18 = tvv.r3d_18


def _as_tuple(p):
    return tuple(float(v) for v in p)


def run_plan(job):
    """Run one reference plan with instrumentation; returns a dict of arrays (the trace)."""
    modname = str(job["module"])
    mod = __import__(modname)
    dim = 3 if modname.startswith("ThreeD") else 2
    neural = "neural" in modname
    start = tuple(int(v) for v in job["start"])
    goal = tuple(int(v) for v in job["goal"])
    if bool(job.get("float_endpoints", False)):
        start, goal = tuple(float(v) for v in start), tuple(float(v) for v in goal)
    max_iter = int(job["max_iterations"])
    thr = float(job["goal_threshold"])

    if "stream_seed" in job:
        from philox_stream import StreamRandom
        rng = StreamRandom(int(job["stream_seed"]), int(job["stream_task"]))
        mod.random = rng
    else:
        rng = None
        mod.random = _py_random
        _py_random.seed(int(job["mt_seed"]))

    if neural:
        planner = mod.RRTStar(job["occ_map"], job["heat_map"], start, goal, max_iter, thr, float(job["neural_bias"]))
    else:
        planner = mod.RRTStar(job["occ_map"], start, goal, max_iter, thr)

    samples, neural_flags, verdicts, radii, pow_hazard = [], [], [], [], []
    state = {"kind": 0, "it": 0}

    orig_nn = planner.find_nearest_neighbor
    orig_cf = planner.is_collision_free
    orig_steer = planner.steer
    orig_connect = planner.can_connect_nodes
    orig_rewire = planner.rewire_tree

    def nn(sample):
        samples.append(_as_tuple(sample))
        radii.append(float(planner.search_radius))
        return orig_nn(sample)

    def steer(from_node, to_point):
        d = [to_point[k] - from_node.position[k] for k in range(dim)]
        hz = any(isinstance(v, float) and (v ** 2 != v * v) for v in d)
        dist = math.sqrt(sum(v ** 2 for v in d))
        if dist > planner.search_radius:
            d2 = [v * planner.search_radius / dist for v in d]
            hz = hz or any((v ** 2 != v * v) for v in d2)
        pow_hazard.append(bool(hz))
        return orig_steer(from_node, to_point)

    def cf(p1, p2):
        r = orig_cf(p1, p2)
        verdicts.append((planner.iteration_no, state["kind"]) + _as_tuple(p1) + _as_tuple(p2) + (1.0 if r else 0.0,))
        return r

    def connect(a, b):
        state["kind"] = 0
        return orig_connect(a, b)

    def rewire(n):
        # loop A / loop B cannot be told apart from outside; tag by argument order inside cf via positions
        state["kind"] = 1
        return orig_rewire(n)

    planner.find_nearest_neighbor = nn
    planner.is_collision_free = cf
    planner.steer = steer
    planner.can_connect_nodes = connect
    planner.rewire_tree = rewire

    if neural:
        orig_ns = planner.generate_neural_sample
        orig_rs = planner.generate_random_sample

        def ns():
            neural_flags.append(1)
            return orig_ns()

        def rs():
            neural_flags.append(0)
            return orig_rs()
        planner.generate_neural_sample = ns
        planner.generate_random_sample = rs

    path, iters = planner.rrt_star()

    nodes = list(planner.nodes)
    index = {id(n): i for i, n in enumerate(nodes)}
    solved = bool(path) and tuple(path[-1]) == tuple(planner.goal) and planner.best_node is not None \
        and planner.best_distance <= thr
    extra = None
    if solved:
        # the goal-reaching node is not in `nodes` (rrt_star.py:203-208); it gets index len(nodes)
        extra = planner.best_node
        if id(extra) not in index:
            index[id(extra)] = len(nodes)
            nodes = nodes + [extra]
    pos = np.array([_as_tuple(n.position) for n in nodes], dtype=np.float64).reshape(len(nodes), dim)
    cost = np.array([float(n.cost) for n in nodes], dtype=np.float64)
    parent = np.array([-1 if n.parent is None else index.get(id(n.parent), -2) for n in nodes], dtype=np.int32)
    best_idx = -1 if planner.best_node is None else index.get(id(planner.best_node), -2)
    out = {
        "pos": pos, "cost": cost, "parent": parent,
        "n_nodes": np.int32(len(planner.nodes)), "iterations": np.int32(iters), "solved": np.int32(solved),
        "best_node": np.int32(best_idx), "best_distance": np.float64(planner.best_distance),
        "path": np.array([_as_tuple(p) for p in path], dtype=np.float64).reshape(len(path), dim),
        "samples": np.array(samples, dtype=np.float64).reshape(len(samples), dim),
        "radii": np.array(radii, dtype=np.float64),
        "verdicts": np.array(verdicts, dtype=np.float64).reshape(len(verdicts), 3 + 2 * dim),
        "neural_flags": np.array(neural_flags, dtype=np.int8),
        "pow_hazard": np.array(pow_hazard, dtype=np.int8),
        "draws": np.int64(rng.draws if rng is not None else -1),
    }
    return out


def run_neural_draws(job):
    """Call generate_neural_sample() `count` times on one heat map under a stream (sampler KAT)."""
    modname = str(job["module"])
    mod = __import__(modname)
    from philox_stream import StreamRandom
    rng = StreamRandom(int(job["stream_seed"]), int(job["stream_task"]))
    mod.random = rng
    dim = 3 if modname.startswith("ThreeD") else 2
    start = (1,) * dim
    planner = mod.RRTStar(job["occ_map"], job["heat_map"], start, start, 10, 1.0, 1.0)
    out = [tuple(int(v) for v in planner.generate_neural_sample()) for _ in range(int(job["count"]))]
    return {"draws_xyz": np.array(out, dtype=np.int32), "draws": np.int64(rng.draws)}


def run_cnn(job):
    """Random-init reference model (torch.manual_seed) forward on CPU fp32."""
    import torch
    _patch_torchvision()
    three_d = bool(job["three_d"])
    torch.manual_seed(int(job["seed"]))
    if three_d:
        import ThreeD_train as T
        model = T.ThreeD_UNet_cooler()
    else:
        import train as T
        model = T.UNet_cooler()
    model.eval()
    if bool(job.get("randomize_bn", False)):
        g = torch.Generator().manual_seed(int(job["seed"]) + 1)
        for m in model.modules():
            if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
                m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    with torch.no_grad():
        y = model(torch.from_numpy(job["x"]), torch.from_numpy(job["coords"]))
    out = {"logits": y.numpy(), "n_keys": np.int32(len(model.state_dict()))}
    if bool(job.get("dump_keys", False)):
        out["keys"] = np.array(list(model.state_dict().keys()))
        out["shapes"] = np.array([str(tuple(v.shape)) for v in model.state_dict().values()])
    return out


def run_train_step(job):
    """One optimisation step of the unmodified reference: UNet_cooler.training_step (train.py:205-210) in train mode, backward,
    and one step of the optimizer configure_optimizers() returns (train.py:227-234) -- what a Lightning Trainer does per batch."""
    import torch
    _patch_torchvision()
    torch.manual_seed(int(job["seed"]))
    import train as T
    model = T.UNet_cooler()
    model.train()
    x, y, coords = (torch.from_numpy(np.asarray(job[k])) for k in ("x", "y", "coords"))
    opt = model.configure_optimizers()["optimizer"]
    opt.zero_grad()
    loss = model.training_step((x, y, coords), 0)
    loss.backward()
    out = {"loss": np.float64(loss.item())}
    names, gnorm, gsum = [], [], []
    keep = set(str(k) for k in job["keep"])
    alias = {"base_model.layer1.": "conv2.", "base_model.layer2.": "conv3.", "base_model.layer3.": "conv4.", "base_model.layer4.": "conv5."}
    params = {}
    for k, p_ in model.named_parameters():   # shared modules are listed once, under base_model.layerN: name them as forward() uses them
        for a, b in alias.items():
            if k.startswith(a):
                k = b + k[len(a):]
        params[k] = p_
    for k, p_ in params.items():
        if p_.grad is None:
            continue
        names.append(k)
        gnorm.append(float(p_.grad.double().norm()))
        gsum.append(float(p_.grad.double().sum()))
        if k in keep:
            out["grad/" + k] = p_.grad.numpy().copy()
    opt.step()
    psum = [float(params[k].detach().double().sum()) for k in names]
    sd = model.state_dict()
    for k in keep:
        out["new/" + k] = sd[k].numpy().copy()
    for k in job["keep_stats"]:
        out["stat/" + str(k)] = sd[str(k)].numpy().copy()
    out.update(names=np.array(names), grad_norm=np.array(gnorm), grad_sum=np.array(gsum), param_sum_after=np.array(psum))
    return out


def run_dataset_cnn(job):
    """The reference's own data path in front of the model: a task file written to disk, read back by `MapsDataset`
    (train.py:38-69 / ThreeD_train.py:38-66: min-max normalisation, coordinates parsed from the file name, ToTensorV2),
    random-init model forward on CPU fp32, then the evaluation harness's glue: `get_occmap_and_coordinates`
    (test_network_and_algorithms.py:37-49 / ThreeD_test_network_and_algorithms.py:37-54) and the post-processing in front
    of the planner (clamp(-3, 1), :109-114 / min-max + 0.96 threshold, :98-105)."""
    import tempfile
    from pathlib import Path
    import torch
    _patch_torchvision()
    three_d = bool(job["three_d"])
    name = str(job["name"])
    occ = np.asarray(job["occ_map"])
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "images"))
        os.makedirs(os.path.join(td, "masks"))
        rng = np.random.default_rng(0)
        if three_d:
            import ThreeD_train as T
            np.save(os.path.join(td, "images", name), occ)
            np.save(os.path.join(td, "masks", name), rng.random(occ.shape))
        else:
            import train as T
            from PIL import Image
            Image.fromarray(occ).save(os.path.join(td, "images", name))
            Image.fromarray((rng.random(occ.shape) * 255).astype(np.uint8)).save(os.path.join(td, "masks", name))
        ds = T.MapsDataset(Path(td), [name], T.MapsDataModule(main_path=Path(td)).transforms)
        image, mask, coords = ds[0]
    image, coords = image[None], coords[None]  # DataLoader(batch_size=1)
    torch.manual_seed(int(job["seed"]))
    model = (T.ThreeD_UNet_cooler() if three_d else T.UNet_cooler()).eval()
    with torch.no_grad():
        output = model(image, coords)
    if three_d:
        occ_map = np.array(image[0].detach().cpu().numpy()).transpose((1, 2, 0))
        c = coords.data.tolist()[0][0]
        start, finish = tuple(int(v) for v in c[0]), tuple(int(v) for v in c[1])
        v = np.array(output[0, 0].detach().cpu().numpy())
        v = ((v - v.min()) / (v.max() - v.min())) * 1.0
        heat = v.copy()
        heat[~(v > 0.96)] = 0
    else:
        occ_map = np.array(image.data.detach().cpu().numpy()).transpose((0, 2, 3, 1))[0]
        c = coords.data.tolist()[0][0]
        start, finish = (int(c[0][1]), int(c[0][0])), (int(c[1][1]), int(c[1][0]))
        heat = torch.clamp(output, min=-3, max=1).detach().cpu().numpy().transpose((0, 2, 3, 1))
    return {"image": image.numpy(), "coords": coords.numpy(), "logits": output.numpy(), "planner_occ_map": occ_map,
            "start": np.array(start), "finish": np.array(finish), "heat_map": heat}


def run_gen(job):
    """generate_maps.create_map + generate_tasks.draw_start_and_finish under random.seed (2D or 3D)."""
    three_d = bool(job["three_d"])
    n_tasks = int(job["n_tasks"])
    if three_d:
        import ThreeD_generate_maps as GM
        import ThreeD_generate_tasks as GT
        GM.random = _py_random
        GT.random = _py_random
        _py_random.seed(int(job["map_seed"]))
        occ = GM.create_map(int(job["w"]), int(job["h"]), int(job["d"]))
        _py_random.seed(int(job["task_seed"]))
        names = [GT.draw_start_and_finish(occ, "blanks/map_0_path_sx_sy_sz_fx_fy_fz.npy", i) for i in range(n_tasks)]
    else:
        import generate_maps as GM
        import generate_tasks as GT
        GM.random = _py_random
        GT.random = _py_random
        GT.visualize_start_finish = lambda *a, **k: None  # writes to a hard-coded sibling dir (generate_tasks.py:61-73)
        _py_random.seed(int(job["map_seed"]))
        occ = GM.create_map(int(job["w"]), int(job["h"]))[:, :, 0]
        _py_random.seed(int(job["task_seed"]))
        names = [GT.draw_start_and_finish(occ, "blanks/map_0_path_sx_sy_fx_fy.png", i) for i in range(n_tasks)]
    return {"occ_map": np.asarray(occ), "names": np.array(names)}


def run_radius(job):
    """compute_search_radius for i = 1..n (rrt_star.py:174-182 / ThreeD_rrt_star.py:183-191)."""
    modname = str(job["module"])
    mod = __import__(modname)
    dim = 3 if modname.startswith("ThreeD") else 2
    shape = tuple(int(v) for v in job["shape"])
    occ = np.ones(shape, dtype=np.uint8) if dim == 2 else np.zeros(shape, dtype=np.float64)
    p = mod.RRTStar(occ, (1,) * dim, (2,) * dim, 10, 1.0)
    out = []
    for i in range(1, int(job["n"]) + 1):
        p.iteration_no = i
        out.append(p.compute_search_radius(dim=dim))
    return {"radius": np.array(out, dtype=np.float64)}


JOBS = {"train_step": run_train_step, "plan": run_plan, "neural_draws": run_neural_draws, "cnn": run_cnn, "gen": run_gen, "radius": run_radius,
        "dataset_cnn": run_dataset_cnn}


def main():
    job = dict(np.load(sys.argv[1], allow_pickle=False))
    job = {k: (v.item() if v.shape == () else v) for k, v in job.items()}
    out = JOBS[str(job["job"])](job)
    np.savez_compressed(sys.argv[2], **out)


if __name__ == "__main__":
    main()
