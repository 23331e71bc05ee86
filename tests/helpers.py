"""Shared helpers for the parity tests: golden loading and oracle-side task setup."""
import os

import numpy as np

from oracle import planner_oracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PLAN_FIXTURES = ["plan2d_uniform_mt_0", "plan2d_uniform_mt_1", "plan2d_uniform_stream", "plan2d_neural_stream",
                 "plan3d_uniform_mt_0", "plan3d_uniform_stream", "plan3d_neural_stream"]
# reference plans at MAX_ITERATIONS = 5000 (uniform sampler on the Philox stream): *_0 run all 5000 iterations (best
# effort), *_1 are solved after > 4000 iterations
LONG_FIXTURES = ["plan2d_long_0", "plan2d_long_1", "plan3d_long_0", "plan3d_long_1"]
STREAM_FIXTURES = [n for n in PLAN_FIXTURES if n.endswith("_stream")] + LONG_FIXTURES
PLAN_FIXTURES = PLAN_FIXTURES + LONG_FIXTURES


def load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        g = {k: z[k] for k in z.files}
    if name in LONG_FIXTURES:  # the map is stored bit-packed (255 = free in 2D / obstacle in 3D)
        shape = tuple(g["occ_shape"])
        g["occ_map"] = (np.unpackbits(g["occ_map"])[:int(np.prod(shape))].reshape(shape) * 255).astype(np.uint8)
    return g


def load_full_cnn(name):
    """cnn2d_512_seed0 / cnn3d_80_seed0 (oracle/gen_golden.py::gen_cnn_full): the occupancy map as the task file holds it
    (uint8, 255 = free in 2D / obstacle in 3D), the tensor the reference's MapsDataset fed to the model, coords, fp32 logits."""
    g = load(name)
    occ_shape, img_shape = tuple(g["occ_shape"]), tuple(g["image_shape"])
    g["occ"] = (np.unpackbits(g["occ_bits"])[:int(np.prod(occ_shape))].reshape(occ_shape) * 255).astype(np.uint8)
    g["x"] = np.unpackbits(g["image_bits"])[:int(np.prod(img_shape))].reshape(img_shape).astype(np.float32)
    return g


def fixture_dim(name):
    return 3 if "3d" in name else 2


def occ_free(name, g):
    return po.occ_free_2d(g["occ_map"]) if fixture_dim(name) == 2 else po.occ_free_3d(g["occ_map"])


def planner_heat(name, g):
    """The planner-side heat map the reference stores: heat_map[0] + 10 (neural_rrt.py:64) / heat_map - 250 (3D :68)."""
    if "heat_map" not in g:
        return None
    if fixture_dim(name) == 2:
        return (g["heat_map"][0] + 10).astype(np.float32)
    return (g["heat_map"] - 250).astype(np.float32)


def padded_samples(g, max_iter, dim):
    s = np.zeros((max_iter, dim), dtype=np.float64)
    it = int(g["ref_iterations"])
    s[:it] = g["ref_samples"]
    return s


def ref_trace_int(g, dim):
    """Reference verdict list as (iteration, verdict) + segments."""
    v = g["ref_verdicts"]
    return v[:, 0].astype(np.int32), v[:, -1].astype(np.int32), v[:, 2:2 + 2 * dim]


def random_tasks_2d(rng, occ, n):
    """Free start/goal pairs at distance >= 100 (generate_tasks.py:27-46 semantics, numpy RNG)."""
    H, W = occ.shape
    out_s, out_g = [], []
    while len(out_s) < n:
        sy, sx, gy, gx = rng.integers(0, H), rng.integers(0, W), rng.integers(0, H), rng.integers(0, W)
        if occ[sy, sx] != 0 and occ[gy, gx] != 0 and np.hypot(sy - gy, sx - gx) >= 100.0:
            out_s.append((sy, sx))
            out_g.append((gy, gx))
    return np.array(out_s, dtype=np.int32), np.array(out_g, dtype=np.int32)


def random_tasks_3d(rng, occ, n):
    S = occ.shape
    out_s, out_g = [], []
    while len(out_s) < n:
        s = tuple(int(rng.integers(0, S[k])) for k in range(3))
        g = tuple(int(rng.integers(0, S[k])) for k in range(3))
        if occ[s] == 0 and occ[g] == 0 and np.linalg.norm(np.subtract(s, g)) >= 20.0:
            out_s.append(s)
            out_g.append(g)
    return np.array(out_s, dtype=np.int32), np.array(out_g, dtype=np.int32)


def load_train_batch(name="train2d_seed0"):
    """The golden training batch as MapsDataset yields it (train.py:38-69): x fp32 [B,3,H,W] in {0,1}, y fp32 [B,1,H,W] in
    [0,1] (min-max of the mask), coords fp32 [B,1,2,2]; plus the raw uint8 masks and the golden record."""
    g = load(name)
    occ = np.unpackbits(g["occ_bits"], axis=-1).astype(np.float32)          # {0, 1}: the min-max normalised map
    x = np.repeat(occ[:, None], 3, axis=1)
    m = g["masks"].astype(np.float64)
    y = np.stack([(k - k.min()) / (k.max() - k.min()) for k in m])[:, None].astype(np.float32)
    return x, y, g["coords"].astype(np.float32), g["masks"], g
