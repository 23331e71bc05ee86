"""Scratch: how much of the planner's time is queue tail?  Runs the 2D uniform batch in task order, then again with the
tasks sorted by the iteration counts of the first run (longest first = the best any difficulty predictor could do)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgr_sampling_cnn_b200 import engine, workload
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wl = workload.make_2d(B)
occ = torch.from_numpy(wl.occ_maps).to(dev)
bits = engine.pack_occupancy(occ, 2, dev)
buf = engine.PlannerBuffers(2, B, 5000, 512, False, 0, dev)
def run(order):
    t2m, st, gl = (torch.from_numpy(np.ascontiguousarray(x[order])).to(dev) for x in (wl.task_to_map, wl.starts, wl.goals))
    ids = torch.from_numpy(order.astype(np.int32)).to(dev)
    best = 1e9
    for _ in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); res = engine.plan_batch(bits, t2m, st, gl, dim=2, shape=(512, 512), seed=1, task_ids=ids, buffers=buf); b.record()
        torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best, res.iterations.cpu().numpy(), res.n_nodes.cpu().numpy()
order = np.arange(B)
t0, it, nn = run(order)
work = (it.astype(np.float64) * nn).sum()
print(f"task order: {t0:.1f} ms; iterations mean {it.mean():.0f}; unsolved(5000) {np.mean(it == 5000):.3f}")
lpt = np.argsort(-(it.astype(np.int64) * nn), kind="stable")
t1, it1, _ = run(lpt)
assert np.array_equal(it1, it[lpt])
print(f"longest first (oracle): {t1:.1f} ms")
rng = np.random.default_rng(0)
t2, _, _ = run(rng.permutation(B))
print(f"random order: {t2:.1f} ms")

# cheap difficulty proxies: straight-line blocked?, start-goal distance, free area in a window around the goal
occ_free = (wl.occ_maps != 0)
def line_blocked(m, s, g):
    t = np.linspace(0.0, 1.0, 64)
    ys = np.round(s[0] + (g[0] - s[0]) * t).astype(int); xs = np.round(s[1] + (g[1] - s[1]) * t).astype(int)
    return not occ_free[m, ys, xs].all()
blocked = np.array([line_blocked(wl.task_to_map[b], wl.starts[b], wl.goals[b]) for b in range(B)])
dist = np.sqrt(((wl.starts - wl.goals).astype(np.float64) ** 2).sum(1))
def goal_free(m, g, r=12):
    y0, y1, x0, x1 = max(g[0] - r, 0), min(g[0] + r + 1, 512), max(g[1] - r, 0), min(g[1] + r + 1, 512)
    return occ_free[m, y0:y1, x0:x1].mean()
gfree = np.array([goal_free(wl.task_to_map[b], wl.goals[b]) for b in range(B)])
from scipy.stats import spearmanr
cost = it.astype(np.float64) * nn
for name, v in (("blocked", blocked.astype(float)), ("dist", dist), ("1-goal_free", 1 - gfree)):
    print(name, "spearman vs cost", round(spearmanr(v, cost).correlation, 3))
for name, key in (("blocked,dist", blocked * 1000 + dist), ("goal_free", (1 - gfree) * 1000 + blocked * 100 + dist / 10)):
    t3, _, _ = run(np.argsort(-key, kind="stable"))
    print(f"proxy order [{name}]: {t3:.1f} ms")
