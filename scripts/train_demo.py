"""End-to-end demonstration of SURVEY 8 f1 + f2 + f3 on one GPU, everything on the device:
maps and tasks (csrc/generate.cu) -> shortest paths (csrc/gridpath.cu) -> training masks (csrc/masks.cu) -> training steps
(trainer.UNetTrainer) -> the trained weights in the guided planner (engine.GuidedPlanner) on held-out maps, against the same
planner with random-init weights and with uniform sampling.
usage: python scripts/train_demo.py [train_maps] [epochs] [eval_tasks]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgr_sampling_cnn_b200 import engine, generate_paths as gp, train, trainer, workload

n_maps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
n_eval = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
dev = torch.device("cuda:0")
B, S = 3, 512


def training_set(first_map, maps):
    wl = workload.make_on_device(2, maps * 10, first_map=first_map, device=dev)
    bits = engine.pack_occupancy(wl.occ_maps, 2, dev)
    _, steps, paths, path_n = engine.grid_paths(bits, wl.task_to_map, wl.starts, wl.goals, (S, S), path_cap=2048)
    masks = gp.path_masks(wl.occ_maps, wl.task_to_map, paths, path_n, dev)
    return wl, masks


def evaluate(model, wl, label):
    planner = engine.GuidedPlanner(model, 2, (S, S), max_iterations=5000)
    t0 = time.time()
    res = planner.plan(wl.occ_maps, wl.task_to_map, wl.starts, wl.goals, wl.coords, seed=7)
    torch.cuda.synchronize()
    st, it = res.status.cpu().numpy(), res.iterations.cpu().numpy()
    out = {"label": label, "solved_fraction": float((st == 0).mean()), "mean_iterations": float(it.mean()), "median_iterations": float(np.median(it)),
           "mean_path_length_solved": float(res.path_len.cpu().numpy()[st == 0].mean()) if (st == 0).any() else None, "seconds": time.time() - t0}
    print(json.dumps(out))
    return out


t0 = time.time()
wl_train, masks = training_set(5000, n_maps)
torch.cuda.synchronize()
n = masks.shape[0]
print("training set: %d tasks on %d maps generated on the device in %.1f s" % (n, n_maps, time.time() - t0))
wl_eval = workload.make_on_device(2, n_eval, first_map=0, device=dev)
torch.manual_seed(0)
model = train.UNet_cooler()
bits_eval = engine.pack_occupancy(wl_eval.occ_maps, 2, dev)
ru = engine.plan_batch(bits_eval, wl_eval.task_to_map, wl_eval.starts, wl_eval.goals, dim=2, shape=(S, S), max_iterations=5000, goal_threshold=5.0,
                       seed=7)
torch.cuda.synchronize()
su, iu = ru.status.cpu().numpy(), ru.iterations.cpu().numpy()
results = [{"label": "uniform sampling (rrt_star.py)", "solved_fraction": float((su == 0).mean()), "mean_iterations": float(iu.mean()),
            "median_iterations": float(np.median(iu)), "mean_path_length_solved": float(ru.path_len.cpu().numpy()[su == 0].mean())}]
print(json.dumps(results[0]))
results.append(evaluate(model.to(dev).eval(), wl_eval, "random-init weights"))
model = model.cpu()
tr = trainer.UNetTrainer(model, B, S, S, dev)
rng = np.random.default_rng(0)
t0 = time.time()
run = []
for step in range(1, n_steps + 1):
    # the reference halves the rate on a validation plateau (ReduceLROnPlateau, train.py:230); here: fixed halvings late in the run
    tr.lr = 1e-3 if step <= 0.6 * n_steps else (5e-4 if step <= 0.8 * n_steps else 2.5e-4)
    idx = torch.from_numpy(rng.choice(n, B, replace=False)).to(dev)
    occ = wl_train.occ_maps[wl_train.task_to_map[idx].long()]
    x = (occ != 0).float()[:, None].expand(-1, 3, -1, -1).contiguous()
    run.append(tr.training_step((x, None, wl_train.coords[idx]), masks_u8=masks[idx]))
    if step % 250 == 0:
        print("step %d: mean loss %.4f, lr %g, loss scale %g, %.1f s" % (step, float(torch.stack(run).mean().item()), tr.lr,
                                                                      float(tr.scale_state[0].item()), time.time() - t0))
        run = []
step = n_steps
torch.cuda.synchronize()
print("trained %d steps in %.1f s (%.1f samples/s)" % (step, time.time() - t0, step * B / (time.time() - t0)))
tr.export_to(model)
# diagnostic: held-out loss in eval mode through the trainer's graph (validation_step) and through the exported model (inference kernels)
wl_val, masks_val = training_set(9000, 3)
model_dev = model.to(dev).eval()
lt, le = [], []
for i in range(0, 30, B):
    idx = torch.arange(i, i + B, device=dev)
    occ = wl_val.occ_maps[wl_val.task_to_map[idx].long()]
    x = (occ != 0).float()[:, None].expand(-1, 3, -1, -1).contiguous()
    lt.append(float(tr.validation_step((x, None, wl_val.coords[idx]), masks_u8=masks_val[idx]).item()))
    with torch.no_grad():
        logits = model_dev(x, wl_val.coords[idx])
    le.append(float(torch.nn.functional.binary_cross_entropy_with_logits(logits[:, 0], tr.target).item()))
print("held-out loss: validation_step (eval mode) %.4f, exported model through the inference kernels %.4f" % (np.mean(lt), np.mean(le)))
model = model.cpu()
del tr
torch.cuda.empty_cache()
results.append(evaluate(model.to(dev).eval(), wl_eval, "trained %d steps" % step))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(results, open("gpurun_out/train_demo.json", "w"), indent=1)
