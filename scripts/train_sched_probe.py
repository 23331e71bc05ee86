"""Scratch: trained-guidance outcome of fit_on_generated under a schedule (argv: maps steps schedule)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgr_sampling_cnn_b200 import engine, train, trainer, workload
n_maps, steps, sched = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
dev = torch.device("cuda:0")
wl = workload.make_on_device(2, 1000, first_map=0, device=dev)
torch.manual_seed(0)
model = train.UNet_cooler()
fit = trainer.fit_on_generated(model, n_maps, steps, device=dev, schedule=sched)
planner = engine.GuidedPlanner(model.to(dev).eval(), 2, (512, 512), max_iterations=5000)
res = planner.plan(wl.occ_maps, wl.task_to_map, wl.starts, wl.goals, wl.coords, seed=7)
torch.cuda.synchronize()
st, it = res.status.cpu().numpy(), res.iterations.cpu().numpy()
print(json.dumps({"schedule": sched, "steps": steps, "final_loss": fit["loss_curve"][-1][1], "val_last": fit["val_loss_last_step"], "val_best": fit["val_loss_best"], "val_kept": fit["val_loss_kept_state_recalibrated"], "val_curve": fit["val_curve"], "train_s": round(fit["train_seconds"], 1),
                  "solved": float((st == 0).mean()), "mean_it": float(it.mean()), "median_it": float(np.median(it)),
                  "len": float(res.path_len.cpu().numpy()[st == 0].mean())}))
