"""Scratch: training trajectory -- loss, loss scale, skipped steps, activation / logit magnitudes every 100 steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgr_sampling_cnn_b200 import engine, generate_paths as gp, train, trainer, workload
dev = torch.device("cuda:0"); B, S = 3, 512
n_maps = 200
wl = workload.make_on_device(2, n_maps * 10, first_map=5000, device=dev)
bits = engine.pack_occupancy(wl.occ_maps, 2, dev)
_, steps, paths, path_n = engine.grid_paths(bits, wl.task_to_map, wl.starts, wl.goals, (S, S), path_cap=2048)
masks = gp.path_masks(wl.occ_maps, wl.task_to_map, paths, path_n, dev)
torch.manual_seed(0)
model = train.UNet_cooler()
lr = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
tr = trainer.UNetTrainer(model, B, S, S, dev, lr=lr)
rng = np.random.default_rng(0)
n = masks.shape[0]
acc = []
for s in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1500):
    idx = torch.from_numpy(rng.choice(n, B, replace=False)).to(dev)
    occ = wl.occ_maps[wl.task_to_map[idx].long()]
    x = (occ != 0).float()[:, None].expand(-1, 3, -1, -1).contiguous()
    loss = tr.training_step((x, None, wl.coords[idx]), masks_u8=masks[idx])
    acc.append(loss)
    if (s + 1) % 100 == 0:
        l = torch.stack(acc).mean().item(); acc = []
        st = tr.scale_state.cpu().tolist()
        print("step %4d loss %.4f | scale %g taken %d | max|x1| %.1f |x5| %.1f |a6| %.1f |a8| %.1f |b8| %.1f |logit| %.1f sigma(logit).mean %.3f target.mean %.3f | |params| %.2f" % (
            s + 1, l, st[0], st[3], float(tr.x1.float().abs().max()), float(tr.c5[..., :512].float().abs().max()), float(tr.a6.float().abs().max()),
            float(tr.a8.float().abs().max()), float(tr.b8.float().abs().max()), float(tr.logits.abs().max()), float(torch.sigmoid(tr.logits).mean()),
            float(tr.target.mean()), float(tr.params.norm())))
