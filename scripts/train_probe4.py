"""Scratch: after some training steps, compare BatchNorm running statistics with the current batch statistics, and the trainer's
train-mode logits with the exported model's eval-mode logits (inference kernels and the torch fp32 oracle)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgr_sampling_cnn_b200 import engine, generate_paths as gp, train, trainer, workload
from oracle import cnn_oracle
dev = torch.device("cuda:0"); B, S = 3, 512
wl = workload.make_on_device(2, 600, first_map=5000, device=dev)
bits = engine.pack_occupancy(wl.occ_maps, 2, dev)
_, steps, paths, path_n = engine.grid_paths(bits, wl.task_to_map, wl.starts, wl.goals, (S, S), path_cap=2048)
masks = gp.path_masks(wl.occ_maps, wl.task_to_map, paths, path_n, dev)
torch.manual_seed(0)
model = train.UNet_cooler()
tr = trainer.UNetTrainer(model, B, S, S, dev)
def batch(i):
    idx = torch.arange(i, i + B, device=dev)
    occ = wl.occ_maps[wl.task_to_map[idx].long()]
    return (occ != 0).float()[:, None].expand(-1, 3, -1, -1).contiguous(), wl.coords[idx], masks[idx]
for s in range(int(sys.argv[1]) if len(sys.argv) > 1 else 150):
    x, c, m = batch((s * B) % 540)
    loss = tr.training_step((x, None, c), masks_u8=m)
print("last train loss", float(loss.item()), "scale state [scale, good, flag, steps]:", tr.scale_state.cpu().tolist(),
      "non-finite params:", int((~torch.isfinite(tr.params)).sum()))
x, c, m = batch(570)
tr.forward(x, c); tr.set_targets(masks_u8=m); tr.loss_and_grad(False)
print("train-mode loss on a fresh batch", float(tr.loss_sum.item()) / tr.M)
lt = tr.logits.clone()
for name in ("enc0.0.c1", "enc0.1.c2", "enc1.0.ds", "enc2.1.c1", "enc3.1.c2"):
    cv = tr.convs[name]
    si, bi = int(name[3]), int(name[5])
    st = tr.blk[si][bi][{"c1": "st1", "c2": "st2", "ds": "std"}[name[7:]]]
    var_b = 1.0 / st[1] ** 2 - 1e-5
    print("%-10s mean: running %.4f batch %.4f (rms diff %.4f) | var: running %.4f batch %.4f" % (
        name, float(cv.bn[2].abs().mean()), float(st[0].abs().mean()), float((cv.bn[2] - st[0]).pow(2).mean().sqrt()),
        float(cv.bn[3].mean()), float(var_b.mean())))
tr.export_to(model)
sd = {k: v.detach().clone().to(dev) for k, v in model.state_dict().items()}
with torch.no_grad():
    lo = cnn_oracle.forward(sd, x, c, 2)[:, 0]
    cnn_oracle._TRAIN[0] = True
    sd2 = {k: v.clone() for k, v in sd.items()}
    lo_train = cnn_oracle.forward(sd2, x, c, 2)[:, 0]
    cnn_oracle._TRAIN[0] = False
    le = model.to(dev).eval()(x, c)[:, 0]
bias = float(tr.params[tr.head_b].item())
bce = lambda l: float(torch.nn.functional.binary_cross_entropy_with_logits(l, tr.target).item())
print("BCE: trainer(train mode) %.4f | oracle train-mode on exported weights %.4f | oracle eval-mode %.4f | inference kernels eval-mode %.4f" % (
    bce(lt), bce(lo_train), bce(lo), bce(le)))
print("max |trainer - oracle train-mode| %.4f ; max |inference - oracle eval| %.4f" % (float((lt - lo_train).abs().max()), float((le - lo).abs().max())))
