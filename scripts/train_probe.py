"""Scratch: gradient errors of the trainer vs the fp32 oracle for several loss scales, and the step time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgr_sampling_cnn_b200 import train, trainer
from oracle import train_oracle
from tests import helpers as H
dev = torch.device("cuda:0")
x, y, coords, masks, g = H.load_train_batch()
torch.manual_seed(0)
model = train.UNet_cooler()
sd = {k: v.clone() for k, v in model.state_dict().items()}
t0 = time.time()
oloss, ograds, ostats, ologits = train_oracle.forward_backward(sd, torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(coords))
print("oracle %.1f s loss %.6f" % (time.time() - t0, oloss))
xd, cd, md = torch.from_numpy(x).to(dev), torch.from_numpy(coords).to(dev), torch.from_numpy(masks).to(dev)
for mult in [float(a) for a in sys.argv[1:]] or [1.0]:
    tr = trainer.UNetTrainer(model, 2, 512, 512, dev, loss_scale=mult * 2 * 512 * 512)
    tr.forward(xd, cd); tr.set_targets(masks_u8=md); tr.loss_and_grad(True); tr.backward()
    torch.cuda.synchronize()
    grads = {k: v.cpu() for k, v in tr.named(tr.grads, 1.0 / tr.loss_scale).items()}
    errs = {k: float((grads[k] - ograds[k]).norm() / ograds[k].norm()) for k in ograds}
    order = sorted(errs, key=errs.get, reverse=True)
    print("loss_scale x%g: loss %.6f logits err %.2e; grad rel-L2 err max %.3g median %.3g; finite %s" % (
        mult, float(tr.loss_sum.item()) / tr.M, float((tr.logits.cpu() - ologits[:, 0]).abs().max()), errs[order[0]],
        float(np.median(list(errs.values()))), all(bool(torch.isfinite(v).all()) for v in grads.values())))
    print("   worst:", ", ".join("%s %.3g" % (k, errs[k]) for k in order[:8]))
    if mult == 1.0:
        for k in ograds:
            print("      %-34s err %.4f  |ref| %.3e  norm ratio %.4f" % (k, errs[k], float(ograds[k].norm()), float(grads[k].norm() / ograds[k].norm())))
    for name in ("dx_c8.2", "dx_c6.0", "dx_blk30", "dx_blk00"):
        t = tr._tmp.get(name)
        if t is not None:
            print("   |%s| max %.3g mean %.3g" % (name, float(t.float().abs().max()), float(t.float().abs().mean())))
    del tr
tr = trainer.UNetTrainer(model, 3, 512, 512, dev)
x3 = torch.from_numpy(np.concatenate([x, x[:1]])).to(dev); c3 = torch.from_numpy(np.concatenate([coords, coords[:1]])).to(dev)
m3 = torch.from_numpy(np.concatenate([masks, masks[:1]])).to(dev)
for it in range(6):
    a, b, c_, d = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    a.record(); tr.forward(x3, c3); tr.set_targets(masks_u8=m3); tr.loss_and_grad(True); b.record(); tr.backward(); c_.record(); tr.optimizer_step(); d.record()
    torch.cuda.synchronize()
    print("step %d (B=3): fwd %.2f ms bwd %.2f ms opt %.2f ms; loss %.5f; %.1f TFLOP/s" % (
        it, a.elapsed_time(b), b.elapsed_time(c_), c_.elapsed_time(d), float(tr.loss_sum.item()) / tr.M, tr.flops_per_step() / a.elapsed_time(d) / 1e9))
