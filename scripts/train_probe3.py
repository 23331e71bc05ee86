"""Scratch: how much do the fp32 gradients move when the forward activations carry fp16-like relative noise (1e-3)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from mgr_sampling_cnn_b200 import train
from oracle import cnn_oracle
from tests import helpers as H
dev = torch.device("cuda:0")
x, y, coords, masks, g = H.load_train_batch()
torch.manual_seed(0)
model = train.UNet_cooler()
base = {k: v.clone().to(dev) for k, v in model.state_dict().items() if not k.startswith("base_model.") and not k.endswith("num_batches_tracked")}
xd, yd, cd = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), torch.from_numpy(coords).to(dev)
orig_conv = cnn_oracle._conv
NOISE = [0.0]
def conv(xx, sd_, p, dims, stride=1, padding=1):
    out = orig_conv(xx, sd_, p, dims, stride, padding)
    if NOISE[0] > 0 and out.numel() > 4096:
        out = out * (1 + NOISE[0] * torch.randn_like(out)).detach()    # multiplicative rounding-like noise on every conv output
    return out
cnn_oracle._conv = conv
def run(noise):
    NOISE[0] = noise
    sd = {k: v.clone().requires_grad_(v.dtype == torch.float32 and not k.endswith(("running_mean", "running_var"))) for k, v in base.items()}
    cnn_oracle._TRAIN[0] = True
    with cnn_oracle.strict_fp32():
        loss = F.binary_cross_entropy_with_logits(cnn_oracle._forward(sd, xd, cd, 2), yd)
        loss.backward()
    cnn_oracle._TRAIN[0] = False
    return {k: v.grad for k, v in sd.items() if v.grad is not None}
g0 = run(0.0)
for noise in (3e-4, 1e-3):
    torch.manual_seed(1)
    g1 = run(noise)
    errs = {k: float((g0[k] - g1[k]).norm() / g0[k].norm()) for k in g0}
    print("noise %.0e:" % noise, ", ".join("%s %.3f" % (k, errs[k]) for k in ("conv1.0.weight", "conv1.2.weight", "conv2.0.conv1.weight", "conv2.0.conv2.weight",
          "conv2.0.bn2.bias", "conv2.1.conv1.weight", "conv3.0.conv1.weight", "conv4.0.conv1.weight", "conv5.0.conv1.weight", "conv5.1.conv2.weight", "conv8.0.weight")))
