"""Scratch: activation gradients of the trainer vs torch autograd (fp32 on the GPU) at the block boundaries of the encoder."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from mgr_sampling_cnn_b200 import train, trainer
from oracle import cnn_oracle
from tests import helpers as H
dev = torch.device("cuda:0")
x, y, coords, masks, g = H.load_train_batch()
torch.manual_seed(0)
model = train.UNet_cooler()
sd = {k: v.clone().to(dev).requires_grad_(v.dtype == torch.float32 and not k.endswith(("running_mean", "running_var")))
      for k, v in model.state_dict().items() if not k.startswith("base_model.") and not k.endswith("num_batches_tracked")}
cap = {}
orig_block, orig_layer = cnn_oracle._block, cnn_oracle._layer
def block(xx, sd_, p, dims, stride):
    out = orig_block(xx, sd_, p, dims, stride)
    out.retain_grad(); cap[p] = out
    return out
def layer(xx, sd_, p, dims, stride):
    if p == "conv2":
        xx.retain_grad(); cap["x1"] = xx
    return orig_layer(xx, sd_, p, dims, stride)
cnn_oracle._block, cnn_oracle._layer = block, layer
cnn_oracle._TRAIN[0] = True
xd, yd, cd = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), torch.from_numpy(coords).to(dev)
with cnn_oracle.strict_fp32():
    logits = cnn_oracle._forward(sd, xd, cd, 2)
    loss = F.binary_cross_entropy_with_logits(logits, yd)
    loss.backward()
cnn_oracle._TRAIN[0] = False
tr = trainer.UNetTrainer(model, 2, 512, 512, dev)
tr.forward(xd, cd); tr.set_targets(masks_u8=torch.from_numpy(masks).to(dev)); tr.loss_and_grad(True)
# forward activations
for si in range(4):
    for bi in range(2):
        out, off = tr.blk[si][bi]["out"]
        ref = cap["conv%d.%d" % (si + 2, bi)].detach().permute(0, 2, 3, 1)
        got = out[..., off:off + ref.shape[-1]].float()
        print("fwd  stage %d block %d: rel err %.3e" % (si, bi, float((got - ref).norm() / ref.norm())))
tr.backward(); torch.cuda.synchronize()
M = tr.loss_scale
def cmp(name, got, ref):
    ref = ref.permute(0, 2, 3, 1)
    got = got.float() / M
    print("bwd  %-10s rel err %.4f  |ref| %.3e  mean(got)/mean(ref) %.4f" % (name, float((got - ref).norm() / ref.norm()), float(ref.norm()),
          float(got.mean() / ref.mean())))
for si in range(4):
    # gradient of block0's output of stage si = dx of block1
    cmp("s%d.b0.out" % si, tr._tmp["dx_blk%d1" % si].view((2,) + tr.lv[si] + (64 << si,)), cap["conv%d.0" % (si + 2)].grad)
    if si > 0:
        cmp("s%d.out" % (si - 1), tr._tmp["dx_blk%d0" % si].view((2,) + tr.lv[si - 1] + (64 << (si - 1),)), cap["conv%d.1" % (si + 1)].grad)
cmp("x1", tr._tmp["dx_blk00"].view(2, 512, 512, 64), cap["x1"].grad)
