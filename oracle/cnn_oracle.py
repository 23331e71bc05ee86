"""TEST INFRASTRUCTURE — plain PyTorch fp32 restatement of the reference's CNN forward passes, used as the checker for
the CUDA forward (a floating-point kernel keeps a torch fp32 reference).  Works from a `state_dict` only.

Follows train.py:171-203 (UNet_cooler.forward) and ThreeD_train.py:169-204 (ThreeD_UNet_cooler.forward) plus
torchvision's BasicBlock (resnet.py) / VideoResNet BasicBlock, with eval-mode BatchNorm.  Pinned against the reference
itself by tests/test_cnn_oracle.py on the golden logits in tests/golden/cnn{2d,3d}_seed0*.npz.
"""
import contextlib

import torch
import torch.nn.functional as F


@contextlib.contextmanager
def strict_fp32():
    """The checker is fp32: on CUDA tensors cuDNN / cuBLAS would otherwise be allowed TF32 (10-bit mantissa) convolutions,
    which with decoder activations of 100-300 is an error of several 1e-3 in the checker itself."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


_TRAIN = [False]   # train-mode BatchNorm (batch statistics, running statistics updated in `sd`): oracle/train_oracle.py


def _bn(x, sd, p):
    if _TRAIN[0]:
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], True, 0.1, 1e-5)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)


def _conv(x, sd, p, dims, stride=1, padding=1):
    f = F.conv2d if dims == 2 else F.conv3d
    return f(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def _block(x, sd, p, dims, stride):
    if dims == 2:
        c1, b1, c2, b2 = p + ".conv1", p + ".bn1", p + ".conv2", p + ".bn2"
    else:
        c1, b1, c2, b2 = p + ".conv1.0", p + ".conv1.1", p + ".conv2.0", p + ".conv2.1"
    out = F.relu(_bn(_conv(x, sd, c1, dims, stride), sd, b1))
    out = _bn(_conv(out, sd, c2, dims), sd, b2)
    if (p + ".downsample.0.weight") in sd:
        x = _bn(_conv(x, sd, p + ".downsample.0", dims, stride, 0), sd, p + ".downsample.1")
    return F.relu(out + x)


def _layer(x, sd, p, dims, stride):
    return _block(_block(x, sd, p + ".0", dims, stride), sd, p + ".1", dims, 1)


def forward(sd, x, coords, dims):
    """sd: state_dict of UNet_cooler (dims=2) / ThreeD_UNet_cooler (dims=3), fp32 CPU or CUDA tensors (TF32 disabled)."""
    with strict_fp32():
        return _forward(sd, x, coords, dims)


def _forward(sd, x, coords, dims):
    mode = "bilinear" if dims == 2 else "trilinear"
    if dims == 2:
        x1 = F.relu(_conv(F.relu(_conv(x, sd, "conv1.0", 2)), sd, "conv1.2", 2))
    else:
        x = x.unsqueeze(1)
        x1 = F.relu(_conv(F.relu(_conv(x, sd, "conv1.0.0", 3)), sd, "conv1.0.2", 3))
        x1 = F.relu(_bn(x1, sd, "conv1.1"))
    x2 = _layer(x1, sd, "conv2", dims, 1)
    x3 = _layer(x2, sd, "conv3", dims, 2)
    x4 = _layer(x3, sd, "conv4", dims, 2)
    x5 = _layer(x4, sd, "conv5", dims, 2)

    def up(c, ref):
        c = c if dims == 2 else c.unsqueeze(-1)
        return F.interpolate(c, size=ref.shape[2:], mode=mode, align_corners=True)

    c64 = F.relu(_conv(coords, sd, "conv_coords_64.0", 2))
    x2 = torch.cat((x2, up(c64, x2)), 1)
    c128 = F.relu(_conv(c64, sd, "conv_coords_128.0", 2))
    x3 = torch.cat((x3, up(c128, x3)), 1)
    c256 = F.relu(_conv(c128, sd, "conv_coords_256.0", 2))
    x4 = torch.cat((x4, up(c256, x4)), 1)
    c512 = F.relu(_conv(c256, sd, "conv_coords_512.0", 2))
    x5 = torch.cat((x5, up(c512, x5)), 1)

    ft = F.conv_transpose2d if dims == 2 else F.conv_transpose3d
    x6 = ft(x5, sd["upconv6.weight"], sd["upconv6.bias"], stride=2)
    x6 = torch.cat((x6, x4), 1)
    x6 = F.relu(_conv(F.relu(_conv(x6, sd, "conv6.0", dims)), sd, "conv6.2", dims))
    x7 = ft(x6, sd["upconv7.weight"], sd["upconv7.bias"], stride=2)
    x7 = torch.cat((x7, x3), 1)
    x7 = F.relu(_conv(F.relu(_conv(x7, sd, "conv7.0", dims)), sd, "conv7.2", dims))
    x8 = ft(x7, sd["upconv8.weight"], sd["upconv8.bias"], stride=2)
    x8 = torch.cat((x8, x2), 1)
    x8 = F.relu(_conv(F.relu(_conv(x8, sd, "conv8.0", dims)), sd, "conv8.2", dims))
    return _conv(x8, sd, "final_conv", dims, 1, 0)


def randomize_bn(model, seed):
    """Same perturbation oracle/run_reference.py applies for the *_bn goldens (non-trivial running stats)."""
    g = torch.Generator().manual_seed(int(seed) + 1)
    for m in model.modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
