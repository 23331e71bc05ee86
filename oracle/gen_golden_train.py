"""TEST INFRASTRUCTURE. Golden training step from the UNMODIFIED reference (train.py:109-234 through oracle/run_reference.py's
`train_step` job): random-init UNet_cooler (torch.manual_seed(0)), train mode, one batch of two 512 x 512 samples built
exactly like MapsDataset builds them (train.py:38-69) from the reference-made maps / A* paths / masks of
tests/golden/masks2d.npz, BCEWithLogitsLoss, backward, one Adam step.

Run in the authoring container only:   python oracle/gen_golden_train.py      -> tests/golden/train2d_seed0.npz
Stored: the batch (maps bit-packed, masks uint8, coords), the loss, per-parameter gradient norms / sums, parameter sums after
the step, and a handful of full gradient / updated tensors and BatchNorm running statistics.
"""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
from gen_golden import GOLDEN, run_job, save  # noqa: E402

KEEP = ["final_conv.weight", "final_conv.bias", "conv1.0.weight", "conv1.0.bias", "conv1.2.bias", "conv8.2.bias", "conv8.0.bias",
        "upconv8.bias", "upconv6.bias", "conv_coords_64.0.weight", "conv_coords_64.0.bias", "conv_coords_512.0.bias",
        "conv2.0.bn1.weight", "conv2.1.bn2.bias", "conv3.0.downsample.1.weight", "conv5.1.bn2.weight", "conv5.0.downsample.0.weight",
        "upconv8.weight", "conv8.2.weight"]
KEEP_STATS = ["conv2.0.bn1.running_mean", "conv2.0.bn1.running_var", "conv5.1.bn2.running_mean", "conv5.1.bn2.running_var",
              "conv3.0.downsample.1.running_var"]


def batch_from_goldens():
    m = np.load(os.path.join(GOLDEN, "masks2d.npz"))
    occs, masks, coords = [], [], []
    for i in (0, 1):
        occ, mask, path = m[f"occ_{i}"], m[f"mask_{i}"], m[f"path_{i}"]
        assert occ.shape == (512, 512)
        occs.append(occ)
        masks.append(mask)
        (sy, sx), (fy, fx) = path[0], path[-1]
        coords.append([[sx, sy], [fx, fy]])                      # train.py:57-60: [[start_x, start_y], [finish_x, finish_y]]
    return np.stack(occs), np.stack(masks), np.array(coords, np.float32)[:, None]


def dataset_tensors(occs, masks):
    """MapsDataset.__getitem__ (train.py:38-52): RGB image and mask min-max normalised in float64, then .float()."""
    x, y = [], []
    for occ, mask in zip(occs, masks):
        img = np.stack([occ] * 3, -1)
        img = (img - img.min()) / (img.max() - img.min())
        mk = (mask - mask.min()) / (mask.max() - mask.min())
        x.append(img.transpose(2, 0, 1).astype(np.float32))
        y.append(mk[None].astype(np.float32))
    return np.stack(x), np.stack(y)


def main():
    occs, masks, coords = batch_from_goldens()
    x, y = dataset_tensors(occs, masks)
    o = run_job(job="train_step", seed=0, x=x, y=y, coords=coords, keep=np.array(KEEP), keep_stats=np.array(KEEP_STATS))
    print("loss", float(o["loss"]), "params with grad", len(o["names"]))
    save("train2d_seed0", occ_bits=np.packbits(occs != 0, axis=-1), masks=masks, coords=coords, seed=np.int64(0), **o)


if __name__ == "__main__":
    main()
