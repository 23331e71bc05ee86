"""Drop-in for the model class of the reference's ThreeD_train.py: `ThreeD_UNet_cooler` (ThreeD_train.py:106-204).

Same constructor, sub-module names and 276-key `state_dict`; `forward(x, coords)`: x (B,D,H,W) fp32 (unsqueezed to one
channel inside, :171), coords (B,1,2,3) fp32 -> logits (B,1,D,H,W) fp32.  The stem keeps the reference's quirk: only
`stem[0]` of torchvision's r3d_18 BasicStem is replaced, so a BatchNorm3d + ReLU follow the two-conv stem (:112-118).
forward() runs the B200 kernels; see train.py in this package for what is out of scope.
"""
import os
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torchvision.models as models

from . import train as _train2d
from .cnn import CudaForwardMixin

MODEL_PATH = "3D_sampling_cnn.pth"


class MapsDataset(_train2d.MapsDataset):
    """ThreeD_train.py:27-66: `images/<name>.npy` voxel map [x][y][z], min-max normalised, moved LAST AXIS FIRST by the
    pipeline's ToTensorV2 (so the network sees [z][x][y]); coords (1, 2, 3) = [[sx, sy, sz], [fx, fy, fz]] from the name
    `map_{m}_path{k}_sx.._sy.._sz.._fx.._fy.._fz...npy`."""

    def __getitem__(self, index: int):
        img_name = self._img_names[index]
        image = self._minmax(np.load(self._main_path / 'images' / img_name))
        mask_path = self._main_path / 'masks' / img_name
        mask = self._minmax(np.load(mask_path)) if mask_path.exists() else np.zeros(image.shape)
        parts = os.path.splitext(img_name)[0].split('_')
        coords = [[int(parts[k][2:]) for k in (3, 4, 5)], [int(parts[k][2:]) for k in (6, 7, 8)]]
        return (torch.from_numpy(np.ascontiguousarray(image.transpose(2, 0, 1))).float(),
                torch.from_numpy(np.ascontiguousarray(mask))[None, ...].float(),
                torch.tensor(coords, dtype=torch.float32)[None])


class MapsDataModule(_train2d.MapsDataModule):
    """ThreeD_train.py:69-103."""
    dataset_cls = MapsDataset

    def __init__(self, main_path=Path('/home/czarek/mgr/3D_data/train'), batch_size: int = 1, test_size=0.15, num_workers=16):
        super().__init__(main_path, batch_size, test_size, num_workers)


class ThreeD_UNet_cooler(CudaForwardMixin, nn.Module):
    _nrrt_dims = 3
    micro_batch = 4

    def __init__(self):
        super(ThreeD_UNet_cooler, self).__init__()
        self.loss = nn.BCEWithLogitsLoss()
        self.base_model = models.video.r3d_18(weights=None)
        self.base_model.stem[0] = nn.Sequential(
            nn.Conv3d(1, 32, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
            nn.Conv3d(32, 64, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
        )

        # Encoder
        self.conv1 = self.base_model.stem
        self.conv2 = self.base_model.layer1
        self.conv3 = self.base_model.layer2
        self.conv4 = self.base_model.layer3
        self.conv5 = self.base_model.layer4

        # Coordinate layers (Conv2d on the (1,2,3) coords tensor)
        self.conv_coords_512 = nn.Sequential(nn.Conv2d(256, 512, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True))
        self.conv_coords_256 = nn.Sequential(nn.Conv2d(128, 256, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True))
        self.conv_coords_128 = nn.Sequential(nn.Conv2d(64, 128, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True))
        self.conv_coords_64 = nn.Sequential(nn.Conv2d(1, 64, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True))

        # Decoder
        self.upconv6 = nn.ConvTranspose3d(1024, 512, kernel_size=2, stride=2)
        self.conv6 = nn.Sequential(
            nn.Conv3d(1024, 512, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
            nn.Conv3d(512, 256, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
        )
        self.upconv7 = nn.ConvTranspose3d(256, 128, kernel_size=2, stride=2)
        self.conv7 = nn.Sequential(
            nn.Conv3d(384, 256, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
            nn.Conv3d(256, 128, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
        )
        self.upconv8 = nn.ConvTranspose3d(128, 64, kernel_size=2, stride=2)
        self.conv8 = nn.Sequential(
            nn.Conv3d(192, 128, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
            nn.Conv3d(128, 64, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
        )
        self.final_conv = nn.Conv3d(64, 1, kernel_size=1)
