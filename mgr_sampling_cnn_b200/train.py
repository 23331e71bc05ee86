"""Drop-in for the model class of the reference's train.py: `UNet_cooler` (train.py:109-203).

Same zero-argument constructor, same sub-module names and therefore the same 268-key `state_dict` (including the unused
ResNet stem/fc under `base_model.*`), same `forward(x, coords)` contract: x (B,3,H,W) fp32, coords (B,1,2,2) fp32 ->
logits (B,1,H,W) fp32.  forward() runs the B200 kernels (cnn.UNetPlan); `training_step` / `configure_optimizers` and the fit loop
live in trainer.UNetTrainer (`test_training()` below drives it like train.py:273-315); `MapsDataset` / `MapsDataModule` read the
reference's task files.  The reference asks torchvision for `pretrained=True` ImageNet weights, which need
a download; here the encoder is random-init (BASELINE.json) and trained weights arrive through `load_state_dict`.
"""
import os
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torchvision.models as models
from torch.utils.data import DataLoader, Dataset

from .cnn import CudaForwardMixin

MODEL_PATH = "sampling_cnn_vol3.pth"


class MapsDataset(Dataset):
    """The reader of the task files (train.py:27-69): `images/<name>.png` (map, read as RGB and min-max normalised to
    [0, 1]), `masks/<name>.png`, start / finish parsed from the file name `map_{m}_path{k}_sx{..}_sy{..}_fx{..}_fy{..}.png`.
    Returns (image (3, H, W) fp32, mask (1, H, W) fp32, coords (1, 2, 2) fp32 = [[sx, sy], [fx, fy]]): the tensors
    `UNet_cooler.forward` takes.  Plain numpy (no albumentations); `transforms` is accepted and ignored beyond ToTensorV2's
    HWC -> CHW move, which is all the reference's pipelines contain (train.py:80-85)."""

    def __init__(self, main_path, img_names, transforms=None):
        super().__init__()
        self._main_path = Path(main_path)
        self._img_names = list(img_names)
        self._transforms = transforms

    def __len__(self):
        return len(self._img_names)

    @staticmethod
    def _minmax(a):
        a = np.asarray(a)
        with np.errstate(invalid="ignore", divide="ignore"):
            return (a - a.min()) / (a.max() - a.min())

    def __getitem__(self, index: int):
        from PIL import Image
        img_name = self._img_names[index]
        image = self._minmax(np.asarray(Image.open(self._main_path / 'images' / img_name).convert('RGB')))
        mask_path = self._main_path / 'masks' / img_name
        mask = self._minmax(np.asarray(Image.open(mask_path))) if mask_path.exists() else np.zeros(image.shape[:2])
        parts = os.path.splitext(img_name)[0].split('_')
        coords = [[int(parts[3][2:]), int(parts[4][2:])], [int(parts[5][2:]), int(parts[6][2:])]]
        return (torch.from_numpy(np.ascontiguousarray(image.transpose(2, 0, 1))).float(),
                torch.from_numpy(np.ascontiguousarray(mask))[None, ...].float(),
                torch.tensor(coords, dtype=torch.float32)[None])


class MapsDataModule(object):
    """train.py:72-106 without Lightning: the same 85 / 15 split (sklearn, random_state=42) over the sorted file names."""
    dataset_cls = MapsDataset

    def __init__(self, main_path=Path('/home/czarek/mgr/data/train'), batch_size: int = 3, test_size=0.15, num_workers=16):
        self._main_path = Path(main_path)
        self._batch_size = batch_size
        self._test_size = test_size
        self._num_workers = num_workers
        self.augmentations = None
        self.transforms = None

    def setup(self, stage: str):
        from sklearn.model_selection import train_test_split
        names = [p.name for p in sorted((self._main_path / 'images').iterdir())]
        train_names, val_names = train_test_split(names, test_size=self._test_size, random_state=42)
        self.train_dataset = self.dataset_cls(self._main_path, train_names, self.augmentations)
        self.val_dataset = self.dataset_cls(self._main_path, val_names, self.transforms)

    def train_dataloader(self):
        return DataLoader(self.train_dataset, batch_size=self._batch_size, num_workers=self._num_workers)

    def val_dataloader(self):
        return DataLoader(self.val_dataset, batch_size=self._batch_size, num_workers=self._num_workers)

    def test_dataloader(self):
        return DataLoader(self.val_dataset, batch_size=self._batch_size)


class UNet_cooler(CudaForwardMixin, nn.Module):
    _nrrt_dims = 2

    def __init__(self):
        super(UNet_cooler, self).__init__()
        self.loss = nn.BCEWithLogitsLoss()
        self.base_model = models.resnet18(weights=None)

        # Encoder
        self.conv1 = nn.Sequential(
            nn.Conv2d(3, 32, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(32, 64, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
        )
        self.conv2 = self.base_model.layer1
        self.conv3 = self.base_model.layer2
        self.conv4 = self.base_model.layer3
        self.conv5 = self.base_model.layer4

        # Coordinate layers (declared 512 -> 64, applied 64 -> 512)
        self.conv_coords_512 = nn.Sequential(nn.Conv2d(256, 512, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True))
        self.conv_coords_256 = nn.Sequential(nn.Conv2d(128, 256, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True))
        self.conv_coords_128 = nn.Sequential(nn.Conv2d(64, 128, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True))
        self.conv_coords_64 = nn.Sequential(nn.Conv2d(1, 64, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True))

        # Decoder
        self.upconv6 = nn.ConvTranspose2d(1024, 512, kernel_size=2, stride=2)
        self.conv6 = nn.Sequential(
            nn.Conv2d(1024, 512, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(512, 256, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
        )
        self.upconv7 = nn.ConvTranspose2d(256, 128, kernel_size=2, stride=2)
        self.conv7 = nn.Sequential(
            nn.Conv2d(384, 256, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(256, 128, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
        )
        self.upconv8 = nn.ConvTranspose2d(128, 64, kernel_size=2, stride=2)
        self.conv8 = nn.Sequential(
            nn.Conv2d(192, 128, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(128, 64, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
        )
        self.final_conv = nn.Conv2d(64, 1, kernel_size=1)


def test_training(data_module=None, max_epochs: int = 100, device="cuda:0", log=print):
    """train.py:273-315 without Lightning / Neptune: fit UNet_cooler on the data module (batch_size 3), early stopping on val_loss
    (patience 10), ReduceLROnPlateau, the best weights saved to MODEL_PATH.  Returns (model, history)."""
    from . import trainer
    data_module = data_module or MapsDataModule()
    data_module.setup("fit")
    model = UNet_cooler()
    sample = data_module.train_dataset[0][0]
    tr = trainer.UNetTrainer(model, data_module._batch_size, int(sample.shape[-2]), int(sample.shape[-1]), device)
    history = tr.fit(data_module.train_dataloader(), data_module.val_dataloader(), max_epochs=max_epochs, patience=10, log=log)
    torch.save(model.state_dict(), MODEL_PATH)
    return model, history
