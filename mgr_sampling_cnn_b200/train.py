"""Drop-in for the model class of the reference's train.py: `UNet_cooler` (train.py:109-203).

Same zero-argument constructor, same sub-module names and therefore the same 268-key `state_dict` (including the unused
ResNet stem/fc under `base_model.*`), same `forward(x, coords)` contract: x (B,3,H,W) fp32, coords (B,1,2,2) fp32 ->
logits (B,1,H,W) fp32.  forward() runs the B200 kernels (cnn.UNetPlan); training, the dataset classes and Lightning
hooks are out of scope.  The reference asks torchvision for `pretrained=True` ImageNet weights, which need
a download; here the encoder is random-init (BASELINE.json) and trained weights arrive through `load_state_dict`.
"""
import torch.nn as nn
import torchvision.models as models

from .cnn import CudaForwardMixin

MODEL_PATH = "sampling_cnn_vol3.pth"


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
