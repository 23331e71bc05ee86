"""Stub of albumentations.pytorch.ToTensorV2: HWC->CHW for image, pass-through for mask (oracle harness only)."""
import numpy as np
import torch


class ToTensorV2(object):
    def __init__(self, *a, **k):
        pass

    def __call__(self, **data):
        out = dict(data)
        if "image" in out:
            img = np.asarray(out["image"])
            if img.ndim == 2:
                img = img[:, :, None]
            out["image"] = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)))
        if "mask" in out:
            out["mask"] = torch.from_numpy(np.ascontiguousarray(np.asarray(out["mask"])))
        return out
