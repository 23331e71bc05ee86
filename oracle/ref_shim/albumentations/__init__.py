"""Stub of albumentations.Compose (oracle harness only)."""


class Compose(object):
    def __init__(self, transforms=None, *a, **k):
        self.transforms = list(transforms or [])

    def __call__(self, **data):
        for t in self.transforms:
            data = t(**data)
        return data
