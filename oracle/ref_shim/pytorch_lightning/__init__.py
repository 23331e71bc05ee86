"""Stub of the pytorch_lightning names the reference touches at import/forward time (oracle harness only)."""
import torch.nn as nn


class LightningModule(nn.Module):
    def log(self, *a, **k):
        return None

    def log_dict(self, *a, **k):
        return None


class LightningDataModule(object):
    def __init__(self, *a, **k):
        pass


class Trainer(object):
    def __init__(self, *a, **k):
        raise RuntimeError("training is out of scope for the oracle harness")


class _Anything(object):
    def __init__(self, *a, **k):
        pass


class callbacks(object):
    ModelCheckpoint = _Anything
    EarlyStopping = _Anything
    LearningRateMonitor = _Anything


class loggers(object):
    NeptuneLogger = _Anything
