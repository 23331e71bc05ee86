"""Stub losses (never called on the hot path)."""


class DiceLoss(object):
    def __init__(self, *a, **k):
        pass


class SoftCrossEntropyLoss(object):
    def __init__(self, *a, **k):
        pass
