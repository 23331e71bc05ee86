"""CPU: pins oracle/cnn_oracle.py (torch fp32 functional restatement) and the drop-in model classes' parameter layout
against the reference model itself: same seed -> same random-init weights -> same logits as the golden run."""
import numpy as np
import pytest
import torch

from oracle import cnn_oracle
from tests import helpers as H


def _model(three_d):
    from mgr_sampling_cnn_b200 import ThreeD_train, train
    torch.manual_seed(0)
    return ThreeD_train.ThreeD_UNet_cooler() if three_d else train.UNet_cooler()


@pytest.mark.parametrize("three_d", [False, True])
def test_state_dict_layout_matches_reference(three_d):
    g = H.load("cnn3d_seed0" if three_d else "cnn2d_seed0")
    sd = _model(three_d).state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in g["shapes"]]
    assert len(sd) == (276 if three_d else 268)


@pytest.mark.parametrize("three_d", [False, True])
@pytest.mark.parametrize("bn", [False, True])
def test_oracle_forward_matches_reference_logits(three_d, bn):
    g = H.load(("cnn3d_seed0" if three_d else "cnn2d_seed0") + ("_bn" if bn else ""))
    model = _model(three_d).eval()
    if bn:
        cnn_oracle.randomize_bn(model, 0)
    with torch.no_grad():
        y = cnn_oracle.forward(model.state_dict(), torch.from_numpy(g["x"]), torch.from_numpy(g["coords"]), 3 if three_d else 2)
    ref = g["logits"]
    assert y.shape == ref.shape
    scale = np.abs(ref).max()
    assert np.abs(y.numpy() - ref).max() <= 1e-5 * max(scale, 1.0)  # same torch CPU kernels: usually bit-identical


@pytest.mark.parametrize("three_d", [False, True])
def test_oracle_forward_matches_reference_at_baseline_size(three_d):
    """512^2 / 80^3 (BASELINE configs): logits of the reference model fed by the reference's own MapsDataset."""
    g = H.load_full_cnn("cnn3d_80_seed0" if three_d else "cnn2d_512_seed0")
    # the dataset's view of the task file: {0,1} normalisation; 2D gray -> 3 equal channels, 3D ToTensorV2 = [z, x, y]
    free01 = (g["occ"] != 0).astype(np.float32)
    x_expected = free01.transpose(2, 0, 1)[None] if three_d else np.repeat(free01[None, None], 3, axis=1)
    assert np.array_equal(g["x"], x_expected)
    model = _model(three_d).eval()
    with torch.no_grad():
        y = cnn_oracle.forward(model.state_dict(), torch.from_numpy(g["x"]), torch.from_numpy(g["coords"]), 3 if three_d else 2)
    assert np.abs(y.numpy() - g["logits"]).max() <= 1e-5 * max(np.abs(g["logits"]).max(), 1.0)


def test_cpu_forward_is_refused():
    model = _model(False).eval()
    from mgr_sampling_cnn_b200 import _lib
    with pytest.raises(_lib.NrrtError):
        model(torch.zeros(1, 3, 64, 64), torch.zeros(1, 1, 2, 2))
