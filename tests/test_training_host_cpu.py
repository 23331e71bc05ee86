"""CPU: the training / mask-maker host layer refuses to run without the CUDA path (no silent fallback), and its host-side
helpers (path packing, shell values, file-name contract) behave like the reference's."""
import numpy as np
import pytest
import torch

from mgr_sampling_cnn_b200 import ThreeD_enlarge_path as ep
from mgr_sampling_cnn_b200 import _lib
from mgr_sampling_cnn_b200 import generate_paths as gp


def test_pack_paths_and_layer_values():
    arr, n = gp.pack_paths([[(1, 2), (3, 4), (5, 6)], [], [(7, 8)]], 2)
    assert arr.shape == (3, 3, 2) and n.tolist() == [3, 0, 1] and arr[0, 2].tolist() == [5, 6] and not arr[1].any()
    assert ep.layer_values().tolist() == [206, 157, 108, 59, 10]          # ThreeD_enlarge_path.py:44-50 with LAYERS = 5
    assert ep.layer_values(3).tolist() == [172, 89, 6]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour of a box without a GPU")
def test_no_cpu_fallback():
    from mgr_sampling_cnn_b200 import train, trainer
    with pytest.raises(_lib.NrrtError):
        trainer.UNetTrainer(train.UNet_cooler(), 1, 512, 512)
    with pytest.raises(_lib.NrrtError):
        gp.path_masks(np.full((1, 64, 64), 255, np.uint8), [0], [[(1, 1)]])
    with pytest.raises(_lib.NrrtError):
        ep.enlarge_paths(np.zeros((1, 8, 8, 8), np.uint8), [0], np.zeros((1, 8, 8, 8), np.uint8))


def test_trainer_rejects_sizes_the_kernels_do_not_cover():
    from mgr_sampling_cnn_b200 import train, trainer
    if not torch.cuda.is_available():
        pytest.skip("the size check sits behind the CUDA check")
    with pytest.raises(ValueError):
        trainer.UNetTrainer(train.UNet_cooler(), 1, 200, 256)


def test_plateau_scheduler_matches_torch():
    """trainer.PlateauScheduler == torch.optim.lr_scheduler.ReduceLROnPlateau(factor=0.5, patience=2) (train.py:230-234)."""
    from mgr_sampling_cnn_b200.trainer import PlateauScheduler
    rng = np.random.default_rng(1)
    for trial in range(5):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.Adam([p], lr=1e-3)
        ref = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=0.5, patience=2)
        mine = PlateauScheduler(1e-3)
        v = 1.0
        for epoch in range(40):
            v = v * (0.97 if rng.random() < 0.4 else 1.0 + 0.02 * rng.random())
            ref.step(v)
            assert abs(mine.step(v) - opt.param_groups[0]["lr"]) <= 1e-15, (trial, epoch)
        d = mine.state_dict()
        again = PlateauScheduler(123.0)
        again.load_state_dict(d)
        assert again.step(v * 2) == mine.step(v * 2)


def test_kernel_layout_round_trip():
    """trainer.to_kernel_layout / from_kernel_layout: the packing of the reference's parameters into the trainer's flat buffers."""
    from mgr_sampling_cnn_b200.trainer import from_kernel_layout, to_kernel_layout
    g = torch.Generator().manual_seed(0)
    w = torch.randn(8, 5, 3, 3, generator=g)                                    # Conv2d weight, channels padded 5 -> 64
    k = to_kernel_layout(w, "w", cin_pad=64)
    assert k.shape == (8, 9, 64) and float(k[:, :, 5:].abs().max()) == 0.0
    assert k[3, 1 * 3 + 2, 4] == w[3, 4, 1, 2]                                  # [cout][tap = ky * 3 + kx][cin]
    assert torch.equal(from_kernel_layout(k.reshape(-1), "w", w.shape, cin_pad=64), w)
    wt = torch.randn(6, 4, 2, 2, generator=g)                                   # ConvTranspose2d weight [cin][cout][2][2]
    kt = to_kernel_layout(wt, "w", convT=True)
    assert kt.shape == (16, 6) and kt[(1 * 2 + 0) * 4 + 3, 5] == wt[5, 3, 1, 0]  # row = (a * 2 + b) * cout + co
    assert torch.equal(from_kernel_layout(kt.reshape(-1), "w", wt.shape, convT=True), wt)
    b = torch.randn(4, generator=g)
    assert torch.equal(to_kernel_layout(b, "b", convT=True), b.repeat(4))
    assert torch.equal(from_kernel_layout(b.repeat(4), "b", b.shape, convT=True), b)
    cw = torch.randn(7, 3, 3, 3, generator=g)                                   # conv_coords weight -> [tap][cin][cout]
    kc = to_kernel_layout(cw, "cw")
    assert kc.shape == (3, 3, 3, 7) and kc[2, 0, 1, 6] == cw[6, 1, 2, 0]
    assert torch.equal(from_kernel_layout(kc.reshape(-1), "cw", cw.shape), cw)
    v = torch.randn(5, generator=g)
    assert torch.equal(from_kernel_layout(to_kernel_layout(v, "v").reshape(-1), "v", v.shape), v)
