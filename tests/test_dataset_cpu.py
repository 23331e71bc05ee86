"""CPU: the drop-in MapsDataset classes read the reference's task files exactly like the reference's own (train.py:27-69,
ThreeD_train.py:27-66): the golden `x` / `coords` are the tensors the reference's MapsDataset produced from the same file."""
import numpy as np

from tests import helpers as H


def test_maps_dataset_2d_matches_reference_reader(tmp_path):
    from PIL import Image
    from mgr_sampling_cnn_b200 import train
    g = H.load_full_cnn("cnn2d_512_seed0")
    name = str(g["file_name"])
    (tmp_path / "images").mkdir()
    (tmp_path / "masks").mkdir()
    Image.fromarray(g["occ"]).save(tmp_path / "images" / name)
    Image.fromarray((np.random.default_rng(0).random(g["occ"].shape) * 255).astype(np.uint8)).save(tmp_path / "masks" / name)
    ds = train.MapsDataset(tmp_path, [name], None)
    image, mask, coords = ds[0]
    assert np.array_equal(image.numpy()[None], g["x"]) and image.dtype.is_floating_point
    assert np.array_equal(coords.numpy()[None], g["coords"]) and mask.shape == (1, 512, 512)
    dm = train.MapsDataModule(main_path=tmp_path, test_size=0.5)
    for k in range(3):
        Image.fromarray(g["occ"]).save(tmp_path / "images" / name.replace("path0", f"path{k + 1}"))
    dm.setup("test")
    assert len(dm.train_dataset) + len(dm.val_dataset) == 4


def test_maps_dataset_3d_matches_reference_reader(tmp_path):
    from mgr_sampling_cnn_b200 import ThreeD_train
    g = H.load_full_cnn("cnn3d_80_seed0")
    name = str(g["file_name"])
    (tmp_path / "images").mkdir()
    (tmp_path / "masks").mkdir()
    np.save(tmp_path / "images" / name, g["occ"])
    np.save(tmp_path / "masks" / name, np.random.default_rng(0).random(g["occ"].shape))
    image, mask, coords = ThreeD_train.MapsDataset(tmp_path, [name], None)[0]
    assert np.array_equal(image.numpy()[None], g["x"])          # last axis first: [z][x][y]
    assert np.array_equal(coords.numpy()[None], g["coords"]) and mask.shape == (1, 80, 80, 80)
