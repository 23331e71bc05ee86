"""Drop-in for the reference's ThreeD_enlarge_path.py (ThreeD_enlarge_path.py:25-65) on the device (libnrrt.so:
nrrt_paths_to_volume, nrrt_enlarge_paths_3d).  There is no CPU fallback."""
import glob

import numpy as np
import torch

from . import _lib
from .engine import _dev, _ptr, _require_cuda, _stream_ptr, check

LAYERS = 5
MAPS_DIRECTORY = '3D_maps/paths/*.npy'


def layer_values(layers: int = LAYERS) -> np.ndarray:
    """`path_value` of shell i as the reference's uint8 array stores it (ThreeD_enlarge_path.py:44-50, Python floats)."""
    vals, path_value = [], 255
    for _ in range(layers):
        path_value -= 255 / layers - 2
        vals.append(np.array([path_value]).astype(np.uint8)[0])
    return np.array(vals, np.uint8)


def path_volumes(paths, path_n, shape, device=None) -> torch.Tensor:
    """ThreeD_generate_paths.save_and_visualize_path:142-150 for B paths: uint8 [B, X, Y, Z], 255 on the path voxels."""
    device = _require_cuda(device)
    pts = _dev(paths, torch.int32, device)
    pn = _dev(path_n, torch.int32, device).reshape(-1)
    B = pn.shape[0]
    vols = torch.empty((B,) + tuple(int(s) for s in shape), dtype=torch.uint8, device=device)
    _lib.count_launch()
    check(_lib.load().nrrt_paths_to_volume(_ptr(pts), _ptr(pn), pts.shape[1], B, int(shape[0]), int(shape[1]), int(shape[2]),
                                           _ptr(vols), _stream_ptr(device)))
    return vols


def enlarge_paths(occ_maps, task_to_map, path_vols, layers: int = LAYERS, device=None) -> torch.Tensor:
    """enlarge_the_path for B path volumes (uint8 [B, X, Y, Z], 255 on the path; occ_maps uint8 [n_maps, X, Y, Z], 255 =
    obstacle): a new tensor, bit-identical to the arrays the reference saves back."""
    device = _require_cuda(device)
    occ = _dev(occ_maps, torch.uint8, device)
    if occ.dim() == 3:
        occ = occ[None]
    vols = _dev(path_vols, torch.uint8, device)
    t2m = _dev(task_to_map, torch.int32, device).reshape(-1)
    B = t2m.shape[0]
    if vols.dim() != 4 or vols.shape[0] != B or vols.shape[1:] != occ.shape[1:]:
        raise ValueError("path_vols must be [B, X, Y, Z] with the maps' shape")
    out = torch.empty_like(vols)
    vals = layer_values(layers)
    _lib.count_launch()
    check(_lib.load().nrrt_enlarge_paths_3d(_ptr(occ), _ptr(t2m), _ptr(vols), B, vols.shape[1], vols.shape[2], vols.shape[3],
                                            layers, vals.ctypes.data, _ptr(out), _stream_ptr(device)))
    return out


def enlarge_the_path(occ_path, path_path):
    """ThreeD_enlarge_path.enlarge_the_path: load both arrays, enlarge, save over the path file."""
    occ_map = np.load(occ_path)
    path = np.load(path_path)
    new_path = enlarge_paths(occ_map[None], [0], path[None])[0].cpu().numpy()
    np.save(path_path, new_path)


def get_blank_maps_list() -> list:
    return sorted(glob.glob(MAPS_DIRECTORY))


def generate_paths():
    for path_path in get_blank_maps_list():
        enlarge_the_path(path_path.replace('paths', 'start_finish'), path_path)
