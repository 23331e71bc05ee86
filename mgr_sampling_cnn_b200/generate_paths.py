"""Drop-in for the reference's generate_paths.py mask maker (generate_paths.py:118-169) on the device.

`path_masks` is the batched engine call (libnrrt.so: nrrt_path_masks_2d); `visualize_path` keeps the reference's
signature and file convention (reads nothing but its arguments, writes the mask next to the task file under `paths/`).
The A* search itself (generate_paths.py:11-74) is engine.grid_path_lengths (lengths / task filter); a path to draw is
supplied by the caller.  There is no CPU fallback."""
import ctypes as C
import glob
import os

import numpy as np
import torch

from . import _lib
from .engine import _dev, _ptr, _require_cuda, _stream_ptr, check

MAPS_DIRECTORY = 'maps/start_finish/*.png'   # the reference hard-codes its author's directory (generate_paths.py:8): set it before use
CIRCLE_RADIUS = 10       # generate_paths.py:131
BLUR_KSIZE = 33          # generate_paths.py:134
BLUR_SIGMA = 3.0         # generate_paths.py:134 passes cv2.BORDER_WRAP (= 3) in the sigmaX position
FREE_ABOVE = 240         # generate_paths.py:148


def pack_paths(paths, dim: int):
    """list of [(y, x), ...] / [(x, y, z), ...] point lists -> (int32 [B, cap, dim], int32 [B]) host arrays."""
    B = len(paths)
    cap = max(1, max((len(p) for p in paths), default=1))
    arr = np.zeros((B, cap, dim), np.int32)
    n = np.zeros(B, np.int32)
    for b, p in enumerate(paths):
        p = np.asarray(p, np.int32).reshape(-1, dim)
        arr[b, :len(p)] = p
        n[b] = len(p)
    return arr, n


def path_masks(occ_maps, task_to_map, paths, path_n=None, device=None, radius: int = CIRCLE_RADIUS, ksize: int = BLUR_KSIZE,
               sigma: float = BLUR_SIGMA, free_above: int = FREE_ABOVE) -> torch.Tensor:
    """Training masks of B tasks: uint8 [B, H, W] on the device, bit-identical to what visualize_path() saves.

    occ_maps uint8 [n_maps, H, W] (255 = free); task_to_map [B]; paths int32 [B, cap, 2] (y, x) + path_n [B], or a list of
    point lists."""
    device = _require_cuda(device)
    if path_n is None:
        paths, path_n = pack_paths(paths, 2)
    occ = _dev(occ_maps, torch.uint8, device)
    if occ.dim() == 2:
        occ = occ[None]
    t2m = _dev(task_to_map, torch.int32, device).reshape(-1)
    pts = _dev(paths, torch.int32, device)
    pn = _dev(path_n, torch.int32, device).reshape(-1)
    B, H, W = t2m.shape[0], occ.shape[1], occ.shape[2]
    if pts.dim() != 3 or pts.shape[0] != B or pts.shape[2] != 2 or pn.shape[0] != B:
        raise ValueError("paths must be [B, cap, 2] and path_n [B]")
    masks = torch.empty((B, H, W), dtype=torch.uint8, device=device)
    _lib.count_launch()
    check(_lib.load().nrrt_path_masks_2d(_ptr(occ), _ptr(t2m), _ptr(pts), _ptr(pn), pts.shape[1], B, H, W, radius, ksize,
                                         C.c_double(sigma), free_above, _ptr(masks), _stream_ptr(device)))
    return masks


def visualize_path(occ_map: np.ndarray, path: list, directory: str):
    """generate_paths.visualize_path: writes the mask of one task to directory.replace('start_finish', 'paths')
    (generate_paths.py:159,162).  The preview image (`paths_with_points`, the lines tagged VISUAL) is not produced."""
    mask = path_masks(occ_map[None], [0], [path])[0].cpu().numpy()
    out = directory.replace('start_finish', 'paths')
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    if out.endswith(".npy"):
        np.save(out, mask)
    else:
        import cv2  # only the PNG writer; the arithmetic is the kernel's
        cv2.imwrite(out, mask)
    return mask


def astar(image: np.ndarray, start: tuple, finish: tuple):
    """generate_paths.astar (generate_paths.py:11-74) for one task on the device: a shortest 8-connected path [(y, x), ...] from
    start to finish (no corner cutting), or None.  Same length and step kinds as the reference's path; among equal-cost paths the
    cells may differ (include/nrrt.h: nrrt_grid_paths_ex)."""
    from . import engine
    device = _require_cuda(None)
    image = np.asarray(image)
    bits = engine.pack_occupancy(image[None], 2, device)
    _, _, paths, n = engine.grid_paths(bits, [0], [list(start)], [list(finish)], image.shape, path_cap=image.shape[0] * image.shape[1])
    n = int(n[0])
    if n <= 0:
        return None
    return [tuple(int(v) for v in p) for p in paths[0, :n].cpu().numpy()]


def get_blank_maps_list() -> list:
    return sorted(glob.glob(MAPS_DIRECTORY))


def get_from_string(path: str, start: str, finish: str) -> str:
    return path[path.find(start) + 3:path.find(finish)]


def get_start_finish_coordinates(path: str) -> tuple:
    x_start = int(get_from_string(path, "_sx", "_sy"))
    y_start = int(get_from_string(path, "_sy", "_fx"))
    x_finish = int(get_from_string(path, "_fx", "_fy"))
    y_finish = int(get_from_string(path, "_fy", ".png"))
    return (y_start, x_start), (y_finish, x_finish)


def generate_paths(batch: int = 256):
    """generate_paths.generate_paths (generate_paths.py:100-115): for every task file, the A* path and its mask under paths/; task
    files without a path are deleted.  Files of one map shape are processed `batch` at a time through the batched kernels."""
    import cv2  # PNG reader / writer only
    from . import engine
    from .ThreeD_generate_paths import _by_shape
    device = _require_cuda(None)
    for chunk, maps in _by_shape(get_blank_maps_list(), lambda f: cv2.imread(f, 0), batch):
        sf = [get_start_finish_coordinates(f) for f in chunk]
        starts, goals = np.array([a for a, _ in sf], np.int32), np.array([b for _, b in sf], np.int32)
        t2m = np.arange(len(chunk), dtype=np.int32)
        bits = engine.pack_occupancy(maps, 2, device)
        _, _, paths, n = engine.grid_paths(bits, t2m, starts, goals, maps.shape[1:], path_cap=4 * max(maps.shape[1:]))
        n_host = n.cpu().numpy()
        if (n_host < 0).any():
            raise _lib.NrrtError("a path is longer than the path buffer")
        masks = path_masks(maps, t2m, paths, n.clamp(min=0), device).cpu().numpy()
        for f, k, m in zip(chunk, n_host, masks):
            if k == 0:
                print("COULDN'T FIND A PATH FOR THIS EXAMPLE:", f)
                os.remove(f)
                continue
            out = f.replace('start_finish', 'paths')
            os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
            cv2.imwrite(out, m)
