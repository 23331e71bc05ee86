"""Drop-in for the reference's ThreeD_generate_paths.py (ThreeD_generate_paths.py:14-152) on the device: the 26-connected A* and the
path volumes it saves (the input of ThreeD_enlarge_path).  Plotting (matplotlib / mayavi) is not reproduced.  No CPU fallback."""
import glob
import os

import numpy as np

from . import ThreeD_enlarge_path as _ep
from . import _lib
from .engine import _require_cuda

MAPS_DIRECTORY = '3D_maps/start_finish/*.npy'   # the reference hard-codes its author's directory (ThreeD_generate_paths.py:11)


def astar(image: np.ndarray, start: tuple, finish: tuple):
    """ThreeD_generate_paths.astar for one task: a shortest path [(x, y, z), ...] (moves of cost 1 / sqrt 2 / sqrt 3 whose three
    axis-projected cells are free; obstacles are the 255 voxels) or None."""
    from . import engine
    device = _require_cuda(None)
    image = np.asarray(image)
    bits = engine.pack_occupancy(image[None], 3, device)
    _, _, paths, n = engine.grid_paths(bits, [0], [list(start)], [list(finish)], image.shape, path_cap=int(np.prod(image.shape)))
    n = int(n[0])
    if n <= 0:
        return None
    return [tuple(int(v) for v in p) for p in paths[0, :n].cpu().numpy()]


def get_blank_maps_list() -> list:
    return sorted(glob.glob(MAPS_DIRECTORY))


def get_from_string(path: str, start: str, finish: str) -> str:
    return path[path.find(start) + 3:path.find(finish)]


def get_start_finish_coordinates(path: str) -> tuple:
    v = [int(get_from_string(path, a, b)) for a, b in (("_sx", "_sy"), ("_sy", "_sz"), ("_sz", "_fx"), ("_fx", "_fy"), ("_fy", "_fz"),
                                                        ("_fz", ".npy"))]
    return tuple(v[:3]), tuple(v[3:])


def save_and_visualize_path(occ_map: np.ndarray, path: list, directory: str, visualize: bool = False, save: bool = True):
    """ThreeD_generate_paths.save_and_visualize_path:141-150: zeros with 255 on the path voxels, saved under paths/."""
    if not save:
        return None
    vol = _ep.path_volumes(np.asarray(path, np.int32)[None], [len(path)], occ_map.shape)[0].cpu().numpy().astype(occ_map.dtype)
    out = directory.replace('start_finish', 'paths')
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    np.save(out, vol)
    return vol


def _by_shape(files, read, batch):
    """Task files grouped by map shape, at most `batch` per group: [(file names, stacked maps), ...]."""
    groups = {}
    for f in files:
        m = read(f)
        groups.setdefault(m.shape, []).append((f, m))
    out = []
    for items in groups.values():
        for i in range(0, len(items), batch):
            out.append(([f for f, _ in items[i:i + batch]], np.stack([m for _, m in items[i:i + batch]])))
    return out


def generate_paths(batch: int = 64):
    """ThreeD_generate_paths.generate_paths:125-138: the path volume of every task file; task files without a path are deleted."""
    from . import engine
    device = _require_cuda(None)
    for chunk, maps in _by_shape(get_blank_maps_list(), np.load, batch):
        sf = [get_start_finish_coordinates(f) for f in chunk]
        starts, goals = np.array([a for a, _ in sf], np.int32), np.array([b for _, b in sf], np.int32)
        t2m = np.arange(len(chunk), dtype=np.int32)
        bits = engine.pack_occupancy(maps, 3, device)
        _, _, paths, n = engine.grid_paths(bits, t2m, starts, goals, maps.shape[1:], path_cap=8 * max(maps.shape[1:]))
        n_host = n.cpu().numpy()
        if (n_host < 0).any():
            raise _lib.NrrtError("a path is longer than the path buffer")
        vols = _ep.path_volumes(paths, n.clamp(min=0), maps.shape[1:], device).cpu().numpy()
        for f, k, v in zip(chunk, n_host, vols):
            if k == 0:
                print("COULDN'T FIND A PATH FOR THIS EXAMPLE:", f)
                os.remove(f)
                continue
            out = f.replace('start_finish', 'paths')
            os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
            np.save(out, v.astype(maps.dtype))
