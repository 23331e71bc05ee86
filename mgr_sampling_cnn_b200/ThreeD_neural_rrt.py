"""Drop-in for the reference's ThreeD_neural_rrt.py (3D CNN-guided RRT*; ThreeD_neural_rrt.py:54-68,233-265) on the
B200 engine.  heat_map: normalised + thresholded network output, stored as `heat_map - 250` (:68); neural samples are
redrawn until they hit a free voxel (:80-108) — bounded here (NRRT_SAMPLER_STUCK) where the reference would spin."""
import random  # noqa: F401
import sys
from pathlib import Path
from time import perf_counter

import numpy as np
import torch

from ._planner_base import Node, RRTStarBase, calculate_length, get_from_string  # noqa: F401
from .ThreeD_train import MapsDataModule, ThreeD_UNet_cooler  # noqa: F401

BASE_PATH = Path('/home/czarek/mgr/3D_eval_data/test/')
MODEL_PATH = "/home/czarek/mgr/models/3D_sampling_cnn_vol2_47.pth"
MAX_ITERATIONS = 5000
GOAL_THRESHOLD = 3.0


def get_blank_maps_list() -> list:
    return [str(image_path) for image_path in sorted((BASE_PATH / 'images').iterdir())]


def get_start_finish_coordinates(path: str) -> tuple:
    v = [int(get_from_string(path, a, b)) for a, b in (("_sx", "_sy"), ("_sy", "_sz"), ("_sz", "_fx"), ("_fx", "_fy"),
                                                       ("_fy", "_fz"), ("_fz", ".npy"))]
    return tuple(v[:3]), tuple(v[3:])


class RRTStar(RRTStarBase):
    _dim, _neural = 3, True
    _module = sys.modules[__name__]

    def __init__(self, occ_map, heat_map, start, goal, max_iterations: int, goal_threshold: float, neural_bias: float):
        self.map_height, self.map_width, self.map_depth = occ_map.shape
        self._init_common(occ_map, np.asarray(heat_map) - 250, start, goal, max_iterations, goal_threshold, neural_bias)

    def search_space_volume(self) -> float:
        return self.map_width * self.map_height * self.map_width

    def generate_neural_sample(self):
        return self._sample_via_kernel(True)


def generate_paths(model=None, index=None):
    """The reference's demo (ThreeD_neural_rrt.py:307-407): trained weights from MODEL_PATH, one random sample of the set
    under BASE_PATH, CNN forward, min-max normalisation + 0.96 threshold (:338-346), guided RRT* (neural_bias 0.75).  The
    planner map is the dataset tensor moved back to [x][y][z] (:330-331); the network output is used un-transposed (:341,
    commented out in the reference).  The model runs on the GPU; the matplotlib figures are not drawn.
    Returns {"neural": (path, iterations, seconds), "start", "finish", "heat_map"}."""
    if model is None:
        model = ThreeD_UNet_cooler()
        model.load_state_dict(torch.load(MODEL_PATH))
    model = model.eval().to("cuda")

    data_module = MapsDataModule(main_path=BASE_PATH)
    data_module.setup("test")
    ds = data_module.train_dataset
    if index is None:
        index = int(torch.randint(len(ds), (1,)).item())
    image, mask, coords = (t[None] for t in ds[index])
    start = tuple(coords.data.tolist()[0][0][0])
    finish = tuple(coords.data.tolist()[0][0][1])

    occ_map = np.array(image[0].detach().cpu().numpy()).transpose((1, 2, 0))

    timer_neural_start = perf_counter()
    with torch.no_grad():
        output = model(image.to("cuda"), coords.to("cuda"))
    visualized_output = np.array(output[0, 0].detach().cpu().numpy())
    visualized_output = ((visualized_output - visualized_output.min()) / (
        visualized_output.max() - visualized_output.min())) * 1.0
    visualized_output_masked = visualized_output.copy()
    visualized_output_masked[~(visualized_output > 0.96)] = 0

    rrt_neural = RRTStar(occ_map=occ_map, heat_map=visualized_output_masked, start=start, goal=finish,
                         max_iterations=MAX_ITERATIONS, goal_threshold=GOAL_THRESHOLD, neural_bias=0.75)
    path, iterations = rrt_neural.rrt_star()
    timer_neural_stop = perf_counter()
    if path:
        print(path)
    else:
        print("COULDN'T FIND A PATH FOR THIS EXAMPLE:", start, finish)
    print(f'Calculation time of neural RRT*: {timer_neural_stop - timer_neural_start}')
    print(f'Start X:{start[0]}, Y:{start[1]}, Z:{start[2]}')
    print(f'Finish X:{finish[0]}, Y:{finish[1]}, Z:{finish[2]}')
    return {"neural": (path, iterations, timer_neural_stop - timer_neural_start), "start": start, "finish": finish,
            "heat_map": visualized_output_masked}


if __name__ == '__main__':
    generate_paths()
