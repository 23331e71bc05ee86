"""Drop-in for the reference's neural_rrt.py (2D CNN-guided RRT*): `RRTStar(occ_map, heat_map, start, goal,
max_iterations, goal_threshold, neural_bias)` (neural_rrt.py:50-64,215-249) on the B200 engine.
occ_map: (H, W, 3) float image, converted to gray like the reference (:59); heat_map: (1, H, W, 1) clamp(-3, 1) network
output, stored as `heat_map[0] + 10` (:64).  The sampler CDF is built once per instance instead of once per draw."""
import random  # noqa: F401
import sys
from pathlib import Path
from time import perf_counter

import numpy as np
import torch

from ._planner_base import Node, RRTStarBase, calculate_length, get_from_string  # noqa: F401
from .train import MapsDataModule, UNet_cooler  # noqa: F401  (the reference module re-exports both)

BASE_PATH = Path('/home/czarek/mgr/eval_data/')
MODEL_PATH = "/home/czarek/mgr/models/sampling_cnn_vol3_32.pth"
MAX_ITERATIONS = 5000
GOAL_THRESHOLD = 5.0


def get_blank_maps_list() -> list:
    return [str(image_path) for image_path in sorted((BASE_PATH / 'images').iterdir())]


def get_start_finish_coordinates(path: str) -> tuple:
    x_start = int(get_from_string(path, "_sx", "_sy"))
    y_start = int(get_from_string(path, "_sy", "_fx"))
    x_finish = int(get_from_string(path, "_fx", "_fy"))
    y_finish = int(get_from_string(path, "_fy", ".png"))
    return (y_start, x_start), (y_finish, x_finish)


class RRTStar(RRTStarBase):
    _dim, _neural = 2, True
    _module = sys.modules[__name__]

    def __init__(self, occ_map, heat_map, start, goal, max_iterations: int, goal_threshold: float, neural_bias: float):
        gray = np.dot(np.asarray(occ_map)[..., :3], [0.2989, 0.5870, 0.1140])
        self.map_height, self.map_width = gray.shape
        self._init_common(gray, np.asarray(heat_map)[0] + 10, start, goal, max_iterations, goal_threshold, neural_bias)

    def search_space_volume(self) -> float:
        return self.map_width * self.map_height

    def generate_neural_sample(self):
        return self._sample_via_kernel(True)


def generate_paths(model=None, index=None):
    """The reference's demo (neural_rrt.py:319-395): trained weights from MODEL_PATH, one random sample of the evaluation
    set under BASE_PATH, CNN forward + clamp(-3, 1), guided RRT* (neural_bias 0.75), then the same class with
    neural_bias = 0 as the plain run; both timed like the reference.  The model runs on the GPU (the reference keeps it on
    the CPU); plotting is a no-op.  `model` / `index` let a caller pass a loaded model / pick the sample.
    Returns {"neural": (path, iterations, seconds), "plain": (path, iterations, seconds), "start", "finish"}."""
    if model is None:
        model = UNet_cooler()
        model.load_state_dict(torch.load(MODEL_PATH))
    model = model.eval().to("cuda")

    data_module = MapsDataModule(main_path=BASE_PATH)
    data_module.setup("test")
    ds = data_module.train_dataset
    if index is None:  # torch.utils.data.RandomSampler(...) + next(iter(dataloader)), batch_size 1
        index = int(torch.randint(len(ds), (1,)).item())
    image, mask, coords = (t[None] for t in ds[index])

    occ_map = image.data.detach().cpu().numpy().transpose((0, 2, 3, 1))[0]
    ideal_mask = mask.data.detach().cpu().numpy().transpose((0, 2, 3, 1))[0]

    timer_neural_start = perf_counter()
    with torch.no_grad():
        output = model(image.to("cuda"), coords.to("cuda"))
        clipped = torch.clamp(output, min=-3, max=1)
    clipped = clipped.detach().cpu().numpy().transpose((0, 2, 3, 1))

    c = coords.data.tolist()[0][0]
    start, finish = (c[0][1], c[0][0]), (c[1][1], c[1][0])
    rrt_neural = RRTStar(occ_map=occ_map, heat_map=clipped, start=start, goal=finish, max_iterations=MAX_ITERATIONS,
                         goal_threshold=GOAL_THRESHOLD, neural_bias=0.75)
    path_n, it_n = rrt_neural.rrt_star()
    timer_neural_stop = perf_counter()
    if path_n:
        rrt_neural.visualize_path(path_n, ideal_mask)
    else:
        print("COULDN'T FIND A PATH FOR THIS EXAMPLE:", start, finish)

    timer_rrt_start = perf_counter()
    rrt = RRTStar(occ_map=occ_map, heat_map=clipped, start=start, goal=finish, max_iterations=MAX_ITERATIONS,
                  goal_threshold=GOAL_THRESHOLD, neural_bias=0.0)
    path_p, it_p = rrt.rrt_star()
    timer_rrt_stop = perf_counter()
    if path_p:
        rrt.visualize_path(path_p, ideal_mask)
    else:
        print("COULDN'T FIND A PATH FOR THIS EXAMPLE:", start, finish)

    print(f'Calculation time of neural RRT*: {timer_neural_stop - timer_neural_start}')
    print(f'Calculation time of RRT*: {timer_rrt_stop - timer_rrt_start}')
    return {"neural": (path_n, it_n, timer_neural_stop - timer_neural_start),
            "plain": (path_p, it_p, timer_rrt_stop - timer_rrt_start), "start": start, "finish": finish}


if __name__ == '__main__':
    generate_paths()
