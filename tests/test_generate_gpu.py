"""GPU: on-device workload generation (csrc/generate.cu) against the reference's own generators: the maps and
task files the UNMODIFIED reference produced for given seeds (tests/golden/maps_tasks.npz), and, at batch size, the host
workload builder (workload.make_2d / make_3d: the repo's drop-ins under CPython's `random` + scipy's labelling)."""
import numpy as np
import pytest
import torch

from mgr_sampling_cnn_b200 import workload
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim", [2, 3])
def test_device_maps_and_tasks_equal_the_reference_run(dim, cuda_device):
    """random.seed(1000 / 1001) -> create_map, random.seed(2000 / 2001) -> 10 x draw_start_and_finish, as recorded from
    the reference (oracle/gen_golden.py::gen_maps): same map bytes, same file names."""
    g = H.load("maps_tasks")
    for i in range(2):
        ms, ts = (int(v) for v in g[f"seeds{dim}d_{i}"])
        wl = workload.make_on_device(dim, 10, first_map=0, filter_unreachable=False, device=cuda_device, map_seed=ms, task_seed=ts)
        assert np.array_equal(wl.occ_maps[0].cpu().numpy(), g[f"occ{dim}d_{i}"])
        names = workload.task_file_names(wl)
        ref = [str(n).split("/")[-1] for n in g[f"names{dim}d_{i}"]]
        assert names == ref, (names[:2], ref[:2])


@pytest.mark.parametrize("dim,n_tasks,first", [(2, 403, 0), (2, 57, 1234), (3, 128, 0), (3, 45, 77)])
def test_device_workload_equals_host_workload(dim, n_tasks, first, cuda_device):
    """Same maps, tasks, coords and redraw count as the host builder, reachability filter included (truncated last map)."""
    host = (workload.make_2d if dim == 2 else workload.make_3d)(n_tasks, first_map=first)
    dev = workload.make_on_device(dim, n_tasks, first_map=first, device=cuda_device)
    d = dev.to_host()
    assert np.array_equal(d.occ_maps, host.occ_maps)
    assert np.array_equal(d.task_to_map, host.task_to_map)
    assert np.array_equal(d.starts, host.starts) and np.array_equal(d.goals, host.goals)
    assert np.array_equal(d.coords, host.coords) and d.coords.dtype == np.float32
    assert d.redrawn == host.redrawn
    if dim == 2 and n_tasks > 400:
        assert host.redrawn > 0  # the filter did reject candidates in this range of maps


def test_small_maps_and_the_name_parsers_round_trip(cuda_device):
    """Other map sizes (the 2D placement is still drawn from 0..512, generate_maps.py:35-36) and the wire format read back
    by the drop-in parsers."""
    from mgr_sampling_cnn_b200 import ThreeD_neural_rrt, neural_rrt
    for dim, size, mod in ((2, 128, neural_rrt), (3, 32, ThreeD_neural_rrt)):
        host = (workload.make_2d if dim == 2 else workload.make_3d)(20, first_map=3, size=size)
        dev = workload.make_on_device(dim, 20, first_map=3, size=size, device=cuda_device)
        assert np.array_equal(dev.occ_maps.cpu().numpy(), host.occ_maps) and np.array_equal(dev.starts.cpu().numpy(), host.starts)
        for b, name in enumerate(workload.task_file_names(dev, first_map=3, directory="/data/start_finish/")):
            s, f = mod.get_start_finish_coordinates(name)
            assert s == tuple(host.starts[b]) and f == tuple(host.goals[b])
            assert name.startswith(f"/data/start_finish/map_{3 + host.task_to_map[b]}_path{b % 10}_")
