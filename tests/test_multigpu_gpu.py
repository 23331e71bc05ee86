"""GPU, 2 ranks over NCCL: the sharded result equals the single-GPU result bit-for-bit -- CNN logits -> CDF ->
draws -> plans do not depend on which rank or which micro-batch a task lands in.  Needs two visible GPUs (skipped on a
one-GPU box; run with `gpurun --gpus 2`)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "multigpu_worker.py")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("dim,interleaved", [(2, False), (2, True), (3, False)])
def test_two_gpu_sharded_result_equals_single_gpu(dim, interleaved, tmp_path, cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n_tasks = 27 if dim == 2 else 23  # ragged shards, maps split across ranks
    one, two = str(tmp_path / "one.npy"), str(tmp_path / "two.npy")
    env = {k: v for k, v in os.environ.items() if k not in ("WORLD_SIZE", "RANK", "LOCAL_RANK")}
    args = [str(dim), str(n_tasks), str(int(interleaved))]
    subprocess.check_call([sys.executable, WORKER, one] + args, env=env, timeout=900)
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                           "127.0.0.1", "--master-port", str(_free_port()), WORKER, two] + args, env=env, timeout=900)
    a, b = np.load(one), np.load(two)
    assert a.shape == (n_tasks, 6) and np.array_equal(a, b, equal_nan=True)
    assert (a[:, 1] >= 1).all() and (a[:, 0] == 0).any()
