"""Worker of tests/test_multigpu_gpu.py::test_two_gpu_data_parallel_training (torch.distributed.run, one process per GPU,
NCCL): two data-parallel training steps of UNetTrainer, each rank on its own synthetic batch.  Every rank writes its parameter
buffer after the steps and, for the first step, its LOCAL gradient buffer (before the all-reduce) to argv[1] + rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mgr_sampling_cnn_b200 import train, trainer  # noqa: E402


def batch(rank, B=1, S=512):
    rng = np.random.default_rng(50 + rank)
    x = torch.from_numpy((rng.random((B, 1, S, S)) > 0.2).astype(np.float32)).repeat(1, 3, 1, 1)
    y = torch.from_numpy((rng.random((B, 1, S, S)) > 0.9).astype(np.float32))
    coords = torch.from_numpy(rng.integers(0, S, (B, 1, 2, 2)).astype(np.float32))
    return x, y, coords


def main():
    out = sys.argv[1]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    tr = trainer.UNetTrainer(train.UNet_cooler(), 1, 512, 512, dev, loss_scale=64.0 * 512 * 512)   # fixed scale: the steps are comparable
    x, y, coords = batch(rank)
    tr.forward(x.to(dev), coords.to(dev))
    tr.set_targets(target=y.to(dev))
    tr.loss_and_grad(True)
    tr.backward()
    torch.cuda.synchronize()
    local_grads = tr.grads.clone()
    tr.optimizer_step()
    loss2 = tr.training_step((x, y, coords))
    torch.cuda.synchronize()
    np.savez(out + str(rank) + ".npz", params=tr.params.cpu().numpy(), grads=local_grads.cpu().numpy(), loss=float(loss2.item()),
             world=world, dp=int(tr.dp))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
