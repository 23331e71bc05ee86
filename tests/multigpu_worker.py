"""Worker of tests/test_multigpu_gpu.py (launched with torch.distributed.run, one process per GPU, NCCL): every rank plans its
shard of one global task list through GuidedPlanner and the records meet on rank 0 through sharding.gather_records.
With WORLD_SIZE unset it is the single-GPU run of the same list.  Writes the gathered table to argv[1]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mgr_sampling_cnn_b200 import ThreeD_train, engine, sharding, train, workload  # noqa: E402


def main():
    out_path, dim, n_tasks, interleaved = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), bool(int(sys.argv[4]))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = workload.make_2d(n_tasks, size=256) if dim == 2 else workload.make_3d(n_tasks, size=32)
    idx = sharding.shard_indices(n_tasks, rank, world, interleaved)
    maps_used, t2m = np.unique(wl.task_to_map[idx], return_inverse=True)
    torch.manual_seed(0)
    model = (train.UNet_cooler() if dim == 2 else ThreeD_train.ThreeD_UNet_cooler()).eval().to(dev)
    gp = engine.GuidedPlanner(model, dim, wl.shape, max_iterations=800)
    res = gp.plan(wl.occ_maps[maps_used], t2m.astype(np.int32), wl.starts[idx], wl.goals[idx], wl.coords[idx], seed=3,
                  task_ids=idx.astype(np.int32))
    rec = torch.stack([res.status.double(), res.iterations.double(), res.n_nodes.double(), res.path_n.double(),
                       res.path_len, res.best_dist], 1)
    table = sharding.gather_records(rec, torch.from_numpy(idx).to(dev), n_tasks, dst=0)
    if rank == 0:
        np.save(out_path, table.cpu().numpy())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
