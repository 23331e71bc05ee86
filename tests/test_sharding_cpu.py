"""CPU, world_size 2 over gloo: shard -> plan -> gather equals the single-rank result bit-for-bit (the planner stand-in
on CPU is the oracle, which is test infrastructure; the sharding/gather code under test is the product's)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mgr_sampling_cnn_b200 import sharding


def _records_for(ids):
    from oracle import planner_oracle as po
    occ = np.full((64, 64), 255, np.uint8)
    occ[20:44, 30:34] = 0
    free = po.occ_free_2d(occ)
    rad = po.radius_table(2, (64, 64), 150)
    out = []
    for t in ids:
        smp, _, _, st = po.draw_samples(2, (64, 64), free, None, 0.0, 7, int(t), 150)
        r = po.plan(2, (64, 64), free, rad, smp.astype(np.float64), (5 + int(t) % 10, 5), (58, 58 - int(t) % 7), 150, 5.0, pow_mode=1)
        out.append([r["status"], r["iterations"], r["n_nodes"], r["path_length"]])
    return torch.tensor(out, dtype=torch.float64).reshape(len(ids), 4)


def _worker(rank, world, port, n_tasks, interleaved, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = sharding.shard_indices(n_tasks, rank, world, interleaved)
    out = sharding.gather_records(_records_for(ids), torch.from_numpy(ids), n_tasks, dst=0)
    if rank == 0:
        q.put(out.numpy())
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_indices_partition():
    for n, w in ((10, 2), (11, 4), (4096, 8), (3, 8)):
        for inter in (True, False):
            parts = [sharding.shard_indices(n, r, w, inter) for r in range(w)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))


def test_two_rank_gather_equals_single_rank():
    n_tasks = 11  # ragged: 6 + 5
    single = sharding.gather_records(_records_for(np.arange(n_tasks)), torch.arange(n_tasks), n_tasks).numpy()
    for interleaved in (True, False):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_tasks, interleaved, q)) for r in range(2)]
        for p in procs:
            p.start()
        got = q.get(timeout=120)
        for p in procs:
            p.join(60)
            assert p.exitcode == 0
        assert np.array_equal(got, single)
