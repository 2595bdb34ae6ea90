"""Multi-process host logic on CPU (gloo, world_size 2): env-axis sharding and the per-episode score all-gather."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ntrtsim_b200.neuroevo import shard_bounds, allgather_scores  # noqa: E402


def test_shard_bounds_partition_the_env_axis():
    for n in (0, 1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, n_items, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard_bounds(n_items, world, rank)
    # fake local scores: score of global member i is (i, -i)
    local = np.stack([np.arange(b, e, dtype=np.float64), -np.arange(b, e, dtype=np.float64)], axis=1)
    full = allgather_scores(local, n_items)
    out[rank] = full.tolist()
    dist.destroy_process_group()


def test_score_allgather_gloo_world2_ragged():
    world, n_items = 2, 7
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(world, n_items, port, out), nprocs=world, join=True)
    expect = np.stack([np.arange(n_items, dtype=np.float64), -np.arange(n_items, dtype=np.float64)], axis=1)
    for r in range(world):
        assert np.array_equal(np.array(out[r]), expect)
