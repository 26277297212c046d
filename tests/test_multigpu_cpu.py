"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: the static split of
the tile rows, the gather of row blocks to rank 0 and the stats merge.  The
per-rank compute call is stood in by the CPU oracle here (tests may use it); the
GPU version of the same flow is tests/test_gpu_rmsds.py::test_parts_tile_the_triangle."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, F, N, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from loos_b200 import multigpu, part_rows
    from loos_b200.synth import synth_frames
    from oracle.oracle import Checker
    X = synth_frames(F, N, seed=11).astype(np.float64)       # replica of the frames on every rank
    full = Checker("port").rmsds_self(X, 2)
    rows = part_rows(F, rank, world)
    block = np.zeros((len(rows), F), np.float32)
    ssum, smax, cnt = 0.0, 0.0, 0
    for r, i in enumerate(rows):
        if i != multigpu.PAD:
            block[r, :i] = full[i, :i]                       # what loosb200_rmsds_self_part returns
            ssum += float(block[r, :i].astype(np.float64).sum()); cnt += int(i)
            smax = max(smax, float(block[r, :i].max()) if i else 0.0)
    M = multigpu.gather_matrix(rows, block, F, rank, world)
    stats = [None] * world
    dist.all_gather_object(stats, {"sum": ssum, "max": smax, "count": cnt})
    if rank == 0:
        merged = multigpu.merge_stats(stats)
        low = full[np.tril_indices(F, -1)].astype(np.float64)
        q.put((bool(np.array_equal(M, full)), merged["count"] == low.size,
               abs(merged["sum"] - low.sum()) < 1e-6, merged["max"] == low.max()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("F", [97, 160])
def test_two_rank_split_and_gather(F):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + F
    procs = [ctx.Process(target=_worker, args=(r, 2, port, F, 23, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(res), res


def test_split_is_a_partition_for_many_world_sizes():
    from loos_b200 import part_rows
    for F in (1, 39, 40, 41, 1000, 4321):
        for world in (1, 2, 3, 4, 8):
            seen = np.concatenate([part_rows(F, r, world) for r in range(world)])
            real = np.sort(seen[seen != 0xFFFFFFFF])
            assert np.array_equal(real, np.arange(F)), (F, world)
