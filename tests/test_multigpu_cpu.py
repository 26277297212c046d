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


def _iter_worker(rank, world, port, F, Na, q):
    """Sharded iterative alignment: frames dealt round-robin to two ranks; the per-shard compute
    (superpose + sum, per-frame rmsd) is the CPU oracle's kabsch here, the loop, the all-reduce of
    the running average and the convergence rule are loos_b200.multigpu's."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from loos_b200 import multigpu
    from loos_b200.synth import synth_frames
    from oracle.oracle import Checker
    chk = Checker("port")
    X = synth_frames(F, Na + 7, seed=21).astype(np.float64)
    A, R = X[:, :Na], X                                   # align selection = first Na atoms, rmsd selection = all
    mine = np.arange(rank, F, world)

    def xform(frame, target):                            # 4x4 W mapping frame onto target (alignment::kabsch)
        return np.asarray(chk.kabsch(frame.reshape(-1), target.reshape(-1))).reshape(4, 4)

    def accumulate(fa, target, fr=None):
        src = A if fr is None else R
        out = np.zeros((src.shape[1], 3))
        for f in mine:
            W = xform(A[f], target)
            out += src[f] @ W[:3, :3].T + W[:3, 3]
        return out

    def rmsd_to(fa, target, fr, ref):
        d, xf = np.zeros(len(mine)), np.zeros((len(mine), 4, 4))
        for k, f in enumerate(mine):
            W = xform(A[f], target)
            y = R[f] @ W[:3, :3].T + W[:3, 3]
            d[k] = np.sqrt(((y - ref) ** 2).sum(axis=1).mean())
            xf[k] = W
        return d, xf

    res = multigpu.rmsd2ref_iterative_sharded("align shard", "rmsd shard", A[0], F, tol=1e-6, maxiter=1000,
                                              accumulate=accumulate, rmsd_to=rmsd_to)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, res["rmsd"], res["avg"], res["iters"]))
    if rank == 0:
        want_d, _, want_avg, want_it, _ = chk.rmsd2ref_iterative(A, R, 1e-6, 1000)
        d = np.zeros(F)
        for idx, dd, _, _ in gathered:
            d[idx] = dd
        q.put((int(res["iters"]) == int(want_it), all(g[3] == res["iters"] for g in gathered),
               float(np.abs(d - want_d).max()) < 1e-9, float(np.abs(res["avg"] - want_avg).max()) < 1e-9,
               all(np.array_equal(g[2], res["avg"]) for g in gathered)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_iterative_alignment():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_iter_worker, args=(r, 2, 29777, 41, 30, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(res), res


def _align_worker(rank, world, port, F1, F2, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from loos_b200 import multigpu
    from loos_b200.synth import synth_frames
    from oracle import oracle
    ck = oracle.Checker("port")
    X = synth_frames(F1, 40, seed=21)
    Y = synth_frames(F2, 40, seed=22)
    sa, sr = np.arange(0, 24, 2), np.arange(10, 40)
    M = multigpu.rmsds_align_sharded(X, sa, sr, Y, sa, sr, rank, world,
                                     compute=lambda a1, r1, a2, r2: oracle.rmsds_align(ck, a1, r1, a2, r2))
    if rank == 0:
        whole = oracle.rmsds_align(ck, X[:, sa], X[:, sr], Y[:, sa], Y[:, sr])
        q.put((M.shape == (F1, F2), float(np.abs(M - whole).max())))
    else:
        assert M is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("F1,F2", [(9, 7), (1, 5)])
def test_two_rank_rmsds_align_stripes(F1, F2):
    """rmsds-align rows striped over two ranks (the second rank's stripe is empty when F1 = 1), gathered on
    rank 0: equals the unsharded matrix.  The per-rank compute is stood in by the oracle's loop."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + F1
    procs = [ctx.Process(target=_align_worker, args=(r, 2, port, F1, F2, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok_shape, err = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok_shape and err < 1e-9


def test_stripes_partition_the_rows():
    from loos_b200.multigpu import stripe
    for n in (0, 1, 7, 8, 100001):
        for world in (1, 2, 3, 8):
            parts = [stripe(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[k][1] == parts[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _cuts(F, nparts, self_, weights=None):
    import ctypes as C
    from loos_b200 import capi
    out = (C.c_uint32 * (nparts + 1))()
    w = None if weights is None else (C.c_double * nparts)(*weights)
    capi.check(capi.lib().loosb200_share_cuts(F, nparts, 1 if self_ else 0, w, out))
    return list(out)


def test_contiguous_shares_of_the_one_process_multi_gpu_calls():
    """loosb200_share_cuts (host logic of multi.cu): contiguous ranges of 40-frame row tiles that cover
    [0, T) exactly once; equal work without weights (row tile t of the triangle costs t + 1 tiles),
    work proportional to the weights with them -- the host-output path weights a device by what its
    PCIe link delivers (10.8 ... 22.6 GB/s on the 8-GPU boxes of this pool)."""
    for F in (1, 40, 41, 1000, 100000):
        T = (F + 39) // 40
        for n in (1, 2, 3, 8):
            for self_ in (True, False):
                c = _cuts(F, n, self_)
                assert c[0] == 0 and c[-1] == T and all(a <= b for a, b in zip(c, c[1:])), (F, n, c)
    work = lambda c, self_: [(b * (b + 1) - a * (a + 1)) / 2 if self_ else b - a for a, b in zip(c, c[1:])]
    c = _cuts(100000, 8, True)
    w = work(c, True)
    assert max(w) / min(w) < 1.02                                    # 2500 row tiles: equal to a row tile
    rates = [10.8, 10.8, 11.0, 11.0, 18.3, 18.2, 19.2, 22.6]
    c = _cuts(100000, 8, True, rates)
    w = work(c, True)
    tot = sum(w)
    for share, r in zip(w, rates):
        assert abs(share / tot - r / sum(rates)) < 0.005, (c, w)
    c = _cuts(100000, 4, False, [1, 1, 2, 4])
    assert [b - a for a, b in zip(c, c[1:])] == pytest.approx([312.5, 312.5, 625, 1250], abs=1.01)
    assert _cuts(1000, 3, True, [0.0, 0.0, 0.0]) == _cuts(1000, 3, True)   # useless weights: equal shares
    assert _cuts(1000, 3, True, [0.0, 1.0, 0.0])[1:3] == [0, 25]           # a part may be empty
