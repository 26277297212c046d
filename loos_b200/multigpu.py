"""Multi-GPU work splitting for the all-to-all path: one process per GPU, static
split of the 40-frame row tiles of the lower triangle, no collective while computing.

    rows, block, stats = local_part(fs, rank, world)      # this rank's packed rows
    M = gather_matrix(rows, block, F, rank, world)        # optional: whole matrix on rank 0

Works with any torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests;
the compute call itself needs a GPU)."""
import numpy as np

from . import api

PAD = 0xFFFFFFFF


def local_part(fs, rank, world, out=None, noout=False, engine="auto"):
    return api.rmsds_part(fs, rank, world, out=out, noout=noout, engine=engine)


def assemble(parts, F):
    """parts: iterable of (rows, block) from every rank -> full symmetric (F, F) float32
    matrix (Fortran order, like the reference's RealMatrix), diagonal 0."""
    M = np.zeros((F, F), dtype=np.float32, order="F")
    for rows, block in parts:
        rows = np.asarray(rows)
        block = np.asarray(block)
        for r, i in enumerate(rows):
            if i == PAD or i == 0:
                continue
            M[i, :i] = block[r, :i]
    M += M.T
    return M


def merge_stats(stats_list):
    tot = sum(s["sum"] for s in stats_list)
    cnt = sum(s["count"] for s in stats_list)
    return {"max": max(s["max"] for s in stats_list), "sum": tot, "count": cnt, "avg": tot / cnt if cnt else 0.0}


def gather_matrix(rows, block, F, rank, world, group=None):
    """Gather every rank's row blocks on rank 0 and assemble the matrix there
    (the only communication of the path: 'final gather of row blocks')."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return assemble([(rows, block)], F)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t_rows = torch.as_tensor(np.asarray(rows).astype(np.int64), device=dev)
    t_blk = torch.as_tensor(np.asarray(block), device=dev) if not hasattr(block, "is_cuda") else block.to(dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([t_rows.numel()], dtype=torch.int64, device=dev), group=group)
    sizes = [int(s.item()) for s in sizes]
    if rank == 0:
        parts = [(t_rows.cpu().numpy(), t_blk.cpu().numpy())]
        for src in range(1, world):
            r = torch.empty(sizes[src], dtype=torch.int64, device=dev)
            b = torch.empty((sizes[src], F), dtype=torch.float32, device=dev)
            dist.recv(r, src=src, group=group)
            dist.recv(b, src=src, group=group)
            parts.append((r.cpu().numpy(), b.cpu().numpy()))
        return assemble([(np.where(r < 0, PAD, r).astype(np.uint32), b) for r, b in parts], F)
    dist.send(t_rows, dst=0, group=group)
    dist.send(t_blk.contiguous(), dst=0, group=group)
    return None


# ---------------------------------------------------------------------------------------------
# Iterative alignment over a trajectory whose frames are sharded across ranks
# (rmsd2ref's default mode, iterativeAlignment src/alignment.cpp:355-404 + averageStructure
# src/ensembles.cpp:91-121).  This is the one place of the whole RMSD path with a real exchange
# step: the running average is a sum over ALL frames, so every iteration all-reduces 3*N doubles.

def _all_reduce_sum(x, group=None):
    """Sum a small float64 numpy array over the ranks (NCCL needs it on the device, gloo on the host)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def rmsd2ref_iterative_sharded(fs_align, fs_rmsd, first_frame, n_frames_total, tol=1e-6, maxiter=1000, group=None,
                               accumulate=None, rmsd_to=None):
    """Each rank holds a contiguous or strided shard of the frames as FrameSets (align selection,
    rmsd selection or None for "same").  `first_frame` is the align selection of GLOBAL frame 0
    ((N_align, 3), identical on every rank: the reference primes the loop with it).

    Returns dict(rmsd, xforms, avg, iters, rms) with `rmsd`/`xforms` for the LOCAL frames and the
    rest identical on every rank.  `accumulate` / `rmsd_to` default to the GPU entry points
    (api.align_accumulate, api.rmsd2ref); the CPU tests inject restatements to cover the loop and
    the collective without a GPU."""
    accumulate = accumulate or api.align_accumulate
    rmsd_to = rmsd_to or (lambda fa, tgt, fr, ref: api.rmsd2ref(fa, tgt, fr, ref, want_xforms=True))
    first = np.asarray(first_frame, dtype=np.float64).reshape(-1, 3)
    target = first - first.mean(axis=0)          # target.centerAtOrigin(), alignment.cpp:372-373
    it, rms, used = 0, 0.0, target
    while True:
        used = target                             # the transforms of the last pass refer to this target
        avg = _all_reduce_sum(accumulate(fs_align, target), group) / float(n_frames_total)
        rms = float(np.sqrt(((avg - target) ** 2).sum(axis=1).mean()))   # avg.rmsd(target), AG_numerical.cpp:301-318
        target = avg
        it += 1
        if not (rms > tol and it <= maxiter):     # alignment.cpp:399
            break
    avg_r = _all_reduce_sum(accumulate(fs_align, used, fs_rmsd), group) / float(n_frames_total)
    rmsd, xf = rmsd_to(fs_align, used, fs_rmsd, avg_r)
    return {"rmsd": rmsd, "xforms": xf, "avg": avg_r, "iters": it, "rms": rms}


# ---------------------------------------------------------------------------------------------
# rmsds-align over several GPUs: rows (frames of trajectory 1) are independent and cost the same,
# so every rank takes a contiguous stripe of them against all of trajectory 2.  No collective while
# computing; the stripes are gathered on rank 0 at the end.

def stripe(n, rank, world):
    """[lo, hi) of rank's share of n rows: contiguous, sizes differ by at most one."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rmsds_align_sharded(X1, sel_align1, sel_rmsd1, X2, sel_align2, sel_rmsd2, rank, world, group=None, compute=None):
    """X1: (F1, natoms1, 3) frames of trajectory 1, X2: (F2, natoms2, 3) of trajectory 2 (every rank
    holds both; only its stripe of X1 is staged).  Returns the (F1, F2) float64 matrix on rank 0, None
    elsewhere.  `compute(A1, R1, A2, R2)` (arrays of selected coordinates) defaults to staging on this
    rank's GPU and api.rmsds_align; the CPU tests inject the oracle's loop."""
    lo, hi = stripe(len(X1), rank, world)

    def on_gpu(A1, R1, A2, R2):
        with api.FrameSet(A1) as a1, api.FrameSet(R1) as r1, api.FrameSet(A2) as a2, api.FrameSet(R2) as r2:
            return api.rmsds_align(a1, r1, a2, r2)

    compute = compute or on_gpu
    X1 = np.asarray(X1)
    X2 = np.asarray(X2)
    block = np.zeros((0, len(X2)))
    if hi > lo:
        block = np.ascontiguousarray(compute(X1[lo:hi][:, sel_align1], X1[lo:hi][:, sel_rmsd1], X2[:, sel_align2], X2[:, sel_rmsd2]),
                                     dtype=np.float64)
    if world == 1:
        return block
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    if rank != 0:
        if hi > lo:
            dist.send(torch.as_tensor(block, device=dev), dst=0, group=group)
        return None
    M = np.zeros((len(X1), len(X2)))
    M[lo:hi] = block
    for src in range(1, world):
        slo, shi = stripe(len(X1), src, world)
        if shi > slo:
            t = torch.empty((shi - slo, len(X2)), dtype=torch.float64, device=dev)
            dist.recv(t, src=src, group=group)
            M[slo:shi] = t.cpu().numpy()
    return M
