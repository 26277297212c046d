"""Sharded iterative alignment on N GPUs (torchrun, NCCL): frames dealt round-robin to the ranks,
one all-reduce of the running average per iteration; rank 0 compares with the single-GPU result.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_iter_nccl.py"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, "/root/repo")
import loos_b200 as lb
from loos_b200 import multigpu
from loos_b200.synth import synth_frames
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
F, N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000, 1500
X = synth_frames(F, N, seed=9)
sel = np.arange(0, N, 3, dtype=np.uint32)
mine = np.arange(rank, F, world)
with lb.FrameSet(X[mine], sel=sel, device=local) as fa, lb.FrameSet(X[mine], device=local) as fr:
    for rep in range(2):
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        res = multigpu.rmsd2ref_iterative_sharded(fa, fr, X[0][sel], F, tol=1e-6, maxiter=1000)
        torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
gathered = [None] * world
dist.all_gather_object(gathered, (mine, res["rmsd"]))
if rank == 0:
    with lb.FrameSet(X, sel=sel, device=0) as fa, lb.FrameSet(X, device=0) as fr:
        t0 = time.perf_counter(); one = lb.rmsd2ref_iterative(fa, fr, tol=1e-6, maxiter=1000); t1 = time.perf_counter() - t0
    d = np.zeros(F)
    for idx, dd in gathered:
        d[idx] = dd
    print("world %d: iters %d (single GPU %d), max |d avg| %.2e, max |d rmsd| %.2e, %.1f ms sharded vs %.1f ms single" % (
        world, res["iters"], one["iters"], np.abs(res["avg"] - one["avg"]).max(), np.abs(d - one["rmsd"]).max(), 1e3 * dt, 1e3 * t1))
    assert res["iters"] == one["iters"] and np.abs(res["avg"] - one["avg"]).max() < 1e-8 and np.abs(d - one["rmsd"]).max() < 2e-6
dist.barrier()
dist.destroy_process_group()
