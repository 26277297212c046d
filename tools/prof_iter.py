"""rmsd2ref default (iterative alignment) mode: wall time per iteration at scale."""
import sys, time, torch, numpy as np
sys.path.insert(0, "/root/repo")
import loos_b200 as lb
from loos_b200.synth import synth_frames_torch
F, N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000, 2000
X = synth_frames_torch(F, N, seed=20261021, device="cuda", chunk=4096)
fs = lb.FrameSet(X); del X; torch.cuda.empty_cache()
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = lb.rmsd2ref_iterative(fs, None, tol=1e-6, maxiter=1000)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("F=%d N=%d iterations %d  wall %.1f ms  -> %.2f ms per iteration, %.0f GB/s per pass-pair" % (
        F, N, r["iters"], 1e3 * (t1 - t0), 1e3 * (t1 - t0) / (r["iters"] + 2), 2 * F * N * 12 / ((t1 - t0) / (r["iters"] + 2)) / 1e9))
fs.close()
