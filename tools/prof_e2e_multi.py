"""One process, G GPUs, C4 host-in / host-out (loosb200_stage_frames_multi + loosb200_rmsds_self_multi):
the two delivery modes of multi.cu (canonical rectangles shipped as column blocks / triangle shares shipped as
strips), shares sized by each device's measured PCIe rate or equal."""
import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
import loos_b200 as lb
from loos_b200 import capi
from loos_b200.synth import synth_frames_torch
G = int(sys.argv[1]) if len(sys.argv) > 1 else capi.device_count()
F, N = int(os.environ.get("E2E_FRAMES", 100000)), 1000
X = synth_frames_torch(F, N, device="cuda:0")
Xh = torch.empty((F, N, 3), dtype=torch.float32, pin_memory=True); Xh.copy_(X); torch.cuda.synchronize()
del X; torch.cuda.empty_cache()
out = lb.PinnedMatrix((F, F))
ref = None
for mode in ("strips", "strips, equal shares", "strips"):
    os.environ["LOOSB200_MULTI_MODE"] = mode.split(",")[0]
    os.environ["LOOSB200_EQUAL_SHARES"] = "1" if "equal" in mode else "0"
    ts = []
    for it in range(3):
        t0 = time.perf_counter()
        reps = lb.stage_multi(Xh.numpy(), list(range(G)))
        t1 = time.perf_counter()
        lb.rmsds_multi(reps, out=out.array, engine="tensor", want_stats=False)
        t2 = time.perf_counter()
        for r in reps: r.close()
        ts.append((t1 - t0, t2 - t1))
    st, call = min(t[0] for t in ts[1:]), min(t[1] for t in ts[1:])
    rng = np.random.default_rng(1); ii, jj = rng.integers(0, F, 4000), rng.integers(0, F, 4000)
    samp = out.array[ii, jj].copy()
    same = True if ref is None else bool(np.array_equal(samp, ref))
    ref = samp if ref is None else ref
    print("%d GPUs, %s: stage %.1f ms, call %.1f ms = %.1f GB/s of matrix to the host; sampled entries identical: %s" % (
        G, mode, 1e3 * st, 1e3 * call, F * F * 4 / call / 1e9, same), flush=True)
