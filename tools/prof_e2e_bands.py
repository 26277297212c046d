"""C4 host-in / host-out on one GPU for several band counts of the result copy (LOOSB200_BANDS)."""
import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
import loos_b200 as lb
from loos_b200.synth import synth_frames_torch
F, N = 100000, 1000
X = synth_frames_torch(F, N, device="cuda")
Xh = torch.empty((F, N, 3), dtype=torch.float32, pin_memory=True); Xh.copy_(X); torch.cuda.synchronize()
del X; torch.cuda.empty_cache()
out = lb.PinnedMatrix((F, F))
for bands in sys.argv[1:]:
    os.environ["LOOSB200_BANDS"] = bands
    ts = []
    for it in range(3):
        t0 = time.perf_counter()
        fs = lb.FrameSet(Xh.numpy()); t1 = time.perf_counter()
        lb.rmsds(fs, out=out.array, engine="tensor"); t2 = time.perf_counter()
        tm = fs.last_timing(); fs.close()
        ts.append((t1 - t0, t2 - t1, tm[0]))
    print("bands %3s: stage %.1f ms, call %s ms (kernels %.1f)" % (bands, 1e3 * ts[-1][0], " / ".join("%.1f" % (1e3 * t[1]) for t in ts), ts[-1][2]), flush=True)
