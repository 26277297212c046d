"""Kernel time of k_pairs_tensor for the TMEM hand-over point / cluster variants (env-selected)."""
import os, sys, torch
sys.path.insert(0, "/root/repo")
import loos_b200 as lb
from loos_b200.synth import synth_frames_torch
for F, N in ((100000, 1000), (20000, 3000)):
    X = synth_frames_torch(F, N, device="cuda")
    out = torch.empty((F, F), dtype=torch.float32, device="cuda")
    fs = lb.FrameSet(X)
    lb.rmsds(fs, out=out, engine="tensor")
    for cl in ("2", "1"):
        for ra in ("1", "0"):
            os.environ["LOOSB200_CLUSTER"] = cl
            os.environ["LOOSB200_RELEASE_AT"] = ra
            ts = []
            for _ in range(2):
                lb.rmsds(fs, out=out, engine="tensor")
                torch.cuda.synchronize()
                ts.append(fs.last_timing()[0])
            print("F=%d N=%d cluster=%s release_at=%s kernel ms %s" % (F, N, cl, ra, ["%.1f" % t for t in ts]), flush=True)
    fs.close()
    del out, X
    torch.cuda.empty_cache()
