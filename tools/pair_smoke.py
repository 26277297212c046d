"""Small correctness run of both launch shapes of k_pairs_tensor against the fp64 engine."""
import os, sys, torch
sys.path.insert(0, "/root/repo")
import loos_b200 as lb
from loos_b200.synth import synth_frames_torch
for F, N in ((97, 200), (1000, 500), (3000, 1000)):
    X = synth_frames_torch(F, N, device="cuda")
    fs = lb.FrameSet(X)
    ref = torch.empty((F, F), dtype=torch.float32, device="cuda")
    lb.rmsds(fs, out=ref, engine="fp64")
    for cl in ("1", "2"):
        os.environ["LOOSB200_CLUSTER"] = cl
        out = torch.full((F, F), -7.0, dtype=torch.float32, device="cuda")
        st = lb.rmsds(fs, out=out, engine="tensor")
        torch.cuda.synchronize()
        print("F=%d N=%d cluster=%s max|d|=%.3e kernel %.3f ms stats %s" % (F, N, cl, (out - ref).abs().max().item(), fs.last_timing()[0], st), flush=True)
    fs.close()
