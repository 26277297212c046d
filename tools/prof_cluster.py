"""One launch of k_pairs_tensor per launch shape (LOOSB200_CLUSTER=2: CTA pairs, 1: independent CTAs),
for an ncu metrics pass: ncu --metrics ... python tools/prof_cluster.py [F]"""
import os, sys, torch
sys.path.insert(0, "/root/repo")
import loos_b200 as lb
from loos_b200.synth import synth_frames_torch
F = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
X = synth_frames_torch(F, 1000, device="cuda")
out = torch.empty((F, F), dtype=torch.float32, device="cuda")
fs = lb.FrameSet(X)
for cl in ("2", "1"):
    os.environ["LOOSB200_CLUSTER"] = cl
    lb.rmsds(fs, out=out, engine="tensor")
    torch.cuda.synchronize()
    print("cluster", cl, "kernel ms", fs.last_timing()[0])
fs.close()
