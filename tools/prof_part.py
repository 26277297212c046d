"""Where a rank's step goes when the C4 triangle is split 8 ways (one GPU plays rank R of 8):
stage / quantise / rmsds_part / close, host wall clock around synchronised calls."""
import os, sys, time, torch
sys.path.insert(0, "/root/repo")
import loos_b200 as lb
from loos_b200.api import rmsds_part
from loos_b200.synth import synth_frames_torch
F, N = 100000, 1000
R, W = int(os.environ.get("PART", 0)), int(os.environ.get("WORLD", 8))
X = synth_frames_torch(F, N, device="cuda")
rows = lb.part_rows(F, R, W)
out = torch.empty((len(rows), F), dtype=torch.float32, device="cuda")
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fs = lb.FrameSet(X); t1 = time.perf_counter()
    _, _, st = rmsds_part(fs, R, W, out=out, engine="tensor"); t2 = time.perf_counter()
    tm = fs.last_timing()
    fs.close(); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("part %d/%d: stage %.2f ms  rmsds_part %.2f ms (kernel %.2f, other on stream %.2f)  close %.2f ms  total %.2f" % (R, W, 1e3*(t1-t0), 1e3*(t2-t1), tm[0], tm[1], 1e3*(t3-t2), 1e3*(t3-t0)))
