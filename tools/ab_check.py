"""Correctness of an A/B build (AB_LIB=path): tensor engine against the fp64 engine at awkward sizes."""
import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
from loos_b200 import capi
if os.environ.get("AB_LIB"):
    capi.LIB_PATH = os.environ["AB_LIB"]
    import ctypes as C
    L = C.CDLL(capi.LIB_PATH)
    for n in list(capi.SIGNATURES):
        if not hasattr(L, n):
            del capi.SIGNATURES[n]
import loos_b200 as lb
from loos_b200.synth import synth_frames
worst = 0.0
for F, N in ((203, 333), (90, 1000), (64, 4200), (50, 64), (45, 50000)):
    X = synth_frames(F, N, seed=F + N)
    with lb.FrameSet(X) as fs:
        Mt, _ = lb.rmsds(fs, engine="tensor")
        Mf, _ = lb.rmsds(fs, engine="fp64")
    d = float(np.abs(Mt.astype(np.float64) - Mf).max())
    worst = max(worst, d)
    print("F=%d N=%d max |tensor - fp64| = %.2e" % (F, N, d), flush=True)
print("OK" if worst < 2e-5 else "MISMATCH")
