import os
import sys, time, torch
sys.path.insert(0, "/root/repo")
from loos_b200 import capi as _capi
if os.environ.get('AB_LIB'):
    _capi.LIB_PATH = os.environ['AB_LIB']   # A/B runs against another build of the library
    import ctypes as _C
    _L = _C.CDLL(_capi.LIB_PATH)
    for _n in list(_capi.SIGNATURES):           # an older build lacks the newer entry points
        if not hasattr(_L, _n):
            del _capi.SIGNATURES[_n]
import loos_b200 as lb
from loos_b200.synth import synth_frames_torch
F, N = int(os.environ.get('PS_FRAMES', 100000)), int(os.environ.get('PS_ATOMS', 1000))
X = synth_frames_torch(F, N, device="cuda")
out = torch.empty((F, F), dtype=torch.float32, device="cuda")
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fs = lb.FrameSet(X); torch.cuda.synchronize(); t1 = time.perf_counter()
    e = fs.unit_log2(); torch.cuda.synchronize(); t2 = time.perf_counter()
    lb.rmsds(fs, out=out, engine="tensor"); torch.cuda.synchronize(); t3 = time.perf_counter()
    tm = fs.last_timing()
    fs.close(); torch.cuda.synchronize(); t4 = time.perf_counter()
    print("stage %.1f ms  quantise %.1f ms  rmsds %.1f ms (kernel %.1f, other %.1f)  close %.1f ms" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), tm[0], tm[1], 1e3*(t4-t3)))
