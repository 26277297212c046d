"""All-to-all RMSD with the matrix text formatted on the GPU: bytes and seconds at scale."""
import ctypes as C, sys, time, torch
sys.path.insert(0, "/root/repo")
import loos_b200 as lb
from loos_b200 import capi
from loos_b200.synth import synth_frames_torch
F = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
X = synth_frames_torch(F, 1000, device="cuda")
fs = lb.FrameSet(X)
count = [0, 0]
def sink(_u, _data, n):
    count[0] += n; count[1] += 1
    return 0
cb = capi.SINK(sink)
st = capi.Stats()
for rep in range(2):
    count[:] = [0, 0]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    capi.check(capi.lib().loosb200_rmsds_self_text(fs._h, 2, cb, None, C.byref(st), capi.ENGINES["tensor"]))
    dt = time.perf_counter() - t0
    print("F=%d: %.2f GB of text in %d pieces, %.2f s total (%.1f GB/s of text, %.2f ns per value)" % (
        F, count[0] / 1e9, count[1], dt, count[0] / dt / 1e9, dt / (F * F) * 1e9), flush=True)
fs.close()
