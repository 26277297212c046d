"""Break the end-to-end (host in, host out) C4 call into its parts.  Run on the GPU box."""
import sys, time, torch
sys.path.insert(0, "/root/repo")
import loos_b200 as lb
from loos_b200.synth import synth_frames_torch

F = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
N = 1000
X = synth_frames_torch(F, N, device="cuda")
Xh = torch.empty((F, N, 3), dtype=torch.float32, pin_memory=True); Xh.copy_(X)
out_t = torch.empty((F, F), dtype=torch.float32, pin_memory=True)
dev = torch.empty((F, F), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()


def t(fn, n=2):
    best = 1e9
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


print("contiguous D2H %.1f GB: %.1f ms" % (F * F * 4 / 1e9, t(lambda: out_t.copy_(dev, non_blocking=True))))
w = F // 30
print("column strip D2H (%d wide, all rows): %.1f ms -> %.1f GB/s" % (
    w, *(lambda ms: (ms, F * w * 4 / ms / 1e6))(t(lambda: out_t[:, F - w:].copy_(dev[:, F - w:], non_blocking=True)))))
print("row strip D2H (%d rows): %.1f ms -> %.1f GB/s" % (
    w, *(lambda ms: (ms, F * w * 4 / ms / 1e6))(t(lambda: out_t[F - w:].copy_(dev[F - w:], non_blocking=True)))))
del dev
torch.cuda.empty_cache()
out_h = out_t.numpy().T
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fs = lb.FrameSet(Xh.numpy()); torch.cuda.synchronize(); t1 = time.perf_counter()
    lb.rmsds(fs, out=out_h, engine="tensor"); torch.cuda.synchronize(); t2 = time.perf_counter()
    tm = fs.last_timing()
    fs.close(); t3 = time.perf_counter()
    print("stage %.1f ms  rmsds(host out) %.1f ms (kernels %.1f)  close %.1f ms  total %.1f" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), tm[0], 1e3 * (t3 - t2), 1e3 * (t3 - t0)))
