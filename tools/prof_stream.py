"""rmsd2ref target mode: kernel time against the atom count (separates the per-frame solve from the streaming)."""
import sys, json, torch, numpy as np
sys.path.insert(0, "/root/repo")
import loos_b200 as lb
from loos_b200.synth import synth_frames_torch
for F, N in ((1000000, 64), (1000000, 256), (1000000, 1000), (1000000, 2000), (250000, 8000)):
    X = synth_frames_torch(F, N, seed=20261021, device="cuda", chunk=4096)
    tgt = X[0].double().cpu().numpy()
    fs = lb.FrameSet(X); del X; torch.cuda.empty_cache()
    lb.rmsd2ref(fs, tgt)
    ms = []
    for _ in range(3):
        lb.rmsd2ref(fs, tgt); ms.append(fs.last_timing()[0])
    ms = float(np.median(ms))
    print("F=%d N=%d kernel %.2f ms  %.1f ns/frame  %.0f GB/s" % (F, N, ms, ms * 1e6 / F, F * (12.0 * N + 8) / ms / 1e6), flush=True)
    fs.close()
