"""The fp64 engine (k_pairs_fp64: selections above the tensor path's 100,000-atom bound, and the on-device
cross-check) at F x N: kernel ms, pairs/s, fp64 TFLOP/s."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loos_b200 as lb  # noqa: E402
from loos_b200.synth import synth_frames  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
X = synth_frames(F, N, seed=5)
with lb.FrameSet(X) as fs:
    lb.rmsds(fs, engine="fp64")
    ms = []
    for _ in range(3):
        M, st = lb.rmsds(fs, engine="fp64")
        ms.append(fs.last_timing()[0])
ms = float(np.median(ms))
pairs = F * (F - 1) / 2.0
print(json.dumps({"measure": "pairs_fp64", "frames": F, "atoms": N, "kernel_ms": ms, "pairs_per_s": pairs / ms * 1e3,
                  "fp64_TFLOPs": 18.0 * N * pairs / ms / 1e9, "avg": st["avg"]}))
