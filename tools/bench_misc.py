"""Side measurements that explain the headline numbers: PCIe copy bandwidth, large
cudaMalloc/cudaFree cost, and the streaming rmsd2ref path (BASELINE config 2) with its
achieved HBM bandwidth.  Prints one JSON line per measurement."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loos_b200 as lb  # noqa: E402
from loos_b200.synth import synth_frames_torch  # noqa: E402


def pcie():
    n = 4 << 30
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, src, dst in (("h2d", h, d), ("d2h", d, h)):
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        print(json.dumps({"measure": "pcie_" + name, "GBps": 3 * n / (time.perf_counter() - t0) / 1e9}))
    del h, d
    t0 = time.perf_counter(); x = torch.empty(40 << 30, dtype=torch.uint8, device="cuda"); torch.cuda.synchronize()
    t1 = time.perf_counter(); del x; torch.cuda.empty_cache(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(json.dumps({"measure": "cudaMalloc_40GiB_ms", "alloc": 1e3 * (t1 - t0), "free": 1e3 * (t2 - t1)}))


def rmsd2ref(F, N, steps=3):
    """C2: F frames x N atoms against one reference; HBM traffic 12*N B read + 8 B written per frame."""
    X = synth_frames_torch(F, N, seed=20261021, device="cuda", chunk=4096)
    tgt = X[0].double().cpu().numpy()
    fs = lb.FrameSet(X)
    del X
    torch.cuda.empty_cache()
    lb.rmsd2ref(fs, tgt)
    ms = []
    for _ in range(steps):
        d = lb.rmsd2ref(fs, tgt)
        ms.append(fs.last_timing()[0])
    ms = float(np.median(ms))
    bytes_ = F * (12.0 * N + 8)
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    print(json.dumps({"measure": "rmsd2ref_stream", "frames": F, "atoms": N, "kernel_ms": ms, "frames_per_s": F / ms * 1e3,
                      "achieved_GBps": bytes_ / ms / 1e6, "hbm_peak_GBps": peaks["hbm_gbs"],
                      "frac": bytes_ / ms / 1e6 / peaks["hbm_gbs"], "rmsd_mean": float(d.mean())}))
    # iterative alignment (rmsd2ref default mode) on a subset
    fs.close()


def rmsd2ref_e2e(F, N, steps=2):
    """C2 end to end: pinned host frames -> stage -> rmsd2ref -> host results (H2D of 12*N*F bytes dominates)."""
    X = synth_frames_torch(F, N, seed=20261021, device="cuda", chunk=4096)
    tgt = X[0].double().cpu().numpy()
    Xh = torch.empty((F, N, 3), dtype=torch.float32, pin_memory=True)
    Xh.copy_(X)
    del X
    torch.cuda.empty_cache()
    xh = Xh.numpy()
    ts = []
    for _ in range(steps + 1):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fs = lb.FrameSet(xh)
        t1 = time.perf_counter()
        d = lb.rmsd2ref(fs, tgt)
        t2 = time.perf_counter()
        fs.close()
        ts.append((time.perf_counter() - t0, t1 - t0, t2 - t1))
    tot, stage, comp = min(ts[1:])
    print(json.dumps({"measure": "rmsd2ref_e2e", "frames": F, "atoms": N, "total_ms": 1e3 * tot, "stage_ms": 1e3 * stage,
                      "rmsd2ref_ms": 1e3 * comp, "frames_per_s": F / tot, "h2d_GBps": F * N * 12 / stage / 1e9,
                      "rmsd_mean": float(d.mean())}))


def rmsds_align(F, Na, Nr, steps=3):
    """rmsds-align: F x F pairs, align on Na atoms, rmsd over Nr others; 18 (Na + Nr) flop per pair on the
    fp64 pipe or the int8 tensor pipe; the two engines are also compared with each other."""
    X = synth_frames_torch(F, Na + Nr, seed=20261022, device="cuda", chunk=4096).cpu().numpy()
    sa, sr = np.arange(Na, dtype=np.uint32), np.arange(Na, Na + Nr, dtype=np.uint32)
    res = {}
    with lb.FrameSet(X, sel=sa) as a, lb.FrameSet(X, sel=sr) as r:
        for eng in ("fp64", "tensor"):
            lb.rmsds_align(a, r, engine=eng)
            ms, wall = [], []
            for _ in range(steps):
                t0 = time.perf_counter()
                M = lb.rmsds_align(a, r, engine=eng)
                wall.append(1e3 * (time.perf_counter() - t0))
                ms.append(a.last_timing()[0])
            res[eng] = M
            ms, wall = float(np.median(ms)), float(np.median(wall))
            pairs = float(F) * F
            print(json.dumps({"measure": "rmsds_align", "engine": eng, "frames": F, "align_atoms": Na, "rmsd_atoms": Nr, "kernel_ms": ms,
                              "call_ms": wall, "pairs_per_s": pairs / ms * 1e3, "algorithmic_TFLOPs": 18.0 * (Na + Nr) * pairs / ms / 1e9,
                              "mean": float(M.mean())}), flush=True)
    d = np.abs(res["fp64"] - res["tensor"])
    print(json.dumps({"measure": "rmsds_align engines", "max_abs_diff": float(d.max()), "mean_abs_diff": float(d.mean())}))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "pcie"):
        pcie()
    if what in ("rmsd2ref_e2e",):
        rmsd2ref_e2e(int(os.environ.get("C2_FRAMES", 1000000)), 2000)
    if what in ("rmsds_align",):
        rmsds_align(int(os.environ.get("ALIGN_FRAMES", 4000)), 500, 500)
    if what in ("all", "rmsd2ref"):
        rmsd2ref(int(os.environ.get("C2_FRAMES", 1000000)), 2000)
