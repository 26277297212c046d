"""Small run of every kernel for `compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`."""
import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import loos_b200 as lb
from loos_b200.api import rmsds_part
from loos_b200.synth import synth_frames
X = synth_frames(203, 333, seed=3)          # odd sizes: partial tiles, Npad > N, Kpad > N
Y = synth_frames(77, 333, seed=4)
for shape in ("1", "2"):
    os.environ["LOOSB200_CLUSTER"] = shape
    with lb.FrameSet(X) as fs, lb.FrameSet(Y) as fy:
        M, st = lb.rmsds(fs, engine="tensor")
        C, _ = lb.rmsds_cross(fs, fy, engine="tensor")
        rows, blk, _ = rmsds_part(fs, 1, 3, engine="tensor")
        print("tensor shape", shape, M.shape, C.shape, blk.shape, st["max"])
with lb.FrameSet(X) as fs:
    M64, _ = lb.rmsds(fs, engine="fp64")
    print("fp64 max |d|", np.abs(M64 - M).max())
    tgt = X[5].astype(np.float64)
    d, xf = lb.rmsd2ref(fs, tgt, want_xforms=True)
    os.environ["LOOSB200_STREAM_EXPLICIT"] = "1"
    d2 = lb.rmsd2ref(fs, tgt)
    print("rmsd2ref", d[:3], np.abs(d - d2).max())
    sel = np.arange(0, 333, 3, dtype=np.uint32)
with lb.FrameSet(X, sel=sel) as fa, lb.FrameSet(X) as fr:
    r = lb.rmsd2ref_iterative(fa, fr, tol=1e-6, maxiter=50)
    print("iterative", r["iters"], r["rmsd"][:3])
# round 2: rmsds-align on both engines (KIND_ROT / KIND_TRACE, views about the align centroid, several bands),
# mirror-canonical rectangles (banded text), strips instead of column blocks, the pair shape picked by size
from loos_b200.api import rmsds_text
sa, sr = np.arange(0, 120, dtype=np.uint32), np.arange(100, 333, dtype=np.uint32)
os.environ["LOOSB200_ALIGN_BAND_TILES"] = "2"
os.environ.pop("LOOSB200_CLUSTER", None)
with lb.FrameSet(X, sel=sa) as a1, lb.FrameSet(X, sel=sr) as r1, lb.FrameSet(Y, sel=sa) as a2, lb.FrameSet(Y, sel=sr) as r2:
    Mt = lb.rmsds_align(a1, r1, a2, r2, engine="tensor")
    Md = lb.rmsds_align(a1, r1, a2, r2, engine="fp64")
    M1 = lb.rmsds_align(a1, r1, engine="tensor")
    print("rmsds_align engines", np.abs(Mt - Md).max(), M1.shape)
with lb.FrameSet(X) as fs:
    whole, _ = rmsds_text(fs, precision=6, engine="tensor")
    os.environ["LOOSB200_TEXT_BAND_ROWS"] = "80"
    banded, _ = rmsds_text(fs, precision=6, engine="tensor")
    os.environ.pop("LOOSB200_TEXT_BAND_ROWS")
    os.environ["LOOSB200_STRIPS"] = "1"
    Ms, _ = lb.rmsds(fs, engine="tensor")
    os.environ.pop("LOOSB200_STRIPS")
    print("banded text identical", banded == whole, "strips == column blocks", np.array_equal(Ms, M))
Z = synth_frames(90, 4200, seed=6)
with lb.FrameSet(Z) as fs:
    Mz, _ = lb.rmsds(fs, engine="tensor")
    print("pair shape by size", Mz.shape, float(Mz.max()))
print("sanitize smoke done")
