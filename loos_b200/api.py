"""Host-side mirror of the reference's function boundary for the RMSD path.

Names and argument meaning follow the reference:
  FrameSet(coords, sel)         ~ readCoords(subset, traj, indices)  (src/ensembles.cpp:291-296)
  rmsds(fs)                     ~ Tools/rmsds.cpp single-trajectory mode -> RealMatrix (float32)
  rmsds_cross(a, b)             ~ Tools/rmsds.cpp two-trajectory mode
  rmsd2ref(fs_align, ...)       ~ Tools/rmsd2ref.cpp with --target
  rmsd2ref_iterative(...)       ~ Tools/rmsd2ref.cpp default mode (iterativeAlignment)
Errors the reference throws (LOOSError on size mismatch) surface as LoosB200Error.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import LoosB200Error, Stats, check, lib


def _as_f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


class FrameSet:
    """A staged set of frames on one GPU (opaque loosb200_frames handle).

    coords: (F, natoms, 3) float32 [layout 'xyz', the readCoords order] or
            (F, 3, natoms) float32 [layout 'planes', the DCD record order]; a
            torch CUDA tensor of the same shapes is staged without touching the host.
    sel:    optional uint32 atom indices into the frame (Atom::index(), model order).
    """

    def __init__(self, coords, sel=None, layout="xyz", device=0):
        L = lib()
        self._h = C.c_void_p()
        self.device = device
        lay = {"xyz": capi.LAYOUT_XYZ, "planes": capi.LAYOUT_PLANES}[layout]
        sel_arr = None if sel is None else np.ascontiguousarray(sel, dtype=np.uint32)
        sel_p = None if sel_arr is None else sel_arr.ctypes.data_as(C.POINTER(C.c_uint32))
        nsel = 0 if sel_arr is None else sel_arr.size
        if hasattr(coords, "is_cuda") and coords.is_cuda:
            import torch
            t = coords.contiguous()
            assert t.dtype == torch.float32 and t.dim() == 3
            F = t.shape[0]
            natoms = t.shape[1] if lay == capi.LAYOUT_XYZ else t.shape[2]
            torch.cuda.current_stream(t.device).synchronize()
            check(L.loosb200_stage_frames_device(C.byref(self._h), C.c_void_p(t.data_ptr()), F, natoms,
                                                 sel_p, nsel, lay, t.device.index or 0))
            self.device = t.device.index or 0
        else:
            a = _as_f32(coords.numpy() if hasattr(coords, "numpy") else coords)
            assert a.ndim == 3
            F = a.shape[0]
            natoms = a.shape[1] if lay == capi.LAYOUT_XYZ else a.shape[2]
            check(L.loosb200_stage_frames(C.byref(self._h), a.ctypes.data_as(C.c_void_p), F, natoms,
                                          sel_p, nsel, lay, device))
        self.F = F
        self.N = nsel if sel_arr is not None else natoms

    @classmethod
    def _wrap(cls, handle, F, N, device):
        self = cls.__new__(cls)
        self._h, self.F, self.N, self.device = handle, F, N, device
        return self

    def replicate(self, device):
        """A copy of this staged set on another GPU (device-to-device, loosb200_replicate)."""
        h = C.c_void_p()
        check(lib().loosb200_replicate(self._h, int(device), C.byref(h)))
        return FrameSet._wrap(h, self.F, self.N, int(device))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().loosb200_free(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def centroids(self):
        out = np.zeros((self.F, 3))
        check(lib().loosb200_get_centroids(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def unit_log2(self):
        e = C.c_int()
        check(lib().loosb200_get_unit_log2(self._h, C.byref(e)))
        return e.value

    def last_timing(self):
        """(ms of the main kernel(s), ms of the rest, kernel launches) of the last call."""
        a, b = C.c_double(), C.c_double()
        n = check(lib().loosb200_last_timing(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value, n


def _stats_tuple(s):
    return {"max": s.max, "sum": s.sum, "count": s.count, "avg": (s.sum / s.count) if s.count else 0.0}


def _dev_ptr(t):
    return C.c_void_p(t.data_ptr())


def rmsds(fs, out=None, noout=False, engine="auto", want_stats=True):
    """All-to-all RMSD of one frame set.  Returns (M, stats): M is (F, F) float32 in
    Fortran order (== the reference's column-major RealMatrix), M[i, j] = M(i, j),
    or None with noout=True (Tools/rmsds.cpp --noout).  `out` may be a
    preallocated Fortran-ordered (F, F) float32 numpy array (pinned for speed) or a
    torch CUDA tensor of shape (F, F) (result stays in HBM; symmetric, so the
    memory order is immaterial)."""
    L = lib()
    st = Stats()
    sp = C.byref(st) if want_stats else None
    eng = capi.ENGINES[engine]
    if noout:
        check(L.loosb200_rmsds_self(fs._h, None, 0, sp, eng))
        return None, _stats_tuple(st)
    if out is not None and hasattr(out, "is_cuda") and out.is_cuda:
        assert out.shape == (fs.F, fs.F) and out.is_contiguous()
        check(L.loosb200_rmsds_self_device(fs._h, _dev_ptr(out), fs.F, sp, eng))
        return out, _stats_tuple(st)
    if out is None:
        out = np.empty((fs.F, fs.F), dtype=np.float32, order="F")
    assert out.dtype == np.float32 and out.shape == (fs.F, fs.F) and out.flags.f_contiguous
    check(L.loosb200_rmsds_self(fs._h, out.ctypes.data_as(C.c_void_p), fs.F, sp, eng))
    return out, _stats_tuple(st)


def rmsds_cross(a, b, out=None, noout=False, engine="auto", want_stats=True):
    """M[i, j] = rmsd(a_i, b_j): Tools/rmsds.cpp with model2/traj2 (DualWorker)."""
    L = lib()
    st = Stats()
    sp = C.byref(st) if want_stats else None
    eng = capi.ENGINES[engine]
    if noout:
        check(L.loosb200_rmsds_cross(a._h, b._h, None, 0, sp, eng))
        return None, _stats_tuple(st)
    if out is None:
        out = np.empty((a.F, b.F), dtype=np.float32, order="F")
    assert out.dtype == np.float32 and out.shape == (a.F, b.F) and out.flags.f_contiguous
    check(L.loosb200_rmsds_cross(a._h, b._h, out.ctypes.data_as(C.c_void_p), a.F, sp, eng))
    return out, _stats_tuple(st)


def part_rows(F, part, nparts):
    """Frame index of every packed row of `part` (0xffffffff marks padding rows)."""
    n = C.c_size_t()
    check(lib().loosb200_part_rows(F, part, nparts, None, C.byref(n)))
    rows = np.empty(n.value, dtype=np.uint32)
    check(lib().loosb200_part_rows(F, part, nparts, rows.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(n)))
    return rows


def rmsds_part(fs, part, nparts, out=None, noout=False, engine="auto"):
    """The rows of the lower triangle owned by `part` of `nparts` (multi-GPU split).
    Returns (rows, block, stats): block[r, j] = M(rows[r], j) for j < rows[r]."""
    L = lib()
    st = Stats()
    eng = capi.ENGINES[engine]
    rows = part_rows(fs.F, part, nparts)
    if noout:
        check(L.loosb200_rmsds_self_part(fs._h, part, nparts, None, 0, C.byref(st), eng))
        return rows, None, _stats_tuple(st)
    if out is not None and hasattr(out, "is_cuda") and out.is_cuda:
        assert out.shape == (len(rows), fs.F) and out.is_contiguous()
        check(L.loosb200_rmsds_self_part_device(fs._h, part, nparts, _dev_ptr(out), fs.F, C.byref(st), eng))
        return rows, out, _stats_tuple(st)
    if out is None:
        out = np.zeros((len(rows), fs.F), dtype=np.float32)
    check(L.loosb200_rmsds_self_part(fs._h, part, nparts, out.ctypes.data_as(C.c_void_p), fs.F, C.byref(st), eng))
    return rows, out, _stats_tuple(st)


def _dptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def rmsd2ref(fs_align, ref_align, fs_rmsd=None, ref_rmsd=None, want_xforms=False, out=None):
    """Per-frame RMSD to a reference after superposition (Tools/rmsd2ref.cpp with
    --target).  ref_align: (N_a, 3) target align-selection coords, or None for no
    alignment; fs_rmsd / ref_rmsd: the rmsd selection (default: the align one).
    out: optional preallocated float64 array of F values (pinned memory, e.g. PinnedMatrix((F,),
    np.float64).array, lets the result copy run at the PCIe rate)."""
    F = fs_align.F
    ra = None if ref_align is None else np.ascontiguousarray(ref_align, dtype=np.float64).reshape(-1)
    rr = None if ref_rmsd is None else np.ascontiguousarray(ref_rmsd, dtype=np.float64).reshape(-1)
    if ra is not None and ra.size != 3 * fs_align.N:
        raise LoosB200Error(capi.ERR_SIZE_MISMATCH, "Cannot superimpose groups with differing numbers of atoms")
    nr = (fs_rmsd or fs_align).N
    if rr is not None and rr.size != 3 * nr:
        raise LoosB200Error(capi.ERR_SIZE_MISMATCH, "Cannot compute RMSD between groups with different sizes")
    if out is None:
        out = np.zeros(F)
    assert out.dtype == np.float64 and out.shape == (F,) and out.flags.c_contiguous
    xf = np.zeros((F, 4, 4)) if want_xforms else None
    check(lib().loosb200_rmsd2ref(fs_align._h, None if fs_rmsd is None else fs_rmsd._h, _dptr(ra), _dptr(rr),
                                  _dptr(out), _dptr(xf)))
    return (out, xf) if want_xforms else out


def rmsd2ref_iterative(fs_align, fs_rmsd=None, tol=1e-6, maxiter=1000):
    """rmsd2ref's default mode.  Returns dict(rmsd, xforms, avg, iters, rms)."""
    F = fs_align.F
    nr = (fs_rmsd or fs_align).N
    out, xf, avg = np.zeros(F), np.zeros((F, 4, 4)), np.zeros((nr, 3))
    it, rms = C.c_int(), C.c_double()
    check(lib().loosb200_rmsd2ref_iterative(fs_align._h, None if fs_rmsd is None else fs_rmsd._h, tol, maxiter,
                                            _dptr(out), _dptr(xf), _dptr(avg), C.byref(it), C.byref(rms)))
    return {"rmsd": out, "xforms": xf, "avg": avg, "iters": it.value, "rms": rms.value}


def align_accumulate(fs_align, target, fs_accum=None, want_xforms=False):
    """One pass of iterativeAlignment's loop body over this frame set: every frame superposed onto
    `target` (N_align x 3), transformed coordinates of `fs_accum` (default: the align set) summed.
    Returns the (N_accum, 3) sum [and the (F, 4, 4) transforms].  Building block of the sharded
    iterative alignment (loos_b200/multigpu.py)."""
    tgt = np.ascontiguousarray(target, dtype=np.float64).reshape(-1)
    if tgt.size != 3 * fs_align.N:
        raise LoosB200Error(capi.ERR_SIZE_MISMATCH, "target has %d coordinates, the frames have %d atoms" % (tgt.size, fs_align.N))
    na = (fs_accum or fs_align).N
    out = np.zeros((na, 3))
    xf = np.zeros((fs_align.F, 4, 4)) if want_xforms else None
    check(lib().loosb200_align_accumulate(fs_align._h, None if fs_accum is None else fs_accum._h, _dptr(tgt), _dptr(out), _dptr(xf)))
    return (out, xf) if want_xforms else out


def rmsds_align(align1, rmsd1=None, align2=None, rmsd2=None, engine="auto"):
    """Pairwise RMSD over one selection after aligning on another
    (Packages/Clustering/rmsds-align.py:236-256): M[i, j] = rmsd of rmsd2[j] against rmsd1[i] moved by
    the superposition of align1[i] onto align2[j].  Second trajectory defaults to the first.
    engine: "fp64" (within 1e-8 A of the reference), "tensor" (int8 tensor path, ~1e-6 A), "auto".
    Returns an (F1, F2) float64 array."""
    if align2 is None:
        align2, rmsd2 = align1, rmsd1
    out = np.zeros((align1.F, align2.F), dtype=np.float64, order="F")
    check(lib().loosb200_rmsds_align_engine(align1._h, None if rmsd1 is None else rmsd1._h, align2._h,
                                            None if rmsd2 is None else rmsd2._h, _dptr(out), max(1, align1.F),
                                            capi.ENGINES[engine]))
    return out


def rmsds_text(fs, fs_b=None, precision=2, engine="auto", write=None):
    """The result matrix as the reference prints it (`cout << setprecision(p) << M`), formatted on
    the GPU and streamed in row order.  `write(bytes)` receives the pieces (default: they are
    collected and returned).  Returns (text or None, stats)."""
    pieces = []
    emit = write or pieces.append

    def _sink(_user, data, n):
        emit(C.string_at(data, n))
        return 0

    cb = capi.SINK(_sink)
    st = Stats()
    if fs_b is None:
        check(lib().loosb200_rmsds_self_text(fs._h, int(precision), cb, None, C.byref(st), capi.ENGINES[engine]))
    else:
        check(lib().loosb200_rmsds_cross_text(fs._h, fs_b._h, int(precision), cb, None, C.byref(st), capi.ENGINES[engine]))
    return (None if write else b"".join(pieces)), _stats_tuple(st)


class PinnedMatrix:
    """Page-locked host memory that is pinned for every GPU of the process (loosb200_host_alloc),
    seen as a numpy array (`.array`).  Keep the object alive while the array is in use."""

    def __init__(self, shape, dtype=np.float32, order="F"):
        count = int(np.prod(shape))
        nbytes = max(1, count * np.dtype(dtype).itemsize)
        self._p = lib().loosb200_host_alloc(nbytes)
        if not self._p:
            raise MemoryError("loosb200_host_alloc(%d)" % nbytes)
        buf = (C.c_char * nbytes).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape, order=order)

    def free(self):
        if getattr(self, "_p", None):
            self.array = None
            lib().loosb200_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:   # interpreter shutdown: the module globals may be gone, the process is ending anyway
            pass


def stage_multi(coords, devices, sel=None, layout="xyz"):
    """FrameSets of the same host frames on several GPUs (loosb200_stage_frames_multi): every device
    uploads its slice over its own PCIe link, the slices are exchanged device-to-device."""
    lay = {"xyz": capi.LAYOUT_XYZ, "planes": capi.LAYOUT_PLANES}[layout]
    a = _as_f32(coords)
    assert a.ndim == 3
    F = a.shape[0]
    natoms = a.shape[1] if lay == capi.LAYOUT_XYZ else a.shape[2]
    sel_arr = None if sel is None else np.ascontiguousarray(sel, dtype=np.uint32)
    sel_p = None if sel_arr is None else sel_arr.ctypes.data_as(C.POINTER(C.c_uint32))
    nsel = 0 if sel_arr is None else sel_arr.size
    n = len(devices)
    hs = (C.c_void_p * n)()
    devs = (C.c_int * n)(*[int(d) for d in devices])
    check(lib().loosb200_stage_frames_multi(hs, devs, n, a.ctypes.data_as(C.c_void_p), F, natoms, sel_p, nsel, lay))
    N = nsel if sel_arr is not None else natoms
    return [FrameSet._wrap(C.c_void_p(hs[k]), F, N, int(devices[k])) for k in range(n)]


def _handles(sets):
    arr = (C.c_void_p * len(sets))(*[fs._h for fs in sets])
    return arr


def rmsds_multi(replicas, out=None, noout=False, engine="auto", want_stats=True):
    """rmsds() on several GPUs of one process: `replicas` are FrameSets of the same frames on
    different devices (fs.replicate(d)); replicas[0] is rank 0.  out: None (a Fortran-ordered numpy
    matrix is returned), a preallocated host matrix, or a torch CUDA tensor on replicas[0]'s device
    (gather of the row blocks into rank 0's HBM over NVLink)."""
    L = lib()
    st = Stats()
    sp = C.byref(st) if want_stats else None
    eng = capi.ENGINES[engine]
    hs = _handles(replicas)
    F = replicas[0].F
    if noout:
        check(L.loosb200_rmsds_self_multi(hs, len(replicas), None, 0, sp, eng))
        return None, _stats_tuple(st)
    if out is not None and hasattr(out, "is_cuda") and out.is_cuda:
        assert out.shape == (F, F) and out.is_contiguous() and (out.device.index or 0) == replicas[0].device
        check(L.loosb200_rmsds_self_multi_device(hs, len(replicas), _dev_ptr(out), F, sp, eng))
        return out, _stats_tuple(st)
    if out is None:
        out = np.empty((F, F), dtype=np.float32, order="F")
    assert out.dtype == np.float32 and out.shape == (F, F) and out.flags.f_contiguous
    check(L.loosb200_rmsds_self_multi(hs, len(replicas), out.ctypes.data_as(C.c_void_p), F, sp, eng))
    return out, _stats_tuple(st)


def rmsds_cross_multi(a_replicas, b_replicas, out=None, noout=False, engine="auto", want_stats=True):
    """rmsds_cross() with the rows of `a` split over several GPUs."""
    L = lib()
    st = Stats()
    sp = C.byref(st) if want_stats else None
    eng = capi.ENGINES[engine]
    ha, hb = _handles(a_replicas), _handles(b_replicas)
    FA, FB = a_replicas[0].F, b_replicas[0].F
    if noout:
        check(L.loosb200_rmsds_cross_multi(ha, hb, len(a_replicas), None, 0, sp, eng))
        return None, _stats_tuple(st)
    if out is None:
        out = np.empty((FA, FB), dtype=np.float32, order="F")
    assert out.dtype == np.float32 and out.shape == (FA, FB) and out.flags.f_contiguous
    check(L.loosb200_rmsds_cross_multi(ha, hb, len(a_replicas), out.ctypes.data_as(C.c_void_p), FA, sp, eng))
    return out, _stats_tuple(st)


def rmsds_text_multi(replicas, b_replicas=None, precision=2, engine="auto", write=None):
    """rmsds_text() with the row bands dealt to several GPUs; same bytes."""
    pieces = []
    emit = write or pieces.append

    def _sink(_user, data, n):
        emit(C.string_at(data, n))
        return 0

    cb = capi.SINK(_sink)
    st = Stats()
    if b_replicas is None:
        check(lib().loosb200_rmsds_self_multi_text(_handles(replicas), len(replicas), int(precision), cb, None,
                                                   C.byref(st), capi.ENGINES[engine]))
    else:
        check(lib().loosb200_rmsds_cross_multi_text(_handles(replicas), _handles(b_replicas), len(replicas),
                                                    int(precision), cb, None, C.byref(st), capi.ENGINES[engine]))
    return (None if write else b"".join(pieces)), _stats_tuple(st)
