"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the two CPU checkers.

* ``port``      -> oracle/liboracle.so, our plain-C restatement (oracle.c)
* ``reference`` -> oracle/_ref/libloosref.so, the reference's own arithmetic
  (src/alignment.cpp:12-281 compiled where it lies, OpenBLAS dgemm_/dgesvj_)

Both expose the same functions under the prefixes ``oracle_`` / ``ref_``; the
``Checker`` class hides the prefix.  Layouts follow the reference: coordinates are
F rows of 3N doubles, xyz interleaved (src/ensembles.cpp:275-279); RMSD matrices
are float32 column-major (src/MatrixOrder.hpp:101-103) -- returned here as numpy
arrays indexed [i, j] (Fortran order), so ``M[i, j] == M(i, j)`` of the reference.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)


def build(quiet=True):
    """(Re)build liboracle.so and, when /root/reference is present, _ref/."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


class Checker:
    def __init__(self, kind="port"):
        self.kind = kind
        if kind == "port":
            path, self.pfx = os.path.join(_HERE, "liboracle.so"), "oracle_"
        elif kind == "reference":
            path, self.pfx = os.path.join(_HERE, "_ref", "libloosref.so"), "ref_"
        else:
            raise ValueError(kind)
        if not os.path.exists(path):
            if kind == "port":
                build()
            else:
                raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        if kind == "reference":
            self.lib.ref_set_blas_threads(1)  # Tools/rmsds.cpp:88-92 advice
        f = self._f
        f("centered_rmsd").restype = C.c_double
        f("centered_rmsd").argtypes = [_dp, _dp, C.c_int]
        f("center_at_origin").argtypes = [_dp, C.c_int, _dp]
        f("kabsch").argtypes = [_dp, _dp, C.c_int, _dp]
        f("kabsch_core_S").argtypes = [_dp, _dp, C.c_int, _dp]
        f("rmsds_self").argtypes = [_dp, C.c_int, C.c_int, _fp, C.c_int]
        f("rmsds_cross").argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_int, _fp, C.c_int]
        f("rmsd2ref_target").argtypes = [_dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, _dp, _dp]
        f("rmsd2ref_iterative").restype = C.c_int
        f("rmsd2ref_iterative").argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double,
                                            C.c_int, _dp, _dp, _dp, _dp]

    def _f(self, name):
        return getattr(self.lib, self.pfx + name)

    # -- per-pair functions (src/alignment.cpp) --
    def center_at_origin(self, v):
        v = np.array(v, dtype=np.float64).ravel().copy()
        c = np.zeros(3)
        self._f("center_at_origin")(v.ctypes.data_as(_dp), v.size, c.ctypes.data_as(_dp))
        return v, c

    def centered_rmsd(self, u, v):
        u, pu = _d(np.ravel(u)); v, pv = _d(np.ravel(v))
        return self._f("centered_rmsd")(pu, pv, u.size)

    def kabsch(self, u, v):
        u, pu = _d(np.ravel(u)); v, pv = _d(np.ravel(v))
        M = np.zeros(16)
        self._f("kabsch")(pu, pv, u.size, M.ctypes.data_as(_dp))
        return M.reshape(4, 4)

    def kabsch_core_S(self, u, v):
        u, pu = _d(np.ravel(u)); v, pv = _d(np.ravel(v))
        S = np.zeros(3)
        self._f("kabsch_core_S")(pu, pv, u.size, S.ctypes.data_as(_dp))
        return S

    # -- tool-level drivers --
    def rmsds_self(self, T, nthreads=1):
        """T: (F, N, 3) or (F, 3N) raw (uncentred) coords -> (F, F) float32."""
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(len(T), -1)
        F, n3 = T.shape
        M = np.zeros((F, F), dtype=np.float32, order="F")
        self._f("rmsds_self")(T.ctypes.data_as(_dp), F, n3, M.ctypes.data_as(_fp), nthreads)
        return M

    def rmsds_cross(self, T1, T2, nthreads=1):
        T1 = np.ascontiguousarray(T1, dtype=np.float64).reshape(len(T1), -1)
        T2 = np.ascontiguousarray(T2, dtype=np.float64).reshape(len(T2), -1)
        assert T1.shape[1] == T2.shape[1]
        M = np.zeros((len(T1), len(T2)), dtype=np.float32, order="F")
        self._f("rmsds_cross")(T1.ctypes.data_as(_dp), len(T1), T2.ctypes.data_as(_dp), len(T2),
                               T1.shape[1], M.ctypes.data_as(_fp), nthreads)
        return M

    def rmsd2ref_target(self, fa, fr, ta, tr):
        """fa: (F, Na, 3) align coords per frame, fr: (F, Nr, 3) rmsd coords per
        frame, ta/tr the target's selections.  Returns (rmsd[F], xforms[F,4,4])."""
        fa = np.ascontiguousarray(fa, dtype=np.float64).reshape(len(fa), -1)
        fr = np.ascontiguousarray(fr, dtype=np.float64).reshape(len(fr), -1)
        ta, pta = _d(np.ravel(ta)); tr, ptr = _d(np.ravel(tr))
        F = len(fa)
        out = np.zeros(F); xf = np.zeros((F, 16))
        self._f("rmsd2ref_target")(fa.ctypes.data_as(_dp), fr.ctypes.data_as(_dp), pta, ptr, F,
                                   fa.shape[1], fr.shape[1], out.ctypes.data_as(_dp),
                                   xf.ctypes.data_as(_dp))
        return out, xf.reshape(F, 4, 4)

    def rmsd2ref_iterative(self, fa, fr, tol=1e-6, maxiter=1000):
        fa = np.ascontiguousarray(fa, dtype=np.float64).reshape(len(fa), -1)
        fr = np.ascontiguousarray(fr, dtype=np.float64).reshape(len(fr), -1)
        F = len(fa)
        out = np.zeros(F); xf = np.zeros((F, 16)); avg = np.zeros(fr.shape[1]); rms = C.c_double()
        it = self._f("rmsd2ref_iterative")(fa.ctypes.data_as(_dp), fr.ctypes.data_as(_dp), F,
                                           fa.shape[1], fr.shape[1], tol, maxiter,
                                           out.ctypes.data_as(_dp), xf.ctypes.data_as(_dp),
                                           avg.ctypes.data_as(_dp), C.byref(rms))
        return out, xf.reshape(F, 4, 4), avg.reshape(-1, 3), it, rms.value


def _apply44(W, X):
    """XForm applied to coordinates: AtomicGroup::applyTransform (src/AG_numerical.cpp:363-370)."""
    return X @ W[:3, :3].T + W[:3, 3]


def rmsds_align(checker, A1, R1, A2, R2):
    """The loop of Packages/Clustering/rmsds-align.py:236-256 on in-memory frames.
    A1: (F1, Na, 3) align selection of trajectory 1, R1: (F1, Nr, 3) its rmsd selection, A2 / R2 the
    same for trajectory 2.  As in the script, the copies of frame i (ref_align, ref_target) KEEP
    the transforms of the earlier j: alignOnto (src/AG_linalg.cpp:188-197 superposition +
    applyTransform) moves ref_align every time.  checker.kabsch is alignment::kabsch
    (src/alignment.cpp:217-233; the reference's own code when checker.kind == "reference").
    Returns the (F1, F2) float64 matrix the script prints (before round())."""
    A1 = np.asarray(A1, dtype=np.float64); R1 = np.asarray(R1, dtype=np.float64)
    A2 = np.asarray(A2, dtype=np.float64); R2 = np.asarray(R2, dtype=np.float64)
    if A1.shape[1] != A2.shape[1]:
        raise ValueError("Cannot superimpose groups with differing numbers of atoms")    # src/AG_linalg.cpp:189-192
    if R1.shape[1] != R2.shape[1]:
        raise ValueError("Cannot compute RMSD between groups with different sizes")      # src/AG_numerical.cpp:303-304
    M = np.zeros((len(A1), len(A2)))
    for i in range(len(A1)):
        ref_align, ref_target = A1[i].copy(), R1[i].copy()
        for j in range(len(A2)):
            W = checker.kabsch(ref_align, A2[j])
            ref_align = _apply44(W, ref_align)
            ref_target = _apply44(W, ref_target)
            d = R2[j] - ref_target
            M[i, j] = np.sqrt((d * d).sum() / R2.shape[1])                                # src/AG_numerical.cpp:301-318
    return M


def py_round_text(M, precision=2):
    """rmsds-align.py:254-256: print(round(v, precision), "", end='') per value, print("") per row."""
    return "".join("".join("%s " % repr(round(float(v), precision)) for v in row) + "\n" for row in M)


def reference_iterative_vecmatrix(checker, frames, tol=1e-6, maxiter=1000):
    """The reference's OWN iterativeAlignment(vecMatrix&, threshold, maxiter) (src/alignment.cpp:288-318, compiled in
    place; reference checker only).  frames: (F, N, 3).  Returns (xforms (F, 4, 4), rms, iterations)."""
    assert checker.kind == "reference"
    X = np.ascontiguousarray(frames, dtype=np.float64).reshape(len(frames), -1)
    xf = np.zeros((len(X), 16))
    rms = C.c_double()
    f = checker.lib.ref_iterative_vecmatrix
    f.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.c_int, _dp, _dp]
    it = f(X.ctypes.data_as(_dp), len(X), X.shape[1], tol, maxiter, xf.ctypes.data_as(_dp), C.byref(rms))
    return xf.reshape(-1, 4, 4), rms.value, it


def have_reference():
    return os.path.exists(os.path.join(_HERE, "_ref", "libloosref.so"))


class RefXTC:
    """The reference's OWN XTC writer and reader (src/xtcwriter.cpp, src/xtc.cpp, src/xdr.hpp compiled where they
    lie: oracle/xtc_ref_wrapper.cpp, oracle/_ref/libloosxtc.so).  Exists only where /root/reference did at build
    time (`available()`); the fixtures it generated travel in tests/golden/xtc_ref.npz."""

    @staticmethod
    def available():
        return os.path.exists(os.path.join(_HERE, "_ref", "libloosxtc.so"))

    def __init__(self):
        self.lib = C.CDLL(os.path.join(_HERE, "_ref", "libloosxtc.so"))
        self.lib.ref_xtc_write.restype = C.c_long
        self.lib.ref_xtc_write.argtypes = [_dp, C.c_int, C.c_int, C.c_float, _dp, C.c_double, C.c_char_p, C.c_long]
        self.lib.ref_xtc_read.argtypes = [C.c_char_p, C.c_long, _dp, C.c_int, C.POINTER(C.c_int)]

    def write(self, frames_angstrom, precision=1000.0, box_angstrom=(50.0, 60.0, 70.0), dt=1.0):
        """XTCWriter::writeFrame over the frames (GCoords are double) -> the bytes of the file."""
        X = np.ascontiguousarray(frames_angstrom, dtype=np.float64)
        F, N, _ = X.shape
        box = np.array(box_angstrom, dtype=np.float64)
        cap = 256 + F * (N * 16 + 256)
        buf = C.create_string_buffer(cap)
        n = self.lib.ref_xtc_write(X.ctypes.data_as(_dp), F, N, precision, box.ctypes.data_as(_dp), dt, buf, cap)
        if n < 0:
            raise RuntimeError("reference XTC writer failed (%d)" % n)
        return buf.raw[:n]

    def read(self, data, max_frames, natoms):
        """XTC::parseFrame over the bytes -> (frames, natoms, 3) float64, the GCoords the reference hands on."""
        out = np.zeros((max_frames, natoms, 3))
        na = C.c_int()
        nf = self.lib.ref_xtc_read(data, len(data), out.ctypes.data_as(_dp), max_frames, C.byref(na))
        if nf < 0:
            raise RuntimeError("reference XTC reader failed (%d)" % nf)
        assert na.value == natoms
        return out[:nf]


class RefMatrix:
    """The reference's OWN matrix text writer and reader (operator<< of src/MatrixImpl.hpp:208-222 under setprecision,
    readAsciiMatrix of src/MatrixRead.hpp:102-143; header-only, included where they lie: oracle/matrix_ref_wrapper.cpp,
    oracle/_ref/libloosmatrix.so).  Exists only where /root/reference did at build time; fixtures from it travel in
    tests/golden/matrix_ref.npz."""

    @staticmethod
    def available():
        return os.path.exists(os.path.join(_HERE, "_ref", "libloosmatrix.so"))

    def __init__(self):
        self.lib = C.CDLL(os.path.join(_HERE, "_ref", "libloosmatrix.so"))
        self.lib.ref_matrix_text.restype = C.c_long
        self.lib.ref_matrix_text.argtypes = [_fp, C.c_int, C.c_int, C.c_long, C.c_int, C.c_char_p, C.c_long]
        self.lib.ref_matrix_read.argtypes = [C.c_char_p, C.c_long, _fp, C.c_long, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_int]

    def text(self, M, precision):
        M = np.asfortranarray(M, dtype=np.float32)
        m, n = M.shape
        cap = 64 + m * n * (precision + 16) + m
        buf = C.create_string_buffer(cap)
        k = self.lib.ref_matrix_text(M.ctypes.data_as(_fp), m, n, m, precision, buf, cap)
        if k < 0:
            raise RuntimeError("reference matrix writer failed (%d)" % k)
        return buf.raw[:k].decode()

    def read(self, text):
        """-> (M, None) or (None, the reference's error message)."""
        data = text.encode() if isinstance(text, str) else text
        out = np.zeros(max(16, len(data)), np.float32)
        m, n = C.c_int(), C.c_int()
        err = C.create_string_buffer(400)
        rc = self.lib.ref_matrix_read(data, len(data), out.ctypes.data_as(_fp), out.size, C.byref(m), C.byref(n), err, 400)
        if rc != 0:
            return None, err.value.decode()
        return out[:m.value * n.value].reshape(n.value, m.value).T.copy(), None


class RefPDB:
    """The reference's OWN PDB atom-record parser (src/pdb.cpp:65-156 with parseStringAs / parseStringAsHybrid36 of
    src/utils.{hpp,cpp}), uniquifyVector and MultiTrajectory indexing (src/MultiTraj.hpp:97-105, MultiTraj.cpp:48-58),
    compiled where they lie: oracle/pdb_ref_wrapper.cpp, oracle/_ref/libloospdb.so.  Fixtures: tests/golden/pdb_ref.npz."""
    FIELDS = ("record", "name", "altloc", "resname", "chain", "icode", "segid", "element")

    @staticmethod
    def available():
        return os.path.exists(os.path.join(_HERE, "_ref", "libloospdb.so"))

    def __init__(self):
        self.lib = C.CDLL(os.path.join(_HERE, "_ref", "libloospdb.so"))
        self.lib.ref_multitraj.restype = C.c_long
        self.lib.ref_uniquify.restype = C.c_long

    def atom(self, line):
        """-> dict of the fields parseAtomRecord stores, or the ParseError text."""
        i, s, r, err = (C.c_int * 2)(), C.create_string_buffer(128), (C.c_double * 5)(), C.create_string_buffer(800)
        if self.lib.ref_pdb_atom(line.encode(), i, s, r, err, 800) != 0:
            return err.value.decode()
        d = {"id": i[0], "resid": i[1], "x": r[0], "y": r[1], "z": r[2], "occupancy": r[3], "bfactor": r[4]}
        for k, name in enumerate(self.FIELDS):
            d[name] = s.raw[16 * k:16 * k + 16].split(b"\0")[0].decode()
        return d

    def hybrid36(self, field):
        return self.lib.ref_hybrid36(field.encode(), len(field))

    def range_item(self, *args):
        """RangeItem(a[, b[, c]]).generate() of src/index_range_parser.cpp:63-102 -> list, or None for 'Invalid range'."""
        self.lib.ref_range_item.restype = C.c_long
        a = int(args[0]); b = int(args[1]) if len(args) > 1 else 0; c = int(args[2]) if len(args) > 2 else 0
        cap = 1 << 16
        out = np.zeros(cap, np.uint32)
        n = self.lib.ref_range_item(len(args), C.c_uint(a), C.c_uint(b), C.c_int(c), out.ctypes.data_as(C.POINTER(C.c_uint32)), cap)
        return None if n < 0 else [int(v) for v in out[:min(n, cap)]]

    def multitraj(self, sizes, skip, stride):
        sizes = np.ascontiguousarray(sizes, dtype=np.uint32)
        cap = int(sizes.sum()) + 1
        lt, lf = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
        P = C.POINTER(C.c_uint32)
        n = self.lib.ref_multitraj(sizes.ctypes.data_as(P), len(sizes), skip, stride, lt.ctypes.data_as(P), lf.ctypes.data_as(P), cap)
        return [(int(a), int(b)) for a, b in zip(lt[:n], lf[:n])]

    def backbone(self, resname, name):
        """BackboneSelector (src/Selectors.cpp:37-132) on an atom with this residue and atom name."""
        return bool(self.lib.ref_backbone(resname.encode(), name.encode()))

    def hydrogen(self, name, mass=None):
        """HydrogenSelector (src/Selectors.cpp:156-164); mass=None: the mass property is not set (PDB input)."""
        self.lib.ref_hydrogen.argtypes = [C.c_char_p, C.c_int, C.c_double]
        return bool(self.lib.ref_hydrogen(name.encode(), 0 if mass is None else 1, 0.0 if mass is None else float(mass)))

    def uniquify(self, v):
        v = np.ascontiguousarray(v, dtype=np.uint32)
        out = np.zeros(max(1, len(v)), np.uint32)
        P = C.POINTER(C.c_uint32)
        n = self.lib.ref_uniquify(v.ctypes.data_as(P), len(v), out.ctypes.data_as(P))
        return out[:n].tolist()


# ---------------------------------------------------------------------------
# The statistics lines of the tools, restated (the GPU tool tests compare the tools' stderr / stdout with these).
def stats_line_half(M):
    """showStatsHalf (Tools/rmsds.cpp:425-439): max and mean over the strict lower triangle, from the float32 entries."""
    M = np.asarray(M, dtype=np.float32)
    low = M[np.tril_indices(len(M), -1)].astype(np.float64)
    return "Max rmsd = %.4f, avg rmsd = %.4f\n" % (low.max() if low.size else 0.0, low.sum() / (len(M) * (len(M) - 1) // 2))


def stats_line_whole(M):
    """showStatsWhole (Tools/rmsds.cpp:442-456) / rms-overlap's showStats (Tools/rms-overlap.cpp:458-473)."""
    r = np.asarray(M, dtype=np.float32).astype(np.float64)
    return "Max rmsd = %.4f, avg rmsd = %.4f\n" % (r.max(), r.sum() / r.size)


def fractional_stats_line(M, cutoff):
    """showFractionalStats (Tools/rms-overlap.cpp:475-521), bug for bug: the `if (R(i,j) > max)` has no braces, so only
    the row index is conditional -- the column index and the "max" value are those of the LAST off-diagonal element
    visited; avg and variance skip i == j but divide by the full size."""
    M32 = np.asarray(M, dtype=np.float32)
    r = M32.astype(np.float64)
    m, n = r.shape
    off = ~np.eye(m, n, dtype=bool)
    avg = r[off].sum() / r.size
    var = (r[off] ** 2).sum() / r.size - avg * avg
    mval, mi, mj = 0.0, 0, 0
    for i in range(m):
        for j in range(n):
            if i != j:
                if r[i, j] > mval:
                    mi = i
                mj = j
                mval = r[i, j]
    return "Max rmsd = %.4f between frames %d, %d, avg rmsd = %.4f, variance = %.4f, frames below %.4f = %d, total = %d\n" % (
        mval, mi, mj, avg, var, float(np.float32(cutoff)), int((M32[off] < np.float32(cutoff)).sum()), r.size)


class RefStats:
    """The reference's OWN showStatsHalf / showStatsWhole (Tools/rmsds.cpp:425-456) and showStats / showFractionalStats
    (Tools/rms-overlap.cpp:458-521), compiled where they lie (oracle/stats_ref_wrapper.cpp, oracle/_ref/libloosstats.so)."""

    @staticmethod
    def available():
        return os.path.exists(os.path.join(_HERE, "_ref", "libloosstats.so"))

    def __init__(self):
        self.lib = C.CDLL(os.path.join(_HERE, "_ref", "libloosstats.so"))
        for name in ("ref_stats_half", "ref_stats_whole", "ref_overlap_stats", "ref_overlap_fractional"):
            getattr(self.lib, name).restype = C.c_long

    def _call(self, name, M, *extra):
        M = np.asfortranarray(M, dtype=np.float32)
        m, n = M.shape
        buf = C.create_string_buffer(1024)
        args = [M.ctypes.data_as(_fp)] + ([C.c_int(m)] if name == "ref_stats_half" else [C.c_int(m), C.c_int(n)]) + [C.c_long(m)] + list(extra) + [buf, C.c_long(1024)]
        k = getattr(self.lib, name)(*args)
        if k < 0:
            raise RuntimeError("%s failed (%d)" % (name, k))
        return buf.raw[:k].decode()

    def half(self, M):
        return self._call("ref_stats_half", M)

    def whole(self, M):
        return self._call("ref_stats_whole", M)

    def overlap(self, M):
        return self._call("ref_overlap_stats", M)

    def fractional(self, M, cutoff, is_noop=True):
        return self._call("ref_overlap_fractional", M, C.c_float(cutoff), C.c_int(1 if is_noop else 0))


class RefSelect:
    """The reference's OWN selection language: its pre-generated flex scanner and bison grammar, the kernel, Atom and the
    backbone / hydrogen predicates compiled unmodified from /root/reference/src (oracle/selection_ref_wrapper.cpp,
    oracle/shim_sel/, oracle/_ref/libloosselect.so).  Fixture: tests/golden/selection_ref.json."""

    @staticmethod
    def available():
        return os.path.exists(os.path.join(_HERE, "_ref", "libloosselect.so"))

    def __init__(self, atoms):
        """atoms: list of dicts with id, resid, name, resname, segid, chain (as oracle/frontend_oracle.py::read_pdb returns)."""
        self.lib = C.CDLL(os.path.join(_HERE, "_ref", "libloosselect.so"))
        self.n = len(atoms)
        self.ids = np.array([a["id"] for a in atoms], np.int32)
        self.resids = np.array([a["resid"] for a in atoms], np.int32)
        self.strs = np.zeros((self.n, 4, 8), np.uint8)
        for i, a in enumerate(atoms):
            for j, key in enumerate(("name", "resname", "segid", "chain")):
                b = a[key].encode()[:7]
                self.strs[i, j, :len(b)] = list(b)

    def select(self, selection):
        """-> (list of selected indices, None) or (None, the exception text)."""
        mask = np.zeros(self.n, np.uint8)
        err = C.create_string_buffer(800)
        ip = C.POINTER(C.c_int)
        rc = self.lib.ref_select(selection.encode(), self.n, self.ids.ctypes.data_as(ip), self.resids.ctypes.data_as(ip),
                                 self.strs.ctypes.data_as(C.c_char_p), mask.ctypes.data_as(C.POINTER(C.c_ubyte)), err, 800)
        if rc != 0:
            return None, err.value.decode()
        return np.nonzero(mask)[0].tolist(), None


class RefDCD:
    """The reference's OWN DCD reader and writer (src/dcd.cpp, src/dcdwriter.cpp compiled where they lie:
    oracle/dcd_ref_wrapper.cpp, oracle/_ref/libloosdcd.so).  Exists only where /root/reference did at build time
    (`available()`); the fixtures it generated travel in tests/golden/dcd_ref.npz."""

    @staticmethod
    def available():
        return os.path.exists(os.path.join(_HERE, "_ref", "libloosdcd.so"))

    def __init__(self):
        self.lib = C.CDLL(os.path.join(_HERE, "_ref", "libloosdcd.so"))
        self.lib.ref_dcd_write.restype = C.c_long
        self.lib.ref_dcd_write.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_char_p, C.c_long]
        self.lib.ref_dcd_read.argtypes = [C.c_char_p, C.c_long, _fp, C.c_long, C.POINTER(C.c_int)]

    def write(self, frames, box=None):
        """DCDWriter::writeFrame over the frames -> the bytes of the file (box: 3 edge lengths or None)."""
        X = np.ascontiguousarray(frames, dtype=np.float64)
        F, N, _ = X.shape
        b = None if box is None else np.array(box, dtype=np.float64)
        cap = 4096 + F * (N * 12 + 256)
        buf = C.create_string_buffer(cap)
        n = self.lib.ref_dcd_write(X.ctypes.data_as(_dp), F, N, None if b is None else b.ctypes.data_as(_dp), buf, cap)
        if n < 0:
            raise RuntimeError("reference DCD writer failed (%d)" % n)
        return buf.raw[:n]

    def read(self, data, max_frames, natoms):
        """DCD::readHeader + seekFrame/parseFrame per frame -> dict(natoms, nframes, crystal, swabbed, xyz (F, natoms, 3) float32)."""
        out = np.zeros((max_frames, 3, natoms), np.float32)
        info = (C.c_int * 4)()
        nf = self.lib.ref_dcd_read(data, len(data), out.ctypes.data_as(_fp), out.size, info)
        if nf < 0:
            raise RuntimeError("reference DCD reader failed (%d)" % nf)
        return {"natoms": info[0], "nframes": info[1], "crystal": bool(info[2]), "swabbed": bool(info[3]),
                "xyz": out[:nf].transpose(0, 2, 1).copy()}


# ---------------------------------------------------------------------------
# Text format of the result matrix: src/MatrixImpl.hpp:210-222 under
# `cout << setprecision(p)` (Tools/rmsds.cpp:566-567).  C++ default floatfield
# with precision p prints like printf("%.{p}g") (p == 0 behaves as 1).
def matrix_text(M, precision=2):
    M = np.asarray(M, dtype=np.float32)
    m, n = M.shape
    lines = ["# %d %d (0)" % (m, n)]
    fmt = "%." + str(max(int(precision), 1) if precision == 0 else int(precision)) + "g"
    for j in range(m):
        lines.append("".join((fmt % float(M[j, i])) + " " for i in range(n)))
    return "\n".join(lines) + "\n"
