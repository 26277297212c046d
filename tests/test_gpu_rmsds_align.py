"""GPU parity tests for the all-to-all RMSD with different alignment and RMSD selections
(loosb200_rmsds_align; Packages/Clustering/rmsds-align.py:236-256).  The checker is the oracle's
restatement of the script's loop on the reference's own alignment::kabsch."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-4      # north_star's bound on the RMSD
# what each engine actually gives on generic structures: fp64 contractions + the fp64 rotation, or exact
# integer contractions of coordinates on a 2^-16..2^-17 A grid (rotation in fp64 as well)
TIGHT = {"fp64": 1e-8, "tensor": 1e-5}
ENGINES = ["fp64", "tensor"]


@pytest.fixture(scope="module")
def lb():
    import loos_b200
    from loos_b200 import capi
    assert capi.device_count() >= 1
    return loos_b200


def _sets(F, natoms, seed):
    from loos_b200.synth import synth_frames
    return synth_frames(F, natoms, seed=seed)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("F1,F2,Na,Nr", [(45, 53, 60, 90), (1, 1, 3, 1), (40, 80, 17, 200), (83, 7, 500, 33), (130, 97, 257, 1025),
                                         (9, 7, 45000, 50000)])   # both selections beyond one accumulation segment
def test_two_trajectories_vs_oracle(lb, checker, F1, F2, Na, Nr, engine):
    from oracle import oracle
    X = _sets(F1, Na + Nr, seed=77 + F1 + Na)
    Y = _sets(F2, Na + Nr, seed=78 + F2 + Nr)
    sa = np.arange(0, Na, dtype=np.uint32)
    sr = np.arange(Na, Na + Nr, dtype=np.uint32)
    want = oracle.rmsds_align(checker, X[:, sa], X[:, sr], Y[:, sa], Y[:, sr])
    with lb.FrameSet(X, sel=sa) as a1, lb.FrameSet(X, sel=sr) as r1, lb.FrameSet(Y, sel=sa) as a2, lb.FrameSet(Y, sel=sr) as r2:
        got = lb.rmsds_align(a1, r1, a2, r2, engine=engine)
    assert got.shape == (F1, F2) and got.dtype == np.float64
    err = np.abs(got - want).max()
    assert err <= TOL
    if Na >= 4:
        assert err < TIGHT[engine], err


@pytest.mark.parametrize("engine", ENGINES)
def test_one_trajectory_overlapping_selections_and_bands(lb, checker, monkeypatch, engine):
    """The one-trajectory call (second pair of handles = the first), selections that overlap, and the
    result produced in several row bands: identical numbers."""
    from oracle import oracle
    F, N = 130, 120
    X = _sets(F, N, seed=4)
    sa = np.arange(0, 80, 2, dtype=np.uint32)          # 40 atoms
    sr = np.arange(30, 120, dtype=np.uint32)           # 90 atoms, overlaps the align selection
    want = oracle.rmsds_align(checker, X[:, sa], X[:, sr], X[:, sa], X[:, sr])
    with lb.FrameSet(X, sel=sa) as a1, lb.FrameSet(X, sel=sr) as r1:
        whole = lb.rmsds_align(a1, r1, engine=engine)
        monkeypatch.setenv("LOOSB200_ALIGN_BAND_TILES", "1")
        banded = lb.rmsds_align(a1, r1, engine=engine)
    assert np.array_equal(whole, banded)
    assert np.abs(np.diag(whole)).max() < 1e-5          # a frame against itself: 0 up to the cancellation floor
    off = ~np.eye(F, dtype=bool)
    assert np.abs(whole - want)[off].max() < TIGHT[engine]
    # one trajectory, the same selections on both sides: Z_ji = Z_ij^T, so the matrix is symmetric
    # (it is not with --align2 / --rmsd2 or a second trajectory)
    assert np.abs(whole - whole.T).max() < 1e-9


@pytest.mark.parametrize("engine", ENGINES)
def test_same_selection_reduces_to_rmsds(lb, engine):
    """With the rmsd selection equal to the align selection the matrix is the minimal RMSD that
    rmsds computes (float32 there)."""
    F, N = 90, 150
    X = _sets(F, N, seed=9)
    Y = _sets(70, N, seed=10)
    with lb.FrameSet(X) as a, lb.FrameSet(Y) as b:
        M = lb.rmsds_align(a, None, b, None, engine=engine)
        ref, _ = lb.rmsds_cross(a, b, engine="fp64")
    assert np.abs(M - ref).max() < (2e-6 if engine == "fp64" else 1e-5)


@pytest.mark.parametrize("engine", ENGINES)
def test_degenerate_align_selection(lb, checker, engine):
    """Two align atoms (rotation about their axis is free in the reference too): the result is still the
    RMSD under A valid optimal superposition -- checked on the part that is determined, the align atoms
    themselves -- and stays finite everywhere."""
    X = _sets(12, 40, seed=21)
    sa = np.array([3, 17], dtype=np.uint32)
    with lb.FrameSet(X, sel=sa) as a1:
        M = lb.rmsds_align(a1, None, engine=engine)
    from oracle import oracle
    want = oracle.rmsds_align(checker, X[:, sa], X[:, sa], X[:, sa], X[:, sa])
    assert np.isfinite(M).all()
    assert np.abs(M - want).max() < (1e-5 if engine == "fp64" else 5e-5)


def test_errors(lb):
    X = _sets(10, 30, seed=1)
    with lb.FrameSet(X, sel=np.arange(10, dtype=np.uint32)) as a10, lb.FrameSet(X, sel=np.arange(12, dtype=np.uint32)) as a12, \
            lb.FrameSet(X[:8], sel=np.arange(10, dtype=np.uint32)) as short:
        with pytest.raises(lb.LoosB200Error):
            lb.rmsds_align(a10, None, a12, None)        # align selections differ in size
        with pytest.raises(lb.LoosB200Error):
            lb.rmsds_align(a10, a10, a10, a12)          # rmsd selections differ in size
        with pytest.raises(lb.LoosB200Error):
            lb.rmsds_align(a10, short, a10, a10)        # align / rmsd sets of one trajectory differ in frames


@pytest.mark.parametrize("engine", ENGINES)
def test_golden(lb, engine):
    """Against the committed fixture made with the reference's own kabsch (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rmsds_align_small.npz"))
    X, Y = g["X"], g["Y"]
    sa, sr = g["sel_a"].astype(np.uint32), g["sel_r"].astype(np.uint32)
    with lb.FrameSet(X, sel=sa) as a1, lb.FrameSet(X, sel=sr) as r1, lb.FrameSet(Y, sel=sa) as a2, lb.FrameSet(Y, sel=sr) as r2:
        M = lb.rmsds_align(a1, r1, a2, r2, engine=engine)
        M1 = lb.rmsds_align(a1, r1, engine=engine)
    assert np.abs(M - g["M"]).max() < TIGHT[engine]
    off = ~np.eye(len(X), dtype=bool)
    assert np.abs(M1 - g["M_self"])[off].max() < TIGHT[engine]
    assert np.abs(np.diag(M1)).max() < 1e-5


def test_tensor_engine_far_from_the_align_centroid(lb, checker):
    """The rmsd selection is put on its fixed-point grid about the ALIGN centroid: a ligand 150 A away from
    a small align selection makes that grid coarser (the unit follows max |x - c_align|), and the engines
    must still agree within the grid."""
    from oracle import oracle
    F, Na, Nr = 50, 40, 64
    X = _sets(F, Na + Nr, seed=123)
    X[:, Na:] += np.float32(150.0)
    sa, sr = np.arange(Na, dtype=np.uint32), np.arange(Na, Na + Nr, dtype=np.uint32)
    want = oracle.rmsds_align(checker, X[:, sa], X[:, sr], X[:, sa], X[:, sr])
    with lb.FrameSet(X, sel=sa) as a1, lb.FrameSet(X, sel=sr) as r1:
        Mt = lb.rmsds_align(a1, r1, engine="tensor")
        Md = lb.rmsds_align(a1, r1, engine="fp64")
    off = ~np.eye(F, dtype=bool)
    assert np.abs(Md - want)[off].max() < 1e-7
    assert np.abs(Mt - want)[off].max() < 1e-4
    assert np.abs(np.diag(Mt)).max() < 1e-4


def test_default_entry_point_is_the_fp64_engine(lb):
    """loosb200_rmsds_align (no engine argument) keeps the double-precision numerics."""
    import ctypes as C
    from loos_b200 import capi
    X = _sets(30, 70, seed=2)
    sa, sr = np.arange(30, dtype=np.uint32), np.arange(30, 70, dtype=np.uint32)
    with lb.FrameSet(X, sel=sa) as a1, lb.FrameSet(X, sel=sr) as r1:
        want = lb.rmsds_align(a1, r1, engine="fp64")
        out = np.zeros((30, 30), dtype=np.float64, order="F")
        capi.check(capi.lib().loosb200_rmsds_align(a1._h, r1._h, a1._h, r1._h, out.ctypes.data_as(C.POINTER(C.c_double)), 30))
    assert np.array_equal(out, want)
