"""Host-side numerics of the per-pair solver (loos_b200/csrc/qcp.cuh compiled with g++).

tests/host/qcp_ff_check.cpp builds exact integer covariances the way the tensor kernel's
epilogue sees them and pushes them through the fp32 / float-float fast path
(coefficients from qcp_poly_ff, or from qcp_poly in fp64 then split as k_pairs_tensor does it
-> fp32 Newton -> qcp_polish_f32, rare cases qcp_root_fp64); the reference is a
cyclic-Jacobi eigen-solve of Horn's 4x4 matrix in fp64.  Scenarios: identical / noisy /
unrelated / mirror-image / planar / collinear structures, N = 2 ... 43000.
Tolerance: 1e-5 Angstrom on the RMSD (the path's budget is 1e-4, BASELINE.json north_star).
"""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def report(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("qcp") / "qcp_ff_check")
    src = os.path.join(HERE, "host", "qcp_ff_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, src], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=600).stdout
    rows = {}
    for line in out.splitlines():
        name, cases, dq0, dy, dr, fb, dr2, fb2 = line.split()
        rows[name] = dict(cases=int(cases), dq0=float(dq0), dy=float(dy), drmsd=float(dr), fallbacks=int(fb),
                          drmsd_fp64coef=float(dr2), fallbacks_fp64coef=int(fb2))
    return rows


def test_all_scenarios_within_tolerance(report):
    assert len(report) >= 24
    for name, r in report.items():
        assert r["drmsd"] < 1e-5, (name, r)
        assert r["drmsd_fp64coef"] < 1e-5, (name, r)  # the variant k_pairs_tensor runs


def test_quartic_constant_term_accuracy(report):
    """q0 = (1-F)^2 - 8 det S - 4 G is the coefficient that cancels: float-float must hold it to 1e-13."""
    for name, r in report.items():
        assert r["dq0"] < 1e-13, (name, r)


def test_fast_path_is_the_common_path(report):
    """Generic structures never need the fp64 fallback; degenerate ones always take it."""
    for name in ["identical_N1000", "noise1e-2_N1000", "noise1_N1000", "noise5_N1000", "planar_N64", "noise1_N30"]:
        assert report[name]["fallbacks"] == 0, name
    for name in ["noise1_N2", "collinear_N64", "unrelated_N1000"]:
        assert report[name]["fallbacks"] == report[name]["cases"], name


# ------------------------------------------------------------------ rmsds-align pair arithmetic
def test_align_pair_arithmetic_matches_the_script(tmp_path, checker):
    """tests/host/align_pair_check.cpp restates what one thread of k_pairs_align computes for a pair
    (same qcp.cuh); the oracle restates the loop of Packages/Clustering/rmsds-align.py on the
    reference's kabsch.  Generic, far-off-centre, identical and near-degenerate cases."""
    import ctypes as C

    import numpy as np
    from oracle import oracle
    so = str(tmp_path / "libalignpair.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "host", "align_pair_check.cpp")], check=True)
    lib = C.CDLL(so)
    lib.align_pair.restype = C.c_double
    fp = C.POINTER(C.c_float)
    rng = np.random.default_rng(3)

    def frames(F, n, noise, shift=0.0):
        base = rng.normal(0, 12, (n, 3))
        out = []
        for _ in range(F):
            q = rng.normal(size=4); q /= np.linalg.norm(q); w, x, y, z = q
            Rm = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                           [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                           [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
            out.append(((base + rng.normal(0, noise, base.shape)) @ Rm.T + rng.normal(0, 20, 3) + shift).astype(np.float32))
        return np.stack(out)

    def run(X, Y, Na):
        A1, R1, A2, R2 = X[:, :Na], X[:, Na:], Y[:, :Na], Y[:, Na:]
        want = oracle.rmsds_align(checker, A1, R1, A2, R2)
        got, fb = np.zeros_like(want), 0
        for i in range(len(X)):
            for j in range(len(Y)):
                f = C.c_int()
                p = [np.ascontiguousarray(v).ctypes.data_as(fp) for v in (A1[i], A2[j], R1[i], R2[j])]
                got[i, j] = lib.align_pair(p[0], p[1], Na, p[2], p[3], X.shape[1] - Na, C.byref(f))
                fb += f.value
        if X is Y:      # a frame against itself is the square root of a fully cancelled sum: ~1e-6, not 1e-12
            assert np.abs(np.diag(got) - np.diag(want)).max() < 1e-5
            np.fill_diagonal(got, 0.0); np.fill_diagonal(want, 0.0)
        return np.abs(got - want).max(), fb

    X = frames(6, 140, 1.0); Y = frames(5, 140, 1.0)
    err, fb = run(X, Y, 50)
    assert err < 1e-9 and fb == 0
    X = frames(4, 60, 0.3, shift=400.0)                       # far from the origin: float32 inputs, fp64 centring
    err, fb = run(X, X, 20)
    assert err < 1e-8 and fb == 0
    X = frames(4, 30, 2.0)
    err, fb = run(X, X, 3)                                    # three align atoms: planar, rotation still unique
    assert err < 1e-8
    X = frames(4, 12, 2.0)
    A = X.copy(); A[:, 2:] = A[:, :2].repeat(5, axis=1)      # two distinct align atoms; rmsd atoms on the same axis
    err, fb = run(A, A, 2)
    assert err < 1e-5 and fb == 16
