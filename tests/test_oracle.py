"""CPU tests: the oracle restatement (oracle/oracle.c) against the reference's own
arithmetic (oracle/_ref) and against the committed golden fixtures."""
import os

import numpy as np
import pytest

from loos_b200.synth import SEED0, synth_frames
from oracle.oracle import matrix_text

G = os.path.join(os.path.dirname(__file__), "golden")


def test_port_matches_golden_matrix(port):
    g = np.load(os.path.join(G, "rmsds_small.npz"))
    M = port.rmsds_self(g["X"].astype(np.float64), nthreads=2)
    assert M.dtype == np.float32 and M.shape == (53, 53)
    assert np.abs(M - g["M"]).max() < 2e-6      # float32 rounding of ~1e-12 differences
    assert np.all(np.diag(M) == 0.0)            # Tools/rmsds.cpp:359-365 never writes it
    assert np.array_equal(M, M.T)
    Mx = port.rmsds_cross(g["X"].astype(np.float64), g["Y"].astype(np.float64), nthreads=2)
    assert np.abs(Mx - g["Mx"]).max() < 2e-6


def test_port_matches_golden_kats(port):
    g = np.load(os.path.join(G, "kat_pairs.npz"))
    for k in [n[:-4] for n in g.files if n.endswith("_xyz")]:
        d = port.rmsds_self(g[k + "_xyz"].astype(np.float64))[1, 0]
        # near-zero cases sit on the reference's own cancellation noise floor
        # (|E0 - 2 ss| of two ~1e6 numbers), hence the absolute tolerance
        assert abs(float(d) - float(g[k + "_rmsd"])) < 5e-6, k


def test_port_matches_golden_rmsd2ref(port):
    g = np.load(os.path.join(G, "rmsd2ref_small.npz"))
    Z = g["Z"].astype(np.float64)
    F = len(Z) - 1
    d, xf = port.rmsd2ref_target(Z[:F, g["sel_a"]], Z[:F, g["sel_r"]], Z[-1][g["sel_a"]], Z[-1][g["sel_r"]])
    assert np.abs(d - g["rmsd"]).max() < 1e-9
    assert np.abs(xf - g["xforms"]).max() < 1e-9
    out, xf, avg, it, rms = port.rmsd2ref_iterative(Z[:F, g["sel_a"]], Z[:F, g["sel_r"]])
    assert it == int(g["it_iters"])
    assert np.abs(out - g["it_rmsd"]).max() < 1e-8
    assert np.abs(avg - g["it_avg"]).max() < 1e-8


def test_port_vs_reference_random(port, reference):
    X = synth_frames(40, 211, seed=SEED0 + 7).astype(np.float64)
    A = port.rmsds_self(X, 2)
    B = reference.rmsds_self(X, 2)
    assert np.abs(A - B).max() < 2e-6
    rng = np.random.default_rng(3)
    for _ in range(50):
        i, j = rng.integers(0, 40, 2)
        u, _ = port.center_at_origin(X[i]); v, _ = port.center_at_origin(X[j])
        su, sv = port.kabsch_core_S(u, v), reference.kabsch_core_S(u, v)
        assert np.abs(su - sv).max() < 1e-9 * max(1.0, np.abs(sv).max())
        assert np.abs(port.kabsch(X[i], X[j]) - reference.kabsch(X[i], X[j])).max() < 1e-9


def test_center_at_origin_matches_reference(port, reference):
    v = synth_frames(1, 77, seed=5)[0].astype(np.float64).ravel()
    a, ca = port.center_at_origin(v)
    b, cb = reference.center_at_origin(v)
    assert np.array_equal(a, b) and np.array_equal(ca, cb)   # same sequential fp64 sums


def test_reflection_branch(port):
    # mirror image: optimal proper rotation cannot undo it -> S[2] flips sign
    a = synth_frames(1, 50, seed=9)[0].astype(np.float64)
    a -= a.mean(0)
    m = a * [1, 1, -1]
    S = port.kabsch_core_S(a.ravel(), m.ravel())
    assert S[2] < 0
    d = port.centered_rmsd(a.ravel(), m.ravel())
    assert d > 1.0


def test_matrix_text_format():
    # src/MatrixImpl.hpp:210-222 under setprecision(p): "%.{p}g " per element
    M = np.array([[0, 1.234567, 123.456], [0.0571, 12, 1e-5]], dtype=np.float32)
    t = matrix_text(M, 2)
    assert t == "# 2 3 (0)\n0 1.2 1.2e+02 \n0.057 12 1e-05 \n"
    g = np.load(os.path.join(G, "rmsds_small.npz"))
    assert str(g["text_p2"]) == matrix_text(g["M"], 2)
    assert str(g["text_p8"]).splitlines()[0] == "# 53 53 (0)"


def test_rmsds_align_restatement_matches_golden(port):
    """oracle.rmsds_align on the C restatement's kabsch against the fixture made with the reference's kabsch
    (tests/golden/make_golden.py::align_golden), and the Python float text of the script."""
    from oracle.oracle import py_round_text, rmsds_align
    g = np.load(os.path.join(G, "rmsds_align_small.npz"))
    X, Y, sa, sr = g["X"], g["Y"], g["sel_a"], g["sel_r"]
    M = rmsds_align(port, X[:, sa], X[:, sr], Y[:, sa], Y[:, sr])
    assert np.abs(M - g["M"]).max() < 1e-9
    assert py_round_text(g["M"], 3) == str(g["text_p3"])
    assert py_round_text(M, 3) == str(g["text_p3"])


def test_restated_iterative_loop_matches_the_references_own_iterative_alignment(reference):
    """rmsd2ref's default mode calls the trajectory-based iterativeAlignment (src/alignment.cpp:355-404), which needs
    AtomicGroup / Trajectory and is therefore RESTATED in oracle/ref_wrapper.cpp on top of the reference's compiled
    kabsch.  The reference's vector-based overload (:288-318) is the same algorithm and IS compiled in place: both must
    give the same iteration count, final rms and transforms."""
    from oracle.oracle import reference_iterative_vecmatrix
    for F, N, seed in [(40, 90, 1), (25, 30, 2), (60, 200, 3), (12, 4, 4)]:
        Z = synth_frames(F, N, seed=seed).astype(np.float64)
        _, xf, _, it, rms = reference.rmsd2ref_iterative(Z, Z, 1e-6, 1000)
        xf2, rms2, it2 = reference_iterative_vecmatrix(reference, Z, 1e-6, 1000)
        assert it == it2
        assert abs(rms - rms2) < 1e-12
        assert np.abs(xf - xf2).max() < 1e-10


def test_stats_lines_against_the_references_own_functions():
    """tests/golden/stats_ref.json: what the reference's own showStatsHalf / showStatsWhole (Tools/rmsds.cpp:425-456) and
    showStats / showFractionalStats (Tools/rms-overlap.cpp:458-521, brace-less `if` included) print for six matrices,
    compiled from the tool sources where they lie (oracle/_ref/libloosstats.so).  The oracle's restatements, which the GPU
    tool tests hold the tools' stderr / stdout to, reproduce every line."""
    import json

    from oracle.oracle import RefStats, fractional_stats_line, stats_line_half, stats_line_whole
    cases = json.load(open(os.path.join(G, "stats_ref.json")))
    live = RefStats() if RefStats.available() else None
    for rec in cases:
        M = np.array(rec["M"], dtype=np.float32)
        assert stats_line_whole(M) == rec["whole"] == rec["overlap"]
        for c, line in rec["fractional"].items():
            assert fractional_stats_line(M, float(c)) == line
            if live:
                assert live.fractional(M, float(c)) == line
        if "half" in rec:
            S = np.array(rec["S"], dtype=np.float32)
            assert stats_line_half(S) == rec["half"]
            if live:
                assert live.half(S) == rec["half"]
