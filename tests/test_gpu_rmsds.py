"""GPU parity tests for the all-to-all path: CUDA kernels (through the C ABI)
against the CPU oracle on the same seeded inputs, the committed golden fixtures,
and size-independent properties.  Tolerance: |dRMSD| <= 1e-4 Angstrom (north_star);
the observed errors are asserted far tighter where the physics allows."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-4
G = os.path.join(os.path.dirname(__file__), "golden")


def _engines():
    from loos_b200 import capi
    return ["fp64", "tensor"] if capi.lib() is not None else []


@pytest.fixture(scope="module")
def lb():
    import loos_b200
    from loos_b200 import capi
    assert capi.device_count() >= 1, "no sm_100 device: GPU tests must not silently pass"
    return loos_b200


@pytest.mark.parametrize("engine", ["fp64", "tensor"])
def test_golden_matrix(lb, engine):
    g = np.load(os.path.join(G, "rmsds_small.npz"))
    with lb.FrameSet(g["X"]) as fs:
        M, st = lb.rmsds(fs, engine=engine)
    assert M.shape == (53, 53) and M.dtype == np.float32
    err = np.abs(M.astype(np.float64) - g["M"].astype(np.float64))
    assert err.max() <= TOL, err.max()
    assert err.max() < 2e-5
    assert np.all(np.diag(M) == 0.0)
    assert np.array_equal(M, M.T)
    lower = M[np.tril_indices(53, -1)].astype(np.float64)
    assert st["count"] == lower.size
    assert abs(st["max"] - lower.max()) < 1e-12
    assert abs(st["sum"] - lower.sum()) < 1e-6
    with lb.FrameSet(g["X"]) as a, lb.FrameSet(g["Y"]) as b:
        Mx, stx = lb.rmsds_cross(a, b, engine=engine)
    assert np.abs(Mx.astype(np.float64) - g["Mx"]).max() <= 2e-5
    assert stx["count"] == 53 * 31


@pytest.mark.parametrize("engine", ["fp64", "tensor"])
def test_golden_kats(lb, engine):
    g = np.load(os.path.join(G, "kat_pairs.npz"))
    for k in [n[:-4] for n in g.files if n.endswith("_xyz")]:
        with lb.FrameSet(g[k + "_xyz"]) as fs:
            M, _ = lb.rmsds(fs, engine=engine)
        assert abs(float(M[1, 0]) - float(g[k + "_rmsd"])) <= TOL, (k, M[1, 0], g[k + "_rmsd"])
        assert M[0, 1] == M[1, 0] and M[0, 0] == 0 and M[1, 1] == 0


@pytest.mark.parametrize("engine", ["fp64", "tensor"])
@pytest.mark.parametrize("F,N", [(1, 5), (2, 1), (41, 64), (80, 65), (121, 333), (200, 1000), (123, 4500)])
def test_vs_oracle_shapes(lb, checker, engine, F, N):
    from loos_b200.synth import SEED0, synth_frames
    X = synth_frames(F, N, seed=SEED0 + F * 7 + N)
    ref = checker.rmsds_self(X.astype(np.float64), nthreads=8)
    with lb.FrameSet(X) as fs:
        M, st = lb.rmsds(fs, engine=engine)
    err = np.abs(M.astype(np.float64) - ref.astype(np.float64))
    assert err.max() <= TOL, (err.max(), err.mean())
    assert np.array_equal(M, M.T) and np.all(np.diag(M) == 0)


@pytest.mark.parametrize("N", [300, 4100])
def test_default_launch_shape(lb, monkeypatch, N):
    """The tensor engine launches CTA pairs by default (pairs_tensor.cu): same bits as the independent-CTA
    shape asked for explicitly, for the matrix, the cross matrix and the banded text (whose mirror-canonical
    rectangles always run as independent CTAs)."""
    from loos_b200.api import rmsds_text
    from loos_b200.synth import synth_frames
    X = synth_frames(97, N, seed=77)
    Y = synth_frames(45, N, seed=78)
    got = []
    for shape in (None, "1"):
        if shape:
            monkeypatch.setenv("LOOSB200_CLUSTER", shape)
        with lb.FrameSet(X) as fs, lb.FrameSet(Y) as fy:
            M, st = lb.rmsds(fs, engine="tensor")
            C, _ = lb.rmsds_cross(fs, fy, engine="tensor")
            monkeypatch.setenv("LOOSB200_TEXT_BAND_ROWS", "40")
            t, _ = rmsds_text(fs, precision=7, engine="tensor")
            monkeypatch.delenv("LOOSB200_TEXT_BAND_ROWS")
        got.append((M, C, t, st["max"]))
    assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1])
    assert got[0][2] == got[1][2] and got[0][3] == got[1][3]


@pytest.mark.parametrize("engine", ["fp64", "tensor"])
def test_config1_full_parity(lb, checker, engine):
    """BASELINE config 1: 1000 frames x 500 atoms, all 499 500 pairs vs the oracle."""
    from loos_b200.synth import SEED0, synth_frames
    from oracle.oracle import matrix_text
    X = synth_frames(1000, 500, seed=SEED0 + 1)
    ref = checker.rmsds_self(X.astype(np.float64), nthreads=0 if checker.kind == "reference" else 8)
    with lb.FrameSet(X) as fs:
        M, st = lb.rmsds(fs, engine=engine)
    err = np.abs(M.astype(np.float64) - ref.astype(np.float64))
    print("config1 %s: max |dRMSD| = %.3e, mean = %.3e" % (engine, err.max(), err.mean()))
    assert err.max() <= TOL
    assert err.max() < 5e-6          # observed ~1e-6: both sides are at float32 resolution
    # text output at the default precision is identical except where a value sits
    # within float32 rounding of a 2-significant-digit boundary
    ta, tb = matrix_text(M, 2).split(), matrix_text(ref, 2).split()
    diff = sum(x != y for x, y in zip(ta, tb))
    assert len(ta) == len(tb) and diff <= 1e-4 * len(ta), diff


@pytest.mark.parametrize("engine", ["fp64", "tensor"])
def test_selection_gather_and_layouts(lb, checker, engine):
    """Staging gathers the selected atoms exactly like updateGroupCoords + readCoords:
    result must be bit-identical whichever way the same coordinates are supplied."""
    from loos_b200.synth import synth_frames
    X = synth_frames(45, 300, seed=77)
    sel = np.array([i for i in range(300) if i % 6 == 1 or i > 280], dtype=np.uint32)
    with lb.FrameSet(X, sel=sel) as fs:
        M1, _ = lb.rmsds(fs, engine=engine)
    with lb.FrameSet(np.ascontiguousarray(X[:, sel])) as fs:
        M2, _ = lb.rmsds(fs, engine=engine)
    with lb.FrameSet(np.ascontiguousarray(X.transpose(0, 2, 1)), sel=sel, layout="planes") as fs:
        M3, _ = lb.rmsds(fs, engine=engine)
    assert np.array_equal(M1, M2) and np.array_equal(M1, M3)
    ref = checker.rmsds_self(X[:, sel].astype(np.float64), 4)
    assert np.abs(M1 - ref).max() <= 2e-5


@pytest.mark.parametrize("engine", ["fp64", "tensor"])
def test_rigid_motion_invariance(lb, engine):
    """Size-independent property: rotating/translating every frame independently
    leaves the matrix unchanged (to float32 input rounding)."""
    from loos_b200.synth import _rot_from_rotvec, synth_frames
    X = synth_frames(90, 400, seed=5).astype(np.float64)
    rng = np.random.default_rng(11)
    R = _rot_from_rotvec(rng.normal(size=(90, 3)))
    Y = (np.einsum("fij,fnj->fni", R, X) + rng.normal(size=(90, 1, 3)) * 20).astype(np.float32)
    with lb.FrameSet(X.astype(np.float32)) as a, lb.FrameSet(Y) as b:
        Ma, _ = lb.rmsds(a, engine=engine)
        Mb, _ = lb.rmsds(b, engine=engine)
    assert np.abs(Ma - Mb).max() < 5e-5


@pytest.mark.parametrize("engine", ["fp64", "tensor"])
def test_parts_tile_the_triangle(lb, engine):
    """Multi-GPU split: the union of the parts' packed rows is exactly the single-GPU
    lower triangle (bit-identical: same kernel, same pair arithmetic)."""
    from loos_b200.api import rmsds_part
    from loos_b200.synth import synth_frames
    F = 171
    X = synth_frames(F, 120, seed=3)
    with lb.FrameSet(X) as fs:
        M, st = lb.rmsds(fs, engine=engine)
        L = np.zeros((F, F), np.float32)
        tot, cnt, mx = 0.0, 0, 0.0
        for p in range(3):
            rows, blk, s = rmsds_part(fs, p, 3, engine=engine)
            for r, i in enumerate(rows):
                if i != 0xFFFFFFFF:
                    L[i, :i] = blk[r, :i]
            tot += s["sum"]; cnt += s["count"]; mx = max(mx, s["max"])
        _, s0 = lb.rmsds(fs, noout=True, engine=engine)
    assert np.array_equal(np.tril(M, -1), L)
    assert cnt == st["count"] == F * (F - 1) // 2 and mx == st["max"]
    assert abs(tot - st["sum"]) < 1e-6
    assert s0["max"] == st["max"] and abs(s0["sum"] - st["sum"]) < 1e-6


def test_tensor_matches_fp64_engine(lb):
    """The two device engines agree to the fixed-point quantisation (~1e-6 A)."""
    from loos_b200.synth import synth_frames
    X = synth_frames(333, 777, seed=21)
    with lb.FrameSet(X) as fs:
        A, _ = lb.rmsds(fs, engine="tensor")
        B, _ = lb.rmsds(fs, engine="fp64")
        e = fs.unit_log2()
    assert -20 <= e <= -14
    assert np.abs(A.astype(np.float64) - B).max() < 2e-5


def test_device_resident_io(lb):
    import torch
    from loos_b200.synth import synth_frames
    X = synth_frames(130, 96, seed=8)
    with lb.FrameSet(X) as fs:
        M, _ = lb.rmsds(fs)
    xd = torch.from_numpy(X).cuda()
    out = torch.empty((130, 130), dtype=torch.float32, device="cuda")
    with lb.FrameSet(xd) as fs:
        lb.rmsds(fs, out=out)
        ms_main, ms_other, n = fs.last_timing()
    assert n >= 1 and ms_main > 0
    assert np.array_equal(out.cpu().numpy(), M)


def test_errors(lb):
    from loos_b200 import capi
    a = lb.FrameSet(np.zeros((4, 10, 3), np.float32) + np.arange(10)[None, :, None])
    b = lb.FrameSet(np.zeros((4, 11, 3), np.float32))
    with pytest.raises(lb.LoosB200Error) as e:
        lb.rmsds_cross(a, b)
    assert e.value.code == capi.ERR_SIZE_MISMATCH      # LOOSError in the reference
    with pytest.raises(lb.LoosB200Error) as e:
        lb.FrameSet(np.zeros((4, 10, 3), np.float32), sel=np.array([1, 10], np.uint32))
    assert e.value.code == capi.ERR_RANGE
    # all-zero frames (every atom at the centroid): RMSD 0, no NaN
    M, _ = lb.rmsds(b)
    assert np.all(M == 0)


def test_incremental_staging_equals_one_shot(lb):
    """loosb200_stage_begin/append/end (decode overlapped with upload) must stage exactly
    what loosb200_stage_frames stages: ragged pieces, selection gather, both layouts."""
    import ctypes as C

    from loos_b200 import capi
    from loos_b200.synth import synth_frames
    L = capi.lib()
    X = synth_frames(157, 90, seed=31)
    sel = np.arange(0, 90, 2, dtype=np.uint32)
    with lb.FrameSet(X, sel=sel) as fs:
        M0, _ = lb.rmsds(fs)
        c0 = fs.centroids()
    h = C.c_void_p()
    selp = sel.ctypes.data_as(C.POINTER(C.c_uint32))
    capi.check(L.loosb200_stage_begin(C.byref(h), 157, 90, selp, sel.size, capi.LAYOUT_XYZ, 0))
    pos = 0
    for n in (1, 40, 3, 100, 13):
        piece = np.ascontiguousarray(X[pos:pos + n])
        capi.check(L.loosb200_stage_append(h, piece.ctypes.data_as(C.c_void_p), n))
        piece[:] = 7.0          # the caller may reuse its buffer immediately
        pos += n
    assert L.loosb200_stage_append(h, X.ctypes.data_as(C.c_void_p), 1) == capi.ERR_RANGE   # more than announced
    capi.check(L.loosb200_stage_end(h))
    fs2 = lb.FrameSet.__new__(lb.FrameSet)
    fs2._h, fs2.F, fs2.N, fs2.device = h, 157, sel.size, 0
    M1, _ = lb.rmsds(fs2)
    assert np.array_equal(M0, M1) and np.array_equal(c0, fs2.centroids())
    fs2.close()
    # ending early is an error, not a silently short set
    capi.check(L.loosb200_stage_begin(C.byref(h), 10, 90, None, 0, capi.LAYOUT_XYZ, 0))
    capi.check(L.loosb200_stage_append(h, X.ctypes.data_as(C.c_void_p), 4))
    assert L.loosb200_stage_end(h) == capi.ERR_INVALID
    L.loosb200_free(h)


@pytest.mark.parametrize("F,N", [(83, 70), (407, 300), (1201, 129)])
def test_cta_pair_launch_shape_is_bit_identical(lb, F, N, monkeypatch):
    """k_pairs_tensor has two launch shapes: independent CTAs (default) and CTA pairs that share
    the "b" operand through tcgen05 cta_group::2.  Exact integer accumulation and the same
    per-pair solve: the matrices, packed parts and the cross matrix must agree bit for bit
    (odd row-tile counts, masked partner tiles and the diagonal included)."""
    from loos_b200.api import rmsds_part
    from loos_b200.synth import synth_frames
    X = synth_frames(F, N, seed=11 + F)
    Y = synth_frames(F // 2 + 3, N, seed=5 + F)
    got = {}
    for shape in ("1", "2"):
        monkeypatch.setenv("LOOSB200_CLUSTER", shape)
        with lb.FrameSet(X) as fs, lb.FrameSet(Y) as fy:
            M, st = lb.rmsds(fs, engine="tensor")
            rows, blk, _ = rmsds_part(fs, 1, 3, engine="tensor")
            C, _ = lb.rmsds_cross(fs, fy, engine="tensor")
        # a packed row holds M(i, j) for j < i only; the rest of the row is not written
        valid = (np.arange(F)[None, :] < np.asarray(rows, dtype=np.int64)[:, None]) & (np.asarray(rows)[:, None] != 0xFFFFFFFF)
        got[shape] = (M, st["max"], np.where(valid, blk, 0), C)
    assert np.array_equal(got["1"][0], got["2"][0])
    assert got["1"][1] == got["2"][1]
    assert np.array_equal(got["1"][2], got["2"][2])
    assert np.array_equal(got["1"][3], got["2"][3])
    assert (np.diag(got["2"][0]) == 0).all() and np.array_equal(got["2"][0], got["2"][0].T)


@pytest.mark.parametrize("F,N,name", [(20000, 3000, "config3"), (100000, 1000, "config4"), (24000, 5000, "config5 at 24,000 of its 200,000 frames")])
def test_large_configs_sampled_parity_and_properties(lb, checker, F, N, name):
    """BASELINE's large shapes in full (C3: 20 000 x 3 000; C4: 100 000 x 1 000, a 40 GB matrix): frames
    are synthesised on the device, the matrix stays in HBM.  Checked: sampled pairs against the
    oracle (random + the near-diagonal band where RMSD is smallest), symmetry and zero diagonal on
    sampled rows/columns, stats consistent with a float64 recount of sampled rows, and invariance
    under an independent rigid motion of every frame."""
    import torch
    from loos_b200.synth import synth_frames_torch
    X = synth_frames_torch(F, N, seed=20261019 + F, device="cuda")
    out = torch.empty((F, F), dtype=torch.float32, device="cuda")
    with lb.FrameSet(X) as fs:
        _, st = lb.rmsds(fs, out=out, engine="tensor")
    rng = np.random.default_rng(F)
    ii = np.concatenate([rng.integers(1, F, 600), rng.integers(1, F, 300)])
    jj = np.concatenate([rng.integers(0, F, 600), np.maximum(ii[600:] - rng.integers(1, 4, 300), 0)])
    keep = ii != jj
    ii, jj = ii[keep], jj[keep]
    frames = np.unique(np.concatenate([ii, jj]))
    Xs = X[torch.as_tensor(frames, device="cuda")].cpu().numpy().astype(np.float64)
    pos = {int(f): k for k, f in enumerate(frames)}
    got = out[torch.as_tensor(jj, device="cuda"), torch.as_tensor(ii, device="cuda")].cpu().numpy()   # column-major M(i,j)
    cen = np.stack([checker.center_at_origin(x)[0] for x in Xs])          # centerAtOrigin + centeredRMSD
    want = np.array([np.float32(checker.centered_rmsd(cen[pos[int(i)]], cen[pos[int(j)]])) for i, j in zip(ii, jj)])
    err = np.abs(got.astype(np.float64) - want)
    print("%s: %d sampled pairs, max |dRMSD| = %.3e, mean %.3e" % (name, len(ii), err.max(), err.mean()))
    assert err.max() <= TOL and err.max() < 1e-5
    rows = torch.as_tensor(rng.integers(0, F, 64), device="cuda")
    assert torch.equal(out[rows, :], out[:, rows].T)                    # symmetric fill
    assert float(out.diagonal().abs().max()) == 0.0                     # diagonal exactly 0
    assert abs(st["max"] - float(out.max())) == 0.0
    tri = torch.tril(out[:4000, :4000].double(), -1)
    # stats over the leading 4000 frames equal a float64 recount (showStatsHalf, Tools/rmsds.cpp:425-439)
    with lb.FrameSet(X[:4000]) as fs4:
        _, st4 = lb.rmsds(fs4, noout=True, engine="tensor")
    assert st4["count"] == 4000 * 3999 // 2
    assert abs(st4["sum"] - float(tri.sum())) < 1e-6 * st4["sum"]
    del tri
    # rigid motions: rotate/translate every frame independently (float32 inputs: 2e-5 A)
    sub = torch.as_tensor(rng.choice(F, 2000, replace=False), device="cuda")
    Y = X[sub].double()
    q = torch.randn(2000, 4, dtype=torch.float64, device="cuda")
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q.unbind(1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1).reshape(2000, 3, 3)
    Ym = (torch.einsum("fij,fnj->fni", R, Y) + 20 * torch.randn(2000, 1, 3, dtype=torch.float64, device="cuda")).float()
    with lb.FrameSet(X[sub].contiguous()) as fa, lb.FrameSet(Ym.contiguous()) as fb:
        A = torch.empty((2000, 2000), dtype=torch.float32, device="cuda")
        B = torch.empty_like(A)
        lb.rmsds(fa, out=A, engine="tensor")
        lb.rmsds(fb, out=B, engine="tensor")
    assert float((A - B).abs().max()) < 5e-5


def test_atom_limit_of_the_tensor_engine(lb, checker):
    """N = 43 000 is the largest selection whose class sums provably fit int32 (3 digit products of
    at most 2^14 per atom).  Worst case for the accumulators: every centred coordinate at +-max, so all
    three digits of every operand are extreme.  The tensor engine must still be exact (compared with
    the fp64 engine and the oracle).  From 43 001 atoms on the contraction runs in several accumulation
    segments (pairs_tensor.cu, SEG) -- the same worst case with one atom more must agree just as well --
    up to 100 000 atoms; 100 001 is refused by the tensor engine and taken by the fp64 engine under
    engine="auto"."""
    from loos_b200 import capi
    rng = np.random.default_rng(43)
    F, N = 44, 43000
    signs = rng.choice([-1.0, 1.0], size=(F, N, 3))
    signs[:, N // 2:, :] = -signs[:, :N // 2, :]            # centroid exactly 0: max |x - c| = L everywhere
    X = (signs * 99.99).astype(np.float32)
    X[3] = X[2]                                              # an identical pair
    X[5] = -X[4]                                             # and an inverted one
    with lb.FrameSet(X) as fs:
        Mt, st = lb.rmsds(fs, engine="tensor")
        Mf, sf = lb.rmsds(fs, engine="fp64")
    # identical frames: exactly 0 while the integer sums stay below 2^53 (any realistic input); here
    # they reach 2^63 and the conversion to fp64 rounds at 1e-16 relative: a few 1e-5 A at most
    # (the fp64 engine, like the reference, carries the cancellation noise of Ga + Gb - 2 lambda:
    # sqrt(1e-16 * 1.3e9 A^2 * few / N) ~ 2e-4 A for this pair)
    assert Mt[3, 2] < 5e-5 and Mf[3, 2] < 5e-4
    D = np.abs(Mt - Mf)
    D[3, 2] = D[2, 3] = 0.0
    assert D.max() <= 2 * np.spacing(np.float32(Mf.max()))    # RMSDs of ~200 A: two float32 ulps = 3e-5
    Xd = X.astype(np.float64)
    cen = [checker.center_at_origin(x)[0] for x in Xd[:6]]
    for i, j in [(1, 0), (5, 4), (5, 0)]:
        assert abs(float(Mt[i, j]) - checker.centered_rmsd(cen[i], cen[j])) < 1e-4
    assert checker.centered_rmsd(cen[3], cen[2]) < 5e-4      # the reference's own noise floor here
    # 43 002 atoms (two more, +L and -L: the centroid stays 0): two segments, 42 880 + 122 atoms
    extra = np.zeros((9, 2, 3), np.float32)
    extra[:, 0, :] = 99.99
    extra[:, 1, :] = -99.99
    X1 = np.concatenate([X[:9], extra], axis=1)
    with lb.FrameSet(X1) as fs:
        Ms, _ = lb.rmsds(fs, engine="tensor")
        Mf1, _ = lb.rmsds(fs, engine="fp64")
        Ma, _ = lb.rmsds(fs, engine="auto")
    D1 = np.abs(Ms - Mf1)
    D1[3, 2] = D1[2, 3] = 0.0
    assert D1.max() <= 2 * np.spacing(np.float32(Mf1.max())) and Ms[3, 2] < 5e-5
    assert np.array_equal(Ma, Ms)
    X2 = np.zeros((6, 100001, 3), np.float32)
    X2[:] = rng.normal(scale=30.0, size=(6, 100001, 3))
    with lb.FrameSet(X2) as fs:
        with pytest.raises(lb.LoosB200Error) as e:
            lb.rmsds(fs, engine="tensor")
        assert e.value.code == capi.ERR_UNSUPPORTED
        Ma, _ = lb.rmsds(fs, engine="auto")
        Mf2, _ = lb.rmsds(fs, engine="fp64")
    assert np.array_equal(Ma, Mf2)


@pytest.mark.parametrize("N", [50000, 86000, 100000])
def test_selections_above_43000_atoms_in_segments(lb, checker, N):
    """Tensor engine beyond one int32 accumulation segment: 2 or 3 segments (42 880 atoms each), the last one
    short or full; against the fp64 engine on the whole matrix, the oracle on a few pairs, the banded
    (mirror-canonical) text against the whole-matrix text, and with frames that exist twice."""
    from loos_b200.api import rmsds_text
    from loos_b200.synth import synth_frames
    F = 83
    X = synth_frames(F, N, seed=N)
    X[70:75] = X[10:15]
    with lb.FrameSet(X) as fs:
        Mt, st = lb.rmsds(fs, engine="tensor")
        Mf, sf = lb.rmsds(fs, engine="fp64")
        whole, _ = rmsds_text(fs, precision=7, engine="tensor")
    assert np.abs(Mt.astype(np.float64) - Mf.astype(np.float64)).max() < 2e-5
    assert np.array_equal(Mt, Mt.T) and np.all(np.diag(Mt) == 0)
    # frames that exist twice: 0 up to the one rounding of the integer sums (> 2^53 at this size) to fp64
    assert Mt[72, 12] < 5e-5 and Mt[70, 10] < 5e-5
    assert st["count"] == F * (F - 1) // 2 and abs(st["sum"] - sf["sum"]) < 1e-3
    Xd = X[:4].astype(np.float64)
    cen = [checker.center_at_origin(x)[0] for x in Xd]
    for i, j in [(1, 0), (3, 2), (3, 0)]:
        assert abs(float(Mt[i, j]) - checker.centered_rmsd(cen[i], cen[j])) < 2e-5
    import os
    os.environ["LOOSB200_TEXT_BAND_ROWS"] = "40"
    try:
        with lb.FrameSet(X) as fs:
            banded, _ = rmsds_text(fs, precision=7, engine="tensor")
    finally:
        del os.environ["LOOSB200_TEXT_BAND_ROWS"]
    assert banded == whole


@pytest.mark.parametrize("precision", [0, 2, 6, 9])
def test_matrix_text_formatted_on_the_gpu(lb, precision):
    """loosb200_rmsds_self_text / _cross_text: the text leaves the GPU already formatted.  Same bytes
    as the reference's `cout << setprecision(p) << M` (checked through the oracle's matrix_text of
    the float32 matrix the library returns), for the symmetric and the rectangular matrix, with
    rows longer than one chunk, zeros on the diagonal and values needing every notation."""
    from loos_b200.api import rmsds_text
    from loos_b200.synth import synth_frames
    from oracle.oracle import matrix_text
    X = synth_frames(1100, 40, seed=5)              # 1100 values per row: two chunks
    X[7] = X[6]                                     # an exact 0 off the diagonal
    X[9] = X[8] * (1 + 1e-6)                        # a tiny RMSD (scientific notation at low precision)
    X[11] = X[10] * 40.0                            # a large one
    Y = synth_frames(37, 40, seed=6)
    with lb.FrameSet(X) as fs, lb.FrameSet(Y) as fy:
        M, st = lb.rmsds(fs, engine="tensor")
        text, st2 = rmsds_text(fs, precision=precision, engine="tensor")
        Cm, _ = lb.rmsds_cross(fs, fy, engine="tensor")
        ctext, _ = rmsds_text(fs, fy, precision=precision, engine="tensor")
    assert st2["max"] == st["max"] and st2["count"] == st["count"]
    assert text.decode() == matrix_text(M, precision)
    assert ctext.decode() == matrix_text(Cm, precision)


@pytest.mark.parametrize("engine", ["tensor", "fp64"])
def test_matrix_text_in_row_bands(lb, monkeypatch, engine):
    """When the F x F float matrix does not fit in HBM the text is produced from bands of rows computed
    as rectangles (LOOSB200_TEXT_BAND_ROWS forces it here): identical bytes, identical stats, exact
    zeros on the diagonal, for the self and the cross matrix, with a ragged last band."""
    from loos_b200.api import rmsds_text
    from loos_b200.synth import synth_frames
    X = synth_frames(333, 50, seed=15)
    # frames met twice (multi-rmsds given one trajectory twice): their RMSD is rounding noise of the
    # solve, and an entry above the diagonal must still be the float its mirror image is
    X[200:260] = X[0:60]
    X[300:320] = X[100:120] + np.float32(3.0)
    Y = synth_frames(61, 50, seed=16)
    with lb.FrameSet(X) as fs, lb.FrameSet(Y) as fy:
        whole, st = rmsds_text(fs, precision=7, engine=engine)
        cwhole, cst = rmsds_text(fs, fy, precision=7, engine=engine)
        monkeypatch.setenv("LOOSB200_TEXT_BAND_ROWS", "120")
        banded, bst = rmsds_text(fs, precision=7, engine=engine)
        cbanded, cbst = rmsds_text(fs, fy, precision=7, engine=engine)
    if banded != whole:
        wl, bl = whole.split(b"\n"), banded.split(b"\n")
        bad = [(r, c, x, y) for r, (u, v) in enumerate(zip(wl, bl)) if u != v for c, (x, y) in enumerate(zip(u.split(), v.split())) if x != y]
        raise AssertionError("banded text differs from the whole matrix at (line, column, whole, banded): %r" % bad[:8])
    assert cbanded == cwhole
    assert bst["count"] == st["count"] and bst["max"] == st["max"] and abs(bst["sum"] - st["sum"]) < 1e-6
    assert cbst["count"] == cst["count"] and cbst["max"] == cst["max"] and abs(cbst["sum"] - cst["sum"]) < 1e-6 * cst["sum"]


@pytest.mark.parametrize("top", [127.9, 63.9, 255.7, 127.4999])
def test_fixed_point_range_top_of_a_power_of_two_band(lb, top):
    """Centred coordinates in the top 0.4 % of a power-of-two band (127.50-127.996 A at unit 2^-16,
    63.75-64 A at 2^-17, ...) exceed what three balanced base-256 digits hold (127 * 65793); the unit
    must step up instead of letting the top digit wrap to -128 (stage.cu, k_unit / k_quantise)."""
    rng = np.random.default_rng(int(top * 1000))
    F, N = 24, 96
    X = rng.normal(scale=12.0, size=(F, N, 3))
    X -= X.mean(axis=1, keepdims=True)
    # two atoms on opposite sides keep the centroid at the origin: centred x of atom 0 is exactly `top`
    X[:, 0, :] = 0.0
    X[:, 1, :] = 0.0
    X[:, 0, 0] = top
    X[:, 1, 0] = -top
    X = X.astype(np.float32)
    with lb.FrameSet(X) as fs:
        e = fs.unit_log2()
        assert top / 2.0 ** e <= 127 * 65793, (top, e)
        Mt, _ = lb.rmsds(fs, engine="tensor")
        Md, _ = lb.rmsds(fs, engine="fp64")
    assert np.abs(Mt.astype(np.float64) - Md.astype(np.float64)).max() <= 2e-5
    # rectangle mode against a set with a smaller extent: the common (coarser) unit is forced on it
    Y = (rng.normal(scale=3.0, size=(F, N, 3))).astype(np.float32)
    with lb.FrameSet(X) as a, lb.FrameSet(Y) as b:
        Ct, _ = lb.rmsds_cross(a, b, engine="tensor")
        Cd, _ = lb.rmsds_cross(a, b, engine="fp64")
    assert np.abs(Ct.astype(np.float64) - Cd.astype(np.float64)).max() <= 2e-5


def test_non_finite_coordinates_are_not_absorbed(lb):
    """A NaN / Inf coordinate has no fixed-point image: the tensor engine refuses the set
    (NumericalError in the reference's dgesvj path, src/alignment.cpp:66-67) and AUTO falls back to the
    fp64 engine, which propagates NaN into every entry that involves the bad frame."""
    from loos_b200 import capi
    from loos_b200.synth import synth_frames
    X = synth_frames(50, 64, seed=8).copy()
    X[17, 5, 1] = np.nan
    with lb.FrameSet(X) as fs:
        with pytest.raises(lb.LoosB200Error) as ei:
            lb.rmsds(fs, engine="tensor")
        assert ei.value.code == capi.ERR_NUMERICAL
        M, _ = lb.rmsds(fs, engine="auto")
    bad = np.isnan(M)
    assert bad[17, :17].all() and bad[:17, 17].all() and bad[18:, 17].all()
    good = np.delete(np.delete(M, 17, axis=0), 17, axis=1)
    assert np.isfinite(good).all()
    X[17, 5, 1] = np.inf
    with lb.FrameSet(X) as fs:
        with pytest.raises(lb.LoosB200Error):
            lb.rmsds(fs, engine="tensor")


def test_centroids_match_center_at_origin(lb, checker):
    """a3: the staged centroids against alignment::centerAtOrigin (src/alignment.cpp:90-109) directly."""
    from loos_b200.synth import synth_frames
    X = synth_frames(200, 333, seed=12)
    sel = np.arange(1, 333, 3, dtype=np.uint32)
    with lb.FrameSet(X, sel=sel) as fs:
        c = fs.centroids()
    want = np.array([checker.center_at_origin(X[f][sel].astype(np.float64))[1] for f in range(200)])
    assert np.abs(c - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
