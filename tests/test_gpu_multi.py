"""Several GPUs in one process (loosb200_rmsds_*_multi): the matrix assembled from the devices' parts
must be BIT-IDENTICAL to the one-device result -- same kernels, same per-pair arithmetic, only the
split and the delivery differ (Tools/rmsds.cpp:387-419: np workers fill one matrix).  With one
visible device the replicas share it, which still exercises the split, the strips, the copies and
the turnstile of the text path; with more, replica k lives on device k."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lb():
    import loos_b200
    from loos_b200 import capi
    assert capi.device_count() >= 1, "no sm_100 device: GPU tests must not silently pass"
    return loos_b200


def _replicas(lb, X, n, **kw):
    from loos_b200 import capi
    nd = capi.device_count()
    first = lb.FrameSet(X, **kw)
    return [first] + [first.replicate(k % nd) for k in range(1, n)]


def _close(sets):
    for s in sets:
        s.close()


@pytest.mark.parametrize("engine", ["tensor", "fp64"])
@pytest.mark.parametrize("n", [2, 3, 8])
def test_multi_host_matrix_is_bit_identical(lb, engine, n):
    from loos_b200.synth import synth_frames
    F, N = (1234, 300) if engine == "tensor" else (333, 120)
    X = synth_frames(F, N, seed=77)
    reps = _replicas(lb, X, n)
    try:
        M1, s1 = lb.rmsds(reps[0], engine=engine)
        pm = lb.PinnedMatrix((F, F))
        pm.array.fill(-7.0)
        Mn, sn = lb.rmsds_multi(reps, out=pm.array, engine=engine)
        assert np.array_equal(M1, Mn)
        assert np.array_equal(Mn, Mn.T) and np.all(np.diag(Mn) == 0.0)
        assert sn["count"] == s1["count"] == F * (F - 1) // 2
        assert sn["max"] == s1["max"]
        assert abs(sn["sum"] - s1["sum"]) <= 1e-9 * s1["sum"]
        _, s0 = lb.rmsds_multi(reps, noout=True, engine=engine)
        assert s0["count"] == s1["count"] and s0["max"] == s1["max"]
        # pageable destination too
        Mp, _ = lb.rmsds_multi(reps, engine=engine)
        assert np.array_equal(M1, Mp)
        pm.free()
    finally:
        _close(reps)


@pytest.mark.parametrize("gather", ["store", "copy"])
def test_multi_gather_into_rank0_hbm(lb, gather, monkeypatch):
    """Row blocks gathered into the HBM of replicas[0]'s device: peer stores from the kernels, or
    strips pushed by the copy engines."""
    import torch
    from loos_b200.synth import synth_frames
    monkeypatch.setenv("LOOSB200_GATHER", gather)
    F, N = 2011, 257
    X = synth_frames(F, N, seed=5)
    reps = _replicas(lb, X, 4)
    try:
        M1, s1 = lb.rmsds(reps[0])
        out = torch.full((F, F), -3.0, dtype=torch.float32, device="cuda:0")
        _, sn = lb.rmsds_multi(reps, out=out)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), M1)  # symmetric: memory order immaterial
        assert sn["count"] == s1["count"] and sn["max"] == s1["max"]
    finally:
        _close(reps)


def test_multi_cross_is_bit_identical(lb):
    from loos_b200.synth import synth_frames
    A = synth_frames(517, 200, seed=1)
    B = synth_frames(301, 200, seed=2)
    ra, rb = _replicas(lb, A, 3), _replicas(lb, B, 3)
    try:
        M1, s1 = lb.rmsds_cross(ra[0], rb[0])
        Mn, sn = lb.rmsds_cross_multi(ra, rb)
        assert np.array_equal(M1, Mn)
        assert sn["count"] == s1["count"] == 517 * 301 and sn["max"] == s1["max"]
    finally:
        _close(ra + rb)


@pytest.mark.parametrize("precision", [2, 6])
def test_multi_text_same_bytes(lb, precision, monkeypatch):
    from loos_b200.synth import synth_frames
    F, N = 423, 150
    X = synth_frames(F, N, seed=9)
    Y = synth_frames(200, N, seed=10)
    reps, rb = _replicas(lb, X, 3), _replicas(lb, Y, 3)
    try:
        t1, s1 = lb.rmsds_text(reps[0], precision=precision)
        monkeypatch.setenv("LOOSB200_TEXT_BAND_ROWS", "80")  # six bands over three devices
        tn, sn = lb.rmsds_text_multi(reps, precision=precision)
        assert tn == t1
        assert sn["count"] == s1["count"] and sn["max"] == s1["max"]
        c1, cs1 = lb.rmsds_text(reps[0], rb[0], precision=precision)
        cn, csn = lb.rmsds_text_multi(reps, rb, precision=precision)
        assert cn == c1
        assert csn["count"] == cs1["count"] == F * 200 and csn["max"] == cs1["max"]
    finally:
        _close(reps + rb)


def test_host_strips_equal_device_matrix(lb):
    """One device: the host-output path (bands of compact strips + 2-D copies) against the matrix the
    kernel writes directly into HBM."""
    import torch
    from loos_b200.synth import synth_frames
    F, N = 1777, 128
    X = synth_frames(F, N, seed=3)
    with lb.FrameSet(X) as fs:
        Mh, _ = lb.rmsds(fs)
        out = torch.empty((F, F), dtype=torch.float32, device="cuda:0")
        lb.rmsds(fs, out=out)
        torch.cuda.synchronize()
        assert np.array_equal(Mh, out.cpu().numpy())
        # leading dimension larger than F
        big = np.full((F + 13, F), -1.0, dtype=np.float32, order="F")
        import ctypes as C
        from loos_b200 import capi
        capi.check(capi.lib().loosb200_rmsds_self(fs._h, big.ctypes.data_as(C.c_void_p), F + 13, None, 0))
        assert np.array_equal(big[:F], Mh) and np.all(big[F:] == -1.0)


def test_tools_gpus_option(lb, tmp_path):
    """rmsds / multi-rmsds --gpus N print the same bytes as --gpus 1 (skipped on a one-GPU box: the
    tools place replica k on device k)."""
    from loos_b200 import capi
    if capi.device_count() < 2:
        pytest.skip("needs two devices")
    from loos_b200.synth import synth_system, write_dcd, write_pdb
    pdb, dcd = str(tmp_path / "m.pdb"), str(tmp_path / "t.dcd")
    atoms, xyz = synth_system(50, 260, seed=31, flavour="rich")
    write_pdb(pdb, atoms, xyz[0])
    write_dcd(dcd, xyz)
    subprocess.run(["make", "-C", os.path.join(HERE, "..", "loos_b200", "frontend")], check=True, stdout=subprocess.DEVNULL)
    tool = os.path.join(HERE, "..", "loos_b200", "bin", "rmsds")
    outs = []
    for g in ("1", "2"):
        r = subprocess.run([tool, "--gpus", g, "--sel1", "!hydrogen", "--stats", "1", "-p", "4", pdb, dcd], capture_output=True, check=True)
        outs.append((r.stdout.split(b"\n", 1)[1], r.stderr))  # first line = the command line itself
        mr = subprocess.run([tool.replace("rmsds", "multi-rmsds"), "--gpus", g, "--stats", "1", pdb, dcd, dcd], capture_output=True, check=True)
        outs[-1] += (mr.stdout.split(b"\n", 1)[1], mr.stderr)
    assert outs[0] == outs[1]


def test_stage_multi_equals_stage_and_replicate(lb):
    """loosb200_stage_frames_multi: every device uploads its own slice of the frames, the slices are
    exchanged device-to-device -- the replicas must be indistinguishable from a staged + replicated set."""
    from loos_b200 import capi
    from loos_b200.synth import synth_frames
    nd = capi.device_count()
    F, natoms = 1500, 400
    X = synth_frames(F, natoms, seed=21)
    sel = np.arange(3, natoms, 2, dtype=np.uint32)
    devices = [k % nd for k in range(3)]
    reps = lb.stage_multi(X, devices, sel=sel)
    try:
        with lb.FrameSet(X, sel=sel) as one:
            M1, s1 = lb.rmsds(one)
            c1 = one.centroids()
        for r in reps:
            assert r.F == F and r.N == len(sel)
            assert np.array_equal(r.centroids(), c1)
        Mn, sn = lb.rmsds_multi(reps)
        assert np.array_equal(M1, Mn) and sn["max"] == s1["max"]
        Ms, _ = lb.rmsds(reps[-1])        # any replica alone gives the same matrix
        assert np.array_equal(M1, Ms)
    finally:
        _close(reps)
    tiny = lb.stage_multi(X[:50], devices)  # too few frames to split: staged once, replicated
    try:
        Mt, _ = lb.rmsds_multi(tiny)
        with lb.FrameSet(X[:50]) as one:
            assert np.array_equal(Mt, lb.rmsds(one)[0])
    finally:
        _close(tiny)
