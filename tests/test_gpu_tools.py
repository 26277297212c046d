"""GPU end-to-end tests of the drop-in tools (loos_b200/bin/rmsds, multi-rmsds,
rmsd2ref): real PDB + DCD files in, the reference's stdout/stderr formats out,
checked against the CPU oracle fed through the Python restatement of the
reference's readers/selection (oracle/frontend_oracle.py)."""
import os
import re
import subprocess

import numpy as np
import pytest

from loos_b200.synth import synth_system, write_dcd, write_pdb
from oracle import frontend_oracle as fo
from oracle.oracle import matrix_text

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "loos_b200", "bin")


@pytest.fixture(scope="module", autouse=True)
def tools_built():
    """The tools are build products (git-ignored): build them if this checkout has none yet."""
    names = ["rmsds", "multi-rmsds", "rmsd2ref", "rms-overlap", "rmsds-align"]
    if not all(os.path.exists(os.path.join(BIN, n)) for n in names):
        subprocess.run(["make", "-C", os.path.join(ROOT, "loos_b200", "frontend")], check=True, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("tools")
    atoms, xyz = synth_system(60, 150, seed=4242, flavour="rich")
    write_pdb(str(d / "model.pdb"), atoms, xyz[0])
    write_dcd(str(d / "a.dcd"), xyz[:70])
    write_dcd(str(d / "b.dcd"), xyz[70:115], with_crystal=True)
    write_dcd(str(d / "c.dcd"), xyz[115:], big_endian=True)
    write_dcd(str(d / "all.dcd"), xyz)
    write_pdb(str(d / "target.pdb"), atoms, xyz[149])
    return dict(d=d, atoms=fo.read_pdb(str(d / "model.pdb")), xyz=xyz)


def run(tool, *args):
    r = subprocess.run([os.path.join(BIN, tool)] + [str(a) for a in args], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return r.stdout, r.stderr


def parse_matrix(out):
    lines = out.splitlines()
    hdr = [k for k, l in enumerate(lines) if re.fullmatch(r"# \d+ \d+ \(0\)", l)][0]
    m, n = map(int, lines[hdr].split()[1:3])
    M = np.array([[float(x) for x in l.split()] for l in lines[hdr + 1:hdr + 1 + m]])
    assert M.shape == (m, n)
    assert all(l.endswith(" ") for l in lines[hdr + 1:hdr + 1 + m])       # "value " per element
    return lines[:hdr], M, "\n".join(lines[hdr:]) + "\n"


def test_rmsds_default_and_precision(files, checker):
    d = files["d"]
    sel = fo.select(files["atoms"], "name == 'CA'")
    ref = checker.rmsds_self(files["xyz"][:, sel].astype(np.float64), 8)
    out, err = run("rmsds", d / "model.pdb", d / "all.dcd")
    head, M, text = parse_matrix(out)
    assert len(head) == 1 and head[0].startswith("# ") and "rmsds" in head[0] and "[loos-b200" in head[0]
    want = matrix_text(ref, 2)
    ta, tb = text.split(), want.split()
    assert len(ta) == len(tb) and sum(x != y for x, y in zip(ta, tb)) <= 2   # 2-digit rounding boundaries only
    assert err == ""
    out8, _ = run("rmsds", "-p", "8", "--sel1", "name == 'CA'", d / "model.pdb", d / "all.dcd")
    _, M8, _ = parse_matrix(out8)
    assert np.abs(M8 - ref).max() <= 1e-4 and np.abs(M8 - ref).max() < 2e-5
    assert np.all(np.diag(M8) == 0)


def test_rmsds_unsorted_range_and_stats(files, checker):
    d = files["d"]
    sel = fo.select(files["atoms"], "backbone")
    frames = fo.assign_frames(150, "40:-7:0,5,5,100:3:120", 0)       # rmsds keeps order and duplicates
    ref = checker.rmsds_self(files["xyz"][frames][:, sel].astype(np.float64), 4)
    out, err = run("rmsds", "--sel1", "backbone", "--range1", "40:-7:0,5,5,100:3:120", "--stats", "1", "-p", "7",
                   d / "model.pdb", d / "all.dcd")
    _, M, _ = parse_matrix(out)
    assert M.shape == (len(frames),) * 2 and np.abs(M - ref).max() < 2e-5
    low = ref[np.tril_indices(len(frames), -1)].astype(np.float64)
    assert err.strip() == "Max rmsd = %.4f, avg rmsd = %.4f" % (low.max(), low.mean())
    out, err = run("rmsds", "-N", "1", "--skip1", "100", d / "model.pdb", d / "all.dcd")
    assert out == "" and err.startswith("Max rmsd = ")


def test_rmsds_two_trajectories(files, checker):
    d = files["d"]
    s1 = fo.select(files["atoms"], "name == 'CA' && resid <= 30")
    s2 = fo.select(files["atoms"], "name == 'CB' && resid > 30")
    ref = checker.rmsds_cross(files["xyz"][:70][:, s1].astype(np.float64), files["xyz"][115:][:, s2].astype(np.float64), 4)
    out, err = run("rmsds", "--sel1", "name == 'CA' && resid <= 30", "--sel2", "name == 'CB' && resid > 30", "-p", "7",
                   "--stats", "1", d / "model.pdb", d / "a.dcd", d / "model.pdb", d / "c.dcd")
    _, M, _ = parse_matrix(out)
    assert M.shape == (70, 35) and np.abs(M - ref).max() < 2e-5
    assert err.strip() == "Max rmsd = %.4f, avg rmsd = %.4f" % (ref.astype(np.float64).max(), ref.astype(np.float64).mean())


def test_multi_rmsds_table_and_matrix(files, checker):
    d = files["d"]
    sizes = [70, 45, 35]
    locs = fo.multitraj_locations(sizes, 4, 3)
    starts = np.cumsum([0] + sizes)
    glob = [starts[t] + f for t, f in locs]
    sel = fo.select(files["atoms"], "name == 'CA'")
    ref = checker.rmsds_self(files["xyz"][glob][:, sel].astype(np.float64), 4)
    out, _ = run("multi-rmsds", "-k", "4", "-i", "3", "-p", "7", d / "model.pdb", d / "a.dcd", d / "b.dcd", d / "c.dcd")
    head, M, _ = parse_matrix(out)
    assert head[1] == "# traj\tstart\tend\tfilename"
    counts = [0 if n <= 4 else (n - 4 + 2) // 3 for n in sizes]
    s = 0
    for k, (n, name) in enumerate(zip(counts, "abc")):
        assert head[2 + k] == "# %d\t%d\t%d\t%s" % (k, s, s + n - 1, d / (name + ".dcd"))
        s += n
    assert M.shape == (sum(counts),) * 2 and np.abs(M - ref).max() < 2e-5
    out, _ = run("multi-rmsds", "-r", "30:-10:0,5,5", "-p", "7", d / "model.pdb", d / "a.dcd", d / "b.dcd")
    head, M, _ = parse_matrix(out)
    assert head[1] == "# Note- composite frame range used was '30:-10:0,5,5'"
    ref = checker.rmsds_self(files["xyz"][[0, 5, 10, 20, 30]][:, sel].astype(np.float64), 2)   # sorted + unique
    assert np.abs(M - ref).max() < 2e-5


def test_rmsd2ref_target_and_iterative(files, checker):
    d = files["d"]
    atoms, xyz = files["atoms"], files["xyz"].astype(np.float64)
    sa = fo.select(atoms, "name == 'CA'")
    sr = fo.select(atoms, "!(hydrogen || segid =~ 'SOLV|BULK')")
    tgt = np.array([a["xyz"] for a in fo.read_pdb(str(d / "target.pdb"))])
    ref, _ = checker.rmsd2ref_target(xyz[:, sa], xyz[:, sr], tgt[sa], tgt[sr])
    out, err = run("rmsd2ref", "--target", d / "target.pdb", d / "model.pdb", d / "all.dcd")
    lines = out.splitlines()
    assert lines[0].startswith("# ") and len(lines) == 151
    got = np.array([float(l.split("\t")[1]) for l in lines[1:]])
    assert [int(l.split("\t")[0]) for l in lines[1:]] == list(range(150))
    assert np.abs(got - ref).max() < 1e-5 * np.maximum(1, ref).max()           # 6 significant digits printed
    assert "Warning: Using --align selection for target" in err and "Computing RMSD vs target" in err
    assert ("Average RMSD was %.3f, std RMSD was %.3f" % (ref.mean(), ref.std(ddof=1))) in err
    # default mode: iterative alignment to the running average, frames 10..59 by --range
    sub = np.arange(10, 60)
    it = checker.rmsd2ref_iterative(xyz[sub][:, sa], xyz[sub][:, sr], 1e-6, 1000)
    out, err = run("rmsd2ref", "-r", "10:59", d / "model.pdb", d / "all.dcd")
    got = np.array([float(l.split("\t")[1]) for l in out.splitlines()[1:]])
    assert np.abs(got - it[0]).max() < 1e-5 * np.maximum(1, it[0]).max()
    assert "Computing using average structure..." in err and "Aligning using 60 atoms" in err
    # skip + range together is an error (src/OptionsFramework.cpp:375-379)
    r = subprocess.run([os.path.join(BIN, "rmsd2ref"), "-k", "3", "-r", "0:5", str(d / "model.pdb"), str(d / "all.dcd")],
                       capture_output=True, text=True)
    assert r.returncode == 255 and "cannot specify both a skip and a frame range" in r.stderr


def test_rms_overlap(files, checker):
    """set A x set B rectangular matrix, two trajectory tables, stats with and
    without --cutoff (including the reference's brace-less `if` in showFractionalStats)."""
    d = files["d"]
    sel = fo.select(files["atoms"], "backbone && !hydrogen")
    A = files["xyz"][:115:2][:, sel].astype(np.float64)          # a.dcd + b.dcd, stride 2 per sub-trajectory
    ia = [f for t, f in fo.multitraj_locations([70, 45], 0, 2)]
    A = np.concatenate([files["xyz"][:70][0::2], files["xyz"][70:115][0::2]])[:, sel].astype(np.float64)
    B = files["xyz"][115:][0::2][:, sel].astype(np.float64)
    ref = checker.rmsds_cross(A, B, 4)
    out, err = run("rms-overlap", "-s", "backbone && !hydrogen", "-i", "2", "-p", "7", "--stats", "1",
                   "-A", "%s %s" % (d / "a.dcd", d / "b.dcd"), "-B", str(d / "c.dcd"), d / "model.pdb")
    head, M, _ = parse_matrix(out)
    assert head[1] == "# traj\tstart\tend\tfilename" and head[2].endswith("a.dcd") and head[3].endswith("b.dcd")
    assert head[4] == "# traj\tstart\tend\tfilename" and head[5] == "# 0\t0\t17\t%s" % (d / "c.dcd")
    assert M.shape == ref.shape == (58, 18) and np.abs(M - ref).max() < 2e-5
    from oracle.oracle import fractional_stats_line, stats_line_whole     # pinned to the reference's own functions (CPU suite)
    assert err.strip() == stats_line_whole(ref).strip()
    out, err = run("rms-overlap", "-s", "backbone && !hydrogen", "-i", "2", "-c", "6.5", "-N", "1",
                   "-A", "%s %s" % (d / "a.dcd", d / "b.dcd"), "-B", str(d / "c.dcd"), d / "model.pdb")
    assert out.strip() == fractional_stats_line(ref, 6.5).strip() and err == ""


def test_amber_netcdf_trajectories_give_the_same_matrices(files, tmp_path):
    """The same frames as Amber NetCDF (.nc, written with scipy's NetCDF-3 writer) and as DCD: rmsds,
    multi-rmsds (mixed formats in one MultiTrajectory) and rmsd2ref print identical numbers."""
    from scipy.io import netcdf_file
    d, xyz = files["d"], files["xyz"]

    def write_nc(path, X, version):
        f = netcdf_file(str(path), "w", version=version)
        f.Conventions = "AMBER"; f.ConventionVersion = "1.0"
        f.createDimension("frame", None); f.createDimension("spatial", 3); f.createDimension("atom", X.shape[1])
        t = f.createVariable("time", "f", ("frame",))
        c = f.createVariable("coordinates", "f", ("frame", "atom", "spatial"))
        for k in range(len(X)):
            c[k] = X[k]; t[k] = k
        f.close()

    write_nc(tmp_path / "all.nc", xyz, 1)
    write_nc(tmp_path / "b.nc", xyz[70:115], 2)
    body = lambda out: "\n".join(l for l in out.splitlines() if not (l.startswith("# ") and "[loos-b200" in l))
    a, _ = run("rmsds", "-p", "6", d / "model.pdb", d / "all.dcd")
    b, _ = run("rmsds", "-p", "6", d / "model.pdb", tmp_path / "all.nc")
    assert body(a) == body(b)
    a, _ = run("multi-rmsds", "-p", "5", d / "model.pdb", d / "a.dcd", d / "b.dcd", d / "c.dcd")
    b, _ = run("multi-rmsds", "-p", "5", d / "model.pdb", d / "a.dcd", tmp_path / "b.nc", d / "c.dcd")
    strip = lambda out: "\n".join(l for l in body(out).splitlines() if not l.startswith("# "))   # trajectory table names differ
    assert strip(a) == strip(b)
    a, _ = run("rmsd2ref", "--target", d / "target.pdb", d / "model.pdb", d / "all.dcd")
    b, _ = run("rmsd2ref", "--target", d / "target.pdb", d / "model.pdb", tmp_path / "all.nc")
    assert body(a) == body(b)


def test_xtc_trajectory_gives_the_same_matrix_as_its_decoded_frames(files, tmp_path):
    """An XTC file (written by the restatement of the reference's writer) and a DCD holding the
    frames the reference's reader arithmetic decodes from it: rmsds and rmsd2ref print the same numbers."""
    d, xyz = files["d"], files["xyz"]
    fo.write_xtc(str(tmp_path / "all.xtc"), xyz)
    write_dcd(str(tmp_path / "decoded.dcd"), fo.read_xtc(str(tmp_path / "all.xtc")))
    body = lambda out: "\n".join(l for l in out.splitlines() if not (l.startswith("# ") and "[loos-b200" in l))
    a, _ = run("rmsds", "-p", "6", d / "model.pdb", tmp_path / "decoded.dcd")
    b, _ = run("rmsds", "-p", "6", d / "model.pdb", tmp_path / "all.xtc")
    assert body(a) == body(b)
    a, _ = run("rmsd2ref", "--target", d / "target.pdb", d / "model.pdb", tmp_path / "decoded.dcd")
    b, _ = run("rmsd2ref", "--target", d / "target.pdb", d / "model.pdb", tmp_path / "all.xtc")
    assert body(a) == body(b)


def test_mdtraj_h5_trajectory_gives_the_same_matrix_as_its_decoded_frames(files, tmp_path):
    """An MDTraj-style HDF5 file (chunked, shuffle + deflate; written by the tests' restatement of the format) and
    a DCD holding 10.0 * its stored nm values: rmsds and multi-rmsds (mixed formats) print the same numbers."""
    d, xyz = files["d"], files["xyz"]
    fo.write_mdtraj_h5(str(tmp_path / "all.h5"), xyz, chunk=(8, 500, 3))
    decoded = (10.0 * (xyz.astype(np.float64) / 10.0).astype(np.float32).astype(np.float64)).astype(np.float32)
    write_dcd(str(tmp_path / "decoded.dcd"), decoded)
    body = lambda out: "\n".join(l for l in out.splitlines() if not (l.startswith("# ") and "[loos-b200" in l))
    a, _ = run("rmsds", "-p", "6", d / "model.pdb", tmp_path / "decoded.dcd")
    b, _ = run("rmsds", "-p", "6", d / "model.pdb", tmp_path / "all.h5")
    assert body(a) == body(b)
    strip = lambda out: "\n".join(l for l in body(out).splitlines() if not l.startswith("# "))
    a, _ = run("multi-rmsds", "-p", "5", d / "model.pdb", tmp_path / "decoded.dcd", d / "b.dcd")
    b, _ = run("multi-rmsds", "-p", "5", d / "model.pdb", tmp_path / "all.h5", d / "b.dcd")
    assert strip(a) == strip(b)


def test_binary_side_channel(files, tmp_path):
    """--binary writes the float32 matrix as .npy (fortran order): identical to the values behind the
    text at -p 9, also with --noout 1, for the self and the cross matrix."""
    d = files["d"]
    npy = tmp_path / "m.npy"
    out, _ = run("rmsds", "-p", "9", "--binary", npy, d / "model.pdb", d / "all.dcd")
    _, M, _ = parse_matrix(out)
    B = np.load(npy)
    assert B.dtype == np.float32 and B.flags["F_CONTIGUOUS"] and B.shape == M.shape
    assert np.array_equal(B, M.astype(np.float32))
    out, err = run("rmsds", "-N", "1", "--binary", npy, "--skip2", "100", d / "model.pdb", d / "a.dcd", d / "model.pdb", d / "all.dcd")
    assert out == "" and err.startswith("Max rmsd")
    C2 = np.load(npy)
    assert C2.shape == (70, 50)
    out, _ = run("rmsds", "-p", "9", "--skip2", "100", d / "model.pdb", d / "a.dcd", d / "model.pdb", d / "all.dcd")
    assert np.array_equal(C2, parse_matrix(out)[1].astype(np.float32))
    run("multi-rmsds", "-N", "1", "--binary", npy, d / "model.pdb", d / "a.dcd", d / "b.dcd")
    assert np.load(npy).shape == (115, 115)


def test_rmsds_align_tool(files, checker, tmp_path):
    """rmsds-align: header, one line per frame of trajectory 1, values as Python prints round(v, p);
    against the oracle's restatement of the script's loop fed through the restated readers/selection.
    One trajectory with different align / rmsd selections, then two trajectories with ranges."""
    import shutil
    from oracle import oracle
    d, atoms, xyz = files["d"], files["atoms"], files["xyz"]
    exe = os.path.join(BIN, "rmsds-align")

    def go(*args):
        r = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        return r

    align, rmsd = 'resid <= 20 && name == "CA"', 'resid > 25 && name =~ "^(C|CA|CB|O|N)$"'
    ia, ir = fo.select(atoms, align), fo.select(atoms, rmsd)
    assert len(ia) == 20 and len(ir) > 100
    X = xyz.astype(np.float64)
    # explicit stop: frames 5, 8, ..., 38
    fr = list(range(5, 40, 3))
    # (--engine fp64: values within 1e-8 A of the script's, so the rounded text is the script's; the default
    # tensor engine is within ~1e-6 A and is held to that below)
    r = go("--engine", "fp64", "--align", align, "--rmsd", rmsd, "--range1", "5:3:40", "--precision", "4", d / "model.pdb", d / "all.dcd")
    want = oracle.rmsds_align(checker, X[fr][:, ia], X[fr][:, ir], X[fr][:, ia], X[fr][:, ir])
    lines = r.stdout.splitlines()
    assert lines[0].startswith("# ") and lines[0].endswith("all.dcd") and "--precision 4" in lines[0]
    assert r.stdout[len(lines[0]) + 1:] == oracle.py_round_text(want, 4)
    rt = go("--align", align, "--rmsd", rmsd, "--range1", "5:3:40", "--precision", "7", d / "model.pdb", d / "all.dcd")
    got = np.array([[float(v) for v in l.split()] for l in rt.stdout.splitlines()[1:]])
    assert got.shape == want.shape and np.abs(got - want).max() < 1e-5
    # the default range stops at len(name), the length of the path as given
    short = tmp_path / "t.dcd"
    shutil.copy(d / "all.dcd", short)
    n = len(str(short))
    r = go("--engine", "fp64", d / "model.pdb", short)
    assert "length of the trajectory's file name" in r.stderr
    body = r.stdout.splitlines()[1:]
    assert len(body) == n and all(len(l.split()) == n for l in body)
    ia, ir = fo.select(atoms, 'name == "CA"'), fo.select(atoms, 'name =~ "^(C|O|N|CA)$"')
    want = oracle.rmsds_align(checker, X[:n][:, ia], X[:n][:, ir], X[:n][:, ia], X[:n][:, ir])
    assert "\n".join(body) + "\n" == oracle.py_round_text(want, 2)
    # two trajectories, different selections of equal size on each side
    a1, a2 = 'resid <= 20 && name == "CA"', 'resid >= 25 && resid <= 44 && name == "CA"'
    r1, r2 = 'resid > 25 && resid < 50 && name == "CB"', 'resid < 25 && name == "CB"'
    r = go("--engine", "fp64", "--align", a1, "--align2", a2, "--rmsd", r1, "--rmsd2", r2, "--model2", d / "model.pdb", "--traj2", d / "b.dcd",
           "--range1", "0:7:70", "--range2", "1:2:30", "--precision", "3", d / "model.pdb", d / "a.dcd")
    f1, f2 = list(range(0, 70, 7)), [70 + k for k in range(1, 30, 2)]          # b.dcd holds frames 70..114
    s = lambda q: fo.select(atoms, q)
    want = oracle.rmsds_align(checker, X[f1][:, s(a1)], X[f1][:, s(r1)], X[f2][:, s(a2)], X[f2][:, s(r2)])
    assert r.stdout.split("\n", 1)[1] == oracle.py_round_text(want, 3)
    # selections of different sizes: the reference's exception text, status 1, nothing on stdout
    r = subprocess.run([exe, "--align2", 'resid <= 10 && name == "CA"', "--range1", "0:1:5", str(d / "model.pdb"), str(d / "all.dcd")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "differing numbers of atoms" in r.stderr and r.stdout == ""
