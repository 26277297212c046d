"""CPU tests of the native front-end (loos_b200/frontend, C++) against the Python
restatement of the reference semantics in oracle/frontend_oracle.py: selection
language, PDB/hybrid-36, frame ranges, MultiTrajectory indexing, DCD decoding and
the Matrix text format -- all integer/byte/text work, so the bar is bit-exact."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from loos_b200.synth import synth_system, write_dcd, write_pdb
from oracle import frontend_oracle as fo
from oracle.oracle import matrix_text

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FE = os.path.join(ROOT, "loos_b200", "lib", "libloosb200fe.so")
_u32p = C.POINTER(C.c_uint32)


@pytest.fixture(scope="module")
def fe():
    if not os.path.exists(FE):
        subprocess.run(["make", "-C", os.path.join(ROOT, "loos_b200", "frontend")], check=True, stdout=subprocess.DEVNULL)
    L = C.CDLL(FE)
    L.lbfe_last_error.restype = C.c_char_p
    L.lbfe_select.restype = C.c_long
    L.lbfe_parse_range.restype = C.c_long
    L.lbfe_assign_frames.restype = C.c_long
    L.lbfe_multitraj.restype = C.c_long
    L.lbfe_matrix_text.restype = C.c_long
    L.lbfe_pdb_atoms.restype = C.c_long
    return L


@pytest.fixture(scope="module")
def system(tmp_path_factory):
    d = tmp_path_factory.mktemp("sys")
    atoms, xyz = synth_system(40, 12, seed=77, flavour="rich")
    # make the model harder: big ids/resids (hybrid-36), a negative resid, odd names
    for k, a in enumerate(atoms):
        a["resid"] = a["resid"] + 9990 if a["segid"] == "PROT" else a["resid"]
    atoms[3]["name"] = "H12'"
    atoms[7]["resname"] = "DA"
    atoms[7]["name"] = "C1'"
    pdb = str(d / "model.pdb")
    write_pdb(pdb, atoms, xyz[0])
    dcd = str(d / "traj.dcd")
    write_dcd(dcd, xyz)
    return dict(dir=d, pdb=pdb, dcd=dcd, atoms=atoms, xyz=xyz)


SELECTIONS = [
    "name == 'CA'", "all", "backbone", "hydrogen", "!hydrogen", "!(hydrogen || segid =~ 'SOLV|BULK')",
    "name =~ '^C'", "name =~ 'c'", "resid >= 10000 && resid < 10010", "resid > 10020 || name == 'N'",
    "name == 'CA' && resid < 10005 || name == 'O'", "name == 'O' || name == 'CA' && resid < 10005",
    "not (name == 'CA')", "segname == 'SOLV'", "segid != 'PROT'", "resname ne 'ALA'", "index < 17", "index >= 200 && index <= 210",
    "id > 100 && id <= 120", "(resid) == 10001", "(name == 'CA')", "((name == 'CA') && (index > 5))",
    "name -> '(\\d+)' == 12", "name -> '(\\d+)' < 13", "name -> 'H(\\d)' >= 1", "chainid == 'W' && name == 'OH2'",
    "resname =~ 'tip'", "name == \"CB\"", "backbone && !hydrogen", "name == 'CA' # trailing comment", "resid+ == 10003",
    'name =~ "^(C|O|N|CA)$"', 'resid <= 10020 && name== "CA"', 'resid > 10025 && name =~ "^(C|CA|CB|O|N)$"',    # rmsds-align.py's defaults / examples
]


@pytest.mark.parametrize("sel", SELECTIONS)
def test_selection_bit_exact(fe, system, sel):
    want = fo.select(fo.read_pdb(system["pdb"]), sel)
    out = np.zeros(len(system["atoms"]) + 8, dtype=np.uint32)
    n = fe.lbfe_select(system["pdb"].encode(), sel.encode(), out.ctypes.data_as(_u32p), len(out))
    assert n >= 0, fe.lbfe_last_error()
    assert n == len(want) and np.array_equal(out[:n], want), (sel, n, len(want))
    if sel in ("name == 'CA'", "backbone", "hydrogen"):
        assert n > 0


@pytest.mark.parametrize("sel", ["name == 5", "name ==", "name === 'CA'", "resid == 'A'", "foo == 1", "(name == 'CA'", "name =~ 3",
                                 "'CA' =~ 'C'", "index"])
def test_selection_errors(fe, system, sel):
    out = np.zeros(8, dtype=np.uint32)
    n = fe.lbfe_select(system["pdb"].encode(), sel.encode(), out.ctypes.data_as(_u32p), 8)
    assert n == -1
    with pytest.raises(Exception):
        fo.select(fo.read_pdb(system["pdb"]), sel)


def test_pdb_fields_and_hybrid36(fe, system):
    atoms = fo.read_pdb(system["pdb"])
    n = len(atoms)
    ids = np.zeros(n, np.int64); resids = np.zeros(n, np.int64); xyz = np.zeros((n, 3)); names = C.create_string_buffer(16 * n)
    got = fe.lbfe_pdb_atoms(system["pdb"].encode(), ids.ctypes.data_as(C.POINTER(C.c_long)),
                            resids.ctypes.data_as(C.POINTER(C.c_long)), xyz.ctypes.data_as(C.POINTER(C.c_double)), names, n)
    assert got == n == len(system["atoms"])
    assert [a["id"] for a in atoms] == list(ids) and [a["resid"] for a in atoms] == list(resids)
    assert max(resids) > 9999                      # the hybrid-36 branch was exercised
    assert np.array_equal(xyz, np.array([a["xyz"] for a in atoms]))
    raw = names.raw
    for i, a in enumerate(atoms):
        assert raw[16 * i:16 * i + 8].rstrip(b"\0").decode() == a["name"]
        assert raw[16 * i + 8:16 * i + 16].rstrip(b"\0").decode() == a["segid"]
    for field, val in [("    1", 1), ("99999", 99999), ("A0000", 100000), ("A0001", 100001), ("ZZZZZ", 43770015),
                       ("a0000", 43770016), ("  -12", None), ("A000", 10000), ("9999", 9999), ("-123", -123)]:
        got = fe.lbfe_hybrid36(field.encode(), len(field))
        assert got == fo.hybrid36(field)
        if val is not None:
            assert got == val


RANGES = ["5", "0:9", "3:", "0:2:20", "0:3:", "10:-1:0", "10:-3:2", ":5", ":-2:10", "1,5,9", "0:4,10:12,7", " 2 : 2 : 8 ", "7:7",
          "0:1:0", "20:-5:", "4,4,4", "15:-1:15"]


@pytest.mark.parametrize("spec", RANGES)
def test_frame_ranges(fe, spec):
    maxsize = 49
    try:
        want = fo.parse_index_range(spec, maxsize)
    except ValueError:
        want = None
    out = np.zeros(4096, dtype=np.uint32)
    n = fe.lbfe_parse_range(spec.encode(), maxsize, out.ctypes.data_as(_u32p), len(out))
    if want is None:
        assert n == -1
    else:
        assert n == len(want) and np.array_equal(out[:n], want), (spec, out[:n], want)


@pytest.mark.parametrize("spec", ["10:0", "0:-1:10", "a:b", "1:2:3:4", "1,,2", "", "5:x", ":2:10", "-3"])
def test_frame_range_errors(fe, spec):
    out = np.zeros(64, dtype=np.uint32)
    assert fe.lbfe_parse_range(spec.encode(), 49, out.ctypes.data_as(_u32p), 64) == -1
    with pytest.raises(ValueError):
        fo.parse_index_range(spec, 49)


def test_range_items_against_the_reference(fe):
    """tests/golden/pdb_ref.json["range_items"]: RangeItem::generate() of the reference itself (src/index_range_parser.cpp:63-102,
    compiled in place: oracle/pdb_ref_wrapper.cpp) for one-, two- and three-argument items -- validation (a step against the
    direction is 'Invalid range'), step 0, and the descending loop's unsigned wrap at 0.  The Boost.Spirit grammar that maps text
    onto these items (:123-130) cannot be compiled here; its five alternatives are spelled out below, so the native parser and the
    Python restatement are held to the reference's own generation semantics through them."""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "pdb_ref.json")))["range_items"]
    maxsize = 49
    out = np.zeros(1 << 16, dtype=np.uint32)
    assert len(g) >= 30
    for item in g:
        a = item["args"]
        if len(a) == 1:
            specs = ["%d" % a[0]]                                    # uint                         -> RangeItem(_1)
        elif len(a) == 2:
            specs = ["%d:%d" % (a[0], a[1])]                         # uint ':' endpoint            -> RangeItem(_1, _2)
        else:
            specs = ["%d:%d:%d" % (a[0], a[2], a[1])]                # uint ':' int ':' endpoint    -> RangeItem(_1, _3, _2)
        if len(a) == 2 and a[1] == maxsize:
            specs.append("%d:" % a[0])                               # the endpoint left out = maxsize
        if len(a) == 3 and a[1] == maxsize:
            specs.append("%d:%d:" % (a[0], a[2]))
        if len(a) == 3 and a[0] == maxsize:
            specs.append(":%d:%d" % (a[2], a[1]))                    # endpoint ':' int ':' uint    -> RangeItem(_1, _3, _2)
            if a[2] == -1:
                specs.append(":%d" % a[1])                           # endpoint ':' uint            -> RangeItem(_1, _2, -1)
        for spec in specs:
            n = fe.lbfe_parse_range(spec.encode(), maxsize, out.ctypes.data_as(_u32p), len(out))
            if item["values"] is None:
                assert n == -1, (spec, n)
                with pytest.raises(ValueError):
                    fo.parse_index_range(spec, maxsize)
            else:
                assert n == len(item["values"]) and list(out[:n]) == item["values"], (spec, list(out[:n])[:8], item["values"][:8])
                assert list(fo.parse_index_range(spec, maxsize)) == item["values"], spec


def test_assign_frames_sorted_vs_raw(fe):
    out = np.zeros(256, dtype=np.uint32)
    # rmsds keeps order and duplicates (Tools/rmsds.cpp:506); multi-rmsds / rmsd2ref sort + dedup
    n = fe.lbfe_assign_frames(30, b"9,3:5,4,20:-10:0", 0, 1, 0, out.ctypes.data_as(_u32p), 256)
    assert list(out[:n]) == [9, 3, 4, 5, 4, 20, 10, 0]
    n = fe.lbfe_assign_frames(30, b"9,3:5,4,20:-10:0", 0, 1, 1, out.ctypes.data_as(_u32p), 256)
    assert list(out[:n]) == [0, 3, 4, 5, 9, 10, 20]
    n = fe.lbfe_assign_frames(30, b"", 7, 4, 0, out.ctypes.data_as(_u32p), 256)
    assert list(out[:n]) == list(fo.assign_frames(30, "", 7, 4)) == [7, 11, 15, 19, 23, 27]


@pytest.mark.parametrize("sizes,skip,stride", [([10, 7, 12], 0, 1), ([10, 7, 12], 3, 2), ([5, 2, 9], 4, 3), ([3], 3, 1), ([100, 1, 50], 10, 7)])
def test_multitraj_indexing(fe, sizes, skip, stride):
    want = fo.multitraj_locations(sizes, skip, stride)
    s = np.array(sizes, dtype=np.uint32)
    lt = np.zeros(512, np.uint32); lf = np.zeros(512, np.uint32)
    n = fe.lbfe_multitraj(s.ctypes.data_as(_u32p), len(sizes), skip, stride, lt.ctypes.data_as(_u32p), lf.ctypes.data_as(_u32p), 512)
    assert n == len(want)
    assert [(int(a), int(b)) for a, b in zip(lt[:n], lf[:n])] == want


@pytest.mark.parametrize("kw", [dict(), dict(with_crystal=True), dict(big_endian=True), dict(header_nframes=0),
                                dict(with_crystal=True, big_endian=True, header_nframes=0)])
def test_dcd_decoding(fe, system, tmp_path, kw):
    path = str(tmp_path / "t.dcd")
    write_dcd(path, system["xyz"], **kw)
    na, nf, xt = C.c_uint32(), C.c_uint32(), C.c_int()
    assert fe.lbfe_dcd_info(path.encode(), C.byref(na), C.byref(nf), C.byref(xt)) == 0, fe.lbfe_last_error()
    F, N, _ = system["xyz"].shape
    assert (na.value, nf.value, bool(xt.value)) == (N, F, bool(kw.get("with_crystal")))
    ref = fo.read_dcd(path)
    assert np.array_equal(ref, system["xyz"].transpose(0, 2, 1))     # byte-exact float planes
    buf = np.zeros((3, N), dtype=np.float32)
    for f in (0, F - 1, F // 2):
        assert fe.lbfe_dcd_read(path.encode(), f, buf.ctypes.data_as(C.POINTER(C.c_float))) == 0
        assert np.array_equal(buf, ref[f])
    assert fe.lbfe_dcd_read(path.encode(), F, buf.ctypes.data_as(C.POINTER(C.c_float))) == -1


@pytest.mark.parametrize("prec", [0, 1, 2, 6, 8])
def test_matrix_text_bytes(fe, prec):
    rng = np.random.default_rng(3)
    M = np.asfortranarray((rng.random((7, 5)) * 10.0 ** rng.integers(-6, 4, (7, 5))).astype(np.float32))
    M[0, 0] = 0.0; M[1, 1] = 123456.0; M[2, 2] = 1e-5; M[3, 3] = 99.5
    cap = 1 << 16
    out = C.create_string_buffer(cap)
    n = fe.lbfe_matrix_text(M.ctypes.data_as(C.POINTER(C.c_float)), 7, 5, 7, prec, out, cap)
    assert out.raw[:n].decode() == matrix_text(M, prec)


@pytest.mark.parametrize("prec", [1, 2, 3, 6, 8, 9])
def test_fast_float_formatter_is_printf_exact(fe, prec):
    """writeMatrix formats floats with exact integer arithmetic instead of snprintf (5x faster; the
    text of a C4 result is 10^10 values).  Same bytes as "%.{p}g" for every float: random values over
    14 decades, exact rounding ties (binary fractions), decade boundaries, zero, +-inf, denormals."""
    fe.lbfe_format_floats.restype = C.c_long
    fe.lbfe_format_floats.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_char_p, C.c_long]
    rng = np.random.default_rng(100 + prec)
    special = [0.0, -0.0, 0.125, 0.375, 2.5, 3.5, 0.5, 1.5, 9.995, 9.95, 99.5, 1e-5, 9.9999e-5, 0.0001, 0.001,
               1e9, 999999999, 123456789, 0.1, 0.2, 0.3, 1.0, 10.0, 100.0, 1e8, 1e-45, 3.4e38, np.inf, -np.inf,
               9.5, 95.0, 0.95, 0.095, 99999.5, 0.99999]
    vals = np.concatenate([rng.uniform(0, 30, 60000), 10.0 ** rng.uniform(-7, 11, 60000), rng.normal(0, 5, 20000),
                           np.array(special), np.arange(0, 4096) / 1024.0, (np.arange(0, 30000) + 0.5) / 1000.0,
                           -(np.arange(0, 3000) + 0.5) / 100.0]).astype(np.float32)
    cap = 48 * len(vals)
    out = C.create_string_buffer(cap)
    n = fe.lbfe_format_floats(vals.ctypes.data, len(vals), prec, out, cap)
    got = out.raw[:n].decode().split("\n")[:-1]
    want = ["%.*g" % (prec, float(x)) for x in vals]
    bad = [(float(x), g, w) for x, g, w in zip(vals, got, want) if g != w]
    assert not bad, bad[:5]


def test_threaded_writer_matches_oracle_on_a_padded_matrix(fe, monkeypatch):
    """Blocks of 16 rows formatted on several threads, ld > m, m not a multiple of 16."""
    rng = np.random.default_rng(8)
    m, n, ld = 203, 67, 211
    buf = np.zeros((n, ld), np.float32)             # column-major storage with padding rows
    buf[:, :m] = rng.gamma(2.0, 3.0, (n, m)).astype(np.float32)
    M = buf[:, :m].T                                  # logical m x n
    texts = []
    for threads in ("1", "5"):
        monkeypatch.setenv("LOOSB200_WRITER_THREADS", threads)
        cap = 1 << 20
        out = C.create_string_buffer(cap)
        k = fe.lbfe_matrix_text(buf.ctypes.data_as(C.POINTER(C.c_float)), m, n, ld, 3, out, cap)
        texts.append(out.raw[:k].decode())
    assert texts[0] == texts[1] == matrix_text(np.asfortranarray(M), 3)


def test_tools_fail_loudly_without_gpu(system):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    exe = os.path.join(ROOT, "loos_b200", "bin", "rmsds")
    r = subprocess.run([exe, system["pdb"], system["dcd"]], capture_output=True, text=True)
    assert r.returncode == 255 and "no usable sm_100" in r.stderr and r.stdout == ""
    r = subprocess.run([exe, system["pdb"]], capture_output=True, text=True)      # traj missing -> usage, exit(-1)
    assert r.returncode == 255 and "Usage-" in r.stderr
    r = subprocess.run([exe, "--sel", "all", system["pdb"], system["dcd"]], capture_output=True, text=True)
    assert r.returncode == 255 and "unrecognised option '--sel'" in r.stderr   # no prefix guessing


# ---------------------------------------------------------------- Amber NetCDF (src/amber_netcdf.cpp)

def _write_amber_nc(path, xyz, version=1, dtype="f", box=True, conventions="AMBER", conv_version="1.0",
                    extra_record_var=True):
    """Amber NetCDF trajectory written with scipy's NetCDF-3 writer (independent of the reader under test)."""
    from scipy.io import netcdf_file
    F, N, _ = xyz.shape
    f = netcdf_file(path, "w", version=version)
    if conventions is not None:
        f.Conventions = conventions
    if conv_version is not None:
        f.ConventionVersion = conv_version
    f.program = "test"
    f.createDimension("frame", None)
    f.createDimension("spatial", 3)
    f.createDimension("atom", N)
    if box:
        f.createDimension("cell_spatial", 3)
    if extra_record_var:
        t = f.createVariable("time", "f", ("frame",))
        t.units = "picosecond"
    c = f.createVariable("coordinates", dtype, ("frame", "atom", "spatial"))
    c.units = "angstrom"
    if box:
        b = f.createVariable("cell_lengths", "d", ("frame", "cell_spatial"))
    for k in range(F):
        c[k] = xyz[k]
        if extra_record_var:
            t[k] = 0.5 * k
        if box:
            b[k] = [50.0, 60.0, 70.0]
    f.close()


@pytest.mark.parametrize("version,dtype,box,extra", [(1, "f", True, True), (2, "f", False, True), (1, "d", True, True),
                                                    (1, "f", False, False)])
def test_amber_netcdf_reader(fe, tmp_path, version, dtype, box, extra):
    """NetCDF-3 classic and 64-bit-offset containers, float and double coordinates, with and without
    other record variables (a lone record variable is stored unpadded): every frame comes back as
    X|Y|Z planes with the exact float32 values."""
    rng = np.random.default_rng(5)
    F, N = 9, 37                                   # N*3*4 not a multiple of 8: exercises record padding rules
    xyz = (rng.normal(size=(F, N, 3)) * 30).astype(np.float32)
    path = str(tmp_path / ("t%d%s.nc" % (version, dtype)))
    _write_amber_nc(path, xyz.astype(np.float64 if dtype == "d" else np.float32), version, dtype, box, extra_record_var=extra)
    na, nf = C.c_uint32(), C.c_uint32()
    assert fe.lbfe_traj_info(path.encode(), N, C.byref(na), C.byref(nf)) == 0, fe.lbfe_last_error()
    assert (na.value, nf.value) == (N, F)
    buf = np.zeros(3 * N, np.float32)
    for k in (0, 4, F - 1):
        assert fe.lbfe_traj_read(path.encode(), N, k, buf.ctypes.data_as(C.POINTER(C.c_float))) == 0
        assert np.array_equal(buf.reshape(3, N), xyz[k].T)
    assert fe.lbfe_traj_read(path.encode(), N, F, buf.ctypes.data_as(C.POINTER(C.c_float))) == -1


def test_amber_netcdf_validation(fe, tmp_path):
    """The checks of AmberNetcdf::init (src/amber_netcdf.cpp:28-45) and the container checks."""
    xyz = np.zeros((2, 5, 3), np.float32)
    na, nf = C.c_uint32(), C.c_uint32()

    def err(path, natoms=5):
        assert fe.lbfe_traj_info(path.encode(), natoms, C.byref(na), C.byref(nf)) == -1
        return fe.lbfe_last_error().decode()

    p = str(tmp_path / "a.nc"); _write_amber_nc(p, xyz, conventions=None)
    assert "Unable to find convention global attributes" in err(p)
    p = str(tmp_path / "b.nc"); _write_amber_nc(p, xyz, conventions="CF-1.0")
    assert "Cannot find AMBER tag" in err(p)
    p = str(tmp_path / "c.nc"); _write_amber_nc(p, xyz, conv_version="2.0")
    assert "only 1.0 is supported" in err(p)
    p = str(tmp_path / "d.nc"); _write_amber_nc(p, xyz)
    assert "different number of atoms than expected" in err(p, natoms=6)
    p = str(tmp_path / "e.nc"); open(p, "wb").write(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    assert "NetCDF-4" in err(p)
    p = str(tmp_path / "f.trr"); open(p, "wb").write(b"\0" * 64)
    assert "unsupported trajectory format" in err(p)
    # the suffix rules of createTrajectory: .crd is NetCDF when it carries the magic
    p = str(tmp_path / "g.crd"); _write_amber_nc(p, xyz)
    assert fe.lbfe_traj_info(p.encode(), 5, C.byref(na), C.byref(nf)) == 0 and nf.value == 2


def test_matrix_write_read_round_trip(fe):
    """writeMatrix -> readMatrix (readAsciiMatrix, src/MatrixRead.hpp:102-143) reproduces every float32
    bit at -p 9 and the rounded values at the default -p 2; header lines before the marker are skipped."""
    rng = np.random.default_rng(12)
    m, n = 41, 29
    M = np.asfortranarray(rng.gamma(2.0, 2.5, (m, n)).astype(np.float32))
    M[3, 4] = 0.0
    fe.lbfe_read_matrix.restype = C.c_int
    for prec in (9, 2):
        cap = 1 << 20
        out = C.create_string_buffer(cap)
        k = fe.lbfe_matrix_text(M.ctypes.data_as(C.POINTER(C.c_float)), m, n, m, prec, out, cap)
        text = b"# rmsds 'model.pdb' 'traj.dcd' - user (date) {cwd} [loos-b200]\n" + out.raw[:k]
        back = np.zeros((n, m), np.float32)
        mm, nn = C.c_long(), C.c_long()
        assert fe.lbfe_read_matrix(text, len(text), back.ctypes.data_as(C.POINTER(C.c_float)), back.size,
                                   C.byref(mm), C.byref(nn)) == 0, fe.lbfe_last_error()
        assert (mm.value, nn.value) == (m, n)
        if prec == 9:
            assert np.array_equal(back.T, M)
        else:
            assert np.array_equal(back.T, np.array([[float("%.2g" % v) for v in row] for row in M], np.float32))
    bad = b"# 2 2 (0)\n1 2 \n3 x \n"
    assert fe.lbfe_read_matrix(bad, len(bad), back.ctypes.data_as(C.POINTER(C.c_float)), back.size, C.byref(mm), C.byref(nn)) == -1
    assert b"Invalid conversion on matrix read at (1,1) of 'x'" in fe.lbfe_last_error()
    assert fe.lbfe_read_matrix(b"1 2 3\n", 6, back.ctypes.data_as(C.POINTER(C.c_float)), back.size, C.byref(mm), C.byref(nn)) == -1
    assert b"Could not find magic marker" in fe.lbfe_last_error()


def test_config_file_option(system, tmp_path):
    """--config file (src/OptionsFramework.cpp:746,765-771, boost parse_config_file): name = value
    lines, comments, positionals by name, command line wins.  Observed through the errors the tool
    raises before it needs a GPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the runs would go through")
    exe = os.path.join(ROOT, "loos_b200", "bin", "rmsds")
    cfg = tmp_path / "opts.cfg"
    cfg.write_text("# options for rmsds\nmodel1 = %s\ntraj1=%s   # the trajectory\n\nsel1 = name == 'NOPE'\nprecision = 5\n" % (system["pdb"], system["dcd"]))
    r = subprocess.run([exe, "--config", str(cfg)], capture_output=True, text=True)
    assert r.returncode == 255 and "name == 'NOPE'" in r.stderr and "matched no atoms" in r.stderr
    r = subprocess.run([exe, "--config", str(cfg), "--sel1", "name == 'CA'"], capture_output=True, text=True)
    assert r.returncode == 255 and "no usable sm_100" in r.stderr            # got past the selection: command line won
    cfg.write_text("model1 = %s\nbogus = 1\n" % system["pdb"])
    r = subprocess.run([exe, "--config", str(cfg)], capture_output=True, text=True)
    assert r.returncode == 255 and "unrecognised option 'bogus'" in r.stderr
    cfg.write_text("model1 %s\n" % system["pdb"])
    r = subprocess.run([exe, "--config", str(cfg)], capture_output=True, text=True)
    assert r.returncode == 255 and "invalid line" in r.stderr
    r = subprocess.run([exe, "--config", str(tmp_path / "missing.cfg"), system["pdb"], system["dcd"]], capture_output=True, text=True)
    assert r.returncode == 255 and "Cannot open options config file" in r.stderr


# ------------------------------------------------------------------ GROMACS XTC reader
def _xtc_read_native(fe, path, N, F):
    na, nf = C.c_uint32(), C.c_uint32()
    assert fe.lbfe_traj_info(path.encode(), N, C.byref(na), C.byref(nf)) == 0, fe.lbfe_last_error()
    assert (na.value, nf.value) == (N, F)
    out = np.zeros((F, 3, N), np.float32)
    for k in range(F):
        assert fe.lbfe_traj_read(path.encode(), N, k, out[k].ctypes.data_as(C.POINTER(C.c_float))) == 0, fe.lbfe_last_error()
    return out.transpose(0, 2, 1)


@pytest.mark.parametrize("case", ["protein", "water", "wide", "tiny", "huge_box"])
def test_xtc_reader(fe, tmp_path, case):
    """XTC frames decoded by the native reader equal, bit for bit, the restatement of the reference's
    reader (readCompressedCoords, src/xtc.cpp:165-332) on files made by the restatement of the
    reference's writer (src/xtcwriter.cpp:387-655); and both are within half a quantum of the input.
    Both restatements are pinned to the reference's own compiled code further down
    (test_xtc_against_the_references_own_writer_and_reader)."""
    from oracle import frontend_oracle as fo
    from loos_b200.synth import synth_system
    rng = np.random.default_rng(5)
    precision = 1000.0
    if case == "protein":
        _, xyz = synth_system(40, 5, seed=3, flavour="rich")
    elif case == "water":       # triplets of close atoms: exercises the swap and the run-length coding
        O = rng.uniform(0, 40, (4, 300, 1, 3))
        xyz = (O + rng.normal(0, 0.4, (4, 300, 3, 3))).reshape(4, 900, 3).astype(np.float32)
    elif case == "wide":        # spread-out atoms: no runs, smallidx walks down and up
        xyz = rng.uniform(-300, 300, (3, 200, 3)).astype(np.float32)
        xyz[:, 50:120] = (xyz[:, 50:51] + rng.normal(0, 0.05, (3, 70, 3))).astype(np.float32)
    elif case == "tiny":        # <= 9 atoms are stored as plain floats
        xyz = rng.uniform(-20, 20, (5, 7, 3)).astype(np.float32)
    else:                       # extent above 2^24 quanta: per-axis bit fields instead of the mixed radix
        xyz = rng.uniform(-1.2e5, 1.2e5, (2, 64, 3)).astype(np.float32)
        xyz[:, 10] = xyz[:, 9] + 1.0      # one close pair: with every atom far apart the reference's writer
                                          # indexes past its table (src/xtcwriter.cpp:513-525), not a case to restate
    F, N, _ = xyz.shape
    path = str(tmp_path / "t.xtc")
    fo.write_xtc(path, xyz, precision=precision)
    want = fo.read_xtc(path)
    got = _xtc_read_native(fe, path, N, F)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    if case == "tiny":
        assert np.abs(got - xyz).max() <= 4e-6 * np.abs(xyz).max()
    else:
        assert np.abs(got.astype(np.float64) - xyz).max() <= 5.0 / precision * 1.001 + np.abs(xyz).max() * 2e-7


def test_xtc_validation(fe, tmp_path):
    from oracle import frontend_oracle as fo
    xyz = np.random.default_rng(1).uniform(0, 30, (3, 20, 3)).astype(np.float32)
    p = str(tmp_path / "a.xtc"); fo.write_xtc(p, xyz)
    na, nf = C.c_uint32(), C.c_uint32()
    # like DCD, the file states its own atom count; the tools compare it with the model
    assert fe.lbfe_traj_info(p.encode(), 21, C.byref(na), C.byref(nf)) == 0 and (na.value, nf.value) == (20, 3)
    raw = open(p, "rb").read()
    q = str(tmp_path / "b.xtc"); open(q, "wb").write(b"\0\0\0\1" + raw[4:])
    assert fe.lbfe_traj_info(q.encode(), 20, C.byref(na), C.byref(nf)) == -1
    assert "magic" in fe.lbfe_last_error().decode().lower()
    r = str(tmp_path / "c.xtc"); open(r, "wb").write(raw[:len(raw) - 10])      # last frame cut short
    assert fe.lbfe_traj_info(r.encode(), 20, C.byref(na), C.byref(nf)) == -1
    assert "truncated" in fe.lbfe_last_error().decode()


# ------------------------------------------------------------------ Python float text (rmsds-align)
def test_py_round_repr_matches_python(fe):
    """pyRoundRepr(v, n) == repr(round(v, n)): the text Packages/Clustering/rmsds-align.py:254 prints.
    Pinned against this interpreter's own float.__round__ / repr."""
    fe.lbfe_py_round_repr.argtypes = [C.c_double, C.c_int, C.c_char_p, C.c_int]
    buf = C.create_string_buffer(400)

    def native(v, n):
        k = fe.lbfe_py_round_repr(v, n, buf, 400)
        assert k >= 0
        return buf.value.decode()

    rng = np.random.default_rng(17)
    vals = list(rng.gamma(2.0, 2.0, 4000)) + list(rng.uniform(0, 1e-3, 500)) + list(rng.uniform(0, 5e4, 500))
    vals += [0.0, -0.0, 0.5, 1.5, 2.5, 0.125, 0.375, 2.675, 1.005, 1e-5, 1.23456e-7, 9.995, 99.995, 0.99999999, 1e15, 1e16, 1.5e16,
             123456789012345.678, 2.0 ** -20, 2.0 ** 60, 1e22, 1e23, float("inf"), float("nan"), -3.14159, -0.004, 5e-324, 1.7976931348623157e308]
    vals += [k / 1000.0 for k in range(0, 4000, 5)] + [k + 0.5 for k in range(20)]
    vals += list(10.0 ** rng.uniform(-6, 16.5, 3000)) + [k / 16.0 + 1 / 32 for k in range(400)]      # around the exact fast path's limits; binary ties
    for n in (0, 1, 2, 3, 4, 6, 9, 12, 15, 16, 17, 20, -1, -2, -5, 400, -400):
        for v in vals:
            v = float(v)
            assert native(v, n) == repr(round(v, n)), (v, n)


def test_rmsds_align_command_line(system, tmp_path):
    """What rmsds-align decides before any GPU work, as Packages/Clustering/rmsds-align.py:118-206 does:
    required positionals (argparse: status 2), the range forms, "Error in range" on stdout with status 0,
    frames beyond the trajectory, and -- without a GPU -- a loud failure with an empty stdout."""
    import shutil
    exe = os.path.join(ROOT, "loos_b200", "bin", "rmsds-align")
    run = lambda *a: subprocess.run([exe] + [str(x) for x in a], capture_output=True, text=True)
    r = run(system["pdb"])
    assert r.returncode == 2 and "the following arguments are required" in r.stderr and r.stdout == ""
    r = run("--range1", "0", system["pdb"], system["dcd"])
    assert r.returncode == 0 and r.stdout == "Error in range\n"
    r = run("--range1", "0:1:2:3", system["pdb"], system["dcd"])
    assert r.returncode == 0 and r.stdout == "Error in range\n"
    r = run("--range1", "0:1:5", "--range2", "0:1", system["pdb"], system["dcd"])      # range_2[2] in the script
    assert r.returncode == 1 and "IndexError" in r.stderr
    r = run("--range1", "0:0:5", system["pdb"], system["dcd"])
    assert r.returncode == 1 and "range() arg 3 must not be zero" in r.stderr
    r = run("--range1", "0:1:100000", system["pdb"], system["dcd"])
    assert r.returncode == 1 and "out of range" in r.stderr and r.stdout == ""
    # the open-ended form stops at len(file name): a long name asks for frames the trajectory does not have
    long_name = tmp_path / ("x" * 150 + ".dcd")
    shutil.copy(system["dcd"], long_name)
    r = run(system["pdb"], long_name)
    assert r.returncode == 1 and "length of the trajectory's file name" in r.stderr and "out of range" in r.stderr
    import torch
    if not torch.cuda.is_available():
        r = run("--range1", "0:1:4", system["pdb"], system["dcd"])
        assert r.returncode == 1 and "no usable sm_100" in r.stderr and r.stdout == ""


@pytest.mark.parametrize("fmt", ["dcd", "dcd_swapped_crystal", "nc", "xtc", "h5"])
def test_trajectory_reads_are_reentrant(fe, tmp_path, fmt):
    """Trajectory::readFramePlanes from 6 threads at once (positional reads, per-thread scratch) gives the
    frames a single thread reads, for every format: the tools decode a staging batch in parallel."""
    from oracle import frontend_oracle as fo
    from loos_b200.synth import write_dcd
    rng = np.random.default_rng(8)
    F, N = 57, 311
    xyz = (rng.normal(0, 15, (1, N, 3)) + rng.normal(0, 1.0, (F, N, 3))).astype(np.float32)
    path = str(tmp_path / ("t." + ("dcd" if fmt.startswith("dcd") else fmt)))
    if fmt == "dcd":
        write_dcd(path, xyz)
    elif fmt == "dcd_swapped_crystal":
        write_dcd(path, xyz, with_crystal=True, big_endian=True)
    elif fmt == "nc":
        _write_amber_nc(path, xyz)
    elif fmt == "h5":
        fo.write_mdtraj_h5(path, xyz, chunk=(5, 100, 3))
    else:
        fo.write_xtc(path, xyz)
    order = rng.permutation(np.arange(F).repeat(3)).astype(np.uint32)       # every frame three times, shuffled
    many = np.zeros((len(order), 3, N), np.float32)
    assert fe.lbfe_traj_read_many(path.encode(), N, order.ctypes.data_as(_u32p), len(order), 6,
                                  many.ctypes.data_as(C.POINTER(C.c_float))) == 0, fe.lbfe_last_error()
    one = np.zeros((3, N), np.float32)
    for k in range(F):
        assert fe.lbfe_traj_read(path.encode(), N, k, one.ctypes.data_as(C.POINTER(C.c_float))) == 0
        for pos in np.nonzero(order == k)[0]:
            assert np.array_equal(many[pos], one)
        if fmt not in ("xtc", "h5"):
            assert np.array_equal(one, xyz[k].T)
    # a frame beyond the end is an error from any thread
    bad = np.array([0, 1, F, 2], dtype=np.uint32)
    assert fe.lbfe_traj_read_many(path.encode(), N, bad.ctypes.data_as(_u32p), 4, 3, many.ctypes.data_as(C.POINTER(C.c_float))) == -1


# ------------------------------------------------------------------ MDTraj HDF5 reader (src/mdtrajtraj.cpp)
H5_CASES = {
    "pytables_like": dict(style="v1"),                                        # chunked, shuffle + deflate, B-tree v1, symbol-table group
    "chunks_split_atoms_and_axes": dict(style="v1", chunk=(3, 50, 2), fletcher=True),
    "uncompressed_chunks": dict(style="v1", compress=False, shuffle=False, chunk=(7, 400, 3)),
    "deep_chunk_btree": dict(style="v1", chunk=(1, 16, 3), node_fanout=3),
    "big_endian_double": dict(style="v1", dtype=">f8"),
    "user_block": dict(style="v1", userblock=1024),
    "latest_contiguous": dict(style="v2", chunk=None),
    "latest_single_chunk": dict(style="v2"),
    "no_box": dict(style="v1", box=False),
}


@pytest.mark.parametrize("case", sorted(H5_CASES))
def test_mdtraj_h5_reader(fe, tmp_path, case):
    """Frames of an MDTraj-style HDF5 file as the reference produces them: 10.0 * (the stored nm value) per
    coordinate (src/mdtrajtraj.cpp:118-123), here rounded to the float32 planes every format is staged as.
    The files come from the tests' own restatement of the HDF5 format (oracle/frontend_oracle.py): parity with
    files written by MDTraj itself is unpinned -- no HDF5 library exists in this image."""
    from oracle import frontend_oracle as fo
    rng = np.random.default_rng(31)
    F, N = 13, 257
    xyz = rng.normal(0, 25, (F, N, 3)).astype(np.float32)
    kw = H5_CASES[case]
    path = str(tmp_path / "t.h5")
    fo.write_mdtraj_h5(path, xyz, **kw)
    stored = (xyz.astype(np.float64) / 10.0).astype(kw.get("dtype", "<f4"))      # what the file holds, nm
    want = (10.0 * stored.astype(np.float64)).astype(np.float32)
    na, nf = C.c_uint32(), C.c_uint32()
    assert fe.lbfe_traj_info(path.encode(), N, C.byref(na), C.byref(nf)) == 0, fe.lbfe_last_error()
    assert (na.value, nf.value) == (N, F)
    buf = np.zeros((3, N), np.float32)
    for k in list(range(F)) + [4, 0]:
        assert fe.lbfe_traj_read(path.encode(), N, k, buf.ctypes.data_as(C.POINTER(C.c_float))) == 0, fe.lbfe_last_error()
        assert np.array_equal(buf, want[k].T), (case, k)
    assert fe.lbfe_traj_read(path.encode(), N, F, buf.ctypes.data_as(C.POINTER(C.c_float))) == -1


def test_mdtraj_h5_validation(fe, tmp_path):
    """The checks of MDTrajTraj::init (src/mdtrajtraj.cpp:52-72) and what this reader refuses to guess at."""
    from oracle import frontend_oracle as fo
    xyz = np.random.default_rng(2).normal(0, 10, (4, 20, 3)).astype(np.float32)
    na, nf = C.c_uint32(), C.c_uint32()

    def err(path, natoms=20):
        assert fe.lbfe_traj_info(path.encode(), natoms, C.byref(na), C.byref(nf)) == -1
        return fe.lbfe_last_error().decode()

    p = str(tmp_path / "a.h5"); fo.write_mdtraj_h5(p, xyz, coords_name="positions")
    assert "No coordinates dataset found in HDF5" in err(p)
    p = str(tmp_path / "b.h5"); fo.write_mdtraj_h5(p, xyz)
    assert "Number of atoms in HDF5 does not match the AtomicGroup" in err(p, natoms=21)
    p = str(tmp_path / "c.h5"); fo.write_mdtraj_h5(p, xyz, extra_filter=32001)
    assert "filter 32001 (blosc)" in err(p)
    p = str(tmp_path / "d.h5"); open(p, "wb").write(b"CDF\x01" + b"\0" * 600)
    assert "not an HDF5 file" in err(p)
    p = str(tmp_path / "e.h5"); fo.write_mdtraj_h5(p, xyz)
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:len(raw) // 2])                                    # cut in half: addresses point past the end
    assert "truncated or corrupt HDF5 file" in err(p)
    p = str(tmp_path / "f.h5"); fo.write_mdtraj_h5(p, xyz[:0])                  # no frames yet: opens, zero frames
    assert fe.lbfe_traj_info(p.encode(), 20, C.byref(na), C.byref(nf)) == 0 and nf.value == 0


@pytest.mark.parametrize("fmt", ["dcd", "nc", "xtc", "h5", "h5v2", "text"])
def test_corrupt_files_never_crash(fe, tmp_path, fmt):
    """120 corrupted copies of a valid file per format (tests/host/fuzz_readers.py): every one is either
    refused with an error or read; the process survives.  (Run with thousands of mutants under
    AddressSanitizer / UBSan while the readers were written; this is the cheap regression guard.)"""
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host", "fuzz_readers.py"), fmt, "0", "120", str(tmp_path)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    assert "survived 120 mutants" in r.stdout


@pytest.mark.parametrize("rows,cols,threads", [(1, 1, 1), (7, 5, 3), (300, 4001, 4), (2500, 9, 8), (1100, 1000, 5)])
def test_py_round_matrix_text(fe, rows, cols, threads):
    """The body of rmsds-align's output (blocks of rows formatted on several threads, written in order) equals
    what the script's print loop produces (oracle.py_round_text = Python's own round/repr), for shapes that
    give one block, many blocks and a ragged last wave."""
    from oracle.oracle import py_round_text
    rng = np.random.default_rng(rows * 31 + cols)
    M = np.asfortranarray(rng.gamma(2.0, 2.0, (rows + 3, cols)))      # ld > rows
    M[0, 0] = 0.0
    fe.lbfe_py_round_matrix.restype = C.c_long
    fe.lbfe_py_round_matrix.argtypes = [C.POINTER(C.c_double), C.c_long, C.c_long, C.c_long, C.c_int, C.c_int, C.c_char_p, C.c_long]
    cap = rows * cols * 12 + rows + 64
    buf = C.create_string_buffer(cap)
    for precision in (2, 5):
        n = fe.lbfe_py_round_matrix(M.ctypes.data_as(C.POINTER(C.c_double)), rows, cols, rows + 3, precision, threads, buf, cap)
        assert n > 0
        assert buf.raw[:n].decode() == py_round_text(M[:rows], precision)


# ------------------------------------------------------------------ XTC pinned to the reference's own code
XTC_GOLDEN = os.path.join(ROOT, "tests", "golden", "xtc_ref.npz")


@pytest.mark.parametrize("case", ["protein", "water", "wide", "tiny", "huge_box"])
def test_xtc_against_the_references_own_writer_and_reader(fe, tmp_path, case):
    """tests/golden/xtc_ref.npz holds files written by the reference's own XTCWriter (src/xtcwriter.cpp) and the
    GCoords its own reader (src/xtc.cpp) decodes from them, both compiled from the sources where they lie
    (oracle/_ref/libloosxtc.so, tests/golden/make_golden.py::xtc_golden).  The native reader reproduces those
    coordinates bit for bit (narrowed to the float32 every format is staged as), and the tests' restatement of the
    writer reproduces the reference's files byte for byte -- so the other XTC tests stand on pinned ground."""
    from oracle import frontend_oracle as fo
    g = np.load(XTC_GOLDEN)
    xyz, data, decoded = g[case + "_xyz"], g[case + "_file"].tobytes(), g[case + "_decoded"]
    F, N, _ = xyz.shape
    path = str(tmp_path / "ref.xtc")
    open(path, "wb").write(data)
    got = _xtc_read_native(fe, path, N, F)
    assert np.array_equal(got.view(np.uint32), decoded.astype(np.float32).view(np.uint32))
    assert np.array_equal(fo.read_xtc(path).view(np.uint32), got.view(np.uint32))
    mine = str(tmp_path / "mine.xtc")
    fo.write_xtc(mine, xyz)
    assert open(mine, "rb").read() == data


def test_xtc_reference_library_live():
    """Where the reference library itself is present (the build container): regenerate the fixture's content and
    compare, plus a case that is not in the fixture."""
    from oracle.oracle import RefXTC
    if not RefXTC.available():
        pytest.skip("oracle/_ref/libloosxtc.so not built (no /root/reference here)")
    from oracle import frontend_oracle as fo
    R = RefXTC()
    g = np.load(XTC_GOLDEN)
    for case in ("protein", "water", "tiny"):
        xyz = g[case + "_xyz"]
        data = R.write(xyz)
        assert data == g[case + "_file"].tobytes()
        assert np.array_equal(R.read(data, len(xyz), xyz.shape[1]), g[case + "_decoded"])
    rng = np.random.default_rng(99)
    xyz = (rng.normal(0, 18, (1, 517, 3)) + rng.normal(0, 0.8, (6, 517, 3))).astype(np.float32)
    data = R.write(xyz, precision=100.0)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "a.xtc")
        fo.write_xtc(p, xyz, precision=100.0)
        assert open(p, "rb").read() == data
        assert np.array_equal(fo.read_xtc(p), R.read(data, 6, 517).astype(np.float32))


# ------------------------------------------------------------------ DCD pinned to the reference's own code
DCD_CASES = ["ref_written", "ref_written_box", "plain", "crystal", "big_endian", "be_crystal", "nframes_from_size"]


@pytest.mark.parametrize("case", DCD_CASES)
def test_dcd_against_the_references_own_reader_and_writer(fe, tmp_path, case):
    """tests/golden/dcd_ref.npz: DCD files written by the reference's own DCDWriter (src/dcdwriter.cpp) and the
    test-suite's own variants (byte-swapped, crystal records, frame count taken from the file size), each with what
    the reference's own reader (src/dcd.cpp readHeader / seekFrameImpl / parseFrame, compiled in place:
    oracle/_ref/libloosdcd.so) makes of it.  The native reader agrees on every count and every coordinate bit."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "dcd_ref.npz"))
    data, want, info = g[case + "_file"].tobytes(), g[case + "_decoded"], g[case + "_info"]
    path = str(tmp_path / "t.dcd")
    open(path, "wb").write(data)
    na, nf, cr = C.c_uint32(), C.c_uint32(), C.c_int()
    assert fe.lbfe_dcd_info(path.encode(), C.byref(na), C.byref(nf), C.byref(cr)) == 0, fe.lbfe_last_error()
    assert (na.value, nf.value, cr.value) == (int(info[0]), int(info[1]), int(info[2]))
    buf = np.zeros((3, na.value), np.float32)
    for k in range(nf.value):
        assert fe.lbfe_traj_read(path.encode(), na.value, k, buf.ctypes.data_as(C.POINTER(C.c_float))) == 0
        assert np.array_equal(buf.T.view(np.uint32), want[k].view(np.uint32))
    assert np.array_equal(want, g["xyz"])                       # and the reference reader returns what was written


def test_dcd_reference_library_live(tmp_path):
    from oracle.oracle import RefDCD
    if not RefDCD.available():
        pytest.skip("oracle/_ref/libloosdcd.so not built (no /root/reference here)")
    from loos_b200.synth import write_dcd
    from oracle import frontend_oracle as fo
    R = RefDCD()
    g = np.load(os.path.join(ROOT, "tests", "golden", "dcd_ref.npz"))
    assert R.write(g["xyz"]) == g["ref_written_file"].tobytes()
    assert R.write(g["xyz"], (61.0, 62.0, 63.0)) == g["ref_written_box_file"].tobytes()
    xyz = np.random.default_rng(12).normal(0, 40, (5, 1003, 3)).astype(np.float32)
    p = str(tmp_path / "a.dcd")
    write_dcd(p, xyz, big_endian=True, with_crystal=True)
    got = R.read(open(p, "rb").read(), 5, 1003)
    assert got["swabbed"] and got["crystal"] and got["nframes"] == 5 and np.array_equal(got["xyz"], xyz)
    assert np.array_equal(fo.read_dcd(p).transpose(0, 2, 1), xyz)     # the Python restatement (planes per frame) agrees too


# ------------------------------------------------------------------ matrix text pinned to the reference's own code
def test_matrix_text_and_reader_against_the_references_own_code(fe):
    """tests/golden/matrix_ref.npz: what the reference's own `os << setprecision(p) << M` (src/MatrixImpl.hpp:208-222)
    prints for a float matrix with awkward values at p = 0 ... 12, and what its own readAsciiMatrix
    (src/MatrixRead.hpp:102-143) returns or throws for good and bad input -- both compiled from the headers where they
    lie (oracle/_ref/libloosmatrix.so).  The oracle's matrix_text, the native writer and the native reader agree with
    them character for character, message for message."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "matrix_ref.npz"))
    M = np.asfortranarray(g["M"])
    m, n = M.shape
    cap = 1 << 20
    out = C.create_string_buffer(cap)
    for p in (0, 1, 2, 3, 6, 8, 9, 12):
        want = str(g["text_p%d" % p])
        assert matrix_text(M, p) == want
        k = fe.lbfe_matrix_text(M.ctypes.data_as(C.POINTER(C.c_float)), m, n, m, p, out, cap)
        assert out.raw[:k].decode() == want
    fe.lbfe_read_matrix.restype = C.c_int
    buf = np.zeros(64, np.float32)
    mm, nn = C.c_long(), C.c_long()
    for text, err, vals in zip(g["bad_inputs"], g["bad_errors"], g["bad_values"]):
        data = str(text).encode()
        rc = fe.lbfe_read_matrix(data, len(data), buf.ctypes.data_as(C.POINTER(C.c_float)), buf.size, C.byref(mm), C.byref(nn))
        if str(err):
            assert rc == -1 and fe.lbfe_last_error().decode() == str(err), (text, fe.lbfe_last_error())
        else:
            assert rc == 0
            got = np.array2string(buf[:mm.value * nn.value], separator=",")
            assert got == str(vals)


def test_matrix_reference_library_live():
    from oracle.oracle import RefMatrix
    if not RefMatrix.available():
        pytest.skip("oracle/_ref/libloosmatrix.so not built (no /root/reference here)")
    R = RefMatrix()
    g = np.load(os.path.join(ROOT, "tests", "golden", "matrix_ref.npz"))
    assert R.text(g["M"], 2) == str(g["text_p2"])
    rng = np.random.default_rng(77)
    M = (10.0 ** rng.uniform(-6, 7, (40, 40))).astype(np.float32)
    for p in (2, 5, 9):
        assert R.text(M, p) == matrix_text(M, p)
    back, err = R.read(matrix_text(M, 9))
    assert err is None and np.array_equal(back, M)


# ------------------------------------------------------------------ PDB records, hybrid-36, MultiTrajectory pinned to the reference's own code
def test_pdb_records_hybrid36_multitraj_against_the_references_own_code(fe, tmp_path):
    """tests/golden/pdb_ref.json: what the reference's own PDB::parseAtomRecord (src/pdb.cpp:65-156, with parseStringAs
    and parseStringAsHybrid36 of src/utils.{hpp,cpp}) stores or throws for 15 ATOM/HETATM lines -- hybrid-36 ids, a
    negative resid with an insertion code, frame-shifted resids, truncated lines (Atom::init defaults survive: occupancy
    1, segid four blanks), exponent notation, unparsable fields --, parseStringAsHybrid36 on 23 fields,
    MultiTrajectory::nframes(i) / frameIndexToLocation (src/MultiTraj.{hpp,cpp}) for 120 random (sizes, skip, stride)
    and uniquifyVector, all compiled from the sources where they lie (oracle/_ref/libloospdb.so).  The native
    front-end and the Python restatement reproduce every field, value and message."""
    import json
    from oracle import frontend_oracle as fo
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "pdb_ref.json")))
    names = ("record", "name", "altloc", "resname", "chain", "icode", "segid", "element")
    for k, rec in enumerate(g["atoms"]):
        line, want = rec["line"], rec["parsed"]
        ints, strs, reals = (C.c_int * 2)(), C.create_string_buffer(128), (C.c_double * 5)()
        rc = fe.lbfe_pdb_atom_line(line.encode(), ints, strs, reals)
        if isinstance(want, str):                                   # the reference threw: same message
            assert rc == -1 and fe.lbfe_last_error().decode() == want, (line, fe.lbfe_last_error())
            continue
        assert rc == 0, (line, fe.lbfe_last_error())
        got = {"id": ints[0], "resid": ints[1], "x": reals[0], "y": reals[1], "z": reals[2], "occupancy": reals[3], "bfactor": reals[4]}
        for j, nm in enumerate(names):
            got[nm] = strs.raw[16 * j:16 * j + 16].split(b"\0")[0].decode()
        assert got == want, (line, got, want)
        p = tmp_path / ("l%d.pdb" % k)                              # and the Python restatement of the reader
        p.write_text(line + "\n")
        try:
            a = fo.read_pdb(str(p))[0]
        except ValueError:                                          # it float()s whole fields: not for the mis-columned lines
            continue
        assert (a["id"], a["resid"], a["name"], a["resname"], a["segid"]) == (want["id"], want["resid"], want["name"], want["resname"], want["segid"])
    for rec in g["hybrid36"]:
        f = rec["field"]
        assert fe.lbfe_hybrid36(f.encode(), len(f)) == rec["value"], f
    P = C.POINTER(C.c_uint32)
    for rec in g["multitraj"]:
        sizes = np.array(rec["sizes"], dtype=np.uint32)
        want = [tuple(x) for x in rec["locations"]]
        lt, lf = np.zeros(len(want) + 1, np.uint32), np.zeros(len(want) + 1, np.uint32)
        n = fe.lbfe_multitraj(sizes.ctypes.data_as(P), len(sizes), rec["skip"], rec["stride"], lt.ctypes.data_as(P), lf.ctypes.data_as(P), len(lt))
        assert n == len(want) and list(zip(lt[:n].tolist(), lf[:n].tolist())) == want
        assert fo.multitraj_locations(rec["sizes"], rec["skip"], rec["stride"]) == want
    fe.lbfe_uniquify.restype = C.c_long
    for rec in g["uniquify"]:
        v = np.array(rec["in"], dtype=np.uint32)
        out = np.zeros(len(v) + 1, np.uint32)
        n = fe.lbfe_uniquify(v.ctypes.data_as(P), len(v), out.ctypes.data_as(P))
        assert out[:n].tolist() == rec["out"]


def test_pdb_reference_library_live():
    from oracle.oracle import RefPDB
    if not RefPDB.available():
        pytest.skip("oracle/_ref/libloospdb.so not built (no /root/reference here)")
    import json
    R = RefPDB()
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "pdb_ref.json")))
    for rec in g["atoms"]:
        assert R.atom(rec["line"]) == rec["parsed"]
    for rec in g["hybrid36"]:
        assert R.hybrid36(rec["field"]) == rec["value"]
    assert [list(x) for x in R.multitraj([5, 0, 9], 2, 3)] == [[0, 2], [2, 2], [2, 5], [2, 8]]


def test_backbone_and_hydrogen_selectors_against_the_references_own_code(fe, tmp_path):
    """tests/golden/pdb_ref.json["selectors"]: the decisions of the reference's own BackboneSelector (its two sorted
    tables, src/Selectors.cpp:37-132) and HydrogenSelector (:156-164), compiled in place, for every pair of 85 residue
    names x 60 atom names (the tables themselves plus near misses: other case, CHARMM/AMBER variants, waters, ions).
    The native `backbone` and `hydrogen` keywords select exactly the same atoms of a PDB holding all the pairs."""
    import json
    from oracle import frontend_oracle as fo
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "pdb_ref.json")))["selectors"]
    resnames, names = g["resnames"], g["names"]
    lines, want_bb, want_h = [], [], []
    k = 0
    for i, r in enumerate(resnames):
        for j, nm in enumerate(names):
            k += 1
            # name in columns 13-16 as given (left-justified), residue name in 18-21; parseStringAs strips blanks
            lines.append("ATOM  %5d %-4s %-4s%1s%4d    %8.3f%8.3f%8.3f  1.00  0.00      SEG " % (k % 100000, nm, r, "A", (k // 60) % 10000, 1.0, 2.0, 3.0))
            want_bb.append(g["backbone"][i][j])
            want_h.append(g["hydrogen"][j])
    path = str(tmp_path / "pairs.pdb")
    open(path, "w").write("\n".join(lines) + "\nEND\n")
    out = np.zeros(len(lines), np.uint32)
    for sel, want in (("backbone", want_bb), ("hydrogen", want_h)):
        n = fe.lbfe_select(path.encode(), sel.encode(), out.ctypes.data_as(_u32p), len(out))
        assert n >= 0, fe.lbfe_last_error()
        got = np.zeros(len(lines), int)
        got[out[:n]] = 1
        assert got.tolist() == want, sel
        atoms = fo.read_pdb(path)
        assert list(fo.select(atoms, sel)) == [i for i, w in enumerate(want) if w]      # and the Python restatement
    assert sum(want_bb) == 48 * 33


# ------------------------------------------------------------------ the selection language pinned to the reference's own code
def test_selection_language_against_the_references_own_scanner_grammar_and_kernel(fe, tmp_path):
    """tests/golden/selection_ref.json: 843 selection strings -- every pairing and triple of a set of predicates under
    && || ! not and parentheses (precedence, associativity), keyword/keyword and literal/keyword comparisons, string
    ordering, regular expressions, the -> operator, comments, odd spacing, wrong case, malformed input -- run through the
    reference's OWN selection language (its pre-generated flex scanner and bison grammar, Kernel*.cpp, Atom.cpp and the
    backbone / hydrogen predicates, compiled unmodified from the sources where they lie: oracle/_ref/libloosselect.so) on
    a 432-atom model.  The native parser selects exactly the same atoms for every accepted string and refuses exactly
    the strings the reference refuses; so does the Python restatement the other tests use."""
    import json
    from oracle import frontend_oracle as fo
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "selection_ref.json")))
    path = str(tmp_path / "model.pdb")
    open(path, "w").write(g["pdb"])
    n = g["natoms"]
    atoms = fo.read_pdb(path)
    assert len(atoms) == n
    out = np.zeros(n, np.uint32)
    refused = 0
    for case in g["cases"]:
        sel = case["selection"]
        k = fe.lbfe_select(path.encode(), sel.encode(), out.ctypes.data_as(_u32p), n)
        if "error" in case:
            refused += 1
            assert k == -1, (sel, "the reference refuses this:", case["error"])
            with pytest.raises(Exception):
                fo.select(atoms, sel)
            continue
        assert k >= 0, (sel, fe.lbfe_last_error())
        want = [i for i, c in enumerate(case["mask"]) if c == "1"]
        assert out[:k].tolist() == want, sel
        assert list(fo.select(atoms, sel)) == want, sel
    assert refused >= 30 and len(g["cases"]) >= 800


def test_selection_reference_library_live(tmp_path):
    from oracle.oracle import RefSelect
    if not RefSelect.available():
        pytest.skip("oracle/_ref/libloosselect.so not built (no /root/reference here)")
    import json
    from oracle import frontend_oracle as fo
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "selection_ref.json")))
    path = str(tmp_path / "model.pdb")
    open(path, "w").write(g["pdb"])
    R = RefSelect(fo.read_pdb(path))
    for case in g["cases"][::7]:
        got, err = R.select(case["selection"])
        if "error" in case:
            assert got is None
        else:
            assert got == [i for i, c in enumerate(case["mask"]) if c == "1"]
