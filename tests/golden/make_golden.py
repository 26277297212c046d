"""Generate tests/golden/*.npz from the REFERENCE's own arithmetic
(oracle/_ref/libloosref.so = /root/reference/src/alignment.cpp:12-281 compiled in
place against OpenBLAS 0.3.15).  Run in the build container, where /root/reference
exists; the fixtures travel to the GPU box, the reference does not.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from loos_b200.synth import SEED0, synth_frames  # noqa: E402
from oracle.oracle import Checker, matrix_text  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    R = Checker("reference")
    # (1) a small all-to-all case with ragged sizes (F, N not multiples of any tile)
    X = synth_frames(53, 137, seed=SEED0 + 101)
    M = R.rmsds_self(X.astype(np.float64), nthreads=4)
    Y = synth_frames(31, 137, seed=SEED0 + 102)
    Mx = R.rmsds_cross(X.astype(np.float64), Y.astype(np.float64), nthreads=4)
    np.savez_compressed(os.path.join(HERE, "rmsds_small.npz"), X=X, Y=Y, M=M, Mx=Mx,
                        text_p2=np.array(matrix_text(M, 2)), text_p8=np.array(matrix_text(M, 8)))
    # (2) analytic / degenerate known-answer cases evaluated by the reference
    rng = np.random.default_rng(SEED0 + 103)
    a = synth_frames(1, 64, seed=SEED0 + 104)[0].astype(np.float64)
    th = 0.7
    Rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    cases = {
        "identical": (a, a.copy()),
        "rigid": (a, (a @ Rz.T + [3.0, -2.0, 9.0]).astype(np.float32).astype(np.float64)),
        "mirror": (a, a * [1, 1, -1]),
        "planar": (a * [1, 1, 0], (a * [1, 1, 0]) @ Rz.T + rng.normal(size=a.shape) * [0.2, 0.2, 0]),
        "collinear": (a * [1, 0, 0], a * [1.1, 0, 0] + 0.5),
        "n1": (a[:1], a[1:2]),
        "n2": (a[:2], a[5:7]),
        "n3": (a[:3], a[7:10]),
        "noisy": (a, a + rng.normal(size=a.shape) * 0.4),
    }
    out = {}
    for k, (u, v) in cases.items():
        u = u.astype(np.float32); v = v.astype(np.float32)
        pair = np.stack([u, v])
        out[k + "_xyz"] = pair
        out[k + "_rmsd"] = np.array(R.rmsds_self(pair.astype(np.float64))[1, 0], dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "kat_pairs.npz"), **out)
    # (3) rmsd2ref with a target, align selection != rmsd selection
    F, N = 40, 90
    Z = synth_frames(F + 1, N, seed=SEED0 + 105).astype(np.float64)
    sel_a = np.arange(0, N, 3)
    sel_r = np.array([i for i in range(N) if i % 5 != 0])
    tgt = Z[-1]
    d, xf = R.rmsd2ref_target(Z[:F, sel_a], Z[:F, sel_r], tgt[sel_a], tgt[sel_r])
    it = R.rmsd2ref_iterative(Z[:F, sel_a], Z[:F, sel_r], 1e-6, 1000)
    np.savez_compressed(os.path.join(HERE, "rmsd2ref_small.npz"), Z=Z.astype(np.float32), sel_a=sel_a,
                        sel_r=sel_r, rmsd=d, xforms=xf, it_rmsd=it[0], it_xforms=it[1], it_avg=it[2],
                        it_iters=np.array(it[3]), it_rms=np.array(it[4]))
    align_golden(R)
    print("golden fixtures written to", HERE)


def align_golden(R):
    """(4) rmsds-align: the script's loop (Packages/Clustering/rmsds-align.py:236-256) on the reference's kabsch,
    two trajectories, align selection != rmsd selection, plus the text the script prints at precision 3."""
    from oracle.oracle import py_round_text, rmsds_align
    F1, F2, N = 14, 9, 70
    X = synth_frames(F1, N, seed=SEED0 + 106)
    Y = synth_frames(F2, N, seed=SEED0 + 107)
    sel_a = np.arange(0, 40, 2)
    sel_r = np.arange(25, N)
    M = rmsds_align(R, X[:, sel_a], X[:, sel_r], Y[:, sel_a], Y[:, sel_r])
    M1 = rmsds_align(R, X[:, sel_a], X[:, sel_r], X[:, sel_a], X[:, sel_r])
    np.savez_compressed(os.path.join(HERE, "rmsds_align_small.npz"), X=X, Y=Y, sel_a=sel_a, sel_r=sel_r, M=M, M_self=M1,
                        text_p3=np.array(py_round_text(M, 3)))


def xtc_cases():
    """The five XTC test shapes (also used by tests/test_frontend_cpu.py): protein, water triplets, a wide box in which
    the small-range index walks both ways, <= 9 atoms (stored as plain floats), extent above 2^24 quanta."""
    from loos_b200.synth import synth_system
    rng = np.random.default_rng(5)
    out = {}
    out["protein"] = synth_system(40, 5, seed=3, flavour="rich")[1]
    O = rng.uniform(0, 40, (4, 300, 1, 3))
    out["water"] = (O + rng.normal(0, 0.4, (4, 300, 3, 3))).reshape(4, 900, 3).astype(np.float32)
    w = rng.uniform(-300, 300, (3, 200, 3)).astype(np.float32)
    w[:, 50:120] = (w[:, 50:51] + rng.normal(0, 0.05, (3, 70, 3))).astype(np.float32)
    out["wide"] = w
    out["tiny"] = rng.uniform(-20, 20, (5, 7, 3)).astype(np.float32)
    h = rng.uniform(-1.2e5, 1.2e5, (2, 64, 3)).astype(np.float32)
    h[:, 10] = h[:, 9] + 1.0
    out["huge_box"] = h
    return out


def xtc_golden():
    """(5) XTC: files written by the reference's own XTCWriter and the coordinates its own XTC reader decodes
    from them (oracle/_ref/libloosxtc.so)."""
    from oracle.oracle import RefXTC
    R = RefXTC()
    out = {}
    for name, xyz in xtc_cases().items():
        data = R.write(xyz)
        out[name + "_xyz"] = xyz
        out[name + "_file"] = np.frombuffer(data, dtype=np.uint8)
        out[name + "_decoded"] = R.read(data, len(xyz), xyz.shape[1])
    np.savez_compressed(os.path.join(HERE, "xtc_ref.npz"), **out)


def dcd_golden():
    """(6) DCD: files written by the reference's own DCDWriter (with and without a periodic box) and what its own DCD
    reader makes of them and of the test-suite's own DCD variants (byte-swapped, crystal records, frame count from the
    file size) -- oracle/_ref/libloosdcd.so."""
    import tempfile

    from loos_b200.synth import write_dcd
    from oracle.oracle import RefDCD
    R = RefDCD()
    rng = np.random.default_rng(3)
    F, N = 11, 97
    xyz = rng.normal(0, 30, (F, N, 3)).astype(np.float32)
    out = {"xyz": xyz}
    for name, box in (("ref_written", None), ("ref_written_box", (61.0, 62.0, 63.0))):
        data = R.write(xyz, box)
        got = R.read(data, F, N)
        out[name + "_file"] = np.frombuffer(data, dtype=np.uint8)
        out[name + "_decoded"] = got["xyz"]
        out[name + "_info"] = np.array([got["natoms"], got["nframes"], got["crystal"], got["swabbed"]])
    with tempfile.TemporaryDirectory() as d:
        for name, kw in (("plain", {}), ("crystal", dict(with_crystal=True)), ("big_endian", dict(big_endian=True)),
                         ("be_crystal", dict(big_endian=True, with_crystal=True)), ("nframes_from_size", dict(header_nframes=0))):
            p = os.path.join(d, "t.dcd")
            write_dcd(p, xyz, **kw)
            data = open(p, "rb").read()
            got = R.read(data, F, N)
            out[name + "_file"] = np.frombuffer(data, dtype=np.uint8)
            out[name + "_decoded"] = got["xyz"]
            out[name + "_info"] = np.array([got["natoms"], got["nframes"], got["crystal"], got["swabbed"]])
    np.savez_compressed(os.path.join(HERE, "dcd_ref.npz"), **out)


def matrix_golden():
    """(7) matrix text: what the reference's own operator<< prints for a float matrix at several precisions, and what its
    own readAsciiMatrix makes of good and bad input (oracle/_ref/libloosmatrix.so)."""
    from oracle.oracle import RefMatrix
    R = RefMatrix()
    rng = np.random.default_rng(4)
    M = np.asfortranarray(rng.gamma(2.0, 2.5, (37, 23)).astype(np.float32))
    for k, v in enumerate([0, 1e-7, 123456.789, 9.9999, 0.00012345, 1e10, 99.5, 0.5, 2.5, 0.25, 1e-5, 3.4e38, 1.17549435e-38, 16777216.0]):
        M[k, k] = v
    out = {"M": M}
    for p in (0, 1, 2, 3, 6, 8, 9, 12):
        out["text_p%d" % p] = np.array(R.text(M, p))
    bad = ["no marker here\n1 2 3\n", "# 2 2 (0)\n1 2 \n3 x \n", "# 2 2 (0)\n1 2 \n3 \n", "# 0 5 (0)\n", "# 2 2 (0)\n1 2 3 4 5 6\n",
           "junk\n# header line\n# 2 3 (0)\n1e-3 2.5 -7 \n4 5 6 \n"]
    out["bad_inputs"] = np.array(bad)
    res = [R.read(b) for b in bad]
    out["bad_errors"] = np.array([e or "" for _, e in res])
    out["bad_values"] = np.array([np.array2string(m.ravel(order="F"), separator=",") if m is not None else "" for m, _ in res])
    np.savez_compressed(os.path.join(HERE, "matrix_ref.npz"), **out)


PDB_LINES = [
    "ATOM      1  N   ALA A   1      11.104   6.134  -6.504  1.00  0.00      PROT N",
    "ATOM  A0000  CA  ALA A9999      -1.500 123.456-100.250  0.50 12.34      PROT C",
    "HETATM99999  OH2 TIP3W a000       0.000   0.000   0.000  1.00  0.00      SOLV O",
    "ATOM     12 HA12 LYS B -12A      1.000   2.000   3.000",
    "ATOM     13  CB  LYS B  13       1.000   2.000   3.000  0.75",
    "ATOM     14  CG  LYS B  14       1.000   2.000   3.000  0.75 20.00",
    "ATOM     15  CD  LYS B10015       1.000   2.000   3.000  1.00  0.00      SEG1 C",
    "ATOM     15  CD  LYS B 10015      1.000   2.000   3.000  1.00  0.00      SEG1",
    "ATOM  1e+04  CE  LYS B  16       1.5e1   2.000   3.000  1.00  0.00      SEG2 C1",
    "ATOM     17  NZ  LYS B  17       abc     2.000   3.000",
    "ATOM     18  C   LYS B  18 ",
    "ATOM  zzzzz  O   LYSXB ZZZZ      9.999  -0.001   0.000  1.00  0.00      Q",
    "HETATM  100 CAL  CAL     7       0.100   0.200   0.300  1.00  0.00          CA",
    "ATOM    200  P     A C  22  .    1.      2.     3.      .5    .25       RNA  P ",
    "ATOM    201  C1' DA  D  23      -0.5     +2.25   3e0     1     2",
]
HYBRID36_FIELDS = ["    1", "99999", "A0000", "ZZZZZ", "a0000", "zzzzz", "  -12", " 1234", "A000", "zzzz", "9999", "-999", "   0", "a00",
                   "Zz", "9", " ", "12 4", "1e3", "-A00", "    ", "00012", "AAAAA"]


def pdb_golden():
    """(8) PDB atom records, hybrid-36 fields, MultiTrajectory indexing and uniquifyVector through the reference's own
    functions (oracle/_ref/libloospdb.so)."""
    import json

    from oracle.oracle import RefPDB
    R = RefPDB()
    rng = np.random.default_rng(0)
    mt = []
    for _ in range(120):
        sizes = rng.integers(0, 40, int(rng.integers(1, 6))).tolist()
        skip, stride = int(rng.integers(0, 12)), int(rng.integers(1, 7))
        mt.append({"sizes": sizes, "skip": skip, "stride": stride, "locations": R.multitraj(sizes, skip, stride)})
    uq = []
    for _ in range(30):
        v = rng.integers(0, 25, int(rng.integers(0, 40))).tolist()
        uq.append({"in": v, "out": R.uniquify(v)})
    # every (residue name, atom name) pair of a candidate set that contains the selector's own tables and near misses
    resnames = ["A", "A3", "A5", "ALA", "ARG", "ASN", "ASP", "C", "C3", "C5", "CYS", "CYX", "DA", "DC", "DG", "DT", "G", "G3", "G5", "GLN", "GLU",
                "GLY", "HID", "HIE", "HIP", "HIS", "HSD", "HSE", "HSP", "ILE", "LEU", "LYS", "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "U", "U3",
                "U5", "VAL", "RA", "RC", "RG", "RU", "DA3", "DA5", "DC3", "DC5", "DG3", "DG5", "DT3", "DT5", "T", "TIP3", "HOH", "POPC", "ala", "Ala",
                "ACE", "NME", "CYM", "ASH", "GLH", "LYN", "MSE", "SOL", "NA", "CL", "ATP", "HEM", "DU", "N", "UNK", "AL", "ALAX", "PTR", "T3", "T5", "SEP",
                "TPO", "HYP", "GLX"]
    names = ["C", "C1'", "C2'", "C3'", "C4'", "C5'", "CA", "H", "H1'", "H12'", "H15'", "H2'", "H2''", "H22'", "H25'", "H3'", "H4'", "H5'", "H5''",
             "HO2'", "HO3'", "HO5'", "N", "O", "O2'", "O3'", "O4'", "O5'", "OP1", "OP2", "OP3", "OXT", "P", "CB", "HA", "HN", "O1P", "O2P", "C1*", "H5'1",
             "OT1", "OT2", "N1", "N9", "C2", "ca", "Ca", "HB1", "1HB", "OH2", "H1", "H2", "HT1", "CG", "SG", "FE", "HG", "He", "c", "O5*"]
    sel = {"resnames": resnames, "names": names,
           "backbone": [[int(R.backbone(r, n)) for n in names] for r in resnames],
           "hydrogen": [int(R.hydrogen(n)) for n in names],
           "hydrogen_mass": [[m, int(R.hydrogen("H1", m)), int(R.hydrogen("C1", m))] for m in (0.5, 1.008, 1.09, 1.1, 2.014, 4.0, 12.0)]}
    # RangeItem(a), (a, b), (a, b, c): what one range of the frame-range grammar generates (src/index_range_parser.cpp:63-102)
    items = [[5], [0], [0, 9], [3, 49], [7, 7], [10, 0], [49, 49], [0, 20, 2], [0, 49, 3], [10, 0, -1], [10, 2, -3], [49, 10, -2],
             [20, 49, -5], [15, 15, -1], [0, 0, 1], [0, 10, -1], [10, 0, 1], [4, 4, 0], [3, 9, 0], [0, 49, 7], [49, 0, -7], [5, 0, -5],
             [6, 0, -4], [1, 0, -1], [0, 1, 1], [2, 8, 2], [8, 2, -2], [0, 3, 5], [3, 0, -5], [40, 49, 1], [49, 40, -1], [49, 10, -1]]
    ranges = [{"args": a, "values": R.range_item(*a)} for a in items]
    out = {"selectors": sel, "range_items": ranges, "atoms": [{"line": ln, "parsed": R.atom(ln)} for ln in PDB_LINES],
           "hybrid36": [{"field": f, "value": R.hybrid36(f)} for f in HYBRID36_FIELDS], "multitraj": mt, "uniquify": uq}
    with open(os.path.join(HERE, "pdb_ref.json"), "w") as fh:
        json.dump(out, fh, indent=0)


def stats_golden():
    """(9) the statistics lines of the tools from the reference's own functions (oracle/_ref/libloosstats.so)."""
    import json

    from oracle.oracle import RefStats
    R = RefStats()
    rng = np.random.default_rng(0)
    cases = []
    for m, n in [(58, 18), (7, 7), (1, 5), (30, 41), (2, 2), (18, 58)]:
        M = rng.gamma(2.0, 2.5, (m, n)).astype(np.float32)
        rec = {"M": M.tolist(), "whole": R.whole(M), "overlap": R.overlap(M),
               "fractional": {str(c): R.fractional(M, c) for c in (6.5, 0.1, 3.0)}}
        if m == n:
            S = np.tril(M, -1)
            S = S + S.T
            rec["S"] = S.tolist()
            rec["half"] = R.half(S)
        cases.append(rec)
    with open(os.path.join(HERE, "stats_ref.json"), "w") as fh:
        json.dump(cases, fh)


def selection_list():
    """~830 selection strings: every pairing / triple of a set of atomic predicates under && || ! not and parentheses
    (precedence and associativity), comparisons between keywords, literals on either side, string ordering, regular
    expressions, the -> operator, comments, odd spacing, wrong case, and a collection of malformed inputs."""
    import itertools
    atomsx = ["name == 'CA'", "resid < 10010", "index >= 100", "hydrogen", "backbone", "segid == 'SOLV'", "name =~ '^H'", "id <= 50", "all",
              "resname == 'ALA'", "chainid == 'A'"]
    sels = []
    for a, b in itertools.product(atomsx[:7], atomsx[:7]):
        sels += ["%s && %s" % (a, b), "%s || %s" % (a, b), "!%s && %s" % (a, b), "!(%s || %s)" % (a, b), "not %s || not %s" % (a, b)]
    for a, b, c in itertools.product(atomsx[:5], atomsx[3:8], atomsx[6:]):
        sels += ["%s && %s || %s" % (a, b, c), "%s || %s && %s" % (a, b, c), "(%s || %s) && %s" % (a, b, c), "%s && (%s || !%s)" % (a, b, c)]
    sels += ["resid == 10001", "10001 == resid", "resid != 10001", "resid > id", "id < index", "index == index", "name == name", "name != resname",
             "name == resname", "'CA' == name", "name < 'CB'", "name >= 'N'", "resname <= 'B'", "name > \"C\"", "resid+1 == 10002", "resid == +10001",
             "  name  ==  'CA'  ", "name=='CA'", "name=~'C'&&resid<10005", "name =~ 'c.'", "name =~ '^[CN]A?$'", "resname =~ 'A.A'", "segid =~ '^S'",
             "chainid =~ 'a'", "name =~ ''", "name == ''", "segid == ''", "name =~ '\\d'", "name =~ \"'\"", "!!hydrogen", "! ! hydrogen", "not not hydrogen",
             "!not hydrogen", "((((all))))", "(hydrogen)", "hydrogen && all", "all || hydrogen", "ALL", "Hydrogen", "NAME == 'CA'", "name == 'ca'",
             "name -> '(\\d+)' == 12", "name -> '(\\d+)' > 1", "name -> '(\\d+)' >= 2", "name -> '(\\d+)' < 2", "name -> '(\\d+)' <= 2",
             "name -> '(\\d+)' != 2", "resname -> '(\\d)' == 3", "segid -> '(\\d+)' > 0", "resid -> '1' == 1", "name -> 'x'",
             "id == 1 || id == 2 || id == 3", "id == 1 && id == 2", "index < 5 # comment || all", "index < 5 || # comment", "resid == 10001 resid == 10002",
             "resid 10001", "name 'CA'", "== 'CA'", "name == 'CA' &&", "&& name == 'CA'", "()", "( )", "name == (('CA'))", "(resid) < (10005)", "index < 1e3",
             "index < 10.5", "index < -1", "index > 99999999999", "name == 'CA", "name == CA", "name == \"CA'", "resid == 0x10", "resid==10001||resid==10002",
             "hydrogenx", "backbone2", "allx", "name=='CA'#c", "resid <= 10002 and name == 'CA'", "resid <= 10002 or name == 'CA'",
             "name != 'CA' && name ne 'CB'", "segname != 'SOLV'", "index <= 10 || index >= 425", "!(index > 10) && !(index < 5)",
             "name == 5", "name ==", "name === 'CA'", "resid == 'A'", "foo == 1", "(name == 'CA'", "name =~ 3", "'CA' =~ 'C'", "index",
             "backbone && !hydrogen", "!(hydrogen || segid =~ 'SOLV|BULK')", "name =~ \"^(C|O|N|CA)$\"", "resname =~ 'tip'", "chainid == 'W' && name == 'OH2'"]
    return sels


def selection_golden():
    """(10) the selection language through the reference's own scanner, grammar and kernel (oracle/_ref/libloosselect.so):
    for every string of selection_list(), the set of selected atoms of a 432-atom model (as a bit string) or the fact that
    the reference refuses it."""
    import json
    import tempfile

    from loos_b200.synth import synth_system, write_pdb
    from oracle import frontend_oracle as fo
    from oracle.oracle import RefSelect
    atoms, xyz = synth_system(40, 3, seed=7, flavour="rich")
    for a in atoms:
        a["resid"] = a["resid"] + 9990 if a["segid"] == "PROT" else a["resid"]
    atoms[3]["name"] = "H12'"
    atoms[7]["resname"] = "DA"
    atoms[7]["name"] = "C1'"
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "m.pdb")
        write_pdb(p, atoms, xyz[0])
        text = open(p).read()
        A = fo.read_pdb(p)
    R = RefSelect(A)
    out = {"pdb": text, "natoms": len(A), "cases": []}
    for sel in selection_list():
        got, err = R.select(sel)
        if got is None:
            out["cases"].append({"selection": sel, "error": err})
        else:
            bits = np.zeros(len(A), int)
            bits[got] = 1
            out["cases"].append({"selection": sel, "mask": "".join(map(str, bits))})
    with open(os.path.join(HERE, "selection_ref.json"), "w") as fh:
        json.dump(out, fh, indent=0)


if __name__ == "__main__":
    if "--only-selection" in sys.argv:
        selection_golden()
    elif "--only-stats" in sys.argv:
        stats_golden()
    elif "--only-pdb" in sys.argv:
        pdb_golden()
    elif "--only-matrix" in sys.argv:
        matrix_golden()
    elif "--only-dcd" in sys.argv:
        dcd_golden()
    elif "--only-align" in sys.argv:
        align_golden(Checker("reference"))
    elif "--only-xtc" in sys.argv:
        xtc_golden()
    else:
        main()
        xtc_golden()
        dcd_golden()
        matrix_golden()
        pdb_golden()
        stats_golden()
        selection_golden()
