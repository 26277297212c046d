"""Corrupt trajectory files must produce an error (or frames), never a crash: every reader is fed copies of a valid
file with a few bytes overwritten, bits flipped, runs of 0xff planted and tails cut off.  Run as a subprocess by
tests/test_frontend_cpu.py::test_corrupt_files_never_crash (a crash would otherwise take pytest down with it).

    python tests/host/fuzz_readers.py <format> <first seed> <count> <workdir>
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from loos_b200.synth import write_dcd  # noqa: E402
from oracle import frontend_oracle as fo  # noqa: E402


def text_parsers(fe, seed0, count, work):
    """Mutated PDB files against valid and scrambled selection strings, random frame-range specs, mutated
    matrix text: lbfe_select / lbfe_parse_range / lbfe_read_matrix must return, whatever they are given."""
    from loos_b200.synth import synth_system, write_pdb
    fe.lbfe_select.restype = C.c_long
    fe.lbfe_parse_range.restype = C.c_long
    atoms, xyz = synth_system(12, 1, seed=3, flavour="rich")
    base = os.path.join(work, "m.pdb")
    write_pdb(base, atoms, xyz[0])
    raw = open(base, "rb").read()
    out = np.zeros(4096, np.uint32)
    sels = ["name == 'CA'", "backbone && !hydrogen", "resid >= 3 && resid < 9 || name =~ '^C'", "name -> '(\\d+)' == 12",
            "!(hydrogen || segid =~ 'SOLV|BULK')", 'name =~ "^(C|O|N|CA)$"']
    alphabet = list("()!&|=<>~'\"-+ #0123456789") + ["name", "resid", "id", "index", "all", "hydrogen", "backbone", "segid", "not",
                                                     "ne", "=~", "->", "resname", "chainid", "\\", "[", "*", "^$", "(?"]
    matrix = b"# rmsds model traj\n# 3 4 (0)\n0 1.5 2.25 3 \n1.5 0 4e-3 5 \n2.25 0.004 0 6.5 \n"
    mout = np.zeros(1 << 16, np.float32)
    m, n = C.c_long(), C.c_long()
    path = os.path.join(work, "mut.pdb")
    for s in range(seed0, seed0 + count):
        r = np.random.default_rng(s)
        b = bytearray(raw)
        for _ in range(int(r.integers(1, 8))):
            pos = int(r.integers(0, len(b)))
            b[pos] = int(r.integers(0, 256)) if r.integers(0, 2) else ord(" ")
        if r.integers(0, 8) == 0:
            b = b[:int(r.integers(1, len(b)))]
        with open(path, "wb") as fh:
            fh.write(bytes(b))
        sel = sels[int(r.integers(0, len(sels)))]
        if r.integers(0, 2):
            toks = [alphabet[int(k)] for k in r.integers(0, len(alphabet), int(r.integers(1, 12)))]
            sel = " ".join(toks) if r.integers(0, 2) else sel[:int(r.integers(0, len(sel)))] + "".join(toks)
        fe.lbfe_select(path.encode(), sel.encode(), out.ctypes.data_as(C.POINTER(C.c_uint32)), len(out))
        spec = "".join(r.choice(list("0123456789:,- "), int(r.integers(1, 14))))
        fe.lbfe_parse_range(spec.encode(), int(r.integers(1, 100000)), out.ctypes.data_as(C.POINTER(C.c_uint32)), len(out))
        t = bytearray(matrix)
        for _ in range(int(r.integers(1, 6))):
            pos = int(r.integers(0, len(t)))
            ch = int(r.choice(list(b"0123456789 .e-+#()\n\0x"))) if r.integers(0, 2) else int(r.integers(0, 256))
            if r.integers(0, 3) == 0:
                t.insert(pos, ch)
            else:
                t[pos] = ch
        if r.integers(0, 10) == 0:
            t = bytearray(b"# 99999999 99999999 (0)\n1 2 3\n")
        fe.lbfe_read_matrix(bytes(t), len(t), mout.ctypes.data_as(C.POINTER(C.c_float)), len(mout), C.byref(m), C.byref(n))
    print("survived %d mutants: text parsers" % count)


def main():
    fmt, seed0, count, work = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    lib = os.environ.get("LOOSB200_FE_LIB", os.path.join(ROOT, "loos_b200", "lib", "libloosb200fe.so"))
    if fmt == "text":
        return text_parsers(C.CDLL(lib), seed0, count, work)
    rng = np.random.default_rng(5)
    F, N = 9, 123
    xyz = rng.normal(0, 20, (F, N, 3)).astype(np.float32)
    suffix = {"h5": "h5", "h5v2": "h5", "xtc": "xtc", "dcd": "dcd", "nc": "nc"}[fmt]
    base = os.path.join(work, "base." + suffix)
    if fmt == "h5":
        fo.write_mdtraj_h5(base, xyz, chunk=(4, 50, 3))
    elif fmt == "h5v2":
        fo.write_mdtraj_h5(base, xyz, style="v2")
    elif fmt == "xtc":
        fo.write_xtc(base, xyz)
    elif fmt == "dcd":
        write_dcd(base, xyz, with_crystal=True)
    else:
        from scipy.io import netcdf_file
        f = netcdf_file(base, "w", version=1)
        f.Conventions = "AMBER"; f.ConventionVersion = "1.0"
        f.createDimension("frame", None); f.createDimension("spatial", 3); f.createDimension("atom", N)
        t = f.createVariable("time", "f", ("frame",))
        c = f.createVariable("coordinates", "f", ("frame", "atom", "spatial"))
        for k in range(F):
            c[k] = xyz[k]; t[k] = k
        f.close()
    raw = open(base, "rb").read()
    fe = C.CDLL(lib)
    out = np.zeros((3, N), np.float32)
    na, nf = C.c_uint32(), C.c_uint32()
    ok = bad = 0
    path = os.path.join(work, "mutant." + suffix)
    for s in range(seed0, seed0 + count):
        r = np.random.default_rng(s)
        b = bytearray(raw)
        for _ in range(int(r.integers(1, 6))):
            pos, mode = int(r.integers(0, len(b))), int(r.integers(0, 3))
            if mode == 0:
                b[pos] = int(r.integers(0, 256))
            elif mode == 1:
                b[pos] ^= 1 << int(r.integers(0, 8))
            else:
                k = int(r.integers(1, 9))
                b[pos:pos + k] = bytes([255] * k)
        if r.integers(0, 10) == 0:
            b = b[:int(r.integers(1, len(b)))]
        with open(path, "wb") as fh:
            fh.write(bytes(b))
        if fe.lbfe_traj_info(path.encode(), N, C.byref(na), C.byref(nf)) != 0:
            bad += 1
            continue
        good = True
        for k in range(min(nf.value, 20)):
            if fe.lbfe_traj_read(path.encode(), N, k, out.ctypes.data_as(C.POINTER(C.c_float))) != 0:
                good = False
                break
        ok += good
        bad += not good
    print("survived %d mutants: %d still readable, %d refused" % (count, ok, bad))


if __name__ == "__main__":
    main()
