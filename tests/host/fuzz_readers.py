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


def main():
    fmt, seed0, count, work = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    lib = os.environ.get("LOOSB200_FE_LIB", os.path.join(ROOT, "loos_b200", "lib", "libloosb200fe.so"))
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
