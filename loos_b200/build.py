"""Build the native pieces of loos_b200 in-tree (no GPU needed: nvcc cross-compiles).

  loos_b200/lib/libloosb200.so   CUDA kernels + C ABI (include/loosb200.h), sm_100a only
  loos_b200/bin/{rmsds,rmsd2ref,multi-rmsds}   native front-end tools (C++), if present

Run as `python -m loos_b200.build` or via __graft_entry__.build().
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
BINDIR = os.path.join(HERE, "bin")
LIB = os.path.join(LIBDIR, "libloosb200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2", "--expt-relaxed-constexpr"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_lib(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    cus = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = cus + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "loosb200.h")]
    if not force and _newer(LIB, deps):
        return LIB
    objs = []
    procs = []
    for cu in cus:
        obj = os.path.join(LIBDIR, os.path.basename(cu)[:-3] + ".o")
        objs.append(obj)
        if not force and _newer(obj, [cu] + deps[len(cus):]):
            continue
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", cu, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.run(cmd, check=True)
    return LIB


def build_tools(force=False):
    src = sorted(glob.glob(os.path.join(HERE, "frontend", "*.cpp")))
    if not src:
        return []
    mk = os.path.join(HERE, "frontend", "Makefile")
    if os.path.exists(mk):
        subprocess.run(["make", "-C", os.path.dirname(mk)] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return sorted(glob.glob(os.path.join(BINDIR, "*")))


def build_all(force=False, verbose=False):
    lib = build_lib(force, verbose)
    tools = build_tools(force)
    return lib, tools


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
