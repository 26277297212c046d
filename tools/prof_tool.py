"""End-to-end timing of the command-line tool on real files (PDB + DCD written here):
rmsds model.pdb traj.dcd with and without the text output."""
import os, subprocess, sys, tempfile, time
import numpy as np
sys.path.insert(0, "/root/repo")
from loos_b200.synth import synth_system, write_dcd, write_pdb
F = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
nres = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
d = tempfile.mkdtemp()
t0 = time.perf_counter()
atoms, xyz = synth_system(nres, F, seed=5, flavour="lean")
write_pdb(os.path.join(d, "m.pdb"), atoms, xyz[0]); write_dcd(os.path.join(d, "t.dcd"), xyz)
print("generated %d frames x %d atoms (%.0f MB DCD) in %.1f s" % (F, xyz.shape[1], os.path.getsize(os.path.join(d, "t.dcd")) / 1e6, time.perf_counter() - t0), flush=True)
BIN = "/root/repo/loos_b200/bin/rmsds"
for args, what in ((["-N", "1"], "no output, stats only"), ([], "matrix text to /dev/null (-p 2)")):
    for rep in range(2):
        t0 = time.perf_counter()
        with open(os.devnull, "w") as devnull:
            r = subprocess.run([BIN] + args + [os.path.join(d, "m.pdb"), os.path.join(d, "t.dcd")], stdout=devnull, stderr=subprocess.PIPE, text=True)
        dt = time.perf_counter() - t0
        print("%-34s %.2f s  rc=%d  %s" % (what, dt, r.returncode, r.stderr.strip()[:80]), flush=True)
