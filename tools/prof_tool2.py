import os, subprocess, sys, tempfile, time
sys.path.insert(0, "/root/repo")
from loos_b200.synth import synth_system, write_dcd, write_pdb
d = tempfile.mkdtemp()
BIN = "/root/repo/loos_b200/bin/rmsds"
for F in (10, 2000, 20000):
    atoms, xyz = synth_system(1000, F, seed=5, flavour="lean")
    write_pdb(os.path.join(d, "m.pdb"), atoms, xyz[0]); write_dcd(os.path.join(d, "t.dcd"), xyz)
    for rep in range(2):
        t0 = time.perf_counter()
        r = subprocess.run([BIN, "-N", "1", os.path.join(d, "m.pdb"), os.path.join(d, "t.dcd")], capture_output=True, text=True)
        print("F=%d: %.3f s %s" % (F, time.perf_counter() - t0, r.stderr.strip()[:60]), flush=True)
t0 = time.perf_counter(); subprocess.run(["nvidia-smi", "-L"], capture_output=True); print("nvidia-smi -L %.3f s" % (time.perf_counter() - t0))
