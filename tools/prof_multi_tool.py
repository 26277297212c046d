"""multi-rmsds / rmsds on several GPUs from real files: a reduced C5 (8 trajectories x F frames x N atoms)
through the command-line tools with --gpus 1 and --gpus G: stats only (-N 1) and matrix text to /dev/null."""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, "/root/repo")
from loos_b200.synth import synth_system, write_dcd, write_pdb
F = int(sys.argv[1]) if len(sys.argv) > 1 else 2500      # frames per trajectory
nres = int(sys.argv[2]) if len(sys.argv) > 2 else 5000   # CA atoms
G = int(sys.argv[3]) if len(sys.argv) > 3 else 2
d = tempfile.mkdtemp()
t0 = time.perf_counter()
atoms, xyz = synth_system(nres, 8 * F, seed=5, flavour="lean")
write_pdb(os.path.join(d, "m.pdb"), atoms, xyz[0])
trajs = []
for k in range(8):
    trajs.append(os.path.join(d, "t%d.dcd" % k))
    write_dcd(trajs[-1], xyz[k * F:(k + 1) * F])
print("generated 8 x %d frames x %d atoms (%d selected) in %.1f s" % (F, xyz.shape[1], nres, time.perf_counter() - t0), flush=True)
BIN = "/root/repo/loos_b200/bin/multi-rmsds"
outs = {}
for g in (1, G):
    for args, what in ((["-N", "1"], "stats only"), ([], "matrix text to a pipe (-p 2)")):
        t0 = time.perf_counter()
        cmd = [BIN, "--gpus", str(g)] + args + [os.path.join(d, "m.pdb")] + trajs
        if args:
            r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        else:
            r = subprocess.run(" ".join(cmd) + " | tail -n +2 | md5sum", shell=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            outs[g] = r.stdout.split()[0] if r.stdout else None
        print("--gpus %d %-30s %.2f s  rc=%d  %s" % (g, what, time.perf_counter() - t0, r.returncode, r.stderr.strip()[-90:]), flush=True)
print("matrix text identical between --gpus 1 and --gpus %d (header line aside): see md5" % G, outs)
