"""bench.py, the part that runs without a GPU: the reference arm (`--impl reference`), i.e. the reference's own CPU
implementation of the path (oracle/_ref when it was compiled here, else the C port) timed on the host cores, and
the keys of its JSON line that the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-sample", "200"] + list(extra), capture_output=True, text=True, env=e, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_line():
    line = json.loads([l for l in _run().splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "frame-pairs/sec all-to-all RMSD" and line["unit"] == "frame-pairs/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["config"]["frames"] == 100000 and line["config"]["atoms"] == 1000 and "first 200 frames" in line["config"]["sample"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_under_torchrun_only_rank_zero_prints():
    assert _run("--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}).strip() == ""
    out = _run("--gpus", "2", env={"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert json.loads(out.strip().splitlines()[-1])["n_gpus"] == 2
