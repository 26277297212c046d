#!/usr/bin/env python
"""Benchmark of the all-to-all optimal-superposition RMSD path (BASELINE.json metric:
frame-pairs/s at 1/2/4/8 B200, fraction of the split-precision tensor-pipe peak,
reference CPU path timed beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c4|c3|c1] [--frames F --atoms N]

One "step" = one pass of the hot path over the workload: stage the frames
(gather, centre, fixed-point split), contract + solve all frame pairs, leave the
float32 matrix where the mode says (HBM for `value`, host for `e2e`).
Under torchrun (N > 1) every rank holds a replica of the frames and computes a
static share of the tile rows of the lower triangle; no collective in the loop.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (frames, atoms, description) -- BASELINE.json configs
    "c1": (1000, 500, "C1 rmsds all-to-all 1,000 frames x 500 atoms"),
    "c3": (20000, 3000, "C3 rmsds all-to-all 20,000 frames x 3,000 atoms"),
    "c4": (100000, 1000, "C4 rmsds all-to-all 100,000 frames x 1,000 atoms"),
    # multi-rmsds of 8 trajectories x 25,000 frames = 200,000 composite frames; the 160 GB result only
    # fits as packed row blocks over >= 2 GPUs (run with --gpus 8 --no-e2e)
    "c5": (200000, 5000, "C5 multi-rmsds all-to-all 8 x 25,000 frames x 5,000 atoms"),
}
METRIC = "frame-pairs/sec all-to-all RMSD"
UNIT = "frame-pairs/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    # fallback stated in B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region, read through NVML in a
    background thread (same counters as `nvidia-smi --query-gpu=clocks.sm,
    clocks_event_reasons.*`, without forking nvidia-smi, which stalls the driver for
    tens of ms per query).  An NVML query that overlaps host-side CUDA calls delays them by
    ~30 ms (measured: free-running 250 ms sampling cost 79 ms per 430 ms step), so samples are
    taken on demand: `kick()` at the start of a step schedules one sample `delay` seconds later,
    i.e. while the step's one long kernel runs and the host is blocked on it anyway."""

    def __init__(self, gpu_index, delay=float(os.environ.get("LOOSB200_CLOCK_DELAY", "0.12"))):
        super().__init__(daemon=True)
        self.idx, self.delay, self.rows, self._stop_ev = gpu_index, delay, [], threading.Event()
        self._go = threading.Event()
        self._ready = threading.Event()
        self.max_mhz = None
        self.query_ms = []

    def kick(self):
        self._go.set()

    def wait_ready(self, timeout=5.0):
        self._ready.wait(timeout)

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.idx]) if vis and vis.split(",")[0].isdigit() else self.idx
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self._ready.set()
            while True:
                self._go.wait()          # blocking: a polling wait costs the main thread GIL hand-offs
                self._go.clear()
                if self._stop_ev.is_set():
                    break
                time.sleep(self.delay)
                t_q = time.perf_counter()
                clk = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((clk, rs))
                self.query_ms.append(1e3 * (time.perf_counter() - t_q))
        except Exception as e:  # pragma: no cover
            self.err = repr(e)
            self._ready.set()

    def stop(self):
        self._stop_ev.set()
        self._go.set()
        self.join(timeout=2.0)
        try:
            import pynvml as nv
            names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        except Exception:
            names = {}
        sm = sorted(c for c, _ in self.rows)
        reasons = sorted(n for n, bit in names.items() if any(r & bit for _, r in self.rows))
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "nvml_query_ms": [round(q, 1) for q in self.query_ms],
                "how": "NVML, one sample %.0f ms into every timed step (while its kernel runs)" % (self.delay * 1e3)}


def dist_setup(ngpus):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world, device):
    import torch
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, world, device):
    import torch
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def reference_checker():
    from oracle.oracle import Checker, have_reference
    kind = "reference" if have_reference() else "port"
    return Checker(kind), kind


def cpu_pairs_per_s(X_sample, threads):
    """The reference's CPU path (oracle/_ref: its own centeredRMSD + our restatement
    of its row-dispatch thread pool) on a bounded sample of the workload."""
    chk, kind = reference_checker()
    S = len(X_sample)
    t0 = time.perf_counter()
    chk.rmsds_self(X_sample.astype(np.float64), nthreads=threads)
    dt = time.perf_counter() - t0
    return (S * (S - 1) // 2) / dt, dt, kind


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the
    host cores, same metric/config, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    F, N, desc = WORKLOADS[args.workload]
    F = args.frames or F
    N = args.atoms or N
    threads = os.cpu_count() or 1
    S = min(F, args.ref_sample)
    from loos_b200.synth import SEED0, synth_frames
    X = synth_frames(S, N, seed=SEED0 + 4)
    times = []
    kind = None
    for it in range(args.warmup + args.steps):
        pps, dt, kind = cpu_pairs_per_s(X, threads)
        if it >= args.warmup:
            times.append(dt)
    pairs = S * (S - 1) // 2
    value = pairs * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": desc, "frames": F, "atoms": N,
                   "sample": "first %d frames (%d pairs) per step" % (S, pairs)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": "first %d frames of the workload, %d pairs per step" % (S, pairs)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "atom_pairs_per_s": value * N,
    }
    print(json.dumps(line))
    return 0


def parity_sample(X_dev, M_get, F, N, owned=None, nrand=3000, nnear=1000):
    """max / mean |dRMSD| of the GPU result against the CPU checker on random and
    near-diagonal pairs (i, j < i) of the rows in `owned` (default: all rows); the
    full-size matrix cannot be recomputed on the CPU."""
    import torch
    chk, kind = reference_checker()
    rng = np.random.default_rng(7)
    rows_ok = np.arange(9, F) if owned is None else np.array(sorted(f for f in owned if f >= 9))
    ii = rng.choice(rows_ok, nrand)
    jj = (rng.random(nrand) * ii).astype(np.int64)
    i2 = rng.choice(rows_ok, nnear)
    j2 = i2 - rng.integers(1, 9, nnear)
    ii, jj = np.concatenate([ii, i2]), np.concatenate([jj, j2])
    frames = np.unique(np.concatenate([ii, jj]))
    sub = X_dev[torch.as_tensor(frames, device=X_dev.device)].cpu().numpy().astype(np.float64)
    pos = {f: k for k, f in enumerate(frames)}
    cen = np.stack([chk.center_at_origin(s)[0] for s in sub])
    errs = []
    for i, j in zip(ii, jj):
        ref = chk.centered_rmsd(cen[pos[i]], cen[pos[j]])
        errs.append(abs(float(M_get(int(i), int(j))) - float(np.float32(ref))))
    errs = np.array(errs)
    return {"pairs": int(errs.size), "max_abs_err": float(errs.max()), "mean_abs_err": float(errs.mean()),
            "tolerance": 1e-4, "checker": kind}


NCU_TRAFFIC = {("c4", 1, "tensor"): 107.047827e9 + 40.017349e9}
NCU_TRAFFIC_SOURCE = "profiles/r1_k_pairs_tensor_F100000_N1000_final_metrics.csv (result matrix 40.0 GB written once, operands 0.92 GB read 114x)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--atoms", type=int, default=0)
    ap.add_argument("--engine", default="tensor", choices=["tensor", "fp64", "auto"])
    ap.add_argument("--ref-sample", type=int, default=3000, help="frames per step for --impl reference")
    ap.add_argument("--cpu-sample", type=int, default=4000, help="frames for the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch

    import loos_b200 as lb
    from loos_b200 import capi
    from loos_b200.api import rmsds_part
    from loos_b200.synth import SEED0, synth_frames_torch

    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    assert capi.device_count() > local, "no sm_100 device: bench.py has no CPU path"
    F, N, desc = WORKLOADS[args.workload]
    F = args.frames or F
    N = args.atoms or N
    pairs_total = F * (F - 1) // 2
    peaks, peaks_kind = load_peaks()
    # the result stays in HBM (F x F float32 on one GPU, packed row blocks otherwise): refuse what cannot fit
    out_bytes = 4 * F * F if world == 1 else 4 * F * (F // world + 80)
    if out_bytes + 12 * F * N + 9 * F * (N + 128) > 170e9:
        raise SystemExit("workload %s needs %.0f GB of HBM per GPU at --gpus %d; use more GPUs" % (args.workload, out_bytes / 1e9, world))

    X = synth_frames_torch(F, N, seed=SEED0 + 4, device=dev)   # same frames on every rank
    torch.cuda.synchronize()

    # ---------------- device-resident leg (`value`) ----------------
    if world == 1:
        rows = None
        out_dev = torch.empty((F, F), dtype=torch.float32, device=dev)
    else:
        rows = lb.part_rows(F, rank, world)
        out_dev = torch.empty((len(rows), F), dtype=torch.float32, device=dev)

    def step_device():
        t_a = time.perf_counter()
        fs = lb.FrameSet(X)
        t_b = time.perf_counter()
        if world == 1:
            _, st = lb.rmsds(fs, out=out_dev, engine=args.engine)
        else:
            _, _, st = rmsds_part(fs, rank, world, out=out_dev, engine=args.engine)
        t = fs.last_timing()
        t_c = time.perf_counter()
        fs.close()
        if os.environ.get("LOOSB200_BENCH_TRACE"):
            sys.stderr.write("  stage %.1f ms, rmsds %.1f ms, close %.1f ms\n" % (1e3 * (t_b - t_a), 1e3 * (t_c - t_b), 1e3 * (time.perf_counter() - t_c)))
        return st, t

    t_est = 0.4
    for _ in range(args.warmup):
        torch.cuda.synchronize()
        t_w = time.perf_counter()
        step_device()
        t_est = time.perf_counter() - t_w
    # one clock sample a third of the way into every timed step
    sampler = ClockSampler(local, delay=min(0.12, t_est / 3)) if rank == 0 and not os.environ.get("LOOSB200_NO_SAMPLER") else None
    if sampler:
        sampler.start()
        sampler.wait_ready()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    kern_ms, launches, st = 0.0, 0, None
    for _ in range(args.steps):
        if sampler:
            sampler.kick()
        t_s = time.perf_counter()
        st, (ms_main, ms_other, nl) = step_device()
        if os.environ.get("LOOSB200_BENCH_TRACE"):
            sys.stderr.write("step wall %.1f ms, kernel %.1f ms, other kernels %.1f ms\n" % (1e3 * (time.perf_counter() - t_s), ms_main, ms_other))
        kern_ms += ms_main
        launches += nl + 3          # + stage, unit, quantise kernels of the FrameSet
    e1.record()
    barrier(world)
    ms_total = max_over_ranks(e0.elapsed_time(e1), world, dev)
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms_total / args.steps
    value = pairs_total / (ms_per_step * 1e-3)
    my_pairs = st["count"]
    kern_ms_step = kern_ms / args.steps
    # roofline of the dominant kernel (k_pairs_tensor): algorithmic flops = 18 N per pair
    # (SURVEY.md 8d), this rank's pairs per launch, duration from CUDA events on the
    # launching stream (loosb200_last_timing).
    achieved = 18.0 * N * my_pairs / (kern_ms_step * 1e-3) / 1e12
    achieved = max_over_ranks(-achieved, world, dev) * -1.0   # slowest rank
    bf16 = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    if args.engine == "fp64":
        peak, basis = 37.0, "B200 fp64 pipe (nominal, not in MEASURED_PEAKS.json)"
    else:
        peak = 2.0 * bf16 / 9.0
        basis = ("int8 dense = 2 x %s bf16 sustained %.1f TF (MEASURED_PEAKS.json: %s), / 9 digit products"
                 % (peaks_kind, bf16, peaks_kind))

    parity = None
    if rank == 0:
        if world == 1:
            parity = parity_sample(X, lambda i, j: out_dev[j, i].item(), F, N)
        else:
            rowpos = {int(f): r for r, f in enumerate(rows) if f != 0xFFFFFFFF}
            parity = parity_sample(X, lambda i, j: out_dev[rowpos[i], j].item(), F, N, owned=list(rowpos))
    del out_dev
    torch.cuda.empty_cache()

    # ---------------- end-to-end leg (`e2e`): host buffers through the C ABI ----------------
    e2e = None
    if not args.no_e2e:
        X_host = torch.empty((F, N, 3), dtype=torch.float32, pin_memory=True)
        X_host.copy_(X)
        torch.cuda.synchronize()
        Xh = X_host.numpy()
        if world == 1:
            out_t = torch.empty((F, F), dtype=torch.float32, pin_memory=True)
            out_h = out_t.numpy().T          # Fortran-ordered view == column-major RealMatrix
            d2h = F * F * 4
        else:
            out_t = torch.empty((len(rows), F), dtype=torch.float32, pin_memory=True)
            out_h = out_t.numpy()
            d2h = len(rows) * F * 4
        h2d = F * N * 12

        def step_e2e():
            fs = lb.FrameSet(Xh, device=local)
            if world == 1:
                lb.rmsds(fs, out=out_h, engine=args.engine)
            else:
                rmsds_part(fs, rank, world, out=out_h, engine=args.engine)
            fs.close()

        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        barrier(world)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        e1.record()
        barrier(world)
        ms_e2e = max_over_ranks(e0.elapsed_time(e1), world, dev) / args.steps
        e2e = {"value": pairs_total / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(sum_over_ranks(h2d, world, dev)),
               "d2h_bytes_per_step": int(sum_over_ranks(d2h, world, dev)),
               "note": "host pinned coords -> loosb200_stage_frames -> loosb200_rmsds_self%s -> host pinned float32 %s"
                       % ("" if world == 1 else "_part", "matrix" if world == 1 else "row blocks")}
        del out_t, X_host

    # ---------------- CPU baseline (rank 0, N == 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        S = min(F, args.cpu_sample)
        threads = os.cpu_count() or 1
        pps, dt, kind = cpu_pairs_per_s(X[:S].cpu().numpy(), threads)
        cpu = {"value": pps, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": "first %d frames of the workload (%d pairs), %.1f s" % (S, S * (S - 1) // 2, dt)}

    # dram__bytes_read.sum + dram__bytes_write.sum of ONE k_pairs_tensor launch from the committed
    # ncu --set full capture (only that workload / engine / 1 GPU has a capture)
    traffic = NCU_TRAFFIC.get((args.workload, world, args.engine)) if not (args.frames or args.atoms) else None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "s8 digits x3 / s32 accumulate (exact) + f64 solve" if args.engine != "fp64" else "f64",
            "data": "synthetic",
            "config": {"workload": desc, "frames": F, "atoms": N, "pairs": pairs_total,
                       "parallelism": "tile-row split x%d, replicated frames, no collective" % world,
                       "l2": "inputs larger than L2 (operands %.0f MB, result %.1f GB)"
                             % (F * N * 9 / 1e6, F * F * 4 / 1e9),
                       "engine": args.engine,
                       "output": "symmetric FxF float32 in HBM" if world == 1 else "lower-triangle row blocks in HBM"},
            "atom_pairs_per_s": value * N,
            "e2e": e2e,
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch",
                         "traffic_source": NCU_TRAFFIC_SOURCE if traffic else None,
                         "kernel": "k_pairs_tensor" if args.engine != "fp64" else "k_pairs_fp64",
                         "kernel_ms_per_launch": kern_ms_step, "peak_basis": basis,
                         "algorithmic": "18*N flop per frame pair x pairs per launch",
                         "shape_ceiling_frac": 0.947,
                         "shape_ceiling_source": "profiles/r1_ubench_mma_n.txt: M=128 kind::i8 MMAs read both operands from "
                                                 "shared memory at 128 B/clk; per K step 3 x N=192 (96 clk) + 3 x N=96 (56 clk) "
                                                 "= 456 clk for 432 clk of MACs"},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "parity": parity,
            "stats": {"max_rmsd": st["max"], "avg_rmsd_this_rank": st["avg"]},
        }
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
