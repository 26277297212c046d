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


NCU_TRAFFIC = {("c4", 1, "tensor"): 27.789879e9 + 40.010743e9}
NCU_TRAFFIC_SOURCE = ("profiles/r2_k_pairs_tensor_C4_final_full_metrics.csv, ncu --set full of one launch (result matrix 40.0 GB written once, "
                      "operands 0.9 GB read 30x: one pass over the columns per 48-row-tile group; 145 GB in round 1)")


class Ctx:
    pass


def host_barrier(ctx):
    """Barrier that keeps the GPUs idle (a NCCL barrier is a kernel that spins on every rank's device)."""
    import torch
    torch.cuda.synchronize()
    if ctx.world > 1:
        import torch.distributed as dist
        dist.barrier(group=ctx.gloo)


def timed_region(ctx, step, steps, sampler_delay):
    """W warm-up steps are the caller's; this times exactly `steps` steps between barriers with CUDA
    events, max over ranks, one NVML clock sample per step (or per ~0.25 s for short steps)."""
    import torch
    sampler = None
    if ctx.rank == 0 and not os.environ.get("LOOSB200_NO_SAMPLER"):
        sampler = ClockSampler(ctx.local, delay=sampler_delay)
        sampler.start()
        sampler.wait_ready()
    barrier(ctx.world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    outs = []
    t_kick = 0.0
    for _ in range(steps):
        if sampler and time.perf_counter() - t_kick > 0.25:
            sampler.kick()
            t_kick = time.perf_counter()
        outs.append(step())
    e1.record()
    barrier(ctx.world)
    ms_total = max_over_ranks(e0.elapsed_time(e1), ctx.world, ctx.dev)
    clocks = sampler.stop() if sampler else None
    return ms_total / steps, outs, clocks


def measure_int8_peak(ctx, seconds):
    """Sustained dense int8 tcgen05 rate of this GPU, measured in this run (loosb200_measure_int8_peak)."""
    import ctypes as C

    from loos_b200 import capi
    tops, ms = C.c_double(), C.c_double()
    sampler = None
    if not os.environ.get("LOOSB200_NO_SAMPLER"):
        sampler = ClockSampler(ctx.local, delay=0.15 + 0.6 * seconds)   # calibration launches first, then the long one
        sampler.start()
        sampler.wait_ready()
        sampler.kick()
    capi.check(capi.lib().loosb200_measure_int8_peak(ctx.local, float(seconds), C.byref(tops), C.byref(ms)))
    clocks = sampler.stop() if sampler else None
    return {"int8_tops": tops.value, "launch_ms": ms.value, "clocks": clocks,
            "how": "tcgen05.mma.cta_group::2.kind::i8 M=256 N=256 K=32 back to back on every SM pair, operands resident "
                   "in shared memory, one launch of %.1f s (loos_b200/csrc/peak.cu)" % (ms.value / 1e3)}


def all_to_all(ctx, args, name, F, N, steps, warmup, want_e2e, want_cpu):
    """One all-to-all workload: device-resident `value`, roofline, sampled parity, optional e2e and CPU legs."""
    import torch

    import loos_b200 as lb
    from loos_b200 import capi
    from loos_b200.api import rmsds_part
    from loos_b200.synth import SEED0, synth_frames_torch

    rank, world, local, dev = ctx.rank, ctx.world, ctx.local, ctx.dev
    desc = WORKLOADS[name][2]
    pairs_total = F * (F - 1) // 2
    # the result stays in HBM (F x F float32 on one GPU, packed row blocks otherwise): refuse what cannot fit
    out_bytes = 4 * F * F if world == 1 else 4 * F * (F // world + 80)
    if out_bytes + 12 * F * N + 9 * F * (N + 128) > 170e9:
        raise SystemExit("workload %s needs %.0f GB of HBM per GPU at --gpus %d; use more GPUs" % (name, out_bytes / 1e9, world))

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
        fs = lb.FrameSet(X)
        if world == 1:
            _, st = lb.rmsds(fs, out=out_dev, engine=args.engine)
        else:
            _, _, st = rmsds_part(fs, rank, world, out=out_dev, engine=args.engine)
        t = fs.last_timing()
        fs.close()
        return st, t

    t_est = 0.4
    for _ in range(warmup):
        torch.cuda.synchronize()
        t_w = time.perf_counter()
        step_device()
        t_est = time.perf_counter() - t_w
    ms_per_step, outs, clocks = timed_region(ctx, step_device, steps, min(0.12, t_est / 3))
    st = outs[-1][0]
    kern_ms_step = sum(o[1][0] for o in outs) / steps
    launches = sum(o[1][2] + 3 for o in outs)          # + stage, unit, quantise kernels of the FrameSet
    value = pairs_total / (ms_per_step * 1e-3)
    my_pairs = st["count"]
    # roofline of the dominant kernel: algorithmic flops = 18 N per pair, this rank's
    # pairs per launch, duration from CUDA events on the launching stream (loosb200_last_timing)
    achieved = 18.0 * N * my_pairs / (kern_ms_step * 1e-3) / 1e12
    achieved = max_over_ranks(-achieved, world, dev) * -1.0   # slowest rank
    bf16 = ctx.peaks.get("bf16_tflops_sustained", ctx.peaks["bf16_tflops"])
    proxy = 2.0 * bf16 / 9.0
    if args.engine == "fp64":
        peak, basis = 37.0, "B200 fp64 pipe (nominal, not in MEASURED_PEAKS.json)"
    elif ctx.int8_peak:
        peak = ctx.int8_peak["int8_tops"] / 9.0
        basis = ("int8 dense tcgen05 rate measured in this run, %.0f TOP/s sustained (SM %s MHz), / 9 digit products"
                 % (ctx.int8_peak["int8_tops"], (ctx.int8_peak.get("clocks") or {}).get("sm_mhz")))
    else:
        peak, basis = proxy, "int8 dense = 2 x bf16 sustained %.1f TF (MEASURED_PEAKS.json: %s), / 9 digit products" % (bf16, ctx.peaks_kind)

    parity = None
    if rank == 0:
        if world == 1:
            parity = parity_sample(X, lambda i, j: out_dev[j, i].item(), F, N)
        else:
            rowpos = {int(f): r for r, f in enumerate(rows) if f != 0xFFFFFFFF}
            parity = parity_sample(X, lambda i, j: out_dev[rowpos[i], j].item(), F, N, owned=list(rowpos))
    del out_dev
    torch.cuda.empty_cache()
    capi.lib().loosb200_trim()

    # ---------------- end-to-end leg (`e2e`): host buffers through the C ABI ----------------
    # One process drives all N devices (loosb200_rmsds_self_multi), as the reference fills one matrix
    # from np threads of one process: host pinned coordinates in, the full symmetric F x F float32
    # matrix in host memory out.  Under torchrun rank 0 makes the call; the other ranks keep their GPUs idle.
    e2e = None
    if want_e2e:
        X_first = X[: min(F, 4000)].cpu().numpy() if want_cpu else None
        X_host = None
        if rank == 0:
            X_host = torch.empty((F, N, 3), dtype=torch.float32, pin_memory=True)
            X_host.copy_(X)
            torch.cuda.synchronize()
        del X
        torch.cuda.empty_cache()
        host_barrier(ctx)
        if rank == 0:
            Xh = X_host.numpy()
            out = lb.PinnedMatrix((F, F))
            devices = list(range(world))

            def step_e2e(dst):
                t0 = time.perf_counter()
                reps = lb.stage_multi(Xh, devices)
                t1 = time.perf_counter()
                lb.rmsds_multi(reps, out=dst, engine=args.engine, want_stats=False)
                t2 = time.perf_counter()
                for r in reps:
                    r.close()
                return t1 - t0, t2 - t1, time.perf_counter() - t0

            for _ in range(max(1, min(warmup, 2))):
                step_e2e(out.array)
            torch.cuda.synchronize()
            ts = [step_e2e(out.array) for _ in range(steps)]
            ms_e2e = 1e3 * sum(t[2] for t in ts) / steps
            ms_stage = 1e3 * sum(t[0] for t in ts) / steps
            ms_call = 1e3 * sum(t[1] for t in ts) / steps
            # sampled check of what arrived on the host: symmetric, zero diagonal, equal to the device-resident result
            rng = np.random.default_rng(3)
            ii, jj = rng.integers(0, F, 2000), rng.integers(0, F, 2000)
            M = out.array
            sym_ok = bool(all(M[i, j] == M[j, i] for i, j in zip(ii, jj)) and all(M[i, i] == 0.0 for i in ii))
            # the same job delivered to rank 0's HBM instead (gather of the row blocks over NVLink)
            hbm = None
            if world > 1 or os.environ.get("LOOSB200_E2E_HBM"):
                out.free()
                out_d0 = torch.empty((F, F), dtype=torch.float32, device="cuda:0")
                step_e2e(out_d0)
                torch.cuda.synchronize()
                th = [step_e2e(out_d0) for _ in range(steps)]
                hbm = {"ms_per_step": 1e3 * sum(t[2] for t in th) / steps, "call_ms": 1e3 * sum(t[1] for t in th) / steps,
                       "value": pairs_total / (sum(t[2] for t in th) / steps), "unit": UNIT,
                       "gather": os.environ.get("LOOSB200_GATHER", "store"),
                       "note": "same call with the F x F matrix gathered in rank 0's HBM over NVLink (loosb200_rmsds_self_multi_device)"}
                del out_d0
            else:
                out.free()
            e2e = {"value": pairs_total / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": F * N * 12, "d2h_bytes_per_step": F * F * 4,
                   "p2p_bytes_per_step": (world - 1) * F * (N + 3) * 12,
                   "stage_ms": ms_stage, "call_ms": ms_call,
                   "d2h_gbs_over_call": F * F * 4 / (ms_call * 1e-3) / 1e9,
                   "delivered": "full symmetric F x F float32 matrix in ONE pinned host buffer" + ("" if sym_ok else " (SYMMETRY CHECK FAILED)"),
                   "to_rank0_hbm": hbm,
                   "note": ("host pinned coords -> loosb200_stage_frames_multi (each of the %d GPUs uploads its slice over its own PCIe link, "
                            "slices exchanged over NVLink) -> loosb200_rmsds_self_multi: each GPU copies its row strips + mirror strips "
                            "into the host matrix over its own PCIe link, shares sized by the links' measured rates" % world) if world > 1 else
                           "host pinned coords -> loosb200_stage_frames_multi -> loosb200_rmsds_self_multi on one device: kernels fill an "
                           "F x F bounce matrix in HBM, 64 bands ship their finished columns as contiguous blocks behind the kernels"}
            del X_host
            torch.cuda.empty_cache()
            capi.lib().loosb200_trim()
        host_barrier(ctx)
    else:
        X_first = X[: min(F, 4000)].cpu().numpy() if want_cpu else None
        del X
        torch.cuda.empty_cache()

    # ---------------- CPU baseline (rank 0, N == 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and want_cpu:
        S = min(F, args.cpu_sample)
        threads = os.cpu_count() or 1
        pps, dt, kind = cpu_pairs_per_s(X_first[:S], threads)
        cpu = {"value": pps, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": "first %d frames of the workload (%d pairs), %.1f s" % (S, S * (S - 1) // 2, dt)}

    # dram__bytes_read.sum + dram__bytes_write.sum of ONE k_pairs_tensor launch from the committed
    # ncu --set full capture (only that workload / engine / 1 GPU has a capture)
    traffic = NCU_TRAFFIC.get((name, world, args.engine)) if (F, N) == WORKLOADS[name][:2] else None
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
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
                     "peak_proxy_2x_bf16_sustained_over_9": proxy, "frac_of_proxy": achieved / proxy,
                     "algorithmic": "18*N flop per frame pair x pairs per launch"},
        "cpu_baseline": cpu,
        "clocks": clocks,
        "parity": parity,
        "stats": {"max_rmsd": st["max"], "avg_rmsd_this_rank": st["avg"]},
    }


def stream_c2(ctx, args, F, N, steps, warmup, want_e2e):
    """C2: rmsd2ref --target, F frames x N atoms against one reference (Tools/rmsd2ref.cpp:205-216,238-245).
    HBM-bound: 12*N bytes read + 8 written per frame."""
    import torch

    import loos_b200 as lb
    from loos_b200 import capi
    from loos_b200.synth import SEED0, synth_frames_torch
    X = synth_frames_torch(F, N, seed=SEED0 + 2, device=ctx.dev, chunk=4096)
    tgt = X[F // 2].double().cpu().numpy()
    fs = lb.FrameSet(X)
    res = lb.PinnedMatrix((F,), np.float64, order="C")   # the per-frame series lands in pinned host memory

    def step():
        d = lb.rmsd2ref(fs, tgt, out=res.array)
        return d, fs.last_timing()

    for _ in range(warmup):
        step()
    ms_per_step, outs, clocks = timed_region(ctx, step, steps, 0.1)
    kern_ms = sum(o[1][0] for o in outs) / steps
    launches = sum(o[1][2] for o in outs)
    d = outs[-1][0]
    bytes_ = F * (12.0 * N + 8)
    # parity on sampled frames against the reference's superposition + rmsd
    chk, kind = reference_checker()
    rng = np.random.default_rng(11)
    idx = np.unique(np.concatenate([rng.integers(0, F, 300), [0, F // 2, F - 1]]))
    sub = X[torch.as_tensor(idx, device=X.device)].cpu().numpy().astype(np.float64)
    want, _ = chk.rmsd2ref_target(sub, sub, tgt, tgt)
    err = np.abs(d[idx] - want)
    parity = {"frames": int(idx.size), "max_abs_err": float(err.max()), "mean_abs_err": float(err.mean()), "tolerance": 1e-4, "checker": kind}
    fs.close()
    del d, outs
    res.free()
    e2e = None
    if want_e2e:
        Xh_t = torch.empty((F, N, 3), dtype=torch.float32, pin_memory=True)
        Xh_t.copy_(X)
        torch.cuda.synchronize()
        del X
        torch.cuda.empty_cache()
        xh = Xh_t.numpy()

        def step_e2e():
            t0 = time.perf_counter()
            f2 = lb.FrameSet(xh, device=ctx.local)
            lb.rmsd2ref(f2, tgt)
            f2.close()
            return time.perf_counter() - t0
        step_e2e()
        ts = [step_e2e() for _ in range(2)]
        e2e = {"value": F / (sum(ts) / len(ts)), "unit": "frames/s", "ms_per_step": 1e3 * sum(ts) / len(ts),
               "h2d_bytes_per_step": F * N * 12, "d2h_bytes_per_step": F * 8,
               "h2d_gbs": F * N * 12 / (sum(ts) / len(ts)) / 1e9,
               "note": "host pinned frames -> loosb200_stage_frames -> loosb200_rmsd2ref -> host doubles; PCIe H2D bound"}
        del Xh_t
    else:
        del X
    torch.cuda.empty_cache()
    capi.lib().loosb200_trim()
    hbm = ctx.peaks["hbm_gbs"]
    ach = bytes_ / (kern_ms * 1e-3) / 1e9
    return {"metric": "frames/sec rmsd2ref single-reference streaming", "value": F / (ms_per_step * 1e-3), "unit": "frames/s",
            "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "dtype": "f32 coordinates, f64 covariance + solve", "data": "synthetic",
            "config": {"workload": "C2 rmsd2ref: %d frames x %d atoms against a single reference (same align and rmsd selection)" % (F, N),
                       "frames": F, "atoms": N, "l2": "inputs larger than L2 (%.1f GB of frames)" % (F * N * 12 / 1e9)},
            "atom_pairs_per_s": F * N / (ms_per_step * 1e-3),
            "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None,
                         "kernel": "k_cov (+ k_solve)", "kernel_ms_per_launch": kern_ms,
                         "peak_basis": "MEASURED_PEAKS.json hbm_gbs (%s)" % ctx.peaks_kind,
                         "algorithmic": "12*N + 8 bytes per frame x frames per launch"},
            "clocks": clocks, "parity": parity}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS) + ["c2"])
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--atoms", type=int, default=0)
    ap.add_argument("--engine", default="tensor", choices=["tensor", "fp64", "auto"])
    ap.add_argument("--ref-sample", type=int, default=3000, help="frames per step for --impl reference")
    ap.add_argument("--cpu-sample", type=int, default=4000, help="frames for the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="only the headline workload (no C2 / C3 / C5 sub-records)")
    ap.add_argument("--also-scale", type=float, default=1.0, help="shrink the frame counts of the sub-records (testing)")
    ap.add_argument("--peak-seconds", type=float, default=2.5, help="duration of the int8 peak micro-benchmark (0: skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.workload == "c2":
            raise SystemExit("--impl reference times the all-to-all path")
        return run_reference(args)

    import torch

    from loos_b200 import capi

    ctx = Ctx()
    ctx.rank, ctx.world, ctx.local = dist_setup(args.gpus)
    ctx.dev = torch.device("cuda", ctx.local)
    ctx.gloo = None
    if ctx.world > 1:
        import torch.distributed as dist
        ctx.gloo = dist.new_group(backend="gloo")
    assert capi.device_count() > ctx.local, "no sm_100 device: bench.py has no CPU path"
    ctx.peaks, ctx.peaks_kind = load_peaks()
    ctx.int8_peak = None
    if args.peak_seconds > 0 and args.engine != "fp64":
        # every rank measures its own GPU at the same time (the boards share the box's power budget
        # in the multi-GPU runs exactly like this); rank 0's figure is the denominator
        ctx.int8_peak = measure_int8_peak(ctx, args.peak_seconds)
        barrier(ctx.world)

    if args.workload == "c2":
        F, N = args.frames or 1000000, args.atoms or 2000
        line = stream_c2(ctx, args, F, N, max(args.steps, 100), args.warmup, not args.no_e2e)
        if ctx.rank == 0:
            print(json.dumps(line))
        return 0

    F, N, _ = WORKLOADS[args.workload]
    F, N = args.frames or F, args.atoms or N
    # the other BASELINE.json configurations that fit this launch, as sub-records with their own clocks and parity.
    # C3 runs BEFORE the headline workload: after that one's end-to-end and CPU legs (40 GB of pinned host memory
    # given back, 16 host threads just finished) the host side of a 40 ms step was seen to cost 8 ms instead of 1.5.
    also = {}
    want_also = not args.no_also and args.workload == "c4" and not (args.frames or args.atoms)
    sc = args.also_scale
    if want_also and ctx.world == 1:
        F3, N3, _ = WORKLOADS["c3"]
        also["c3"] = all_to_all(ctx, args, "c3", max(2000, int(F3 * sc)), N3, max(args.steps, 10), max(args.warmup, 5), False, False)
    line = all_to_all(ctx, args, args.workload, F, N, args.steps, args.warmup, not args.no_e2e, not args.no_cpu)
    line["int8_peak"] = ctx.int8_peak
    if want_also:
        if ctx.world == 1 and ctx.rank == 0:
            also["c2"] = stream_c2(ctx, args, max(20000, int(1000000 * sc)), 2000, 200, max(args.warmup, 5), not args.no_e2e)
        if ctx.world == 8:
            F5, N5, _ = WORKLOADS["c5"]
            also["c5"] = all_to_all(ctx, args, "c5", max(8000, int(F5 * sc)), N5, args.steps, args.warmup, False, False)
    if ctx.rank == 0:
        line["also"] = also or None
        print(json.dumps(line))
    if ctx.world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
