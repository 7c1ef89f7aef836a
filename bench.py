#!/usr/bin/env python
"""Headline benchmark: Mvoxel-steps/s of the k-space time loop on synthetic heterogeneous media (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 512]

One JSON line on stdout (rank 0).  `value` is device-resident throughput (inputs in HBM when the timed region starts,
CUDA events on the solver stream); `e2e` is the same metric driven step by step through the C ABI with host buffers
(per-step H2D of the source row from pinned memory, per-step D2H of the sampled sensor row).  `roofline` describes the
dominant kernel of the step (live CUDA-event timing inside the library, algorithmic bytes from DESIGN.md), and
`cpu_baseline` is the NumPy/SciPy oracle port timed on the host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES = {  # algorithmic bytes per voxel-step (SURVEY.md 8(d) / DESIGN.md)
    (1, 1): 292.0, (0, 1): 284.0, (1, 0): 208.0, (0, 0): 204.0,
}  # (nonlinear, absorbing)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.t0 = self.t1 = None  # timed window (perf_counter); samples outside it are dropped

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)  # fmt: skip
            self.thread = threading.Thread(target=lambda: [self.lines.append((time.perf_counter(), l)) for l in self.proc.stdout], daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, l in self.lines:
            if self.t0 is not None and not (self.t0 <= ts <= self.t1 + 0.05):
                continue
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_throughput(size, budget_s=15.0, nonlinear=True, absorbing=True):
    """The oracle (NumPy/SciPy port of the reference's step) timed on the host cores; bounded sample."""
    from oracle import kspace_oracle as ko

    kw = importlib.import_module("k-wave-fluid-cuda_b200")
    sfft = importlib.import_module("scipy.fft")
    cores = os.cpu_count() or 1
    cfg, arrays = kw.synth.make_case(size, nt=1000, nonlinear=nonlinear, absorbing=absorbing, source="p_plane")
    with sfft.set_workers(cores):
        o = ko.KSpaceOracle(cfg, arrays, dtype=np.float32)
        o.step()  # warm-up (plans, page faults)
        t0, n = time.perf_counter(), 0
        while True:
            o.step()
            n += 1
            el = time.perf_counter() - t0
            if el >= budget_s or n >= 200:
                break
    v = size**3 * n / el / 1e6
    return {"value": v, "unit": "Mvoxel-steps/s", "cores": cores, "kind": "port",
            "sample": f"{size}^3 nonlinear+absorbing heterogeneous, {n} steps in {el:.1f} s, FP32 NumPy/SciPy (pocketfft, {cores} threads)"}  # fmt: skip


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_kspace")


def run_reference_binary(args):
    """The reference's own cuFFT CUDA build (its unmodified sources over the minih5 HDF5 shim, oracle/ref_build) on the
    same workload and the same box.  The reference stops its loop timer without a device sync, so it is timed
    from two runs (--benchmark n1, n2): T = its own simulation-phase + post-processing-phase timers (output-file header; the
    first device-to-host copy of post-processing drains the launch queue), per step = (T2 - T1) / (n2 - n1).  Whole-process
    wall times are recorded too but are dominated by file loading and pinning (tens of seconds at 512^3)."""
    import shutil
    import tempfile

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import kwh5

    kw = importlib.import_module("k-wave-fluid-cuda_b200")
    N, K, W = args.size, args.steps, max(3, args.warmup)
    cores = os.cpu_count() or 1
    tmp = tempfile.mkdtemp(prefix="kw_ref_")
    try:
        cfg, arrays = kw.synth.make_case(N, nt=W + K + 8, nonlinear=True, absorbing=True, source="p_plane", sensor="full_cuboid",
                                         pml_size=20 if N >= 128 else None)
        fin = os.path.join(tmp, "in.h5")
        kwh5.write_input(fin, cfg, arrays)
        del arrays
        times, phase = {}, {}
        for n in (W, W + K):
            fout = os.path.join(tmp, f"out_{n}.h5")
            cmd = [REF_BIN, "-i", fin, "-o", fout, "-t", str(cores), "--verbose", "0", "--benchmark", str(n), "--p_max_all", "--p_rms"]
            t0 = time.perf_counter()
            r = subprocess.run(cmd, capture_output=True, text=True)
            times[n] = time.perf_counter() - t0
            if r.returncode != 0:
                raise RuntimeError(f"reference binary failed: {r.stdout[-500:]} {r.stderr[-500:]}")
            at = kwh5.read_root_attrs(fout)
            phase[n] = float(at["simulation_phase_execution_time"].strip().rstrip("s")) + float(
                at["post-processing_phase_execution_time"].strip().rstrip("s"))
            os.remove(fout)
        per_step = (phase[W + K] - phase[W]) / K
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    value = N**3 / per_step / 1e6
    return {
        "impl": "reference", "metric": "Mvoxel-steps/s", "value": value, "unit": "Mvoxel-steps/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{N}^3 synthetic heterogeneous medium, nonlinear (BonA) + power-law absorption, PML 20, plane pressure source, "
                               f"--p_max_all --p_rms over a full-domain cuboid (BASELINE.json configs[3]); the reference's own cuFFT build "
                               f"(sm_100, cuFFT 11.4, KWH5 files through minih5), timed as (T[{W + K}] - T[{W}]) / {K}, T = its simulation + "
                               f"post-processing phase timers",
                   "grid": [N, N, N], "phase_s": {str(k): v for k, v in phase.items()}, "wall_s": {str(k): v for k, v in times.items()}},
        "cpu_baseline": {"value": value, "unit": "Mvoxel-steps/s", "cores": cores, "kind": "reference",
                         "sample": "the reference has no CPU solver: this is its GPU (cuFFT) build; host cores only do file I/O, pre-processing"},
        "e2e": {"value": value, "unit": "Mvoxel-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip


def run_reference(args):
    """--impl reference.  The reference has no CPU solver (BASELINE.json): the reported baseline is its own cuFFT CUDA build
    on the same box when it was compiled (oracle/_ref/ref_kspace) and a GPU is present, otherwise the oracle port on all
    host cores (bounded sample)."""
    have_gpu = False
    try:
        have_gpu = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).returncode == 0
    except Exception:
        pass
    if have_gpu and os.path.exists(REF_BIN):
        try:
            print(json.dumps(run_reference_binary(args)), flush=True)
            return
        except Exception as e:  # fall through to the CPU port, but say why
            print(f"bench.py: reference binary run failed ({e}); falling back to the CPU port", file=sys.stderr)
    size = min(args.size, 128)
    steps = max(1, args.steps)
    t_budget = min(120.0, 3.0 * steps)
    cb = cpu_port_throughput(size, budget_s=t_budget)
    line = {
        "impl": "reference", "metric": "Mvoxel-steps/s", "value": cb["value"], "unit": "Mvoxel-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": size**3 / cb["value"] / 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{size}^3 synthetic heterogeneous nonlinear absorbing medium, PML 20 (bounded CPU sample of the "
                               f"{args.size}^3 workload; the reference has no CPU solver: oracle port on host cores)"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "Mvoxel-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=0, help="grid edge; default 512 on one GPU (configs[3]), 1024 slab-decomposed on N > 1 (configs[4])")
    ap.add_argument("--replicas", action="store_true", help="N > 1: run N independent single-GPU replicas instead of one slab-decomposed grid")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    sharded = world > 1 and not args.replicas
    if not args.size:
        args.size = 1024 if sharded else 512
    if args.impl == "reference":
        if rank == 0:
            if world > 1:
                args.size = min(args.size, 512)  # the reference is single-GPU: its largest one-GPU configuration
            run_reference(args)
        return

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    kw = importlib.import_module("k-wave-fluid-cuda_b200")

    N, K, W = args.size, args.steps, max(3, args.warmup)
    nonlinear, absorbing = 1, 1
    KP = min(K, 10)  # profiled steps (per-kernel CUDA events) after the timed region
    nt = 2 * (K + W + KP) + 8
    # ---- workload: BASELINE.json configs[3]: 512^3 heterogeneous nonlinear absorbing, whole-domain p_max / p_rms
    # ---- N > 1: BASELINE.json configs[4]: ONE grid slab-decomposed along z over the ranks, all-to-all FFT transposes
    pml = 20 if N >= 128 else None
    slab_kw, sim_kw = {}, {}
    if N >= 1024 and not sharded:
        slab_kw = dict(medium="waves")  # the analytic medium: the low-passed noise of 1024^3 needs > 100 GB of host memory and minutes

    def fresh_comm_id():  # one ncclUniqueId per context: created on rank 0, handed to every rank
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(kw.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    if sharded:
        z0, nzl = kw.slab.slab_extent(N, rank, world)
        slab_kw = dict(medium="waves", z_range=(z0, nzl))
        sim_kw = dict(rank=rank, nranks=world, nccl_id=fresh_comm_id())
    cfg, arrays = kw.synth.make_case(N, nt=nt, nonlinear=True, absorbing=True, source="p_plane", sensor="full_cuboid", pml_size=pml, **slab_kw)
    streams = ["KW_S_P_RMS", "KW_S_P_MAX_ALL"]
    sim = kw.Simulation(cfg, arrays, streams=streams, device=local_rank, **sim_kw)
    del arrays

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with ClockSampler(local_rank) as clk:  # nvidia-smi needs ~1 s to start: launched before the warm-up, windowed below
        sim.run(W)  # warm-up steps (also the first-launch attribute setup)
        l0 = sim.launch_count()
        barrier()
        t0 = time.perf_counter()
        sim.run(K, sync=False)  # timed region: K steps, CUDA events on the solver stream around them
        host_enqueue_ms = (time.perf_counter() - t0) * 1e3 / K  # how long the host needs to enqueue one step
        sim.synchronize()
        barrier()
        t1 = time.perf_counter()
        clk.window(t0, t1)
        wall_ms = (t1 - t0) * 1e3
        dev_ms = sim.last_run_ms()
        launches = sim.launch_count() - l0
        time.sleep(0.1)
    clocks = clk.summary()
    # per-kernel timing for the roofline: a separate, profiled run of KP steps (events around every launch)
    sim.profile(True, True)
    sim.run(KP, sync=True)
    prof = sim.profile_report()
    sim.profile(False, False)
    comm_bytes = sim.comm_bytes() if sharded else 0.0
    comm_mode = sim.comm_mode()
    comm_steps = W + K + KP
    sim.close()
    if world > 1:
        tmax = torch.tensor([dev_ms], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dev_ms = float(tmax.item())
    ms_per_step = dev_ms / K
    jobs = 1 if sharded else world  # one decomposed grid, or `world` replicas of the grid
    value = jobs * N**3 * K / (dev_ms * 1e-3) / 1e6

    # ---- end to end: host-driven loop through the C ABI, host buffers in the timed region
    e2e = None
    if not args.no_e2e:
        cfg2, arrays2 = kw.synth.make_case(N, nt=nt, nonlinear=True, absorbing=True, source="p_many", sensor="index", n_sensor=4096, pml_size=pml, **slab_kw)
        nsrc = arrays2["p_source_index"].size
        sig = torch.from_numpy(np.ascontiguousarray(arrays2["p_source_input"]).reshape(nt, nsrc)).pin_memory()
        if sharded:
            sim_kw["nccl_id"] = fresh_comm_id()
        s2 = kw.Simulation(cfg2, arrays2, streams=["KW_S_P_RAW", "KW_S_P_RMS", "KW_S_P_MAX_ALL"], raw_rows_capacity=4, device=local_rank, **sim_kw)
        del arrays2
        out_rows = torch.empty((nt, 4096), dtype=torch.float32).pin_memory()
        import ctypes as C

        def one_step(t):
            row = sig[t]
            kw.capi._check(s2.lib.kw_set_source_row(s2.ctx, kw.ARRAY_IDS["KW_P_SOURCE_INPUT"], t, row.data_ptr(), nsrc))
            done = C.c_uint64()
            kw.capi._check(s2.lib.kw_run(s2.ctx, 1, C.byref(done), 0))
            got = C.c_uint64()
            kw.capi._check(s2.lib.kw_stream_fetch(s2.ctx, kw.STREAM_IDS["KW_S_P_RAW"], out_rows[t].data_ptr(), 4096, C.byref(got)))

        for t in range(W):
            one_step(t)
        barrier()
        t0 = time.perf_counter()
        for t in range(W, W + K):
            one_step(t)
        barrier()
        e2e_s = time.perf_counter() - t0
        s2.close()
        if world > 1:
            tm = torch.tensor([e2e_s], device="cuda")
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            e2e_s = float(tm.item())
        e2e = {"value": jobs * N**3 * K / e2e_s / 1e6, "unit": "Mvoxel-steps/s", "h2d_bytes_per_step": int(nsrc * 4),
               "d2h_bytes_per_step": 4096 * 4,
               "how": "one kw_set_source_row + kw_run(1) + kw_stream_fetch(p_raw row) per step from pinned host buffers (on every rank); wall clock, max over ranks"}  # fmt: skip

    if rank != 0:
        return
    peak, peak_src = measured_peak()
    # dominant KERNEL (the NCCL exchange of sharded runs is reported separately under "nvlink")
    def is_kernel(name):  # exchanges and the waits of the solver stream are reported, but they are not kernels
        return not (name.startswith("all_to_all") or name.startswith("idle_before_"))

    kern = {k: v for k, v in prof.items() if is_kernel(k)}
    top = max(kern.items(), key=lambda kv: kv[1]["ms"]) if kern else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if top and os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(str(N), {}).get(top[0])
        except Exception:
            traffic = None
    roofline = None
    if top:
        name, st = top
        ach = st["bytes"] / st["launches"] / (st["ms"] / st["launches"] * 1e-3) / 1e9
        alg = ALG_BYTES[(nonlinear, absorbing)] + 16.0  # + p_max_all and full-cuboid p_rms
        step_gbs = alg * N**3 / world / (ms_per_step * 1e-3) / 1e9 if sharded else alg * N**3 / (ms_per_step * 1e-3) / 1e9  # per GPU
        roofline = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "peak_source": peak_src,
                    "kernel_share_of_step": st["ms"] / sum(v["ms"] for v in kern.values()),
                    "step": {"algorithmic_bytes_per_voxel_step": alg, "achieved": step_gbs, "frac": step_gbs / peak},
                    "profiled_ms_per_step": sum(v["ms"] for v in kern.values()) / KP,
                    "solver_stream_idle_ms_per_step": sum(v["ms"] for k, v in prof.items() if k.startswith("idle_before_")) / KP,
                    "kernels": {k: {"launches_per_step": v["launches"] / KP, "ms_per_step": v["ms"] / KP,
                                    "GBps": v["bytes"] / max(v["ms"], 1e-9) / 1e6} for k, v in sorted(prof.items())}}  # fmt: skip
    line = {
        "metric": "Mvoxel-steps/s", "value": value, "unit": "Mvoxel-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{N}^3 synthetic heterogeneous medium, nonlinear (BonA) + power-law absorption, PML 20, plane pressure "
                               f"source, whole-domain p_max_all + p_rms over a full-domain cuboid (BASELINE.json "
                               + (f"configs[4]: ONE grid slab-decomposed along z over {world} GPUs, NCCL all-to-all per 3-D transform)" if sharded
                                  else "configs[3])"),
                   "grid": [N, N, N], "l2_policy": "inputs larger than L2 (every field >= 512 MiB per GPU)" if N**3 // (world if sharded else 1) >= 512**3 else
                   "working set partly L2 resident at this size",
                   "parallelism": "1 GPU" if world == 1 else (f"z-slabs over {world} GPUs" if sharded else f"{world} independent replicas"),
                   "wall_ms_timed_region": wall_ms, "host_enqueue_ms_per_step": host_enqueue_ms,
                   "same_grid_on_one_gpu": ("1024^3 on ONE B200 (python bench.py --size 1024): 113.3 ms/step = 9475 Mvoxel-steps/s, "
                                            "profiles/r01_o_bench_1024_1gpu.json" if sharded and N == 1024 else None)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
    }  # fmt: skip
    if sharded and "all_to_all" in prof:
        a2a = prof["all_to_all"]
        sent_per_step = comm_bytes / comm_steps  # bytes this rank sent per step (one direction)
        line["nvlink"] = {"all_to_all_ms_per_step": a2a["ms"] / KP, "exchanges_per_step": a2a["launches"] / KP,
                          "credit_wait_ms_per_step": prof.get("all_to_all_credit_wait", {}).get("ms", 0.0) / KP,
                          "arrival_wait_ms_per_step": prof.get("all_to_all_arrival_wait", {}).get("ms", 0.0) / KP,
                          "sent_bytes_per_gpu_per_step": sent_per_step,
                          "GBps_per_gpu_per_direction": sent_per_step / (a2a["ms"] / KP * 1e-3) / 1e9, "path": comm_mode,
                          "how": "rank 0: bytes sent per step / CUDA-event time of the pushes (peer path: the copies alone; waits for credits and "
                                 "arrivals are listed beside them) on the communication stream"}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_port_throughput(128, budget_s=12.0)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
