#!/usr/bin/env python
"""Headline benchmark: Mvoxel-steps/s of the k-space time loop on synthetic heterogeneous media (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5] [--size S]

One JSON line on stdout (rank 0).  `value` is device-resident throughput (inputs in HBM when the timed region starts,
CUDA events on the solver stream, max over ranks); `e2e` is the same metric on the SAME workload driven step by step
through the C ABI with host buffers (per-step H2D of the step's source row from pinned memory, per-step D2H of a result
row).  `roofline` describes the dominant kernel of the step (live CUDA-event timing inside the library, algorithmic bytes
of SURVEY.md 8(d) on the UNPADDED spectrum), `cpu_baseline` is the NumPy/SciPy oracle port timed on the host cores on a
bounded sample.

Workloads (BASELINE.json `configs`):
  --config 1  128^3 linear lossless, p0, two cuboids, raw p                       (stand-in for the bundled cuboid file)
  --config 2  128^3 linear lossless, index mask, --p_c --I_avg_c                  (stand-in for the bundled index file)
  --config 3  256^3 nonlinear + power-law absorption, PML 20, index mask, -p --p_max --p_rms
  --config 4  512^3 same physics, whole-domain p_max_all + p_rms over a full-domain cuboid      [default on 1 GPU]
  --config 5  the same physics slab-decomposed over the ranks                                   [default on N > 1 GPUs]
              default: WEAK scaling, 512^3 voxels per GPU (2: 512x512x1024, 4: 512x1024x1024, 8: 1024^3 = configs[4]);
              --size 1024: the 1024^3 grid on every N (strong scaling).
`--impl reference` runs the reference's own cuFFT build (oracle/_ref/ref_kspace) on the same workload, timed from repeated
pairs of --benchmark runs (its own phase timers have 10 ms resolution: the pair is >= 100 steps apart).
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

ALG_BYTES = {  # algorithmic bytes per voxel-step of the time loop itself (SURVEY.md 8(d) / DESIGN.md): (nonlinear, absorbing)
    (1, 1): 292.0, (0, 1): 284.0, (1, 0): 208.0, (0, 0): 204.0,
}  # fmt: skip


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


# ---- workloads -------------------------------------------------------------------------------------------------------
def grid_for(config, world, size):
    """(nx, ny, nz, scaling label)"""
    if config in (1, 2):
        n = size or 128
        return n, n, n, "weak"
    if config == 3:
        n = size or 256
        return n, n, n, "weak"
    if config == 4:
        n = size or 512
        return n, n, n, "weak"
    if size:  # config 5, explicit edge: one cubic grid on every N
        return size, size, size, "strong"
    f = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}.get(world)
    if f is None:
        raise SystemExit(f"config 5 (weak scaling) is defined for 1, 2, 4, 8 ranks; got {world} (use --size)")
    return 512 * f[0], 512 * f[1], 512 * f[2], "weak"


def workload(config, nx, ny, nz):
    """make_case kwargs, streams of our arm, flags of the reference arm, compression settings, description."""
    pml = 20 if min(nx, ny, nz) >= 128 else None
    g = f"{nx}x{ny}x{nz}" if not (nx == ny == nz) else f"{nx}^3"
    if config == 1:
        return dict(
            case=dict(nonlinear=False, absorbing=False, source="p0", sensor="cuboid", pml_size=10 if min(nx, ny, nz) >= 64 else None),
            streams=["KW_S_P_RAW"], ref_flags=["-p"], compression=None, physics=(0, 0), extra_bytes=0.0, raw="KW_S_P_RAW",
            text=f"{g} synthetic heterogeneous medium, linear lossless, initial pressure p0, two sensor cuboids, raw p series (BASELINE.json "
                 f"configs[0] stand-in: the bundled input_data_128_128_128_het_cuboid.h5 is not in the reference mount)")  # fmt: skip
    if config == 2:
        return dict(
            case=dict(nonlinear=False, absorbing=False, source="p_plane", sensor="index", n_sensor=4096, period=50, shifts=True, pml_size=pml),
            streams=["KW_S_P_C", "KW_S_IX_AVG_C", "KW_S_IY_AVG_C", "KW_S_IZ_AVG_C"], ref_flags=["--p_c", "--I_avg_c", "--period", "50", "--mos", "1", "--harmonics", "2"],
            compression=dict(period=50.0, mos=1, harmonics=2), physics=(0, 0), extra_bytes=0.0, raw="KW_S_P_C",
            text=f"{g} synthetic heterogeneous medium, linear lossless, plane tone-burst source (period 50 steps), index sensor mask of 4096 points, "
                 f"on-the-fly compressed p (2 harmonics) + time-averaged intensity (BASELINE.json configs[1] stand-in)")  # fmt: skip
    if config == 3:
        return dict(
            case=dict(nonlinear=True, absorbing=True, source="p_plane", sensor="index", n_sensor=4096, pml_size=pml),
            streams=["KW_S_P_RAW", "KW_S_P_MAX", "KW_S_P_RMS"], ref_flags=["-p", "--p_max", "--p_rms"], compression=None, physics=(1, 1),
            extra_bytes=0.0, raw="KW_S_P_RAW",
            text=f"{g} synthetic heterogeneous medium, nonlinear (BonA) + power-law absorption, PML 20, plane pressure source, index sensor mask "
                 f"of 4096 points, raw p + p_max + p_rms (BASELINE.json configs[2])")  # fmt: skip
    return dict(
        case=dict(nonlinear=True, absorbing=True, source="p_many", sensor="full_cuboid", pml_size=pml),
        streams=["KW_S_P_RMS", "KW_S_P_MAX_ALL"], ref_flags=["--p_max_all", "--p_rms"], compression=None, physics=(1, 1), extra_bytes=16.0, raw=None,
        text=f"{g} synthetic heterogeneous medium, nonlinear (BonA) + power-law absorption, PML 20, apodised plane pressure source (one signal per "
             f"source point), whole-domain p_max_all + p_rms over a full-domain cuboid (BASELINE.json configs[{3 if config == 4 else 4}])")  # fmt: skip


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


def run_reference_binary(args, config, nx, ny, nz):
    """The reference's own cuFFT CUDA build (its unmodified sources over the minih5 HDF5 shim, oracle/ref_build) on the
    same workload and the same box.  The reference stops its loop timer without a device sync and prints its
    phase timers with 10 ms resolution, so it is timed from PAIRS of runs (--benchmark n1, --benchmark n2, n2 - n1 >= 100
    steps): T = its own simulation-phase + post-processing-phase timers (the first device-to-host copy of post-processing
    drains the launch queue), per step = (T2 - T1) / (n2 - n1).  Every run is repeated `reps` times; the line reports the
    median pair and the min / max over all pairs, plus the same difference taken from whole-process wall clocks."""
    import shutil
    import tempfile

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import kwh5

    kw = importlib.import_module("k-wave-fluid-cuda_b200")
    K, W = args.steps, max(3, args.warmup)
    wl = workload(config, nx, ny, nz)
    nvox = nx * ny * nz
    span = max(K, 100 if nvox >= 256**3 else 1000)  # steps between the two runs of a pair
    reps = args.ref_reps
    cores = os.cpu_count() or 1
    tmp = tempfile.mkdtemp(prefix="kw_ref_")
    try:
        medium = dict(medium="waves") if nvox > 512**3 else {}
        cfg, arrays = kw.synth.make_case(nx, ny, nz, nt=W + span + 8, **wl["case"], **medium)
        fin = os.path.join(tmp, "in.h5")
        kwh5.write_input(fin, cfg, arrays)
        del arrays
        phase, wall = {W: [], W + span: []}, {W: [], W + span: []}
        for rep in range(reps):
            for n in (W, W + span):
                fout = os.path.join(tmp, f"out_{n}.h5")
                cmd = [REF_BIN, "-i", fin, "-o", fout, "-t", str(cores), "--verbose", "0", "--benchmark", str(n)] + wl["ref_flags"]
                t0 = time.perf_counter()
                r = subprocess.run(cmd, capture_output=True, text=True)
                wall[n].append(time.perf_counter() - t0)
                if r.returncode != 0:
                    raise RuntimeError(f"reference binary failed: {r.stdout[-500:]} {r.stderr[-500:]}")
                at = kwh5.read_root_attrs(fout)
                phase[n].append(float(at["simulation_phase_execution_time"].strip().rstrip("s")) + float(
                    at["post-processing_phase_execution_time"].strip().rstrip("s")))
                os.remove(fout)
        pairs = [(b - a) / span for a in phase[W] for b in phase[W + span]]
        per_step = (statistics.median(phase[W + span]) - statistics.median(phase[W])) / span
        wall_per_step = (statistics.median(wall[W + span]) - statistics.median(wall[W])) / span
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    value = nvox / per_step / 1e6
    return {
        "impl": "reference", "metric": "Mvoxel-steps/s", "value": value, "unit": "Mvoxel-steps/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["text"] + f"; the reference's own cuFFT build (sm_100, cuFFT 11.4, KWH5 files through minih5), flags "
                               f"{' '.join(wl['ref_flags'])}",
                   "grid": [nx, ny, nz],
                   "timing": f"(T[{W + span}] - T[{W}]) / {span} with T = its simulation + post-processing phase timers (10 ms resolution), "
                             f"median of {reps} repetitions of each run",
                   "ms_per_step_min": min(pairs) * 1e3, "ms_per_step_max": max(pairs) * 1e3,
                   "ms_per_step_from_process_wall_clock": wall_per_step * 1e3,
                   "phase_s": {str(k): v for k, v in phase.items()}, "wall_s": {str(k): v for k, v in wall.items()}},
        "cpu_baseline": {"value": value, "unit": "Mvoxel-steps/s", "cores": cores, "kind": "reference",
                         "sample": "the reference has no CPU solver: this is its GPU (cuFFT) build; host cores only do file I/O, pre-processing"},
        "e2e": {"value": value, "unit": "Mvoxel-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip


def run_reference(args, config, nx, ny, nz):
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
            print(json.dumps(run_reference_binary(args, config, nx, ny, nz)), flush=True)
            return
        except Exception as e:  # fall through to the CPU port, but say why
            print(f"bench.py: reference binary run failed ({e}); falling back to the CPU port", file=sys.stderr)
    size = min(nx, 128)
    steps = max(1, args.steps)
    t_budget = min(120.0, 3.0 * steps)
    cb = cpu_port_throughput(size, budget_s=t_budget)
    line = {
        "impl": "reference", "metric": "Mvoxel-steps/s", "value": cb["value"], "unit": "Mvoxel-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": size**3 / cb["value"] / 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{size}^3 synthetic heterogeneous nonlinear absorbing medium, PML 20 (bounded CPU sample of the "
                               f"{nx}x{ny}x{nz} workload; the reference has no CPU solver: oracle port on host cores)"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "Mvoxel-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line), flush=True)


def sharded_check(kw, dist, torch, rank, world, local_rank, fresh_comm_id):
    """Start-up correctness bit of a sharded run: a small case through the same slab-decomposed code path, compared on rank 0
    with the single-GPU run of the same input -- bit for bit (p_max_all over the whole grid, final pressure)."""
    n, nt = 64, 12
    kwargs = dict(nonlinear=True, absorbing=True, source="p_plane", sensor="full_cuboid")
    cfg, arrays = kw.synth.make_case(n, nt=nt, **kwargs)
    sim = kw.Simulation(cfg, arrays, streams=["KW_S_P_MAX_ALL", "KW_S_P_RMS"], device=local_rank, rank=rank, nranks=world, nccl_id=fresh_comm_id())
    sim.run(nt)
    sim.finish()
    mine = np.concatenate([sim.fetch("KW_S_P_MAX_ALL")[0], sim.get_array("KW_P").ravel()])
    sim.close()
    t = torch.from_numpy(mine).cuda()
    parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, parts, dst=0)
    if rank != 0:
        return None
    s1 = kw.Simulation(cfg, arrays, streams=["KW_S_P_MAX_ALL", "KW_S_P_RMS"], device=local_rank)
    s1.run(nt)
    s1.finish()
    one_max, one_p = s1.fetch("KW_S_P_MAX_ALL")[0], s1.get_array("KW_P").ravel()
    s1.close()
    half = n * n * (n // world)
    got_max = np.concatenate([p.cpu().numpy()[:half] for p in parts])
    got_p = np.concatenate([p.cpu().numpy()[half:] for p in parts])
    same = bool(np.array_equal(got_max.view(np.uint32), one_max.view(np.uint32)) and np.array_equal(got_p.view(np.uint32), one_p.view(np.uint32)))
    rel = float(np.linalg.norm(got_p.astype(np.float64) - one_p) / max(np.linalg.norm(one_p.astype(np.float64)), 1e-300))
    return {"grid": [n, n, n], "steps": nt, "ranks": world, "bit_identical_to_one_gpu": same, "rel_l2_p_final": rel}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=0, choices=[0, 1, 2, 3, 4, 5], help="BASELINE.json workload (see the module docstring); default 4 on one GPU, 5 on several")
    ap.add_argument("--size", type=int, default=0, help="grid edge override (config 5: one cubic grid on every N = strong scaling)")
    ap.add_argument("--replicas", action="store_true", help="N > 1: run N independent single-GPU replicas instead of one slab-decomposed grid")
    ap.add_argument("--ref-reps", type=int, default=3, help="--impl reference: repetitions of each --benchmark run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the sharded start-up correctness check")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    sharded = world > 1 and not args.replicas
    config = args.config or (5 if sharded else 4)
    if config == 5 and not sharded:
        config = 4
    if sharded and config != 5:
        raise SystemExit("slab-decomposed runs use --config 5")
    if args.impl == "reference":
        if rank == 0:
            # the reference is single-GPU: on N > 1 it runs its largest one-GPU configuration (configs[3], 512^3)
            rc = 4 if config == 5 else config
            nx, ny, nz, _ = grid_for(rc, 1, 0 if config == 5 else args.size)
            run_reference(args, rc, nx, ny, nz)
        return

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    kw = importlib.import_module("k-wave-fluid-cuda_b200")
    import ctypes as C

    nx, ny, nz, scaling = grid_for(config, world if sharded else 1, args.size)
    wl = workload(config, nx, ny, nz)
    nonlinear, absorbing = wl["physics"]
    K, W = args.steps, max(3, args.warmup)
    KP = min(K, 10)  # profiled steps (per-kernel CUDA events) after the timed region
    nt = 2 * (K + W + KP) + 8
    nvox = nx * ny * nz
    slab_kw, sim_kw = {}, {}
    if nvox > 512**3 and not sharded:
        slab_kw = dict(medium="waves")  # the analytic medium: low-passed noise beyond 512^3 needs > 100 GB of host memory and minutes

    def fresh_comm_id():  # one ncclUniqueId per context: created on rank 0, handed to every rank
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(kw.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    check = None
    if sharded and not args.no_check:
        check = sharded_check(kw, dist, torch, rank, world, local_rank, fresh_comm_id)
    if sharded:
        z0, nzl = kw.slab.slab_extent(nz, rank, world)
        slab_kw = dict(medium="waves", z_range=(z0, nzl))
        sim_kw = dict(rank=rank, nranks=world, nccl_id=fresh_comm_id())
    cfg, arrays = kw.synth.make_case(nx, ny, nz, nt=nt, **wl["case"], **slab_kw)
    many = bool(cfg.get("p_source_many"))
    nsrc = int(arrays["p_source_index"].size) if many else 0
    sig = None
    if many and not args.no_e2e:  # the source signal of the e2e leg stays on the host (pinned) and travels row by row
        sig = torch.from_numpy(np.ascontiguousarray(arrays["p_source_input"]).reshape(nt, nsrc)).pin_memory()

    def make_sim(rows_capacity):
        if sharded:
            sim_kw["nccl_id"] = fresh_comm_id()
        return kw.Simulation(cfg, arrays, streams=wl["streams"], device=local_rank, compression=wl["compression"],
                             raw_rows_capacity=rows_capacity, **sim_kw)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sim = make_sim(nt if wl["raw"] else 0)
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
    jobs = 1 if sharded or world == 1 else world  # one decomposed grid, or `world` replicas of the grid
    value = jobs * nvox * K / (dev_ms * 1e-3) / 1e6

    # ---- end to end: the SAME workload driven step by step through the C ABI, host buffers inside the timed region:
    #      H2D of the step's source row (pinned), kw_run(1), D2H of a result row (the raw / compressed rows of the step when the
    #      workload stores a series, otherwise 4096 values of the running p_max_all aggregate)
    e2e = None
    if not args.no_e2e:
        s2 = make_sim(4)
        peek = torch.empty(4096, dtype=torch.float32).pin_memory()
        raw_sid = kw.STREAM_IDS[wl["raw"]] if wl["raw"] else None
        row_floats = 0
        if raw_sid is not None:
            rf, rr = C.c_uint64(), C.c_uint64()
            kw.capi._check(s2.lib.kw_stream_info(s2.ctx, raw_sid, C.byref(rf), C.byref(rr)))
            row_floats = int(rf.value)
        rowbuf = torch.empty(max(4 * row_floats, 1), dtype=torch.float32).pin_memory()
        d2h = [0]

        def one_step(t):
            if many:
                kw.capi._check(s2.lib.kw_set_source_row(s2.ctx, kw.ARRAY_IDS["KW_P_SOURCE_INPUT"], t, sig[t].data_ptr(), nsrc))
            done = C.c_uint64()
            kw.capi._check(s2.lib.kw_run(s2.ctx, 1, C.byref(done), 0))
            got = C.c_uint64()
            if raw_sid is not None:
                kw.capi._check(s2.lib.kw_stream_fetch(s2.ctx, raw_sid, rowbuf.data_ptr(), rowbuf.numel(), C.byref(got)))
                d2h[0] += int(got.value) * row_floats * 4
            else:
                kw.capi._check(s2.lib.kw_stream_peek(s2.ctx, kw.STREAM_IDS["KW_S_P_MAX_ALL"], 0, peek.data_ptr(), peek.numel()))
                d2h[0] += peek.numel() * 4

        for t in range(W):
            one_step(t)
        barrier()
        d2h[0] = 0
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
        e2e = {"value": jobs * nvox * K / e2e_s / 1e6, "unit": "Mvoxel-steps/s", "h2d_bytes_per_step": int(nsrc * 4), "d2h_bytes_per_step": d2h[0] / K,
               "how": "same workload and streams as `value`; per step: kw_set_source_row (H2D of the step's source row from pinned memory) + kw_run(1) + "
                      + ("kw_stream_fetch of the rows the step produced" if raw_sid is not None else "kw_stream_peek (D2H of 4096 values of the running p_max_all)")
                      + " on every rank; wall clock, max over ranks"}  # fmt: skip

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()

    def is_kernel(name):  # exchanges and the waits of the solver stream are reported, but they are not kernels
        return not (name.startswith("all_to_all") or name.startswith("idle_before_"))

    kern = {k: v for k, v in prof.items() if is_kernel(k)}
    top = max(kern.items(), key=lambda kv: kv[1]["ms"]) if kern else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if top and os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(str(nx) if nx == ny == nz else f"{nx}x{ny}x{nz}", {}).get(top[0])
        except Exception:
            traffic = None
    roofline = None
    if top:
        name, st = top
        ach = st["bytes"] / st["launches"] / (st["ms"] / st["launches"] * 1e-3) / 1e9
        alg = ALG_BYTES[(nonlinear, absorbing)] + wl["extra_bytes"]
        step_gbs = alg * nvox / (world if sharded else 1) / (ms_per_step * 1e-3) / 1e9  # per GPU
        roofline = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes": "SURVEY.md 8(d) per-unit figures on the unpadded half spectrum (Nx/2+1 complex values per row)",
                    "kernel_share_of_step": st["ms"] / sum(v["ms"] for v in kern.values()),
                    "step": {"algorithmic_bytes_per_voxel_step": alg, "achieved": step_gbs, "frac": step_gbs / peak},
                    "profiled_ms_per_step": sum(v["ms"] for v in kern.values()) / KP,
                    "solver_stream_idle_ms_per_step": sum(v["ms"] for k, v in prof.items() if k.startswith("idle_before_")) / KP,
                    "kernels": {k: {"launches_per_step": v["launches"] / KP, "ms_per_step": v["ms"] / KP,
                                    "GBps": v["bytes"] / max(v["ms"], 1e-9) / 1e6} for k, v in sorted(prof.items())}}  # fmt: skip
    per_gpu_vox = nvox // (world if sharded else 1)
    line = {
        "metric": "Mvoxel-steps/s", "value": value, "unit": "Mvoxel-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": wl["text"] + (f"; ONE grid slab-decomposed along z over {world} GPUs, all-to-all per 3-D transform over NVLink "
                                             f"({'weak scaling: 512^3 voxels per GPU' if scaling == 'weak' else 'strong scaling'})" if sharded else ""),
                   "grid": [nx, ny, nz],
                   "l2_policy": "inputs larger than L2 (every field >= 512 MiB per GPU)" if per_gpu_vox >= 512**3 else
                   "working set partly L2 resident at this size (fields of %d MiB)" % (per_gpu_vox * 4 >> 20),
                   "parallelism": "1 GPU" if world == 1 else (f"z-slabs over {world} GPUs" if sharded else f"{world} independent replicas"),
                   "wall_ms_timed_region": wall_ms, "host_enqueue_ms_per_step": host_enqueue_ms},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
    }  # fmt: skip
    if check is not None:
        line["sharded_check"] = check
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
