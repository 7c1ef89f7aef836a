#!/usr/bin/env python
"""Run the REFERENCE's own solver (oracle/_ref/ref_kspace: its unmodified sources + cuFFT, built by oracle/ref_build)
on the synthetic cases of the parity tests and keep its outputs as small fixtures.  Needs a GPU:

    gpurun -- python oracle/make_ref_goldens.py        # writes gpurun_out/golden/ref_*.npz
    cp gpurun_out/golden/ref_*.npz tests/golden/

The fixtures pin the oracle (tests/test_oracle_cpu.py::test_oracle_matches_reference_run) and are compared with the CUDA
path directly (tests/test_solver_gpu.py::test_matches_reference_fixture).  TEST INFRASTRUCTURE ONLY.
"""
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import kwh5  # noqa: E402

synth = importlib.import_module("k-wave-fluid-cuda_b200.synth")
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_kspace")

# name -> (shape, make_case kwargs, nt, command-line flags, datasets to keep)
CASES = {
    "nonlinear_absorbing_index": ((32, 32, 32), dict(nonlinear=True, absorbing=True, source="p_plane", n_sensor=64), 60,
                                  ["-p", "--p_max", "--p_rms", "--u_raw", "--p_final"], ["p", "p_max", "p_rms", "ux", "uy", "uz", "p_final"]),
    "linear_lossless_p0_cuboid": ((32, 32, 32), dict(nonlinear=False, absorbing=False, source="p0", sensor="cuboid"), 60,
                                  ["-p", "--p_max_all", "--p_min_all"], ["p/1", "p/2", "p_max_all", "p_min_all"]),
    "nonlinear_lossless_u_plane": ((32, 32, 32), dict(nonlinear=True, absorbing=False, source="u_plane", n_sensor=64), 60,
                                   ["-p", "--u_max"], ["p", "ux_max"]),
    "linear_absorbing_transducer": ((32, 32, 32), dict(nonlinear=False, absorbing=True, source="transducer", n_sensor=64), 60,
                                    ["-p"], ["p"]),
    "additive_p_source": ((32, 32, 32), dict(nonlinear=True, absorbing=True, source="p_plane", source_mode=2, n_sensor=64), 60,
                          ["-p"], ["p"]),
    "dirichlet_many": ((32, 32, 32), dict(nonlinear=True, absorbing=True, source="p_many", source_mode=0, n_sensor=64), 60,
                       ["-p"], ["p"]),
    "homogeneous_scalars": ((32, 32, 32), dict(nonlinear=True, absorbing=True, heterogeneous=False, source="p_plane", n_sensor=64), 60,
                            ["-p"], ["p"]),
    "non_cubic_shuffled": ((64, 32, 16), dict(nonlinear=True, absorbing=True, source="p_plane", shuffle_sensor=True, n_sensor=64), 60,
                           ["-p"], ["p"]),
    "n64_long": ((64, 64, 64), dict(nonlinear=True, absorbing=True, source="p_plane", n_sensor=64), 300,
                 ["-p", "--p_rms"], ["p", "p_rms"]),
    "two_d_p_source": ((64, 32, 1), dict(nonlinear=True, absorbing=True, source="p_plane", n_sensor=64), 150,
                       ["-p", "--p_rms", "-u"], ["p", "p_rms", "ux"]),
    "two_d_p0_cuboid": ((32, 32, 1), dict(nonlinear=False, absorbing=False, source="p0", sensor="cuboid"), 100,
                        ["-p", "--p_max"], ["p/1", "p/2", "p_max/1"]),
    "compressed_p_and_intensity": ((32, 32, 32), dict(nonlinear=False, absorbing=False, source="p_plane", n_sensor=64, period=20, shifts=True), 200,
                                   ["--p_c", "--u_non_staggered_c", "--I_avg_c", "--u_non_staggered_raw", "--period", "20", "--mos", "1", "--harmonics", "2"],
                                   ["p_c", "ux_non_staggered_c", "Ix_avg_c", "Iy_avg_c", "Iz_avg_c", "ux_non_staggered", "uy_non_staggered", "uz_non_staggered"]),
}  # fmt: skip


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = sys.argv[1:]
    for name, (shape, kwargs, nt, flags, keep) in CASES.items():
        if only and name not in only:
            continue
        cfg, arrays = synth.make_case(*shape, nt=nt, **kwargs)
        with tempfile.TemporaryDirectory() as tmp:
            fin, fout = os.path.join(tmp, "in.h5"), os.path.join(tmp, "out.h5")
            kwh5.write_input(fin, cfg, arrays)
            cmd = [BIN, "-i", fin, "-o", fout, "-t", "4", "--verbose", "1"] + flags
            t0 = time.time()
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                print(f"[{name}] reference failed rc={r.returncode}\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}")
                continue
            res = kwh5.read_output(fout)
            print(f"[{name}] ok in {time.time() - t0:.1f}s; datasets: {sorted(res)[:40]}")
            data = {k.replace("/", "__"): res[k] for k in keep if k in res}
            missing = [k for k in keep if k not in res]
            if missing:
                print(f"[{name}] missing datasets: {missing}")
            np.savez_compressed(os.path.join(out_dir, f"ref_{name}.npz"), make_case=json.dumps(dict(kwargs, shape=list(shape))),
                                nt=nt, flags=json.dumps(flags), **data)


if __name__ == "__main__":
    main()
