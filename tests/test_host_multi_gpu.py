"""kspaceFirstOrder-B200 --gpus N (host/Team.h: one process per GPU, slab decomposition along z, rank 0 owns the files): the same input
file run on 1 GPU and on 2 GPUs gives the same output file, bit for bit -- raw and compressed series assembled in mask order, aggregates,
whole-domain maxima, final fields, intensities, Q term; also with cuboid masks and across checkpoint legs.  Needs >= 2 GPUs
(gpurun --gpus 2); no counterpart in the single-GPU reference (main.cpp:840-966)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import kwh5  # noqa: E402

OURS = os.path.join(ROOT, "k-wave-fluid-cuda_b200", "kspaceFirstOrder-B200")
pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


def run(fin, fout, flags, gpus):
    r = subprocess.run([OURS, "-i", fin, "-o", fout, "-t", "4", "--verbose", "0", "--gpus", str(gpus)] + flags, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, f"--gpus {gpus} failed:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"
    return kwh5.read_file(fout)


def same_bits(a, b):
    assert set(a) == set(b), sorted(set(a) ^ set(b))
    for p, o in a.items():
        if o["kind"] == "group":
            continue
        x, y = o["data"], b[p]["data"]
        assert x.shape == y.shape, p
        assert np.array_equal(x.view(np.uint32) if o["kind"] == "f32" else x, y.view(np.uint32) if o["kind"] == "f32" else y), p


CASES = {
    "index": (dict(nonlinear=True, absorbing=True, source="p_plane", n_sensor=200, shuffle_sensor=True, period=20, shifts=True),
              ["-p", "-u", "--p_rms", "--p_max", "--p_min", "--p_max_all", "--u_min_all", "--p_final", "--u_final", "--p_c", "--u_non_staggered_raw",
               "--I_avg_c", "--Q_term_c", "--period", "20", "--harmonics", "2", "-s", "5"]),
    "cuboid": (dict(nonlinear=False, absorbing=False, source="p0", sensor="cuboid"), ["-p", "--p_rms", "--p_max", "--u_max", "--p_max_all", "--copy_sensor_mask"]),
}  # fmt: skip


@pytest.mark.parametrize("name", list(CASES))
def test_two_gpus_write_the_file_one_gpu_writes(synth, tmp_path, name):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    kwargs, flags = CASES[name]
    cfg, arrays = synth.make_case(32, nt=90, **kwargs)
    fin = str(tmp_path / "in.h5")
    kwh5.write_input(fin, cfg, arrays)
    one = run(fin, str(tmp_path / "one.h5"), flags, 1)
    two = run(fin, str(tmp_path / "two.h5"), flags, 2)
    same_bits(one, two)


def test_two_gpu_run_resumes_from_its_checkpoint(synth, tmp_path):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    kwargs, flags = CASES["index"]
    cfg, arrays = synth.make_case(32, nt=90, **kwargs)
    fin = str(tmp_path / "in.h5")
    kwh5.write_input(fin, cfg, arrays)
    whole = run(fin, str(tmp_path / "whole.h5"), flags, 2)
    ck, out = str(tmp_path / "ck.h5"), str(tmp_path / "legs.h5")
    legs = 0
    while True:
        run(fin, out, flags + ["--checkpoint_file", ck, "--checkpoint_timesteps", "40"], 2)
        legs += 1
        if not os.path.exists(ck):
            break
        assert legs < 5
    assert legs == 3
    same_bits(whole, kwh5.read_file(out))


@pytest.mark.parametrize("sensor", ["index", "cuboid"])
def test_intensities_q_term_and_post_on_two_gpus(synth, tmp_path, sensor):
    """--I_avg / --Q_term (post-processing of the stored raw series, KSpaceFirstOrderSolver.cpp:1231-1534, computeQTerm :1783-2080) and --post
    (:231-239) on a slab-decomposed run: rank 0 forms the intensities from its file, the Q term's 3-D transforms run on every rank.
    --gpus 2 writes the datasets --gpus 1 writes, bit for bit; --post --gpus 2 on a stored file equals --post --gpus 1."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    kwargs = dict(nonlinear=True, absorbing=True, source="p_plane", n_sensor=150, period=20, shifts=True, shuffle_sensor=(sensor == "index"))
    if sensor == "cuboid":
        kwargs["sensor"] = "cuboid"
    cfg, arrays = synth.make_case(32, nt=80, **kwargs)
    fin = str(tmp_path / "in.h5")
    kwh5.write_input(fin, cfg, arrays)
    comp = ["--period", "20", "--harmonics", "2"]
    flags = ["--I_avg", "--Q_term", "--I_avg_c", "--Q_term_c", "-p"] + comp
    one = run(fin, str(tmp_path / "one.h5"), flags, 1)
    two = run(fin, str(tmp_path / "two.h5"), flags, 2)
    assert any(n.startswith("/Q_term") for n in one) and any(n.startswith("/Ix_avg") for n in one)
    same_bits(one, two)
    # --post: the same stored file post-processed by one and by two GPUs
    import shutil

    stored = str(tmp_path / "stored.h5")
    run(fin, stored, ["-p", "--u_non_staggered_raw", "--p_c", "--u_non_staggered_c"] + comp, 2)
    shutil.copy(stored, str(tmp_path / "stored1.h5"))
    post = ["--post", "--I_avg", "--Q_term", "--I_avg_c", "--Q_term_c"] + comp
    p1 = run(fin, str(tmp_path / "stored1.h5"), post, 1)
    p2 = run(fin, stored, post, 2)
    assert any(n.startswith("/Q_term_c") for n in p2)
    same_bits(p1, p2)


def test_512_cubed_from_a_file_on_two_gpus(synth, tmp_path):
    """BASELINE.json configs[3] physics (512^3 heterogeneous nonlinear absorbing, PML 20, whole-domain p_max_all + p_rms) from an input FILE:
    every rank reads its z-slab of the 0.5 GB datasets through the shared mapping, rank 0 assembles the whole-domain outputs.  --gpus 2 = --gpus 1,
    bit for bit.  (The 1024^3 configuration needs a 30 GB file; its slab code path is the one exercised here and in tests/test_slab_gpu.py.)"""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    cfg, arrays = synth.make_case(512, nt=12, nonlinear=True, absorbing=True, source="p_many", sensor="full_cuboid", pml_size=20, medium="waves")
    fin = str(tmp_path / "in.h5")
    kwh5.write_input(fin, cfg, arrays)
    del arrays
    flags = ["--p_max_all", "--p_rms", "--p_final"]
    one = run(fin, str(tmp_path / "one.h5"), flags, 1)
    two = run(fin, str(tmp_path / "two.h5"), flags, 2)
    os.remove(fin)
    assert one["/p_max_all"]["data"].shape == (512, 512, 512) and float(np.abs(one["/p_final"]["data"]).max()) > 0
    same_bits(one, two)
