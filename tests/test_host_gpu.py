"""The C++ host (k-wave-fluid-cuda_b200/host: KSpaceFirstOrderSolver interface, command line and file layout of the
reference) end to end: the same input FILE and the same FLAGS given to kspaceFirstOrder-B200 and to the reference's own
binary (its unmodified sources + cuFFT, oracle/_ref/ref_kspace), every dataset of the two output files compared --
names and shapes exactly, values to rel-L2 <= 1e-5 (max-abs printed).  For the components of a vector quantity (ux, uy, uz
and their derived outputs) the error of each component is taken relative to the largest component of that quantity: the
three travel through the same FP32 transforms, so a component two orders below the main one carries the same absolute
rounding noise in both codes (e.g. uz at 4 % of ux differs by 1.7e-8, exactly like ux)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import kwh5  # noqa: E402

OURS = os.path.join(ROOT, "k-wave-fluid-cuda_b200", "kspaceFirstOrder-B200")
REF = os.path.join(ROOT, "oracle", "_ref", "ref_kspace")
SCALARS = {"Nx", "Ny", "Nz", "Nt", "dt", "dx", "dy", "dz", "c_ref", "t_index"}

CASES = {
    "index_raw_and_aggregates": (dict(nonlinear=True, absorbing=True, source="p_plane", n_sensor=96, shuffle_sensor=True),
                                 ["-p", "--p_rms", "--p_max", "--p_min", "--p_max_all", "--p_min_all", "--p_final", "-u", "--u_max", "--u_final", "-s", "6"]),
    "cuboids": (dict(nonlinear=False, absorbing=False, source="p0", sensor="cuboid"), ["-p", "--p_rms", "--p_max", "--u_min_all", "--copy_sensor_mask"]),
    "compressed": (dict(nonlinear=False, absorbing=False, source="p_plane", n_sensor=64, period=20, shifts=True),
                   ["--p_c", "--u_non_staggered_c", "--I_avg_c", "--u_non_staggered_raw", "--period", "20", "--mos", "1", "--harmonics", "2"]),
    "transducer_u_sources": (dict(nonlinear=True, absorbing=True, source="transducer", n_sensor=32), ["-p", "-u", "--u_rms"]),
    "no_output_flags": (dict(nonlinear=False, absorbing=False, source="p0", n_sensor=8), []),
    "2d": (dict(ny=64, nz=1, nonlinear=True, absorbing=True, source="p_plane", n_sensor=80, period=20, shifts=True),
           ["-p", "--p_rms", "--p_max_all", "--p_final", "-u", "--u_final", "--u_non_staggered_raw", "--p_c", "--I_avg_c", "--period", "20", "--harmonics", "2"]),
    "q_term_c": (dict(nonlinear=True, absorbing=True, source="p_plane", n_sensor=300, period=20, shifts=True),
                 ["--Q_term_c", "--I_avg_c", "--period", "20", "--harmonics", "2"]),
    # (cuboid masks with --Q_term_c / --I_avg_c: the reference's own binary aborts or writes NaN there, so that combination is
    #  checked against the oracle instead: tests/test_streams_gpu.py::test_q_term_c_matches_oracle[cuboid])
    "i_avg_q_term_raw": (dict(nonlinear=True, absorbing=True, source="p_plane", n_sensor=120, shifts=True), ["--I_avg", "--Q_term", "--block_size", "50"]),
    "cuboid_compressed": (dict(nonlinear=True, absorbing=True, source="p_plane", sensor="cuboid", period=20, shifts=True),
                          ["--p_c", "--u_c", "--u_non_staggered_c", "--period", "20", "--harmonics", "2", "-s", "3"]),
    "2d_p0": (dict(ny=32, nz=1, nonlinear=False, absorbing=False, source="p0", sensor="cuboid"), ["-p", "--p_min", "--u_max_all"]),
}


def run(binary, fin, fout, flags, may_fail=False):
    r = subprocess.run([binary, "-i", fin, "-o", fout, "-t", "4", "--verbose", "0"] + flags, capture_output=True, text=True)
    if may_fail and r.returncode != 0:
        return None
    assert r.returncode == 0, f"{os.path.basename(binary)} failed:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"
    return kwh5.read_file(fout)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_same_file_same_flags_same_output(synth, tmp_path, name):
    if not os.path.exists(REF):
        pytest.skip("reference binary not built (oracle/ref_build)")
    assert os.path.exists(OURS), "kspaceFirstOrder-B200 not built (__graft_entry__.build())"
    kwargs, flags = CASES[name]
    nt = 120
    kwargs = dict(kwargs)
    ny, nz = kwargs.pop("ny", None), kwargs.pop("nz", None)
    cfg, arrays = synth.make_case(32, ny, nz, nt=nt, **kwargs)
    fin = str(tmp_path / "in.h5")
    kwh5.write_input(fin, cfg, arrays)
    got = run(OURS, fin, str(tmp_path / "out.h5"), flags)
    ref = run(REF, fin, str(tmp_path / "ref.h5"), flags, may_fail=True)
    if ref is None:  # the reference's own binary aborts on some flag / mask combinations: ours ran, nothing to compare with
        assert all(np.isfinite(o["data"]).all() for o in got.values() if o["kind"] == "f32")
        pytest.skip("the reference binary fails on this configuration")
    ref_ds = {p: o for p, o in ref.items() if o["kind"] != "group"}
    got_ds = {p: o for p, o in got.items() if o["kind"] != "group"}
    assert set(ref_ds) == set(got_ds), (sorted(set(ref_ds) ^ set(got_ds)))
    assert {p for p, o in ref.items() if o["kind"] == "group"} == {p for p, o in got.items() if o["kind"] == "group"}
    compared = 0
    for p, o in sorted(ref_ds.items()):
        a, b = got_ds[p]["data"], o["data"]
        assert a.shape == b.shape and got_ds[p]["kind"] == o["kind"], (p, a.shape, b.shape)
        for k in ("data_type", "domain_type", "c_harmonics", "c_type", "c_mos", "c_shift", "c_max_exp", "c_period", "c_complex_size"):
            assert got_ds[p]["attrs"].get(k) == o["attrs"].get(k), (p, k)
        if o["kind"] == "u64" or p.strip("/") in SCALARS or a.size == 1:
            assert np.array_equal(a, b), p
            continue
        if not np.isfinite(b).all():  # the reference itself produced NaN/inf here: nothing to match, ours must stay finite
            print(f"{name}: {p}: the reference output is not finite ({int((~np.isfinite(b)).sum())} values); ours checked for finiteness only")
            assert np.isfinite(a).all(), p
            continue
        nb = np.linalg.norm(b.astype(np.float64).ravel())
        base = p.strip("/")
        if base[:2] in ("ux", "uy", "uz", "Ix", "Iy", "Iz"):  # vector quantity: norm of its largest component
            sib = ["/" + base[0] + c + base[2:] for c in "xyz"]
            nb = max(np.linalg.norm(ref_ds[q]["data"].astype(np.float64).ravel()) for q in sib if q in ref_ds)
        err = np.linalg.norm((a.astype(np.float64) - b).ravel()) / max(nb, 1e-300)
        print(f"{name}: {p} {a.shape}: rel-L2 {err:.3e}, max-abs {np.abs(a - b).max():.3e} (scale {np.abs(b).max():.3e})")
        assert err <= 1e-5, (p, err)
        compared += 1
    assert compared >= 1 or not flags
    for k in ("file_type", "major_version", "minor_version"):
        assert got["/"]["attrs"].get(k) == ref["/"]["attrs"].get(k)


def test_command_line_errors_exit_like_the_reference(tmp_path):
    """Boxed message on stderr + EXIT_FAILURE (Logger/Logger.cpp:82-89); -s is 1-based; unsupported features say so."""
    assert os.path.exists(OURS), "kspaceFirstOrder-B200 not built (__graft_entry__.build())"
    for args, needle in (([], "Input file was not specified"), (["-i", "a"], "Output file was not specified"),
                         (["-i", "a", "-o", "b", "--checkpoint_interval", "5"], "Checkpoint file was not specified"),
                         (["-i", "a", "-o", "b", "--checkpoint_file", "c"], "Checkpoint interval or the number of time steps"),
                         (["-i", "a", "-o", "b", "--post"], "--post needs at least one"), (["-i", "a", "-o", "b", "--Q_term_c"], "--period or --frequency"),
                         (["-i", "a", "-o", "b", "--p_c"], "--period or --frequency"), (["-i", "a", "-o", "b", "-s", "0"], "Invalid value"),
                         (["-i", "a", "-o", "b", "-c", "12"], "Invalid value"),
                         (["-i", str(tmp_path / "missing.h5"), "-o", "b"], "could not be opened")):  # fmt: skip
        r = subprocess.run([OURS] + args, capture_output=True, text=True)
        assert r.returncode == 1, args
        assert "K-Wave experienced a fatal error" in r.stderr and needle in r.stderr, (args, r.stderr)
    r = subprocess.run([OURS, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--p_max_all" in r.stdout


@pytest.mark.gpu
def test_device_selection_follows_the_reference(synth, tmp_path):
    """-g N must name an existing device (CudaParameters::selectDevice, Parameters/CudaParameters.cpp:81-167: "Wrong CUDA device id");
    without -g the first usable device is taken."""
    cfg, arrays = synth.make_case(16, nt=4, source="p0")
    fin = str(tmp_path / "in.h5")
    kwh5.write_input(fin, cfg, arrays)
    r = subprocess.run([OURS, "-i", fin, "-o", str(tmp_path / "o.h5"), "-g", "97"], capture_output=True, text=True)
    assert r.returncode == 1 and "Wrong CUDA device id 97. Allowed devices <0," in r.stderr, r.stderr
    r = subprocess.run([OURS, "-i", fin, "-o", str(tmp_path / "o.h5"), "-g", "0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([OURS, "-i", fin, "-o", str(tmp_path / "o2.h5")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_no_gpu_means_error_exit_not_cpu_fallback(synth, tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg, arrays = synth.make_case(16, nt=4, source="p0")
    fin = str(tmp_path / "in.h5")
    kwh5.write_input(fin, cfg, arrays)
    r = subprocess.run([OURS, "-i", fin, "-o", str(tmp_path / "o.h5")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


def test_no_gpu_multi_process_run_exits_without_hanging(synth, tmp_path):
    """--gpus 2 (host/Team.h: one forked process per GPU): when the ranks cannot get a device every process reports the error and the
    parent exits with EXIT_FAILURE -- no rank is left waiting on a socket."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg, arrays = synth.make_case(16, nt=4, source="p0")
    fin = str(tmp_path / "in.h5")
    kwh5.write_input(fin, cfg, arrays)
    r = subprocess.run([OURS, "-i", fin, "-o", str(tmp_path / "o.h5"), "--gpus", "2"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and r.stderr.count("no CUDA device") >= 1
    # rank counts other than 1, 2, 4, 8 are rejected on the command line, before any process is forked
    r = subprocess.run([OURS, "-i", fin, "-o", str(tmp_path / "o.h5"), "--gpus", "3"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Invalid value of --gpus" in r.stderr, r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("sensor", ["index", "cuboid"])
def test_checkpoint_restart_is_bit_identical(synth, tmp_path, sensor):
    """--checkpoint_file / --checkpoint_timesteps (KSpaceFirstOrderSolver.cpp:1176-1224, :186-228): a run interrupted twice and
    resumed with the same command line writes the same output file, bit for bit, as the uninterrupted run -- raw series,
    aggregates, compressed frames (accumulators live across the interruption), intensities, final fields."""
    assert os.path.exists(OURS), "kspaceFirstOrder-B200 not built (__graft_entry__.build())"
    nt = 130
    kwargs = dict(nonlinear=True, absorbing=True, source="p_plane", n_sensor=48, period=20, shifts=True)
    if sensor == "cuboid":
        kwargs["sensor"] = "cuboid"
    cfg, arrays = synth.make_case(32, nt=nt, **kwargs)
    fin = str(tmp_path / "in.h5")
    kwh5.write_input(fin, cfg, arrays)
    flags = ["-p", "--p_rms", "--p_max", "--p_max_all", "--p_final", "--u_final", "--p_c", "--I_avg_c", "--u_non_staggered_raw", "--u_max",
             "--period", "20", "--harmonics", "2", "-s", "4"]  # fmt: skip
    whole = run(OURS, fin, str(tmp_path / "whole.h5"), flags)
    ck, out = str(tmp_path / "ck.h5"), str(tmp_path / "legs.h5")
    legs = 0
    while True:
        r = subprocess.run([OURS, "-i", fin, "-o", out, "-t", "4", "--verbose", "0", "--checkpoint_file", ck, "--checkpoint_timesteps", "47"] + flags,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        legs += 1
        if not os.path.exists(ck):
            break
        assert legs < 5
        assert kwh5.read_root_attrs(ck)["file_type"] == "checkpoint"
    assert legs == 3  # 47 + 47 + 36 steps
    got = kwh5.read_file(out)
    assert set(got) == set(whole)
    for p, o in whole.items():
        if o["kind"] == "group":
            continue
        assert np.array_equal(got[p]["data"].view(np.uint32) if o["kind"] == "f32" else got[p]["data"],
                              o["data"].view(np.uint32) if o["kind"] == "f32" else o["data"]), p


def _compare_files(got, ref, tol=1e-5, label=""):
    ref_ds = {p: o for p, o in ref.items() if o["kind"] != "group"}
    got_ds = {p: o for p, o in got.items() if o["kind"] != "group"}
    assert set(ref_ds) == set(got_ds), sorted(set(ref_ds) ^ set(got_ds))
    for p, o in sorted(ref_ds.items()):
        a, b = got_ds[p]["data"], o["data"]
        assert a.shape == b.shape, (p, a.shape, b.shape)
        if o["kind"] == "u64" or a.size == 1:
            assert np.array_equal(a, b), p
            continue
        nb = np.linalg.norm(b.astype(np.float64).ravel())
        base = p.strip("/")
        if base[:2] in ("ux", "uy", "uz", "Ix", "Iy", "Iz"):
            sib = ["/" + base[0] + c + base[2:] for c in "xyz"]
            nb = max(np.linalg.norm(ref_ds[q]["data"].astype(np.float64).ravel()) for q in sib if q in ref_ds)
        err = np.linalg.norm((a.astype(np.float64) - b).ravel()) / max(nb, 1e-300)
        print(f"{label}: {p} {a.shape}: rel-L2 {err:.3e}, max-abs {np.abs(a - b).max():.3e}")
        assert err <= tol, (label, p, err)


@pytest.mark.gpu
@pytest.mark.parametrize("first", ["reference", "ours"])
def test_checkpoint_files_are_interchangeable_with_the_reference(synth, tmp_path, first):
    """The checkpoint layout is the reference's (KSpaceFirstOrderSolver.cpp:1176-1224, :186-228; BaseOutputStream.cpp:528-606): the seven
    state arrays + t_index + Nx, Ny, Nz + header in the checkpoint file, compression accumulators as Temp_<name>_1 / _2, running
    intensities as Temp_<name>, aggregate accumulators flushed into the output file.  A run interrupted by ONE code is resumed by the
    OTHER code and must end with the output of the reference's uninterrupted run (rel-L2 <= 1e-5).
    Two defects of the reference shape the test: its index streams read min / max attributes on reopen that it never stores
    (IndexOutputStream.cpp:244 vs :551-555) -- our checkpoints carry placeholders, so it can resume OURS but not its own; and after a
    reopen its I_avg_c stream flushes at row t_index - start of a one-row dataset (IndexOutputStream.cpp:203-207, :583-591), so the
    reference can never finish a resumed run that has --I_avg_c -- the direction in which IT resumes runs without that flag."""
    if not os.path.exists(REF):
        pytest.skip("reference binary not built (oracle/ref_build)")
    assert os.path.exists(OURS), "kspaceFirstOrder-B200 not built (__graft_entry__.build())"
    nt = 130
    cfg, arrays = synth.make_case(32, nt=nt, nonlinear=True, absorbing=True, source="p_plane", n_sensor=48, period=20, shifts=True)
    fin = str(tmp_path / "in.h5")
    kwh5.write_input(fin, cfg, arrays)
    flags = ["-p", "--p_rms", "--p_max", "--p_min", "--p_max_all", "--p_final", "--u_final", "--p_c", "--I_avg_c", "--u_non_staggered_raw", "--u_max",
             "--u_min_all", "--period", "20", "--harmonics", "2", "-s", "4"]  # fmt: skip
    if first == "ours":
        flags.remove("--I_avg_c")
    whole = run(REF, fin, str(tmp_path / "whole.h5"), flags)
    ck, out = str(tmp_path / "ck.h5"), str(tmp_path / "legs.h5")
    order = [REF, OURS] if first == "reference" else [OURS, REF]
    every = "47" if first == "reference" else "70"  # reference first: legs of 47 + 47 + 36 steps; ours first: 70 + 60
    legs = 0
    while True:
        binary = order[min(legs, 1)]  # first leg by one code, every later leg by the other
        r = subprocess.run([binary, "-i", fin, "-o", out, "-t", "4", "--verbose", "0", "--checkpoint_file", ck, "--checkpoint_timesteps", every] + flags,
                           capture_output=True, text=True)
        assert r.returncode == 0, f"leg {legs} ({os.path.basename(binary)}):\n{r.stdout[-1500:]}\n{r.stderr[-1500:]}"
        legs += 1
        if not os.path.exists(ck):
            break
        assert legs < 5
        assert kwh5.read_root_attrs(ck)["file_type"] == "checkpoint"
        if legs == 1:  # what the first leg left behind has the reference's object names
            names = set(kwh5.read_file(ck))
            want = {"/p", "/ux_sgx", "/uy_sgy", "/uz_sgz", "/rhox", "/rhoy", "/rhoz", "/t_index", "/Nx", "/Ny", "/Nz", "/Temp_p_c_1", "/Temp_p_c_2"}
            if "--I_avg_c" in flags:
                want |= {"/Temp_ux_non_staggered_c_1", "/Temp_Ix_avg_c"}
            assert want <= names, sorted(names)
    assert legs == (3 if first == "reference" else 2)
    _compare_files(kwh5.read_file(out), whole, label=f"{first} first")


@pytest.mark.gpu
@pytest.mark.parametrize("sensor", ["index", "cuboid"])
def test_post_processing_of_an_existing_output_file(synth, tmp_path, sensor):
    """--post (KSpaceFirstOrderSolver.cpp:231-239, :975-1030, computeAverageIntensitiesC :1543-1775): a first run stores the raw series and the
    compression coefficients; a second invocation with --post computes I_avg / Q_term from the stored series and I_avg_c / Q_term_c from the
    stored coefficients, without a time loop.  The results equal those of a run that computed them directly -- ours (<= 1e-6) and,
    for the index mask, the reference's (<= 1e-5; its binary fails on these flags with cuboid masks)."""
    assert os.path.exists(OURS), "kspaceFirstOrder-B200 not built (__graft_entry__.build())"
    nt = 120
    kwargs = dict(nonlinear=True, absorbing=True, source="p_plane", n_sensor=150, period=20, shifts=True)
    if sensor == "cuboid":
        kwargs["sensor"] = "cuboid"
    cfg, arrays = synth.make_case(32, nt=nt, **kwargs)
    fin = str(tmp_path / "in.h5")
    kwh5.write_input(fin, cfg, arrays)
    comp = ["--period", "20", "--harmonics", "2"]
    stored = str(tmp_path / "stored.h5")
    run(OURS, fin, stored, ["-p", "--u_non_staggered_raw", "--p_c", "--u_non_staggered_c"] + comp)
    before = set(kwh5.read_file(stored))
    assert not any(n.startswith(("/Ix_avg", "/Q_term")) for n in before)
    post = run(OURS, fin, stored, ["--post", "--I_avg", "--Q_term", "--I_avg_c", "--Q_term_c"] + comp)
    direct = run(OURS, fin, str(tmp_path / "direct.h5"), ["--I_avg", "--Q_term", "--I_avg_c", "--Q_term_c"] + comp)
    want = [n for n in direct if n.strip("/").split("/")[0] in ("Ix_avg", "Iy_avg", "Iz_avg", "Q_term", "Ix_avg_c", "Iy_avg_c", "Iz_avg_c", "Q_term_c")
            and direct[n]["kind"] == "f32"]
    assert len(want) >= 8
    for n in want:
        a, b = post[n]["data"].astype(np.float64), direct[n]["data"].astype(np.float64)
        err = np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)
        print(f"--post vs direct run ({sensor}): {n}: rel-L2 {err:.3e}")
        assert err <= 1e-6, (n, err)
    assert before <= set(post)  # everything the first run stored is still there
    if sensor == "index" and os.path.exists(REF):
        ref = run(REF, fin, str(tmp_path / "ref.h5"), ["--I_avg", "--Q_term", "--I_avg_c", "--Q_term_c"] + comp, may_fail=True)
        if ref is not None:
            for n in want:
                a, b = post[n]["data"].astype(np.float64), ref[n]["data"].astype(np.float64)
                base = n.strip("/")
                nb = np.linalg.norm(b.ravel())
                if base[:2] in ("Ix", "Iy", "Iz"):
                    nb = max(np.linalg.norm(ref["/" + base[0] + c + base[2:]]["data"].astype(np.float64).ravel()) for c in "xyz")
                err = np.linalg.norm((a - b).ravel()) / max(nb, 1e-300)
                print(f"--post vs the reference's direct run: {n}: rel-L2 {err:.3e}")
                assert err <= 1e-5, (n, err)
