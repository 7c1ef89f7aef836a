"""Parity of the CUDA time loop (through the C ABI) against the FP64 oracle on the same seeded inputs.
Bar (BASELINE.json north_star): relative L2 <= 1e-5 on sampled p/u after the full run, max-abs stated;
sensor ordering / mask handling bit-exact (checked through shuffled masks and cuboid ordering)."""
import numpy as np
import pytest

import fixtures
from oracle import kspace_oracle as ko

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def run_cuda(kw, cfg, arrays, nt, streams, start_index=0):
    sim = kw.Simulation(cfg, arrays, streams=streams, start_index=start_index, raw_rows_capacity=nt)
    done = sim.run(nt)
    assert done == nt
    sim.finish()
    out = {s: sim.fetch(s) for s in streams}
    out["p_final"] = sim.get_array("KW_P")
    out["ux_final"] = sim.get_array("KW_UX_SGX")
    out["launches"] = sim.launch_count()
    sim.close()
    return out


CASES = {
    # name: (grid, make_case kwargs)
    "nonlinear_absorbing_het_index": ((32, 32, 32), dict(nonlinear=True, absorbing=True, source="p_plane")),
    "linear_lossless_p0_cuboid": ((32, 32, 32), dict(nonlinear=False, absorbing=False, source="p0", sensor="cuboid")),
    "nonlinear_lossless_u_plane": ((32, 32, 32), dict(nonlinear=True, absorbing=False, source="u_plane")),
    "linear_absorbing_transducer": ((32, 32, 32), dict(nonlinear=False, absorbing=True, source="transducer")),
    "additive_p_source": ((32, 32, 32), dict(nonlinear=True, absorbing=True, source="p_plane", source_mode=2)),
    "additive_u_source": ((32, 32, 32), dict(nonlinear=False, absorbing=False, source="u_plane", source_mode=2)),
    "dirichlet_many": ((32, 32, 32), dict(nonlinear=True, absorbing=True, source="p_many", source_mode=0)),
    "homogeneous_scalars": ((32, 32, 32), dict(nonlinear=True, absorbing=True, heterogeneous=False, source="p_plane")),
    "non_cubic": ((64, 32, 16), dict(nonlinear=True, absorbing=True, source="p_plane", shuffle_sensor=True)),
    "n64": ((64, 64, 64), dict(nonlinear=True, absorbing=True, source="p_plane")),
    # 2-D simulations (Nz == 1, Parameters.h:88-94): no z arrays in the input
    "2d_nonlinear_absorbing": ((64, 32, 1), dict(nonlinear=True, absorbing=True, source="p_plane", n_sensor=100, shuffle_sensor=True)),
    "2d_p0_cuboid": ((64, 64, 1), dict(nonlinear=False, absorbing=False, source="p0", sensor="cuboid")),
    "2d_additive_u_source": ((32, 64, 1), dict(nonlinear=True, absorbing=False, source="u_plane", source_mode=2, n_sensor=64)),
}


@pytest.mark.parametrize("name", list(CASES))
def test_time_loop_matches_oracle(kw, synth, name):
    (nx, ny, nz), kwargs = CASES[name]
    nt = 60
    cfg, arrays = synth.make_case(nx, ny, nz, nt=nt, **kwargs)
    ref = ko.run(cfg, arrays, nt=nt, dtype=np.float64, record=("p_raw", "p_max", "p_rms", "u_raw", "p_final"))
    got = run_cuda(kw, cfg, arrays, nt, ["KW_S_P_RAW", "KW_S_P_MAX", "KW_S_P_RMS", "KW_S_UX_RAW"])
    assert got["KW_S_P_RAW"].shape == ref["p"].shape
    for a, b, what in (
        (got["KW_S_P_RAW"], ref["p"], "p raw"),
        (got["KW_S_UX_RAW"], ref["ux"], "ux raw"),
        (got["KW_S_P_MAX"][0], ref["p_max"], "p max"),
        (got["KW_S_P_RMS"][0], ref["p_rms"], "p rms"),
        (got["p_final"], ref["p_final"], "p final"),
    ):
        err = rel_l2(a, b)
        print(f"{name}: {what}: rel-L2 {err:.3e}, max-abs {np.abs(a - b).max():.3e} (scale {np.abs(b).max():.3e})")
        assert err <= TOL, (name, what, err)


def test_whole_domain_aggregates(kw, synth):
    nt = 40
    cfg, arrays = synth.make_case(32, nt=nt, source="p0", sensor="full_cuboid")
    ref = ko.run(cfg, arrays, nt=nt, dtype=np.float64, record=("p_max_all", "p_min_all", "p_rms", "u_max_all"))
    got = run_cuda(kw, cfg, arrays, nt, ["KW_S_P_MAX_ALL", "KW_S_P_MIN_ALL", "KW_S_P_RMS", "KW_S_UX_MAX_ALL"])
    assert rel_l2(got["KW_S_P_MAX_ALL"][0], ref["p_max_all"]) <= TOL
    assert rel_l2(got["KW_S_P_MIN_ALL"][0], ref["p_min_all"]) <= TOL
    assert rel_l2(got["KW_S_P_RMS"][0], ref["p_rms"]) <= TOL
    assert rel_l2(got["KW_S_UX_MAX_ALL"][0], ref["ux_max_all"]) <= TOL


def test_sampling_start_and_chunked_fetch(kw, synth):
    """-s semantics (CommandLineParameters.cpp:424) and KW_ERR_STREAM_FULL draining give the same rows."""
    nt, start = 30, 7
    cfg, arrays = synth.make_case(32, nt=nt, source="p_plane", shuffle_sensor=True)
    ref = ko.run(cfg, arrays, nt=nt, dtype=np.float64, record=("p_raw",), start_index=start)
    sim = kw.Simulation(cfg, arrays, streams=["KW_S_P_RAW"], start_index=start, raw_rows_capacity=5)
    rows = []
    while sim.t_index < nt:
        try:
            sim.run(nt)
        except kw.KwError as e:
            assert e.code == -5
        rows.append(sim.fetch("KW_S_P_RAW"))
    got = np.concatenate(rows)
    sim.close()
    assert got.shape == ref["p"].shape == (nt - start, arrays["sensor_mask_index"].size)
    assert rel_l2(got, ref["p"]) <= TOL


@pytest.mark.parametrize("name", [n for n in fixtures.fixture_names() if "compressed" not in n])
def test_matches_reference_fixture(kw, synth, name):
    """The CUDA path against the outputs of the reference's own solver (cuFFT build run on a B200, tests/golden/ref_*.npz)
    on the same input: rel-L2 <= 1e-5 on every sampled series / aggregate kept in the fixture."""
    shape, kwargs, nt, flags, data = fixtures.load_fixture(name)
    cfg, arrays = synth.make_case(*shape, nt=nt, **kwargs)
    streams = ["KW_S_P_RAW"]
    want = {"p_max": "KW_S_P_MAX", "p_rms": "KW_S_P_RMS", "p_max_all": "KW_S_P_MAX_ALL", "p_min_all": "KW_S_P_MIN_ALL",
            "ux": "KW_S_UX_RAW", "ux_max": "KW_S_UX_MAX"}
    streams += [v for k, v in want.items() if k in data]
    got = run_cuda(kw, cfg, arrays, nt, streams)
    for key, (a, b) in fixtures.time_series_views(got["KW_S_P_RAW"], data, nt).items():
        err = fixtures.rel_l2(a, b)
        print(f"{name}: {key}: rel-L2 {err:.3e}, max-abs {np.abs(a - b).max():.3e} (scale {np.abs(b).max():.3e})")
        assert err <= TOL, (name, key, err)
    for k, sid in want.items():
        if k in data:
            a = got[sid] if k == "ux" else got[sid][0]
            err = fixtures.rel_l2(a, data[k])
            print(f"{name}: {k}: rel-L2 {err:.3e}")
            assert err <= TOL, (name, k, err)
    if "p_final" in data:
        assert fixtures.rel_l2(got["p_final"], data["p_final"]) <= TOL



def test_asynchronous_output_returns_the_same_rows(kw, synth):
    """kw_stream_async: a full device row buffer travels to pinned host memory on its own stream while the loop samples into the second
    buffer (the reference's zero-copy + one-step-delayed flush, IndexOutputStream.cpp:253-263, :583-591, has the same purpose); the
    rows fetched are bit-identical to the synchronous path, in the same order."""
    nt = 40
    cfg, arrays = synth.make_case(32, nt=nt, source="p_plane", shuffle_sensor=True)
    sync = kw.Simulation(cfg, arrays, streams=["KW_S_P_RAW", "KW_S_UX_RAW"], raw_rows_capacity=nt)
    sync.run(nt)
    want = {s: sync.fetch(s) for s in ("KW_S_P_RAW", "KW_S_UX_RAW")}
    sync.close()
    sim = kw.Simulation(cfg, arrays, streams=["KW_S_P_RAW", "KW_S_UX_RAW"], raw_rows_capacity=6, async_output=True)
    rows = {"KW_S_P_RAW": [], "KW_S_UX_RAW": []}
    saw_pending, stalls = False, 0
    while sim.t_index < nt:
        try:
            sim.run(9)  # 9 steps per call: the first buffer (6 rows) fills inside a call and leaves without stopping the loop
        except kw.KwError as e:
            assert e.code == -5  # both buffers occupied: drain
            stalls += 1
        for s in rows:
            while True:
                saw_pending |= sim.pending(s) > 0
                got = sim.fetch(s)
                if got.shape[0] == 0:
                    break
                rows[s].append(got)
    for s in rows:
        got = np.concatenate(rows[s])
        assert got.shape == want[s].shape
        assert np.array_equal(got.view(np.uint32), want[s].view(np.uint32)), s
    assert saw_pending
    sim.close()
