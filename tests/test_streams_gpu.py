"""In-step output streams beyond raw / aggregate sampling: non-staggered velocity (computeShiftedVelocity,
KSpaceFirstOrderSolver.cpp:2714-2735), on-the-fly harmonic compression (IndexOutputStream.cpp:373-470,
CuboidOutputStream.cpp:431-532) incl. the 40-bit packing (CompressHelper.cpp:224-389) and the compressed time-averaged
intensity (IndexOutputStream.cpp:299-342, :477-490) -- against the reference's own run (tests/golden) and the oracle.
Bar: rel-L2 <= 1e-5 for coefficients and intensities; the 40-bit packed bytes bit-exact for the same sampled series."""
import numpy as np
import pytest

import fixtures
from oracle import compress_oracle as co
from oracle import kspace_oracle as ko

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel_l2(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def as_complex(frames, nsens, harm):
    f = np.asarray(frames).reshape(-1, nsens, harm, 2)
    return f[..., 0] + 1j * f[..., 1]


def test_compressed_streams_match_reference_run(kw, synth):
    """p_c, ux_non_staggered_c, I{x,y,z}_avg_c and the raw non-staggered velocity written by the reference's own solver."""
    shape, kwargs, nt, flags, data = fixtures.load_fixture("compressed_p_and_intensity")
    cfg, arrays = synth.make_case(*shape, nt=nt, **kwargs)
    period, harm = 20.0, 2
    streams = ["KW_S_P_C", "KW_S_UX_NS_C", "KW_S_UY_NS_C", "KW_S_UZ_NS_C", "KW_S_IX_AVG_C", "KW_S_IY_AVG_C", "KW_S_IZ_AVG_C",
               "KW_S_UX_NS_RAW", "KW_S_UY_NS_RAW", "KW_S_UZ_NS_RAW"]  # fmt: skip
    sim = kw.Simulation(cfg, arrays, streams=streams, raw_rows_capacity=nt, compression=dict(period=period, mos=1, harmonics=harm))
    assert sim.run(nt) == nt
    sim.finish()
    got = {s: sim.fetch(s) for s in streams}
    sim.close()
    nsens = arrays["sensor_mask_index"].size
    for key, sid in (("p_c", "KW_S_P_C"), ("ux_non_staggered_c", "KW_S_UX_NS_C")):
        ref = as_complex(data[key], nsens, harm)
        mine = as_complex(got[sid], nsens, harm)
        assert mine.shape == ref.shape, (key, mine.shape, ref.shape)
        err = rel_l2(mine, ref)
        print(f"{key}: {mine.shape[0]} frames, rel-L2 {err:.3e}, max-abs {np.abs(mine - ref).max():.3e} (scale {np.abs(ref).max():.3e})")
        assert err <= TOL, (key, err)
    for a in "xyz":  # the y/z components are small (plane wave along x, scattered by the heterogeneities) but not noise
        err = rel_l2(got[f"KW_S_U{a.upper()}_NS_RAW"], data[f"u{a}_non_staggered"].reshape(nt, nsens))
        print(f"u{a}_non_staggered raw: rel-L2 {err:.3e}")
        assert err <= (TOL if a == "x" else 5 * TOL)
        err = rel_l2(got[f"KW_S_I{a.upper()}_AVG_C"][0], data[f"I{a}_avg_c"].reshape(-1))
        print(f"I{a}_avg_c: rel-L2 {err:.3e}")
        assert err <= (TOL if a == "x" else 5 * TOL)


@pytest.mark.parametrize("variant", ["index_mos2", "cuboid", "no_overlap", "short_run", "staggered_u"])
def test_compression_matches_oracle(kw, synth, variant):
    """State machine variants against CompressedStream (FP64 accumulators) fed with the oracle's sampled series."""
    nt, period, mos, harm, no_overlap, start = 130, 16.0, 1, 3, False, 0
    kwargs = dict(nonlinear=True, absorbing=True, source="p_plane", n_sensor=48, period=16, shifts=True)
    if variant == "index_mos2":
        mos, start = 2, 5
        kwargs["shuffle_sensor"] = True
    elif variant == "cuboid":
        kwargs["sensor"] = "cuboid"
    elif variant == "no_overlap":
        no_overlap = True
    elif variant == "short_run":
        nt = 12  # fewer steps than one period: overlap is switched off (Parameters.cpp:141-145), one short last frame
    cfg, arrays = synth.make_case(32, nt=nt, **kwargs)
    ustream = "KW_S_UX_C" if variant == "staggered_u" else "KW_S_UX_NS_C"
    streams = ["KW_S_P_C", ustream, "KW_S_IX_AVG_C"]
    sim = kw.Simulation(cfg, arrays, streams=streams, start_index=start, raw_rows_capacity=nt,
                        compression=dict(period=period, mos=mos, harmonics=harm, no_overlap=no_overlap))
    assert sim.run(nt) == nt
    sim.finish()
    got = {s: sim.fetch(s) for s in streams}
    sim.close()
    rec = ("p_raw", "u_raw", "u_non_staggered_raw")
    out = ko.run(cfg, arrays, nt=nt, dtype=np.float64, record=rec, start_index=start)
    nsens = out["p"].shape[1]
    nsteps = nt - start
    eff_no_overlap = no_overlap or period >= nsteps
    sp = co.CompressedStream(nsens, period, mos, harm, shifted=False, no_overlap=eff_no_overlap, nsteps_total=nsteps, dtype=np.complex128)
    su = co.CompressedStream(nsens, period, mos, harm, shifted=(variant != "staggered_u"), no_overlap=eff_no_overlap, nsteps_total=nsteps, dtype=np.complex128)
    su_ns = co.CompressedStream(nsens, period, mos, harm, shifted=True, no_overlap=eff_no_overlap, nsteps_total=nsteps, dtype=np.complex128)
    useries = out["ux"] if variant == "staggered_u" else out["ux_non_staggered"]
    pf, uf, inten, nfr = [], [], np.zeros(nsens), 0
    for t in range(nsteps):
        a, b, b_ns = sp.feed(out["p"][t]), su.feed(useries[t]), su_ns.feed(out["ux_non_staggered"][t])
        if a is not None:
            pf.append(a), uf.append(b)
            inten += co.intensity_frame(a, b_ns)
            nfr += 1
    assert nfr >= 1
    for frames, ref, what in ((got["KW_S_P_C"], pf, "p_c"), (got[ustream], uf, ustream)):
        mine = as_complex(frames, nsens, harm)
        assert mine.shape[0] == len(ref), (what, mine.shape, len(ref))
        err = rel_l2(mine, np.stack(ref))
        print(f"{variant}: {what}: {len(ref)} frames, rel-L2 {err:.3e}")
        assert err <= TOL, (variant, what, err)
    err = rel_l2(got["KW_S_IX_AVG_C"][0], inten / nfr)
    print(f"{variant}: Ix_avg_c rel-L2 {err:.3e}")
    assert err <= TOL


@pytest.mark.parametrize("no_overlap", [False, True])
def test_40bit_frames_bit_exact(kw, synth, no_overlap):
    """--40-bit_complex: accumulators are kept packed and re-quantised at every step.  The packed frames must equal, byte
    for byte, the oracle's restatement of the same integer codec fed with the SAME sampled series (the raw p stream of
    this run), and decode to the float frames within the quantisation step."""
    nt, period, harm = 50, 10.0, 2
    cfg, arrays = synth.make_case(32, nt=nt, nonlinear=True, absorbing=True, source="p_plane", n_sensor=24, period=10, shifts=True)
    comp = dict(period=period, mos=1, harmonics=harm, c40=1, no_overlap=int(no_overlap))
    streams = ["KW_S_P_RAW", "KW_S_P_C", "KW_S_UX_NS_RAW", "KW_S_UX_NS_C", "KW_S_IX_AVG_C"]
    sim = kw.Simulation(cfg, arrays, streams=streams, raw_rows_capacity=nt, compression=comp)
    assert sim.run(nt) == nt
    sim.finish()
    got = {s: sim.fetch(s) for s in streams}
    # the bases depend on libm's cosf/sinf to an ulp (oracle vs C++ host): the bit-exact comparison uses the context's own,
    # after checking them against the oracle's
    bases = {}
    for shifted in (False, True):
        osz, bsz, be, be1 = sim.compression_bases(shifted)
        o_os, o_bs, o_be, o_be1 = co.generate_bases(period, 1, harm, True, shifted)
        assert (osz, bsz) == (o_os, o_bs)
        assert np.abs(be - o_be).max() <= 4e-7 * np.abs(o_be).max() and np.abs(be1 - o_be1).max() <= 4e-7 * np.abs(o_be1).max()
        bases[shifted] = (be, be1)
    sim.close()
    nsens = arrays["sensor_mask_index"].size
    row = int(np.ceil(np.float32(nsens) * np.float32(1.25))) * harm
    inten, nfr = np.zeros(nsens, np.float32), 0
    frames = {}
    for sid, raw, shifted in (("KW_S_P_C", "KW_S_P_RAW", False), ("KW_S_UX_NS_C", "KW_S_UX_NS_RAW", True)):
        assert got[sid].shape[1] == row
        s = co.CompressedStream(nsens, period, 1, harm, shifted=shifted, no_overlap=no_overlap, nsteps_total=nt, c40=True, bases=bases[shifted])
        ref = [f for f in (s.feed(got[raw][t]) for t in range(nt)) if f is not None]
        mine = got[sid].view(np.uint8).reshape(len(got[sid]), -1)[:, : nsens * harm * 5].reshape(-1, nsens, harm, 5)
        assert mine.shape[0] == len(ref)
        assert np.array_equal(mine, np.stack(ref)), f"{sid}: packed frames differ"
        frames[sid] = [co.unpack40(f, s.e) for f in ref]
    for a, b in zip(frames["KW_S_P_C"], frames["KW_S_UX_NS_C"]):
        for ih in range(harm):
            P, U = a[:, ih], b[:, ih]
            inten = (inten + (P.real * U.real + P.imag * U.imag).astype(np.float32) / np.float32(2.0)).astype(np.float32)
        nfr += 1
    ref_i = inten / np.float32(nfr)
    assert rel_l2(got["KW_S_IX_AVG_C"][0], ref_i) <= 1e-6


def test_internal_streams_are_not_fetchable(kw, synth):
    cfg, arrays = synth.make_case(32, nt=30, source="p_plane", n_sensor=16, period=10, shifts=True)
    sim = kw.Simulation(cfg, arrays, streams=["KW_S_IX_AVG_C"], compression=dict(period=10.0, harmonics=1))
    sim.run(30)
    sim.finish()
    assert sim.fetch("KW_S_IX_AVG_C").shape == (1, 16)
    with pytest.raises(kw.KwError):
        sim.fetch("KW_S_P_C")  # exists only as an input of I_avg_c (do-not-save stream of the reference)
    sim.close()
    with pytest.raises(kw.KwError):  # the shift operators are required for non-staggered outputs
        cfg2, arrays2 = synth.make_case(32, nt=4, source="p_plane", n_sensor=16)
        kw.Simulation(cfg2, arrays2, streams=["KW_S_UX_NS_RAW"])


def test_40bit_codec_bit_exact_against_reference_vectors(kw):
    """The device codec against the vectors produced by the reference's own CompressHelper.cpp (tests/golden/compress_ref.json)."""
    import json
    import os

    gold = json.load(open(os.path.join(fixtures.GOLD, "compress_ref.json")))["codec"]
    for e in sorted({c["e"] for c in gold}):
        cs = [c for c in gold if c["e"] == e]
        bits = np.array([[c["re"], c["im"]] for c in cs], dtype=np.uint32)
        vals = bits.view(np.float32).view(np.complex64).reshape(-1)
        packed = kw.c40_encode(vals, e)
        want = np.array([c["bytes"] for c in cs], dtype=np.uint8)
        assert np.array_equal(packed, want)
        dec = kw.c40_decode(want, e)
        want_bits = np.array([[c["dre"], c["dim"]] for c in cs], dtype=np.uint32)
        assert np.array_equal(dec.view(np.uint32).reshape(-1, 2), want_bits)


def cuboid_linear_indices(corners_1based, nx, ny):
    """0-based linear indices of all cuboid points: x fastest inside a cuboid, cuboids concatenated (the row order)."""
    out = []
    for x0, y0, z0, x1, y1, z1 in np.asarray(corners_1based, np.int64).reshape(-1, 6) - 1:
        z, y, x = np.meshgrid(np.arange(z0, z1 + 1), np.arange(y0, y1 + 1), np.arange(x0, x1 + 1), indexing="ij")
        out.append(((z * ny + y) * nx + x).reshape(-1))
    return np.concatenate(out)


@pytest.mark.parametrize("shape,sensor", [((32, 32, 32), "index"), ((64, 32, 1), "index"), ((32, 32, 32), "cuboid")], ids=["3d", "2d", "cuboid"])
def test_q_term_c_matches_oracle(kw, synth, shape, sensor):
    """--Q_term_c (cpp:1013-1020, :1783-2080): the divergence of the compressed time-averaged intensity, 3-D and 2-D, against
    the NumPy restatement applied to this run's own I_avg_c."""
    nt = 100
    cfg, arrays = synth.make_case(*shape, nt=nt, nonlinear=True, absorbing=True, source="p_plane", n_sensor=200, period=20, shifts=True, sensor=sensor)
    comps = ["X", "Y"] + (["Z"] if shape[2] > 1 else [])
    streams = [f"KW_S_I{a}_AVG_C" for a in comps] + ["KW_S_Q_TERM_C"]
    sim = kw.Simulation(cfg, arrays, streams=streams, compression=dict(period=20.0, harmonics=2))
    assert sim.run(nt) == nt
    sim.finish()
    got = {s: sim.fetch(s)[0] for s in streams}
    sim.close()
    if sensor == "index":
        idx = arrays["sensor_mask_index"].astype(np.int64) - 1
    else:
        idx = cuboid_linear_indices(arrays["sensor_mask_corners"], shape[0], shape[1])
    ref = co.q_term(cfg, [got[f"KW_S_I{a}_AVG_C"] for a in comps], idx)
    err = rel_l2(got["KW_S_Q_TERM_C"], ref)
    print(f"Q_term_c {shape}: rel-L2 {err:.3e}, scale {np.abs(ref).max():.3e}")
    assert np.abs(ref).max() > 0 and err <= TOL


@pytest.mark.parametrize("shape", [(32, 32, 32), (64, 32, 1)], ids=["3d", "2d"])
def test_raw_series_intensity_and_q_term(kw, synth, shape):
    """--I_avg / --Q_term building blocks (computeAverageIntensities cpp:1231-1534, computeQTerm :1783-2080): the half-step
    temporal shift of the stored velocity series (any number of steps: 77 here) and the Q term, against the NumPy restatement."""
    nt = 77
    cfg, arrays = synth.make_case(*shape, nt=nt, nonlinear=True, absorbing=True, source="p_plane", n_sensor=150, shifts=True)
    comps = ["X", "Y"] + (["Z"] if shape[2] > 1 else [])
    streams = ["KW_S_P_RAW"] + [f"KW_S_U{a}_NS_RAW" for a in comps]
    sim = kw.Simulation(cfg, arrays, streams=streams, raw_rows_capacity=nt)
    assert sim.run(nt) == nt
    sim.finish()
    got = {s: sim.fetch(s) for s in streams}
    p = got["KW_S_P_RAW"]
    us = [got[f"KW_S_U{a}_NS_RAW"] for a in comps]
    mine = kw.intensity_avg_block(p, us)
    scale = max(np.abs(co.intensity_avg(p, u)).max() for u in us)
    for a, m, u in zip(comps, mine, us):
        ref = co.intensity_avg(p, u)
        err = float(np.linalg.norm(m - ref) / max(np.linalg.norm(ref), 1e-300))
        print(f"I{a.lower()}_avg {shape}: rel-L2 {err:.3e}, max-abs {np.abs(m - ref).max():.3e} (scale {scale:.3e})")
        assert np.abs(m - ref).max() <= 2e-5 * scale
    q = sim.q_term(mine)
    sim.close()
    idx = arrays["sensor_mask_index"].astype(np.int64) - 1
    ref_q = co.q_term(cfg, mine, idx)
    err = rel_l2(q, ref_q)
    print(f"Q_term {shape}: rel-L2 {err:.3e}")
    assert err <= TOL
