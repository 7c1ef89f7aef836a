"""Localises where x -> 2x stops being exact on the GPU (round-1 open question).  Every FP32 operation of the linear step is
homogeneous unless an intermediate is subnormal, so each stage is probed on its own:
  1. the standalone 3-D R2C / C2R transforms on normal-range random data and on data with a subnormal tail;
  2. the fused z pass;
  3. the time loop after 1, 2, 3 steps (full fields), with the Gaussian p0 as generated and with p0 floored at 1e-3.
usage: probe_homogeneity.py [n]     (default 256)"""
import importlib
import sys

import numpy as np

sys.path.insert(0, ".")
kw = importlib.import_module("k-wave-fluid-cuda_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256


def report(tag, a1, a2):
    d = a2.astype(np.float64) - 2.0 * a1.astype(np.float64)
    nz = d != 0
    print(f"{tag}: differing {int(nz.sum())} of {d.size}, max |diff| {np.abs(d).max():.3e}, scale {np.abs(a1).max():.3e}", flush=True)


rng = np.random.default_rng(0)
x = rng.standard_normal((n, n, n), dtype=np.float32)
report("r2c random", kw.fft_r2c_3d(x).view(np.float32), kw.fft_r2c_3d(2 * x).view(np.float32))
xk = kw.fft_r2c_3d(x)
report("c2r random", kw.fft_c2r_3d(xk, n), kw.fft_c2r_3d(2 * xk, n))
zz, yy, xx = np.ogrid[:n, :n, :n]
g = (1.0e5 * np.exp(-((xx - n // 2) ** 2 + (yy - n // 2) ** 2 + (zz - n // 2) ** 2) / 128.0)).astype(np.float32)
print("gaussian: subnormal values", int(((g != 0) & (np.abs(g) < 1.1754944e-38)).sum()), "zeros", int((g == 0).sum()))
report("r2c gaussian", kw.fft_r2c_3d(g).view(np.float32), kw.fft_r2c_3d(2 * g).view(np.float32))
gf = np.maximum(g, np.float32(1e-3))
report("r2c gaussian floored", kw.fft_r2c_3d(gf).view(np.float32), kw.fft_r2c_3d(2 * gf).view(np.float32))
del x, xk, g, gf

m = 64
z = (rng.standard_normal((n, 16, 17)) + 1j * rng.standard_normal((n, 16, 17))).astype(np.complex64)
mul = rng.uniform(0.5, 1.5, size=(n, 16, 17)).astype(np.float32)
vz = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
report("zmid axis 2", kw.fft_zmid(z, 2, mul=mul, scal=1.0 / n, vec_z=vz).view(np.float32), kw.fft_zmid(2 * z, 2, mul=mul, scal=1.0 / n, vec_z=vz).view(np.float32))

for floor in (None, 1e-3):
    for absorbing in (False, True):
        for nt in (1, 2, 3):
            cfg, arrays = kw.synth.make_case(n, nt=nt, nonlinear=False, absorbing=absorbing, source="p0", sensor="index", n_sensor=64, medium="waves", pml_size=20)
            if floor:
                arrays["p0_source_input"] = np.maximum(arrays["p0_source_input"], np.float32(floor))
            outs = []
            for scale in (1.0, 2.0):
                a = dict(arrays)
                a["p0_source_input"] = (arrays["p0_source_input"] * np.float32(scale)).astype(np.float32)
                sim = kw.Simulation(cfg, a, streams=["KW_S_P_RAW"], raw_rows_capacity=nt)
                sim.run(nt)
                sim.finish()
                outs.append({k: sim.get_array(k) for k in ("KW_P", "KW_UX_SGX", "KW_UY_SGY", "KW_UZ_SGZ", "KW_RHOX", "KW_RHOZ")})
                sim.close()
            for k in outs[0]:
                report(f"floor={floor} absorbing={absorbing} nt={nt} {k}", outs[0][k], outs[1][k])
