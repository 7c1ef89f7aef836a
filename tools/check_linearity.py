"""Open question of round 1: for a linear medium, doubling p0 should double every output EXACTLY in FP32 (the FP32 NumPy oracle\ndoes); the CUDA path at 256^3 did not (exact: False, magnitude not yet measured).  Run on a GPU to quantify."""
import importlib, numpy as np, sys, time
sys.path.insert(0, ".")
kw = importlib.import_module("k-wave-fluid-cuda_b200")
n, nt = 256, 4
t0 = time.time()
cfg, arrays = kw.synth.make_case(n, nt=nt, nonlinear=False, absorbing=True, source="p0", sensor="index", n_sensor=4096, medium="waves", pml_size=20)
outs = []
for scale in (1.0, 1.0, 2.0):  # the first two runs also answer: is a 256^3 run bit-reproducible?
    a = dict(arrays); a["p0_source_input"] = (arrays["p0_source_input"] * np.float32(scale)).astype(np.float32)
    sim = kw.Simulation(cfg, a, streams=["KW_S_P_RAW", "KW_S_P_MAX_ALL", "KW_S_UX_RAW"], raw_rows_capacity=nt)
    sim.run(nt); sim.finish()
    outs.append({s: sim.fetch(s) for s in ("KW_S_P_RAW", "KW_S_P_MAX_ALL", "KW_S_UX_RAW")}); sim.close()
for s in outs[0]:
    print(s, "run-to-run identical:", bool(np.array_equal(outs[0][s], outs[1][s])))
for s in outs[0]:
    d = outs[2][s].astype(np.float64) - 2.0 * outs[0][s]
    print(s, "scale", np.abs(outs[0][s]).max(), "exact:", bool((d == 0).all()), "differing", int((d != 0).sum()), "of", d.size,
          "max |diff|", np.abs(d).max(), "rel-L2", np.linalg.norm(d) / np.linalg.norm(outs[2][s]))
print("sec", time.time() - t0)
