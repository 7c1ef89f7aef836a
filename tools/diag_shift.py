"""Diagnostic: u_non_staggered_raw of the CUDA path vs the FP64 oracle for long x / y / z axes."""
import importlib, sys
import numpy as np
sys.path.insert(0, ".")
kw = importlib.import_module("k-wave-fluid-cuda_b200")
from oracle import kspace_oracle as ko
for shape in [(64, 32, 32), (128, 32, 32), (256, 32, 32), (512, 32, 32), (32, 128, 32), (32, 32, 128), (128, 128, 128)]:
    nt = 12
    cfg, arrays = kw.synth.make_case(*shape, nt=nt, nonlinear=False, absorbing=False, source="p_plane", n_sensor=shape[0] * shape[1], shifts=True)
    ref = ko.run(cfg, arrays, nt=nt, dtype=np.float64, record=("u_non_staggered_raw", "u_raw"))
    sim = kw.Simulation(cfg, arrays, streams=["KW_S_UX_NS_RAW", "KW_S_UY_NS_RAW", "KW_S_UZ_NS_RAW", "KW_S_UX_RAW"], raw_rows_capacity=nt)
    sim.run(nt); sim.finish()
    for sid, key in (("KW_S_UX_RAW", "ux"), ("KW_S_UX_NS_RAW", "ux_non_staggered"), ("KW_S_UY_NS_RAW", "uy_non_staggered"), ("KW_S_UZ_NS_RAW", "uz_non_staggered")):
        a, b = sim.fetch(sid).astype(np.float64), ref[key]
        print(shape, key, f"rel-L2 {np.linalg.norm(a-b)/max(np.linalg.norm(b),1e-300):.3e} max-abs {np.abs(a-b).max():.3e} scale {np.abs(b).max():.3e}", flush=True)
    sim.close()
