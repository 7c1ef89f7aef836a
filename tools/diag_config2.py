"""Diagnostic: config-2 style runs (compressed p + I_avg_c, index mask) of kspaceFirstOrder-B200 vs the reference binary over a few
variations, every dataset's error printed (no asserts)."""
import importlib, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import kwh5
kw = importlib.import_module("k-wave-fluid-cuda_b200")
OURS = os.path.join(ROOT, "k-wave-fluid-cuda_b200", "kspaceFirstOrder-B200"); REF = os.path.join(ROOT, "oracle", "_ref", "ref_kspace")
def run(n, nt, threads, extra_flags=(), **kwargs):
    base = dict(nonlinear=False, absorbing=False, source="p_plane", sensor="index", n_sensor=4096, period=50, shifts=True, shuffle_sensor=True)
    base.update(kwargs)
    cfg, arrays = kw.synth.make_case(n, nt=nt, **base)
    d = tempfile.mkdtemp()
    fin = os.path.join(d, "in.h5"); kwh5.write_input(fin, cfg, arrays)
    flags = ["--p_c", "--I_avg_c", "--u_non_staggered_c", "--period", str(base["period"]), "--mos", "1", "--harmonics", "2"] + list(extra_flags)
    outs = {}
    for name, b in (("ours", OURS), ("ref", REF)):
        fo = os.path.join(d, name + ".h5")
        r = subprocess.run([b, "-i", fin, "-o", fo, "-t", str(threads), "--verbose", "0"] + flags, capture_output=True, text=True)
        if r.returncode: print(name, "FAILED", r.stderr[-300:]); return
        outs[name] = kwh5.read_file(fo)
    print(f"--- n={n} nt={nt} threads={threads} {kwargs} {extra_flags}")
    for p, o in sorted(outs["ref"].items()):
        if o["kind"] != "f32" or o["data"].size <= 1: continue
        a, b = outs["ours"][p]["data"].astype(np.float64), o["data"].astype(np.float64)
        print(f"   {p:28s} {str(a.shape):18s} rel-L2 {np.linalg.norm(a-b)/max(np.linalg.norm(b),1e-300):.3e} max-abs {np.abs(a-b).max():.3e} scale {np.abs(b).max():.3e}")
run(128, 400, os.cpu_count())
run(128, 400, 1)
run(128, 400, 4, shuffle_sensor=False)
run(128, 200, 4)
run(64, 400, 4, n_sensor=1024)
run(128, 400, 4, period=20)
run(128, 400, 4, n_sensor=512)
