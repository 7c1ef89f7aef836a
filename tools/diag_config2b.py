"""Diagnostic: raw vs compressed non-staggered velocity, ours vs reference binary vs the compression oracle fed with either raw series."""
import importlib, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import kwh5
from oracle import compress_oracle as co
kw = importlib.import_module("k-wave-fluid-cuda_b200")
OURS = os.path.join(ROOT, "k-wave-fluid-cuda_b200", "kspaceFirstOrder-B200"); REF = os.path.join(ROOT, "oracle", "_ref", "ref_kspace")
def rel(a, b): return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)
for n, nt in ((64, 200), (128, 200)):
    cfg, arrays = kw.synth.make_case(n, nt=nt, nonlinear=False, absorbing=False, source="p_plane", sensor="index", n_sensor=1024, period=50, shifts=True)
    d = tempfile.mkdtemp(); fin = os.path.join(d, "in.h5"); kwh5.write_input(fin, cfg, arrays)
    flags = ["--u_non_staggered_raw", "--u_non_staggered_c", "-u", "--u_c", "--period", "50", "--mos", "1", "--harmonics", "2"]
    outs = {}
    for name, b in (("ours", OURS), ("ref", REF)):
        fo = os.path.join(d, name + ".h5")
        r = subprocess.run([b, "-i", fin, "-o", fo, "-t", "4", "--verbose", "0"] + flags, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-300:]
        outs[name] = kwh5.read_file(fo)
    print(f"--- n={n}")
    for p in ("/ux", "/ux_non_staggered", "/ux_c", "/ux_non_staggered_c", "/uy_non_staggered", "/uy_non_staggered_c"):
        a, b = outs["ours"][p]["data"].astype(np.float64), outs["ref"][p]["data"].astype(np.float64)
        print(f"   ours vs ref {p:24s} {str(a.shape):16s} rel-L2 {rel(a,b):.3e} max-abs {np.abs(a-b).max():.3e} scale {np.abs(b).max():.3e}")
    for who in ("ours", "ref"):
        raw = outs[who]["/ux_non_staggered"]["data"].reshape(nt, -1)
        st = co.CompressedStream(raw.shape[1], 50.0, 1, 2, shifted=True, nsteps_total=nt)
        frames = [f for f in (st.feed(raw[t]) for t in range(nt)) if f is not None]
        want = np.stack(frames).astype(np.complex64).view(np.float32).reshape(len(frames), -1)
        for other in ("ours", "ref"):
            got = outs[other]["/ux_non_staggered_c"]["data"].reshape(len(frames), -1)
            print(f"   oracle-compress({who} raw) vs {other} ux_non_staggered_c: rel-L2 {rel(got.astype(np.float64), want.astype(np.float64)):.3e}")
    # per-frame error
    a, b = outs["ours"]["/ux_non_staggered_c"]["data"].reshape(-1, 1024, 2, 2).astype(np.float64), outs["ref"]["/ux_non_staggered_c"]["data"].reshape(-1, 1024, 2, 2).astype(np.float64)
    for fr in range(a.shape[0]):
        print(f"   frame {fr}: h1 rel {rel(a[fr,:,0], b[fr,:,0]):.3e}  h2 rel {rel(a[fr,:,1], b[fr,:,1]):.3e}  |h1| {np.abs(b[fr,:,0]).max():.3e} |h2| {np.abs(b[fr,:,1]).max():.3e}")
