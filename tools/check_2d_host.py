import importlib, numpy as np, sys, os, subprocess, tempfile
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import kwh5
kw = importlib.import_module("k-wave-fluid-cuda_b200")
tmp = tempfile.mkdtemp()
cfg, arrays = kw.synth.make_case(32, 64, 1, nt=120, nonlinear=True, absorbing=True, source="p_plane", n_sensor=80, period=20, shifts=True)
fin = os.path.join(tmp, "in.h5"); kwh5.write_input(fin, cfg, arrays)
flags = sys.argv[1:] or ["-p", "--p_c", "--I_avg_c", "--u_non_staggered_raw", "--period", "20", "--harmonics", "2"]
outs = {}
for name, binary in (("ours", "k-wave-fluid-cuda_b200/kspaceFirstOrder-B200"), ("ref", "oracle/_ref/ref_kspace")):
    fo = os.path.join(tmp, name + ".h5")
    r = subprocess.run([binary, "-i", fin, "-o", fo, "--verbose", "0"] + flags, capture_output=True, text=True)
    print(name, "rc", r.returncode, r.stderr[-300:])
    outs[name] = kwh5.read_output(fo)
for k in sorted(outs["ours"]):
    a = outs["ours"][k]; b = outs["ref"].get(k)
    if a.size < 2: continue
    err = None if b is None else float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))
    print(k, a.shape, "rel-L2", err, "max ours", float(np.nanmax(np.abs(a))), "max ref", None if b is None else float(np.nanmax(np.abs(b))))
    if k == "p_c" and b is not None:
        fa, fb = a.reshape(a.shape[1], -1), b.reshape(b.shape[1], -1)
        print("   per-frame rel-L2", [float(np.linalg.norm(fa[i] - fb[i]) / np.linalg.norm(fb[i])) for i in range(fa.shape[0])])
