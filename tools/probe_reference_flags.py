"""Run the reference's own binary (oracle/_ref/ref_kspace) on a small synthetic input with the given flags and print its
exit code and messages: used to document flag / mask combinations on which the reference itself fails."""
import importlib, os, subprocess, sys, tempfile
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import kwh5
kw = importlib.import_module("k-wave-fluid-cuda_b200")
tmp = tempfile.mkdtemp()
sensor = sys.argv[1]
cfg, arrays = kw.synth.make_case(32, nt=60, nonlinear=True, absorbing=True, source="p_plane", sensor=sensor, n_sensor=64, period=20, shifts=True)
fin = os.path.join(tmp, "in.h5"); kwh5.write_input(fin, cfg, arrays)
for flags in [f.split() for f in sys.argv[2:]]:
    r = subprocess.run(["oracle/_ref/ref_kspace", "-i", fin, "-o", os.path.join(tmp, "o.h5"), "--verbose", "0"] + flags, capture_output=True, text=True)
    print("FLAGS", flags, "rc", r.returncode, "|", (r.stderr.strip().splitlines() or [""])[-4:], (r.stdout.strip().splitlines() or [""])[-2:])
