"""Helpers shared by the CPU and GPU parity tests: the reference-run fixtures under tests/golden/ref_*.npz
(produced by oracle/make_ref_goldens.py from the reference's own cuFFT build on a B200)."""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# quantities that are pure rounding noise in a given case are not compared (plane source along x: uy, uz ~ 0 at the sensors)
SKIP = {"uy", "uz"}


def fixture_names():
    return sorted(f[4:-4] for f in os.listdir(GOLD) if f.startswith("ref_") and f.endswith(".npz"))


def load_fixture(name):
    g = np.load(os.path.join(GOLD, f"ref_{name}.npz"))
    kwargs = json.loads(str(g["make_case"]))
    shape = kwargs.pop("shape")
    data = {k.replace("__", "/"): np.asarray(g[k]) for k in g.files if k not in ("make_case", "nt", "flags")}
    return shape, kwargs, int(g["nt"]), json.loads(str(g["flags"])), data


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def time_series_views(series, data, nt):
    """Map reference datasets of raw series onto a (steps, Nsens) array sampled in mask order: index masks give one
    dataset (1, Nt, Nsens); cuboid masks give one 4-D dataset per cuboid, concatenated in the series columns."""
    out = {}
    if "p" in data:
        out["p"] = (series, data["p"].reshape(nt, -1))
    else:
        col = 0
        k = 1
        while f"p/{k}" in data:
            r = data[f"p/{k}"].reshape(nt, -1)
            out[f"p/{k}"] = (series[:, col : col + r.shape[1]], r)
            col += r.shape[1]
            k += 1
    return out
