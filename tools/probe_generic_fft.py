import sys, importlib, numpy as np
sys.path.insert(0, ".")
kw = importlib.import_module("k-wave-fluid-cuda_b200")
x = np.random.default_rng(0).standard_normal((128, 480, 480)).astype(np.float32)
k = kw.fft_r2c_3d(x)
y = kw.fft_c2r_3d(k, 480)
print(float(np.abs(y / x.size - x).max()))
