"""Micro-benchmark of the column passes (y vs z access pattern, plain vs fused z pass)."""
import ctypes as C, importlib, sys
sys.path.insert(0, ".")
kw = importlib.import_module("k-wave-fluid-cuda_b200")
lib = kw.load_library()
lib.kw_bench_col.argtypes = [C.c_uint64] * 3 + [C.c_int] * 3 + [C.POINTER(C.c_float)]
for n in (256, 512):
    for axis, fused, name in ((1, 0, "y plain"), (2, 0, "z plain"), (2, 1, "z fused fwd*inv")):
        ms = C.c_float()
        rc = lib.kw_bench_col(n, n, n, axis, fused, 20, C.byref(ms))
        nc = (n // 2 + 1 + 15) // 16 * 16 * n * n
        gb = nc * (16 + (4 if fused else 0)) / 1e9
        print(f"N={n} {name:16s} rc={rc} {ms.value:.4f} ms/pass  {gb/ms.value*1e3:.0f} GB/s")
