"""Micro-benchmark of the column passes (y vs z access pattern, plain vs fused z pass).
usage: bench_col.py [nx,ny,nz ...]   (default: 256^3 and 512^3)"""
import ctypes as C, importlib, sys
sys.path.insert(0, ".")
kw = importlib.import_module("k-wave-fluid-cuda_b200")
lib = kw.load_library()
lib.kw_bench_col.argtypes = [C.c_uint64] * 3 + [C.c_int] * 3 + [C.POINTER(C.c_float)]
shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(256, 256, 256), (512, 512, 512)]
for nx, ny, nz in shapes:
    for axis, fused, name in ((1, 0, "y plain"), (2, 0, "z plain"), (2, 1, "z fused fwd*inv")):
        ms = C.c_float()
        rc = lib.kw_bench_col(nx, ny, nz, axis, fused, 20, C.byref(ms))
        nc = (nx // 2 + 1 + 15) // 16 * 16 * ny * nz
        gb = nc * (16 + (4 if fused else 0)) / 1e9
        print(f"{nx}x{ny}x{nz} {name:16s} rc={rc} {ms.value:.4f} ms/pass  {gb/ms.value*1e3:.0f} GB/s")
