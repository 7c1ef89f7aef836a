"""Synthetic k-Wave inputs in the reference's input-file schema (main.cpp:446-563).

The arrays returned here are exactly what ``MatrixContainer::loadDataFromInputFile`` would load
(Containers/MatrixContainer.cpp:94-410): file conventions, i.e. 1-based indices, unsquared ``c0``, undivided
``rho0_sg*``.  Both the CUDA path and the oracle consume the same dictionaries, so the generator only needs to be
self-consistent (k-Wave formulas for the k-space vectors and the PML are followed, MATLAB bit-identity is not needed).
NumPy/SciPy only; no GPU code.
"""
from __future__ import annotations

import numpy as np
import scipy.fft as sfft

F32 = np.float32


def kvec(n, d):
    """ifftshift'ed wavenumber vector of k-Wave's kgrid for even/odd n."""
    if n % 2 == 0:
        nn = np.arange(-n // 2, n // 2) / n
    else:
        nn = np.arange(-(n - 1) // 2, (n - 1) // 2 + 1) / n
    return np.fft.ifftshift(2 * np.pi / d * nn)


def shift_ops(n, d):
    k = kvec(n, d)
    pos = 1j * k * np.exp(1j * k * d / 2)
    neg = 1j * k * np.exp(-1j * k * d / 2)
    sh = np.exp(-1j * k * d / 2)
    return pos.astype(np.complex64), neg.astype(np.complex64), sh.astype(np.complex64)


def pml_vec(n, d, dt, c_ref, size, alpha, staggered):
    x = np.arange(1, size + 1, dtype=np.float64)
    if staggered:
        x = x + 0.5
    left = alpha * (c_ref / d) * ((x - size - 1) / (0 - size)) ** 4
    right = alpha * (c_ref / d) * (x / size) ** 4
    v = np.ones(n)
    v[:size] = np.exp(-left * dt / 2)
    v[n - size :] = np.exp(-right * dt / 2)
    return v.astype(F32)


def smooth_noise(shape, seed, sigma=8.0):
    """Gaussian white noise low-passed by a sigma-voxel Gaussian, normalised to unit max-abs (FP32)."""
    rng = np.random.default_rng(seed)
    w = rng.standard_normal(shape, dtype=F32)
    wk = sfft.rfftn(w, workers=-1)
    for ax, n in enumerate(shape):
        f = np.fft.rfftfreq(n) if ax == len(shape) - 1 else np.fft.fftfreq(n)
        g = np.exp(-2 * (np.pi * f * sigma) ** 2).astype(F32)
        sh = [1] * len(shape)
        sh[ax] = -1
        wk *= g.reshape(sh)
    w = sfft.irfftn(wk, s=shape, workers=-1).astype(F32)
    w /= np.abs(w).max()
    return w


def wave_field(shape, seed, z_range=None):
    """Analytic smooth heterogeneity in [-1, 1]: mean of six products of low-frequency cosines with seeded phases.
    Unlike smooth_noise it can be evaluated for any z-range on its own (slab-wise generation of 1024^3 inputs)."""
    nz, ny, nx = shape
    z0, nzl = (0, nz) if z_range is None else z_range
    rng = np.random.default_rng(seed)
    out = np.zeros((nzl, ny, nx), dtype=F32)
    for _ in range(6):
        f = rng.integers(1, 5, size=3)
        ph = rng.uniform(0, 2 * np.pi, size=3)
        cz = np.cos(2 * np.pi * f[0] * (np.arange(z0, z0 + nzl) / nz) + ph[0]).astype(F32)
        cy = np.cos(2 * np.pi * f[1] * (np.arange(ny) / ny) + ph[1]).astype(F32)
        cx = np.cos(2 * np.pi * f[2] * (np.arange(nx) / nx) + ph[2]).astype(F32)
        out += cz[:, None, None] * cy[None, :, None] * cx[None, None, :]
    out *= F32(1.0 / 6.0)
    return out


def make_case(
    nx,
    ny=None,
    nz=None,
    *,
    nt=100,
    nonlinear=True,
    absorbing=True,
    heterogeneous=True,
    source="p_plane",  # p_plane | p_many | p0 | u_plane | transducer | none
    source_mode=1,
    sensor="index",  # index | cuboid | full_cuboid
    n_sensor=4096,
    pml_size=None,
    seed=1234,
    period=50,
    shuffle_sensor=False,
    shifts=False,
    medium="noise",  # noise: low-passed white noise; waves: analytic, can be generated slab by slab
    z_range=None,  # (z0, nzl): full-grid arrays hold these planes only (medium="waves"); everything else stays global
):
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    shape = (nz, ny, nx)
    if z_range is not None and medium != "waves":
        raise ValueError("slab-wise generation needs medium='waves'")
    zs0, zsl = (0, nz) if z_range is None else z_range
    n = nx * ny * nz
    dx = dy = dz = 1e-4
    two_d = nz == 1  # 2-D simulation: the file carries no z arrays (main.cpp:446-563)
    ext = min(shape[1:]) if two_d else min(shape)
    pml_size = pml_size if pml_size is not None else (20 if ext >= 128 else max(2, ext // 8))
    sig = 8.0 if ext >= 64 else max(1.5, ext / 8)
    arrays = {}
    if heterogeneous and medium == "waves":
        zr = (zs0, zsl)
        c0 = (1500.0 * (1.0 + 0.05 * wave_field(shape, seed, zr))).astype(F32)
        zz, yy, xx = np.ogrid[zs0 : zs0 + zsl, :ny, :nx]
        r = min(shape) // 8
        c0[(xx - nx // 2) ** 2 + (yy - ny // 2) ** 2 + (zz - nz // 2) ** 2 <= r * r] = 1600.0
        # one extra plane for the z-staggered density (the last plane of the grid copies itself)
        ext = wave_field(shape, seed + 1, (zs0, min(zsl + 1, nz - zs0)))
        rho_ext = (1000.0 * (1.0 + 0.05 * ext)).astype(F32)
        if rho_ext.shape[0] == zsl:
            rho_ext = np.concatenate([rho_ext, rho_ext[-1:]], axis=0)
        rho0 = np.ascontiguousarray(rho_ext[:zsl])
        sg = []
        for ax in (2, 1):
            nb = np.concatenate([np.take(rho0, range(1, shape[ax]), axis=ax), np.take(rho0, [-1], axis=ax)], axis=ax)
            sg.append(((rho0 + nb) * F32(0.5)).astype(F32))
        sg.append(((rho0 + rho_ext[1:]) * F32(0.5)).astype(F32))
        del rho_ext, ext
        arrays.update(c0=c0, rho0=rho0, rho0_sgx=sg[0], rho0_sgy=sg[1], rho0_sgz=sg[2])
        if nonlinear:
            arrays["BonA"] = (6.0 + wave_field(shape, seed + 2, zr)).astype(F32)
        if absorbing:
            arrays["alpha_coeff"] = (0.75 + 0.25 * wave_field(shape, seed + 3, zr)).astype(F32)
        c_ref = 1600.0  # global bound of c0 (1500 * 1.05 < 1600 = the ball), known without seeing the other slabs
    elif heterogeneous:
        g0, g1 = smooth_noise(shape, seed, sig), smooth_noise(shape, seed + 1, sig)
        c0 = (1500.0 * (1.0 + 0.05 * g0)).astype(F32)
        zz, yy, xx = np.ogrid[:nz, :ny, :nx]
        r = ext // 8
        ball = (xx - nx // 2) ** 2 + (yy - ny // 2) ** 2 + (zz - nz // 2) ** 2 <= r * r
        c0[ball] = 1600.0
        rho0 = (1000.0 * (1.0 + 0.05 * g1)).astype(F32)
        sg = []
        for ax in (2, 1, 0):
            nb = np.concatenate([np.take(rho0, range(1, shape[ax]), axis=ax), np.take(rho0, [-1], axis=ax)], axis=ax)
            sg.append(((rho0 + nb) * F32(0.5)).astype(F32))
        arrays.update(c0=c0, rho0=rho0, rho0_sgx=sg[0], rho0_sgy=sg[1], rho0_sgz=sg[2])
        if nonlinear:
            arrays["BonA"] = (6.0 + smooth_noise(shape, seed + 2, sig)).astype(F32)
        if absorbing:
            arrays["alpha_coeff"] = (0.75 + 0.25 * smooth_noise(shape, seed + 3, sig)).astype(F32)
        c_ref = float(c0.max())
    else:
        one = lambda v: np.array([v], dtype=F32)  # noqa: E731
        arrays.update(c0=one(1500.0), rho0=one(1000.0), rho0_sgx=one(1000.0), rho0_sgy=one(1000.0), rho0_sgz=one(1000.0))
        if nonlinear:
            arrays["BonA"] = one(6.0)
        if absorbing:
            arrays["alpha_coeff"] = one(0.75)
        c_ref = 1500.0
    dt = float(F32(0.3 * dx / c_ref))
    cfg = dict(
        Nx=nx, Ny=ny, Nz=nz, Nt=nt, dt=dt, dx=dx, dy=dy, dz=dz, c_ref=float(F32(c_ref)),
        pml_x_size=pml_size, pml_y_size=pml_size, pml_z_size=pml_size,
        pml_x_alpha=2.0, pml_y_alpha=2.0, pml_z_alpha=2.0,
        nonlinear_flag=int(nonlinear), absorbing_flag=int(absorbing), alpha_power=1.5, nonuniform_grid_flag=0,
        p_source_flag=0, p_source_mode=0, p_source_many=0, p0_source_flag=0, transducer_source_flag=0,
        ux_source_flag=0, uy_source_flag=0, uz_source_flag=0, u_source_mode=0, u_source_many=0,
        sensor_mask_type=0 if sensor == "index" else 1,
    )  # fmt: skip
    # k-space vectors
    px, nxg, sx = shift_ops(nx, dx)
    py, nyg, sy = shift_ops(ny, dy)
    pz, nzg, sz = shift_ops(nz, dz)
    arrays.update(
        ddx_k_shift_pos_r=px[: nx // 2 + 1], ddy_k_shift_pos=py, ddz_k_shift_pos=pz,
        ddx_k_shift_neg_r=nxg[: nx // 2 + 1], ddy_k_shift_neg=nyg, ddz_k_shift_neg=nzg,
    )  # fmt: skip
    if shifts:
        arrays.update(x_shift_neg_r=sx[: nx // 2 + 1], y_shift_neg_r=sy[: ny // 2 + 1], z_shift_neg_r=sz[: nz // 2 + 1])
    for a, nn, d in (("x", nx, dx), ("y", ny, dy), ("z", nz, dz)):
        if two_d and a == "z":
            continue
        arrays[f"pml_{a}_sg{a}"] = pml_vec(nn, d, dt, c_ref, pml_size, 2.0, True)
        arrays[f"pml_{a}"] = pml_vec(nn, d, dt, c_ref, pml_size, 2.0, False)
    if two_d:  # what a 2-D input file does not contain
        for k in ("ddz_k_shift_pos", "ddz_k_shift_neg", "z_shift_neg_r", "rho0_sgz"):
            arrays.pop(k, None)
        for k in ("pml_z_size", "pml_z_alpha", "uz_source_flag"):  # dz stays in cfg for the callers' convenience; nothing reads it
            cfg.pop(k, None)
    # sources
    tt = np.arange(nt, dtype=np.float64)
    tone = np.sin(2 * np.pi * tt / period)
    ramp = np.minimum(1.0, tt / (2.0 * period))
    xs = pml_size + 2
    zz, yy = np.meshgrid(np.arange(nz), np.arange(ny), indexing="ij")
    plane = ((zz * ny + yy) * nx + xs).reshape(-1).astype(np.uint64) + 1  # 1-based
    if source in ("p_plane", "p_many"):
        cfg.update(p_source_flag=nt, p_source_mode=source_mode, p_source_many=int(source == "p_many"))
        arrays["p_source_index"] = plane
        amp = 1.0e5 if source_mode == 0 else 2.0e4
        # pressure sources are scaled to density units by k-Wave before being written (p / (3 c0^2) style factors);
        # here the signal is used as-is by both implementations, amplitude chosen to give O(1e5 Pa) fields.
        sig_t = (amp / (3.0 * 1500.0**2) * tone * ramp).astype(F32)
        if source == "p_many":
            apod = (0.5 + 0.5 * np.cos(np.linspace(-np.pi, np.pi, plane.size))).astype(F32)
            arrays["p_source_input"] = (sig_t[:, None] * apod[None, :]).reshape(-1)  # time-major: t*Nsrc + i
        else:
            arrays["p_source_input"] = sig_t
    elif source == "u_plane":
        cfg.update(ux_source_flag=nt, u_source_mode=source_mode, u_source_many=0)
        arrays["u_source_index"] = plane
        arrays["ux_source_input"] = (0.05 * tone * ramp).astype(F32)
    elif source == "transducer":
        cfg.update(transducer_source_flag=nt)
        arrays["u_source_index"] = plane
        delay = ((yy + zz) % 8).reshape(-1).astype(np.uint64)
        arrays["delay_mask"] = delay + 1  # 1-based in the file (MatrixContainer.cpp:224-225 loads it as an index matrix)
        tt2 = np.arange(nt + 8, dtype=np.float64)
        arrays["transducer_source_input"] = (0.05 * np.sin(2 * np.pi * tt2 / period)).astype(F32)
    elif source == "p0":
        cfg.update(p0_source_flag=1)
        zz3, yy3, xx3 = np.ogrid[zs0 : zs0 + zsl, :ny, :nx]
        s2 = 2.0 * (max(2.0, min(shape) / 32.0)) ** 2
        r2 = (xx3 - nx // 2) ** 2 + (yy3 - ny // 2) ** 2 + (zz3 - nz // 2) ** 2
        arrays["p0_source_input"] = (1.0e5 * np.exp(-r2 / s2)).astype(F32)
    # sensors
    if sensor == "index":
        zpl = nz // 2
        total = nx * ny
        stride = max(1, total // n_sensor)
        idx = (zpl * nx * ny + np.arange(0, total, stride)[:n_sensor]).astype(np.uint64)
        if shuffle_sensor:
            np.random.default_rng(7).shuffle(idx)
        arrays["sensor_mask_index"] = idx + 1
    elif sensor == "cuboid":
        a0, a1 = int(0.3 * nx), int(0.7 * nx) - 1
        b0, b1 = max(1, nx // 12), max(2, nx // 6)
        zb0, zb1 = (0, 0) if two_d else (b0, nz - b0 - 1)
        arrays["sensor_mask_corners"] = np.array(
            [[a0, a0 * ny // nx, a0 * nz // nx, a1, a1 * ny // nx, a1 * nz // nx], [b0, b0, zb0, b1, b1, zb1]],
            dtype=np.uint64,
        ) + 1
    elif sensor == "full_cuboid":
        arrays["sensor_mask_corners"] = np.array([[0, 0, 0, nx - 1, ny - 1, nz - 1]], dtype=np.uint64) + 1
    return cfg, arrays
