"""CPU restatement (NumPy/SciPy) of the per-timestep hot path of kspaceFirstOrder-CUDA.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it,
and there only as the checker or the timed CPU baseline.  The product path is the CUDA library behind
``include/kwave_b200.h`` and fails loudly without it.

Pinning status (see DESIGN.md "Oracle"):
  * compression bases + 40-bit codec: pinned against the reference's own ``Compression/CompressHelper.cpp`` compiled
    on the CPU (``oracle/Makefile`` -> ``oracle/_ref/compress_ref``; fixtures in ``tests/golden/compress_*.json``).
  * solver step: pinned against the reference's own solver sources (cuFFT build, ``oracle/_ref/ref_kspace``) run on
    a B200 through ``gpurun``; fixtures in ``tests/golden/ref_*.npz`` with the generating script
    ``oracle/make_ref_goldens.sh``.  Until those fixtures exist for a case, that case is "parity unpinned".

All citations are file:line under the reference tree (klepo/k-Wave-Fluid-CUDA).

Array convention: every 3-D field is a NumPy array of shape (Nz, Ny, Nx), C-order, i.e. x fastest
(``i = (z*Ny + y)*Nx + x``, KSpaceSolver/KSpaceFirstOrderSolver.cpp:2916-2920); half spectra are
``rfftn`` outputs of shape (Nz, Ny, Nx/2+1), which is cuFFT's R2C layout (MatrixClasses/CufftComplexMatrix.cpp:87-91).
"""
from __future__ import annotations

import numpy as np
import scipy.fft as sfft

F32 = np.float32


# ---------------------------------------------------------------------------------------------------------------------
# pre-processing (host, once)  --  KSpaceFirstOrderSolver.cpp:784-857
# ---------------------------------------------------------------------------------------------------------------------
def _k_parts(n, d2rec):
    """0.5 - |0.5 - i/N| folding, squared, times 1/d^2 (KSpaceFirstOrderSolver.cpp:2424-2444), all FP32."""
    i = np.arange(n, dtype=F32)
    nrec = F32(1.0) / F32(n)
    part = F32(0.5) - np.abs(F32(0.5) - i * nrec)
    return (part * part) * F32(d2rec)


def generate_kappa(cfg):
    """kappa only (lossless case).  KSpaceFirstOrderSolver.cpp:2404-2452."""
    nx, ny, nz = cfg["Nx"], cfg["Ny"], cfg["Nz"]
    dx2 = F32(1.0) / (F32(cfg["dx"]) * F32(cfg["dx"]))
    dy2 = F32(1.0) / (F32(cfg["dy"]) * F32(cfg["dy"]))
    dz2 = F32(1.0) / (F32(cfg["dz"]) * F32(cfg["dz"]))
    c_ref_dt_pi = F32(cfg["c_ref"]) * F32(cfg["dt"]) * F32(np.pi)
    xp = _k_parts(nx, dx2)[: nx // 2 + 1]
    yp = _k_parts(ny, dy2)
    zp = _k_parts(nz, dz2)
    yz = zp[:, None] + yp[None, :]
    k = c_ref_dt_pi * np.sqrt(xp[None, None, :] + yz[:, :, None], dtype=F32)
    with np.errstate(invalid="ignore", divide="ignore"):
        kappa = np.where(k == 0, F32(1.0), np.sin(k, dtype=F32) / k)
    return kappa.astype(F32)


def generate_source_kappa(cfg):
    """cos(c_ref*dt*pi*|k|).  KSpaceFirstOrderSolver.cpp:2460-2506."""
    nx, ny, nz = cfg["Nx"], cfg["Ny"], cfg["Nz"]
    dx2 = F32(1.0) / (F32(cfg["dx"]) * F32(cfg["dx"]))
    dy2 = F32(1.0) / (F32(cfg["dy"]) * F32(cfg["dy"]))
    dz2 = F32(1.0) / (F32(cfg["dz"]) * F32(cfg["dz"]))
    c_ref_dt_pi = F32(cfg["c_ref"]) * F32(cfg["dt"]) * F32(np.pi)
    xp = _k_parts(nx, dx2)[: nx // 2 + 1]
    yp = _k_parts(ny, dy2)
    zp = _k_parts(nz, dz2)
    yz = zp[:, None] + yp[None, :]
    k = c_ref_dt_pi * np.sqrt(xp[None, None, :] + yz[:, :, None], dtype=F32)
    return np.cos(k, dtype=F32)


def generate_kappa_and_nablas(cfg):
    """kappa, absorb_nabla1, absorb_nabla2.  KSpaceFirstOrderSolver.cpp:2514-2577."""
    nx, ny, nz = cfg["Nx"], cfg["Ny"], cfg["Nz"]
    dx2 = F32(1.0) / (F32(cfg["dx"]) * F32(cfg["dx"]))
    dy2 = F32(1.0) / (F32(cfg["dy"]) * F32(cfg["dy"]))
    dz2 = F32(1.0) / (F32(cfg["dz"]) * F32(cfg["dz"]))
    c_ref_dt2 = F32(cfg["c_ref"]) * F32(cfg["dt"]) * F32(0.5)
    pi2 = F32(np.pi) * F32(2.0)
    y = F32(cfg["alpha_power"])
    xp = _k_parts(nx, dx2)[: nx // 2 + 1]
    yp = _k_parts(ny, dy2)
    zp = _k_parts(nz, dz2)
    yz = zp[:, None] + yp[None, :]
    k = pi2 * np.sqrt(xp[None, None, :] + yz[:, :, None], dtype=F32)
    crefk = c_ref_dt2 * k
    with np.errstate(invalid="ignore", divide="ignore"):
        kappa = np.where(crefk == 0, F32(1.0), np.sin(crefk, dtype=F32) / crefk).astype(F32)
        n1 = np.power(k, y - F32(2.0), dtype=F32)
        n2 = np.power(k, y - F32(1.0), dtype=F32)
    n1[np.isposinf(n1)] = 0
    n2[np.isposinf(n2)] = 0
    return kappa, n1.astype(F32), n2.astype(F32)


def generate_tau_eta(cfg, c0, alpha_coeff):
    """absorb_tau, absorb_eta (scalar or matrix).  KSpaceFirstOrderSolver.cpp:2584-2643.  c0 is the *unsquared* c0."""
    y = F32(cfg["alpha_power"])
    tan_term = np.tan(F32(np.pi / 2) * y, dtype=F32)
    alpha_neper = (F32(100.0) * np.power(F32(1.0e-6) / (F32(2.0) * F32(np.pi)), y, dtype=F32)) / (
        F32(20.0) * F32(np.log10(np.e))
    )
    a2 = F32(2.0) * alpha_neper * np.asarray(alpha_coeff, dtype=F32)
    c0 = np.asarray(c0, dtype=F32)
    tau = (-a2) * np.power(c0, y - F32(1.0), dtype=F32)
    eta = a2 * np.power(c0, y, dtype=F32) * tan_term
    return tau.astype(F32), eta.astype(F32)


# ---------------------------------------------------------------------------------------------------------------------
# the solver
# ---------------------------------------------------------------------------------------------------------------------
class KSpaceOracle:
    """One object == one simulation.  ``arrays`` holds what MatrixContainer loads from the input file
    (Containers/MatrixContainer.cpp:94-410) with the file's own conventions: indices 1-based, c0 unsquared,
    rho0_sg* undivided.  Scalars may be given as 0-d / 1-element arrays (homogeneous medium,
    Parameters/Parameters.cpp:426-459)."""

    def __init__(self, cfg, arrays, dtype=np.float64):
        self.cfg = dict(cfg)
        self.dt = np.dtype(dtype)
        self.cdt = np.complex128 if self.dt == np.float64 else np.complex64
        c = self.cfg
        self.nx, self.ny, self.nz = c["Nx"], c["Ny"], c["Nz"]
        self.shape = (self.nz, self.ny, self.nx)
        self.n = self.nx * self.ny * self.nz
        self.t = 0
        a = {k: np.asarray(v) for k, v in arrays.items()}
        T = self.dt.type
        if self.nz == 1:
            # 2-D simulation (Parameters.h:88-94): the input file has no z arrays; every z term of the 3-D formulas is dropped
            # by the reference's k2D branches (e.g. SolverCudaKernels.cu:197,1147).  Neutral z operators reproduce that exactly:
            # the z gradient is 0, u_z and rho_z stay 0 and add exact zeros to the sums.
            a.setdefault("ddz_k_shift_pos", np.zeros(1, np.complex64))
            a.setdefault("ddz_k_shift_neg", np.zeros(1, np.complex64))
            a.setdefault("pml_z", np.ones(1, F32))
            a.setdefault("pml_z_sgz", np.ones(1, F32))
            a.setdefault("rho0_sgz", np.ones(1, F32))
            if "x_shift_neg_r" in a:
                a.setdefault("z_shift_neg_r", np.ones(1, np.complex64))

        def field(x):
            x = np.asarray(x, dtype=F32)
            return x.reshape(self.shape).astype(self.dt) if x.size == self.n else T(x.reshape(-1)[0])

        # --- A.1 step 1: index shift (KSpaceFirstOrderSolver.cpp:787-813)
        self.sensor_index = a["sensor_mask_index"].astype(np.int64).reshape(-1) - 1 if "sensor_mask_index" in a else None
        self.sensor_corners = (
            a["sensor_mask_corners"].astype(np.int64).reshape(-1, 6) - 1 if "sensor_mask_corners" in a else None
        )
        self.p_src_idx = a["p_source_index"].astype(np.int64).reshape(-1) - 1 if "p_source_index" in a else None
        self.u_src_idx = a["u_source_index"].astype(np.int64).reshape(-1) - 1 if "u_source_index" in a else None
        self.delay_mask = a["delay_mask"].astype(np.int64).reshape(-1) - 1 if "delay_mask" in a else None
        # --- step 2: dt / rho0_sg (KSpaceFirstOrderSolver.cpp:825-830; BaseFloatMatrix.cpp:86-93) in FP32
        dt32 = F32(c["dt"])
        self.dtrho = [field(dt32 / np.asarray(a[k], dtype=F32)) for k in ("rho0_sgx", "rho0_sgy", "rho0_sgz")]
        # --- step 3: k-space operators (FP32 on the host)
        c0 = np.asarray(a["c0"], dtype=F32)
        if c["absorbing_flag"]:
            kappa, n1, n2 = generate_kappa_and_nablas(c)
            tau, eta = generate_tau_eta(c, c0, a["alpha_coeff"])
            self.nabla1, self.nabla2 = n1.astype(self.dt), n2.astype(self.dt)
            self.tau, self.eta = field(tau), field(eta)
        else:
            kappa = generate_kappa(c)
        self.kappa = kappa.astype(self.dt)
        # --- step 4: source kappa for additive sources (KSpaceFirstOrderSolver.cpp:846-852)
        need_sk = (c.get("p_source_flag", 0) and c.get("p_source_mode", 0) == 2) or (
            (c.get("ux_source_flag", 0) or c.get("uy_source_flag", 0) or c.get("uz_source_flag", 0))
            and c.get("u_source_mode", 0) == 2
        )
        self.source_kappa = generate_source_kappa(c).astype(self.dt) if need_sk else None
        # --- step 5: c2 = c0^2 (KSpaceFirstOrderSolver.cpp:2690-2703) in FP32
        self.c2 = field(c0 * c0)
        self.rho0 = field(a["rho0"])
        self.bona = field(a["BonA"]) if c["nonlinear_flag"] else None
        # 1-D operators
        cx = lambda k: np.asarray(a[k]).reshape(-1).astype(self.cdt)  # noqa: E731
        self.ddx_pos, self.ddy_pos, self.ddz_pos = cx("ddx_k_shift_pos_r"), cx("ddy_k_shift_pos"), cx("ddz_k_shift_pos")
        self.ddx_neg, self.ddy_neg, self.ddz_neg = cx("ddx_k_shift_neg_r"), cx("ddy_k_shift_neg"), cx("ddz_k_shift_neg")
        rv = lambda k: np.asarray(a[k], dtype=F32).reshape(-1).astype(self.dt)  # noqa: E731
        self.pml_sg = [rv("pml_x_sgx"), rv("pml_y_sgy"), rv("pml_z_sgz")]
        self.pml = [rv("pml_x"), rv("pml_y"), rv("pml_z")]
        self.shift_neg = (
            [cx("x_shift_neg_r"), cx("y_shift_neg_r"), cx("z_shift_neg_r")] if "x_shift_neg_r" in a else None
        )
        # sources
        g = lambda k: np.asarray(a[k], dtype=F32).reshape(-1).astype(self.dt) if k in a else None  # noqa: E731
        self.p_src_in = g("p_source_input")
        self.u_src_in = [g("ux_source_input"), g("uy_source_input"), g("uz_source_input")]
        self.transducer_in = g("transducer_source_input")
        self.p0 = field(a["p0_source_input"]) if c.get("p0_source_flag", 0) else None
        # state (starts at zero, BaseFloatMatrix.cpp:144-145)
        z = lambda: np.zeros(self.shape, dtype=self.dt)  # noqa: E731
        self.p = z()
        self.u = [z(), z(), z()]
        self.rho = [z(), z(), z()]
        self.fd = T(F32(1.0) / F32(self.n))  # fftDivider, CudaParameters.cpp:259

    # -- helpers ------------------------------------------------------------------------------------------------------
    def _fft(self, x):
        return sfft.rfftn(x, axes=(0, 1, 2)).astype(self.cdt, copy=False)

    def _ifft(self, xk):
        # unnormalised C2R (cuFFT semantics)
        return (sfft.irfftn(xk, s=self.shape, axes=(0, 1, 2)) * self.dt.type(self.n)).astype(self.dt, copy=False)

    def _bx(self, v):
        return v[None, None, :]

    def _by(self, v):
        return v[None, :, None]

    def _bz(self, v):
        return v[:, None, None]

    def _pressure_gradient(self):
        """F[p]*kappa * ddk_pos -> three unnormalised inverse transforms.  SolverCudaKernels.cu:1139-1157."""
        e = self._fft(self.p) * self.kappa
        return (
            self._ifft(e * self._bx(self.ddx_pos)),
            self._ifft(e * self._by(self.ddy_pos)),
            self._ifft(e * self._bz(self.ddz_pos)),
        )

    def _scale_source(self, values, idx):
        """Additive (k-space corrected) source.  KSpaceFirstOrderSolver.cpp:2339-2352; SolverCudaKernels.cu:679-745."""
        tmp = np.zeros(self.n, dtype=self.dt)
        tmp[idx] = values
        tk = self._fft(tmp.reshape(self.shape)) * (self.source_kappa * self.fd)
        return self._ifft(tk)

    def _src_values(self, src_in, nsrc, many):
        # SolverCudaKernels.cu:509-511,577-579: index2D = t (single) or t*Nsrc (many, time-major)
        t = self.t
        return src_in[t * nsrc : (t + 1) * nsrc] if many else np.full(nsrc, src_in[t], dtype=self.dt)

    # -- one time step  (KSpaceFirstOrderSolver.cpp:885-935) ----------------------------------------------------------
    def step(self):
        c, t, fd = self.cfg, self.t, self.fd
        bc = (self._bx, self._by, self._bz)
        # computeVelocity (cpp:2087-2119; SolverCudaKernels.cu:199-212)
        g = self._pressure_gradient()
        for i in range(3):
            pml = bc[i](self.pml_sg[i])
            self.u[i] = (self.u[i] * pml - (fd * g[i]) * self.dtrho[i]) * pml
        # addVelocitySource (cpp:2252-2303; SolverCudaKernels.cu:504-527)
        umode, umany = c.get("u_source_mode", 0), c.get("u_source_many", 0)
        for i, key in enumerate(("ux_source_flag", "uy_source_flag", "uz_source_flag")):
            if c.get(key, 0) > t:
                s = self._src_values(self.u_src_in[i], self.u_src_idx.size, umany)
                flat = self.u[i].reshape(-1)
                if umode == 0:
                    flat[self.u_src_idx] = s
                elif umode == 1:
                    flat[self.u_src_idx] += s
                else:
                    self.u[i] = self.u[i] + self._scale_source(s, self.u_src_idx)
        # transducer (cpp:894-897; SolverCudaKernels.cu:463-471)
        if c.get("transducer_source_flag", 0) > t:
            self.u[0].reshape(-1)[self.u_src_idx] += self.transducer_in[self.delay_mask + t]
        # computeVelocityGradient (cpp:2126-2150; SolverCudaKernels.cu:1210-1239)
        kfd = self.kappa * fd
        neg = (self._bx(self.ddx_neg), self._by(self.ddy_neg), self._bz(self.ddz_neg))
        du = [self._ifft((self._fft(self.u[i]) * kfd) * neg[i]) for i in range(3)]
        # computeDensity (SolverCudaKernels.cu:1358-1393 nonlinear, :1470-1497 linear)
        dtT = self.dt.type(F32(c["dt"]))
        if c["nonlinear_flag"]:
            s = (2 * (self.rho[0] + self.rho[1] + self.rho[2]) + self.rho0) * dtT
            for i in range(3):
                pml = bc[i](self.pml[i])
                self.rho[i] = pml * ((pml * self.rho[i]) - s * du[i])
        else:
            dtrho0 = dtT * self.rho0
            for i in range(3):
                pml = bc[i](self.pml[i])
                self.rho[i] = pml * (pml * self.rho[i] - dtrho0 * du[i])
        # addPressureSource (cpp:2310-2334; SolverCudaKernels.cu:570-629, :795-807)
        if c.get("p_source_flag", 0) > t:
            pmode, pmany = c.get("p_source_mode", 0), c.get("p_source_many", 0)
            s = self._src_values(self.p_src_in, self.p_src_idx.size, pmany)
            ndim = 2 if self.nz == 1 else 3  # 2-D: rhox and rhoy only (SolverCudaKernels.cu:570-629, :795-807)
            if pmode == 2:
                f = self._scale_source(s, self.p_src_idx)
                for i in range(ndim):
                    self.rho[i] = self.rho[i] + f
            else:
                for i in range(ndim):
                    flat = self.rho[i].reshape(-1)
                    if pmode == 0:
                        flat[self.p_src_idx] = s
                    else:
                        flat[self.p_src_idx] += s
        # computePressure (cpp:2180-2246)
        rs = self.rho[0] + self.rho[1] + self.rho[2]
        if c["absorbing_flag"]:
            a_term = self.rho0 * (du[0] + du[1] + du[2])  # SolverCudaKernels.cu:1588-1601 / :1733-1741
            ta = self._ifft(self._fft(a_term) * self.nabla1)  # :1816-1819
            tb = self._ifft(self._fft(rs) * self.nabla2)
            if c["nonlinear_flag"]:
                nl = ((self.bona * rs * rs) / (2 * self.rho0)) + rs
                self.p = self.c2 * (nl + fd * ((ta * self.tau) - (tb * self.eta)))  # :1872-1878
            else:
                self.p = self.c2 * (rs + fd * (ta * self.tau - tb * self.eta))  # :1973-1979
        else:
            if c["nonlinear_flag"]:
                self.p = self.c2 * (rs + (self.bona * (rs * rs) / (2 * self.rho0)))  # :2079-2082
            else:
                self.p = self.c2 * rs  # :2229-2235
        # addInitialPressureSource (cpp:2359-2396; SolverCudaKernels.cu:870-883, :971-980)
        if t == 0 and c.get("p0_source_flag", 0) == 1:
            self.p = self.p0.copy() if isinstance(self.p0, np.ndarray) else np.full(self.shape, self.p0, self.dt)
            if self.nz == 1:  # SolverCudaKernels.cu:873: rho_x = rho_y = p0 / (2 c^2), no z component
                r = self.p / (2 * self.c2)
                self.rho = [r.copy(), r.copy(), np.zeros_like(r)]
            else:
                r = self.p / (3 * self.c2)
                self.rho = [r.copy(), r.copy(), r.copy()]
            g = self._pressure_gradient()
            half_fd = fd * self.dt.type(0.5)
            for i in range(3):
                self.u[i] = g[i] * (self.dtrho[i] * half_fd)
        self.t += 1

    # -- non-staggered velocity (cpp:2714-2735; SolverCudaKernels.cu:2617-2689) -----------------------------------------
    def shifted_velocity(self):
        out = []
        for i, ax in enumerate((2, 1, 0)):
            n = self.shape[ax]
            uk = sfft.rfft(self.u[i], axis=ax).astype(self.cdt, copy=False)
            sh = self.shift_neg[i][: n // 2 + 1]
            shp = [1, 1, 1]
            shp[ax] = -1
            uk = (uk * sh.reshape(shp)) * self.dt.type(F32(1.0) / F32(n))
            out.append((sfft.irfft(uk, n=n, axis=ax) * n).astype(self.dt, copy=False))
        return out

    # -- sampling (OutputStreams/OutputStreamsCudaKernels.cu:83-316) --------------------------------------------------
    def sample_index(self, fieldarr):
        return fieldarr.reshape(-1)[self.sensor_index]

    def cuboid_indices(self):
        """Linear indices of every cuboid, x fastest inside a cuboid, cuboids concatenated
        (OutputStreamsCudaKernels.cu:164-188; CuboidOutputStream.cpp:265-338)."""
        out = []
        for cc in self.sensor_corners:
            x0, y0, z0, x1, y1, z1 = cc
            zz, yy, xx = np.meshgrid(
                np.arange(z0, z1 + 1), np.arange(y0, y1 + 1), np.arange(x0, x1 + 1), indexing="ij"
            )
            out.append(((zz * self.ny + yy) * self.nx + xx).reshape(-1))
        return out

    def sample_cuboids(self, fieldarr):
        flat = fieldarr.reshape(-1)
        return [flat[i] for i in self.cuboid_indices()]


# ---------------------------------------------------------------------------------------------------------------------
# running a whole simulation with the reference's output-stream semantics
# ---------------------------------------------------------------------------------------------------------------------
def run(cfg, arrays, nt=None, dtype=np.float64, record=("p_raw",), start_index=0):
    """Run ``nt`` steps and return a dict of outputs named as in the output file (Utils/MatrixNames.h).
    ``start_index`` is the 0-based sampling start (``-s`` minus one, CommandLineParameters.cpp:424).
    Supported ``record`` entries: p_raw p_rms p_max p_min p_max_all p_min_all p_final u_raw u_rms u_max u_min
    u_max_all u_min_all u_final u_non_staggered_raw."""
    o = KSpaceOracle(cfg, arrays, dtype)
    nt = cfg["Nt"] if nt is None else nt
    rec = set(record)
    out = {}
    idx_mode = o.sensor_index is not None
    if not idx_mode and o.sensor_corners is not None:
        cub = np.concatenate(o.cuboid_indices())
    else:
        cub = None
    sens = o.sensor_index if idx_mode else cub
    FMAX = np.finfo(np.float32).max

    def samp(f):
        return f.reshape(-1)[sens]

    series = {k: [] for k in ("p", "ux", "uy", "uz", "ux_non_staggered", "uy_non_staggered", "uz_non_staggered")}
    agg = {}

    def agg_update(name, f, full=False):
        x = f.reshape(-1) if full else samp(f)
        if name.endswith("rms"):
            agg[name] = agg.get(name, 0) + x * x
        elif "max" in name:
            agg[name] = np.maximum(agg.get(name, -FMAX), x)
        else:
            agg[name] = np.minimum(agg.get(name, FMAX), x)

    for t in range(nt):
        o.step()
        if t < start_index:
            continue
        if "p_raw" in rec:
            series["p"].append(samp(o.p).copy())
        for nm in ("p_rms", "p_max", "p_min"):
            if nm in rec:
                agg_update(nm, o.p)
        for nm in ("p_max_all", "p_min_all"):
            if nm in rec:
                agg_update(nm, o.p, full=True)
        if "u_non_staggered_raw" in rec:
            us = o.shifted_velocity()
            for i, a in enumerate("xyz"):
                series[f"u{a}_non_staggered"].append(samp(us[i]).copy())
        for i, a in enumerate("xyz"):
            if "u_raw" in rec:
                series[f"u{a}"].append(samp(o.u[i]).copy())
            for kind in ("rms", "max", "min"):
                if f"u_{kind}" in rec:
                    agg_update(f"u{a}_{kind}", o.u[i])
            for kind in ("max_all", "min_all"):
                if f"u_{kind}" in rec:
                    agg_update(f"u{a}_{kind}", o.u[i], full=True)
    nsamp = nt - start_index
    for k, v in series.items():
        if v:
            out[k] = np.stack(v)  # (steps, Nsens): one row per step (IndexOutputStream.cpp:583-591)
    for k, v in agg.items():
        # rms post-processing: sqrt(buf * 1/(Nt-s))  (OutputStreamsCudaKernels.cu:359-363)
        out[k] = np.sqrt(v * o.dt.type(F32(1.0) / F32(nsamp))) if k.endswith("rms") else v
    if "p_final" in rec:
        out["p_final"] = o.p.copy()
    if "u_final" in rec:
        for i, a in enumerate("xyz"):
            out[f"u{a}_final"] = o.u[i].copy()
    out["_oracle"] = o
    return out
