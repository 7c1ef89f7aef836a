"""CPU restatement of the reference's on-the-fly harmonic compression (TEST INFRASTRUCTURE ONLY; see kspace_oracle.py).

Follows Compression/CompressHelper.cpp (:48-65 init, :672-778 basis generation, :224-389 40-bit codec) and the host-side
accumulation of OutputStreams/IndexOutputStream.cpp (:373-470 flushRaw, :299-342 postSample) and
BaseOutputStream.cpp (:117-133 postSample2).  Pinned against the reference's own CompressHelper.cpp compiled on the CPU
(oracle/_ref/compress_ref -> tests/golden/compress_ref.json): bases to 1 ulp of libm cosf/sinf, codec bit-exact.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
K_MAX_EXP_P, K_MAX_EXP_U = 138, 114  # CompressHelper.h:91-92


# ---------------------------------------------------------------------------------------------------------------------
def triangular(osize):
    """CompressHelper.cpp:700-710."""
    x = np.arange(2 * osize + 1, dtype=F32)
    w = np.where(x < osize, x / F32(osize), F32(2.0) - x / F32(osize))
    return w.astype(F32)


def generate_bases(period, mos, harmonics, normalize=True, shift=False):
    """bE, bE_1 as complex64 arrays of shape (harmonics, bSize).  CompressHelper.cpp:48-65, :672-778."""
    period = F32(period)
    osize = int(period * F32(mos))
    bsize = 2 * osize + 1
    b = triangular(osize)
    x = np.arange(bsize, dtype=F32)
    be = np.zeros((harmonics, bsize), np.complex64)
    be1 = np.zeros((harmonics, bsize), np.complex64)
    for ih in range(harmonics):
        h = F32(ih + 1)
        # e = exp(-i * (2 pi / (period / h)) * x) [* exp(+i pi / (period / h)) when shifted], all in FP32 (:733-746)
        w = F32(2.0) * F32(np.pi) / (period / h)
        ang = (w * x).astype(F32)
        e = (np.cos(ang, dtype=F32) - 1j * np.sin(ang, dtype=F32)).astype(np.complex64)
        if shift:
            s = F32(np.pi) / (period / h)
            e = (e * np.complex64(np.cos(s, dtype=F32) + 1j * np.sin(s, dtype=F32))).astype(np.complex64)
        idx = (np.arange(bsize) + osize) % (bsize - 1)
        be[ih] = (b * e).astype(np.complex64)
        be1[ih] = (b[idx] * e[idx]).astype(np.complex64)
        if normalize:
            sc = F32(2.0) / F32(osize)
            be[ih] = (be[ih] * sc).astype(np.complex64)
            be1[ih] = (be1[ih] * sc).astype(np.complex64)
    return osize, bsize, be, be1


# ---------------------------------------------------------------------------------------------------------------------
def _bits(f):
    return int(np.array(f, dtype=F32).view(np.uint32))


def _float(u):
    return np.array(u & 0xFFFFFFFF, dtype=np.uint32).view(F32)[()]


def encode40(re, im, e):
    """convertFloatCTo40b (CompressHelper.cpp:292-389) -> 5 bytes."""
    mr, mi = _bits(re), _bits(im)
    sr, si = mr >> 31, mi >> 31
    ers = ((mr & 0x7F800000) >> 23) - e
    eis = ((mi & 0x7F800000) >> 23) - e
    es = ers
    mr &= 0x007FFFFF
    mi &= 0x007FFFFF
    rsr = rsi = 6
    if ers > eis:
        rsi += ers - eis
        es = ers
    elif eis > ers:
        rsr += eis - ers
        es = eis
    if es < 0:
        rsr += -es
        rsi += -es
        es = 0
    # the reference keeps the shifts in uint8_t: wrap like it does before clamping
    rsr &= 0xFF
    rsi &= 0xFF
    rsr, rsi = min(rsr, 23), min(rsi, 23)
    mr >>= rsr
    mi >>= rsi
    if mr > 0 and mr != (0x7FFFFF >> rsr):
        mr += 1
    if mi > 0 and mi != (0x7FFFFF >> rsi):
        mi += 1
    mr |= 1 << (23 - rsr)
    mr >>= 1
    mi |= 1 << (23 - rsi)
    mi >>= 1
    if es > 0xF:
        mr = mi = 0xFFFF
        es = 0xF
    b0 = ((sr << 7) | (si << 6) | ((mr & 0x10000) >> 11) | ((mi & 0x10000) >> 12) | (es & 0xF)) & 0xFF
    return bytes([b0, mr & 0xFF, (mr >> 8) & 0xFF, mi & 0xFF, (mi >> 8) & 0xFF])


def decode40(b, e):
    """convert40bToFloatC (CompressHelper.cpp:224-290) -> (re, im) float32."""
    b0 = b[0]
    mr = ((b0 & 0x20) << 11) | (b[1] | (b[2] << 8))
    mi = ((b0 & 0x10) << 12) | (b[3] | (b[4] << 8))
    sr, si, es = b0 >> 7, (b0 & 0x40) >> 6, b0 & 0xF
    mr <<= 6
    mi <<= 6
    er = ei = es + e

    def norm(m, ex):
        if m != 0:
            index = m.bit_length() - 1
            m = (m << (23 - index)) & 0xFFFFFFFF
            ex -= 22 - index
        else:
            ex = 0
        return m, ex

    mr, er = norm(mr, er)
    mi, ei = norm(mi, ei)
    ccr = ((sr << 31) | ((er << 23) & 0xFFFFFFFF) | (mr & 0x007FFFFF)) & 0xFFFFFFFF
    cci = ((si << 31) | ((ei << 23) & 0xFFFFFFFF) | (mi & 0x007FFFFF)) & 0xFFFFFFFF
    return _float(ccr), _float(cci)


# ---------------------------------------------------------------------------------------------------------------------
class CompressedStream:
    """Host-side state machine of one `*_c` stream (IndexOutputStream::flushRaw, IndexOutputStream.cpp:373-470).
    feed(x) takes the Nsens raw samples of one sampled step and returns the frame that the reference would write to the
    file at that step (complex64 array of shape (Nsens, harmonics)), or None."""

    def __init__(self, nsens, period, mos=1, harmonics=1, shifted=False, no_overlap=False, nsteps_total=None, dtype=np.complex64,
                 c40=False, bases=None):
        self.osize, self.bsize, self.be, self.be1 = generate_bases(period, mos, harmonics, True, shifted)
        if bases is not None:  # bases produced elsewhere (they depend on the libm's cosf/sinf to an ulp)
            self.be, self.be1 = (np.asarray(b, np.complex64).reshape(harmonics, self.bsize) for b in bases)
        self.h = harmonics
        self.no_overlap = no_overlap
        self.c40 = c40
        self.e = K_MAX_EXP_U if shifted else K_MAX_EXP_P  # BaseOutputStream.cpp:63-102
        if c40:  # the accumulators are kept packed (5 bytes per value) and re-quantised at every step
            self.q1 = np.zeros((nsens, harmonics, 5), np.uint8)
            self.q2 = self.q1 if no_overlap else np.zeros((nsens, harmonics, 5), np.uint8)
        self.buf1 = np.zeros((nsens, harmonics), dtype)
        self.buf2 = self.buf1 if no_overlap else np.zeros((nsens, harmonics), dtype)  # BaseOutputStream.cpp:246-257
        self.sampled = 0
        self.compressed = 0
        self.nsteps_total = nsteps_total
        self.dtype = dtype

    def _feed40(self, x, step_local, mirror):
        """40-bit branch of flushRaw (IndexOutputStream.cpp:412-441): FP32, products and sums rounded separately."""
        x = np.asarray(x, dtype=F32)
        for i in range(self.q1.shape[0]):
            for ih in range(self.h):
                e, e1 = self.be[ih, step_local], self.be1[ih, step_local]
                c1 = decode40(bytes(self.q1[i, ih]), self.e)
                if self.no_overlap:
                    re = F32(c1[0]) + (F32(e.real) * x[i] + F32(e1.real) * x[i])
                    im = F32(c1[1]) + (F32(e.imag) * x[i] + F32(e1.imag) * x[i])
                    self.q1[i, ih] = np.frombuffer(encode40(re, im, self.e), np.uint8)
                    continue
                c2 = decode40(bytes(self.q2[i, ih]), self.e)
                r1, i1 = F32(c1[0]) + F32(e.real) * x[i], F32(c1[1]) + F32(e.imag) * x[i]
                r2, i2 = F32(c2[0]) + F32(e1.real) * x[i], F32(c2[1]) + F32(e1.imag) * x[i]
                self.q1[i, ih] = np.frombuffer(encode40(r1, i1, self.e), np.uint8)
                if mirror:
                    r2, i2 = r2 + r1, i2 + i1
                self.q2[i, ih] = np.frombuffer(encode40(r2, i2, self.e), np.uint8)

    def feed(self, x):
        step_local = self.sampled % (self.bsize - 1)
        saving = (step_local + 1) % self.osize == 0
        odd = (self.compressed + 1) % 2 == 0
        mirror = self.compressed == 0 and saving and not self.no_overlap
        if self.c40:
            self._feed40(x, step_local, mirror)
        else:
            x = np.asarray(x).astype(self.buf1.real.dtype)[:, None]
            # NOTE with no_overlap buf1 is buf2: both updates land in the same accumulator, as in the reference
            self.buf1 += self.be[:, step_local].astype(self.dtype)[None, :] * x
            self.buf2 += self.be1[:, step_local].astype(self.dtype)[None, :] * x
            if mirror:
                self.buf2 += self.buf1
        out = None
        last = (
            self.nsteps_total is not None
            and (self.nsteps_total - self.sampled == 1)
            and self.nsteps_total <= self.osize
        )
        if saving or last:
            cur = (self.q1 if odd else self.q2) if self.c40 else (self.buf1 if odd else self.buf2)
            out = cur.copy()  # c40: packed bytes (Nsens, harmonics, 5)
            self.compressed += 1
            if saving:
                cur[...] = 0  # postSample2 (BaseOutputStream.cpp:117-133): only after a regular saving step
        self.sampled += 1
        return out


def intensity_frame(pc, uc):
    """IndexOutputStream::postSample (:299-342): sum over harmonics of Re(P * conj(U)) / 2 for one saved frame."""
    return (pc * np.conj(uc)).real.sum(axis=1) / 2.0


def unpack40(frame_bytes, e):
    """Packed frame (Nsens, harmonics, 5) uint8 -> complex64 (Nsens, harmonics)."""
    out = np.zeros(frame_bytes.shape[:2], np.complex64)
    for i in range(frame_bytes.shape[0]):
        for ih in range(frame_bytes.shape[1]):
            re, im = decode40(bytes(frame_bytes[i, ih]), e)
            out[i, ih] = complex(re, im)
    return out


def q_term(cfg, intensities, lin_index):
    """computeQTerm (KSpaceFirstOrderSolver.cpp:1783-2080): Q = -(dIx/dx + dIy/dy [+ dIz/dz]) with the intensities scattered
    onto a zero grid at the (0-based linear) sensor indices, each differentiated spectrally along its own axis
    (1-D R2C, * i*k / N, C2R; k = 2 pi / d * shift / N, shift = (i + N/2) % N - N/2), gathered at the same indices."""
    nx, ny, nz = cfg["Nx"], cfg["Ny"], cfg["Nz"]
    idx = np.asarray(lin_index, dtype=np.int64)
    total = np.zeros((nz, ny, nx), np.float64)
    for comp, (ax, n, d) in enumerate(((2, nx, cfg["dx"]), (1, ny, cfg["dy"]), (0, nz, cfg.get("dz", 1.0)))):
        if comp >= len(intensities) or (nz == 1 and comp == 2):
            break
        grid = np.zeros(nx * ny * nz, np.float64)
        grid[idx] = np.asarray(intensities[comp], np.float64)
        i = np.arange(n // 2 + 1)
        shift = (i + n // 2) % n - n // 2
        ik = 1j * (2.0 * np.pi / float(F32(d))) * (shift / n)
        shp = [1, 1, 1]
        shp[ax] = -1
        xk = np.fft.rfft(grid.reshape(nz, ny, nx), axis=ax) * ik.reshape(shp)
        total += np.fft.irfft(xk, n=n, axis=ax)
    return (-total).reshape(-1)[idx]


def intensity_avg(p, u):
    """computeAverageIntensities (KSpaceFirstOrderSolver.cpp:1231-1534) for series of shape (steps, n): the velocity is shifted by
    half a time step spectrally -- R2C along time, * exp(i*pi*shift/steps) / steps with shift = (k + steps/2) % steps - steps/2,
    C2R (which ignores the imaginary part of the Nyquist bin) -- and I = sum_t p * u_shifted / steps."""
    p = np.asarray(p, np.float64)
    u = np.asarray(u, np.float64)
    steps = p.shape[0]
    k = np.arange(steps // 2 + 1)
    shift = (k + steps // 2) % steps - steps // 2
    kx = np.exp(1j * np.pi * shift / steps)
    us = np.fft.irfft(np.fft.rfft(u, axis=0) * kx[:, None], n=steps, axis=0)
    return (p * us).sum(axis=0) / steps


def half_step_kernel(steps):
    """The same shift as a circular convolution kernel: u_shifted[t] = sum_s u[s] * h[(t - s) % steps]."""
    m = np.arange(steps)
    kk = np.arange(1, (steps - 1) // 2 + 1)
    return (1.0 + 2.0 * np.cos(2.0 * np.pi * np.outer(m, kk) / steps + np.pi * kk / steps).sum(axis=1)) / steps
