"""Host-side logic of slab-decomposed runs (no reference counterpart: the reference is single-GPU,
Containers/MatrixContainer.cpp:418-476 holds whole arrays).

Rank r of P owns the planes z in [r*Nz/P, (r+1)*Nz/P) of every real-space array and, after the all-to-all of a 3-D
transform, the ky range [r*Ny/P, (r+1)*Ny/P) of every half spectrum for all kz.  This module holds what the host has
to do around the C ABI -- cut inputs to slabs, know which sensor / source points a rank keeps and where its columns
belong in the row of the undecomposed run -- plus a NumPy statement of the exchange layout the CUDA kernels use, so that
the decomposition can be tested on CPU ranks (gloo) without a GPU.  NumPy only; nothing here is on the compute path.
"""
from __future__ import annotations

import numpy as np

# datasets of the input file that are Nx*Ny*Nz grids when heterogeneous (Containers/MatrixContainer.cpp:94-410)
FULL_GRID = ("c0", "rho0", "rho0_sgx", "rho0_sgy", "rho0_sgz", "BonA", "alpha_coeff", "p0_source_input",
             "KW_P", "KW_RHOX", "KW_RHOY", "KW_RHOZ", "KW_UX_SGX", "KW_UY_SGY", "KW_UZ_SGZ")  # fmt: skip


def slab_extent(n, rank, nranks):
    """(begin, count) of rank's share of an axis of length n."""
    if nranks < 1 or n % nranks:
        raise ValueError(f"axis of length {n} is not divisible by {nranks} ranks")
    c = n // nranks
    return rank * c, c


def slice_arrays(cfg, arrays, rank, nranks):
    """Inputs of one rank: full-grid arrays cut to its z-slab (complete grids and slabs are both accepted), everything
    else (1-D operators, signals, index lists with GLOBAL 1-based indices) unchanged."""
    nx, ny, nz = cfg["Nx"], cfg["Ny"], cfg["Nz"]
    z0, nzl = slab_extent(nz, rank, nranks)
    out = {}
    for name, a in arrays.items():
        a = np.asarray(a)
        if name in FULL_GRID and a.size == nx * ny * nz:
            out[name] = np.ascontiguousarray(a.reshape(nz, ny, nx)[z0 : z0 + nzl])
        elif name in FULL_GRID and a.size not in (1, nx * ny * nzl):
            raise ValueError(f"{name}: expected 1, {nx * ny * nz} (grid) or {nx * ny * nzl} (slab) elements, got {a.size}")
        else:
            out[name] = a
    return out


def index_partition(index_1based, cfg, rank, nranks):
    """Which entries of a 1-based linear index list (sensor mask, source index) rank keeps: returns (positions in the
    original list, local 0-based voxel indices), both in list order (bit-exact ordering requirement)."""
    nx, ny, nz = cfg["Nx"], cfg["Ny"], cfg["Nz"]
    z0, nzl = slab_extent(nz, rank, nranks)
    idx = np.asarray(index_1based, dtype=np.uint64).reshape(-1) - np.uint64(1)
    lo = np.uint64(z0 * nx * ny)
    hi = lo + np.uint64(nzl * nx * ny)
    keep = np.nonzero((idx >= lo) & (idx < hi))[0]
    return keep.astype(np.uint64), (idx[keep] - lo).astype(np.uint64)


def cuboid_partition(corners_1based, cfg, rank, nranks):
    """Cuboid sensor masks: every cuboid is clipped to the slab; a z-range of a cuboid is a contiguous range of its
    x-fastest buffer (OutputStreams/CuboidOutputStream.cpp:265-338).  Returns (total points of all cuboids,
    positions of this rank's points in the concatenated row, local clipped corners 0-based (ncub, 6) with -1 rows for
    cuboids that miss the slab)."""
    nz = cfg["Nz"]
    z0, nzl = slab_extent(nz, rank, nranks)
    q = np.asarray(corners_1based, dtype=np.int64).reshape(-1, 6) - 1
    pos, loc, goff = [], [], 0
    for x0, y0, c0, x1, y1, c1 in q:
        cxy = (x1 - x0 + 1) * (y1 - y0 + 1)
        zlo, zhi = max(c0, z0), min(c1, z0 + nzl - 1)
        if zlo <= zhi:
            pos.append(np.arange(goff + cxy * (zlo - c0), goff + cxy * (zhi - c0 + 1), dtype=np.uint64))
            loc.append([x0, y0, zlo - z0, x1, y1, zhi - z0])
        else:
            loc.append([-1] * 6)
        goff += cxy * (c1 - c0 + 1)
    return goff, (np.concatenate(pos) if pos else np.empty(0, np.uint64)), np.asarray(loc, dtype=np.int64)


def assemble_rows(total, parts):
    """Rows of the undecomposed run from per-rank rows: parts = [(positions, rows[nrows, nlocal]), ...]."""
    parts = [(np.asarray(p, dtype=np.int64), np.asarray(r)) for p, r in parts]
    nrows = max((r.shape[0] for _, r in parts), default=0)
    dtype = parts[0][1].dtype if parts else np.float32
    out = np.zeros((nrows, int(total)), dtype=dtype)
    seen = np.zeros(int(total), dtype=bool)
    for p, r in parts:
        if p.size == 0:
            continue
        if seen[p].any():
            raise ValueError("a sensor point is claimed by two ranks")
        seen[p] = True
        out[:, p] = r
    if not seen.all():
        raise ValueError(f"{int((~seen).sum())} sensor points are claimed by no rank")
    return out


# ---------------------------------------------------------------------------------------------------------------------
# The exchange layout, stated in NumPy (what k_xfwd / k_col<BLOCKED> write and k_zmid reads; csrc/fft_kernels.cuh RowMap)
def to_exchange_layout(spec_zyx, nranks):
    """x/y-local half spectrum [nzl][Ny][NX] -> y-blocked [q][nzl][nyl][NX]; block q goes to rank q."""
    nzl, ny, nxc = spec_zyx.shape
    nyl = ny // nranks
    return np.ascontiguousarray(spec_zyx.reshape(nzl, nranks, nyl, nxc).transpose(1, 0, 2, 3))


def from_exchange_layout(blocks, nranks):
    """inverse of to_exchange_layout."""
    q, nzl, nyl, nxc = blocks.shape
    return np.ascontiguousarray(blocks.transpose(1, 0, 2, 3).reshape(nzl, q * nyl, nxc))


def slab_rfftn(x_local, nranks, all_to_all):
    """Unnormalised 3-D R2C of a z-slab: local x and y transforms, ONE all-to-all, local z transform.  Returns this rank's
    z-local spectrum [Nz][nyl][Nx/2+1].  `all_to_all(list of P blocks) -> list of P blocks` (block q to/from rank q)."""
    s = np.fft.fft(np.fft.rfft(x_local, axis=2), axis=1)
    send = to_exchange_layout(s, nranks)
    recv = all_to_all([send[q] for q in range(nranks)])
    zl = np.concatenate(recv, axis=0)  # block r holds the planes of rank r: [Nz][nyl][NX]
    return np.fft.fft(zl, axis=0)


def slab_irfftn(spec_zlocal, nx, nranks, all_to_all):
    """Unnormalised 3-D C2R back to this rank's z-slab [nzl][Ny][Nx] (cuFFT convention: no 1/N)."""
    nz, nyl, nxc = spec_zlocal.shape
    nzl = nz // nranks
    zl = np.fft.ifft(spec_zlocal, axis=0) * nz
    recv = all_to_all([np.ascontiguousarray(zl[q * nzl : (q + 1) * nzl]) for q in range(nranks)])
    s = from_exchange_layout(np.stack(recv, axis=0), nranks)
    s = np.fft.ifft(s, axis=1) * (nyl * nranks)
    return np.fft.irfft(s, n=nx, axis=2) * nx
