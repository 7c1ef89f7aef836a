"""Host logic of slab-decomposed runs on CPU ranks (gloo, world_size 2 and 4): the exchange layout of the 3-D transform
(one all-to-all per transform), the partition of sensor / source index lists and cuboids, and the re-assembly of rows in
mask order (bit-exact ordering requirement)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
slab = importlib.import_module("k-wave-fluid-cuda_b200.slab")
synth = importlib.import_module("k-wave-fluid-cuda_b200.synth")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shape, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sl = importlib.import_module("k-wave-fluid-cuda_b200.slab")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    def all_to_all(blocks):
        send = [torch.from_numpy(np.ascontiguousarray(b).view(np.float64).copy()) for b in blocks]
        recv = [torch.empty_like(s) for s in send]
        # gloo has no all_to_all: pairwise exchange (what ncclSend/ncclRecv in a group do on the GPU side)
        ops = []
        for peer in range(world):
            if peer == rank:
                recv[peer].copy_(send[peer])
                continue
            ops.append(dist.isend(send[peer], peer))
            ops.append(dist.irecv(recv[peer], peer))
        for o in ops:
            o.wait()
        return [r.numpy().view(np.complex128).reshape(blocks[0].shape) for r in recv]

    nz, ny, nx = shape
    x = np.random.default_rng(5).standard_normal(shape)
    z0, nzl = sl.slab_extent(nz, rank, world)
    y0, nyl = sl.slab_extent(ny, rank, world)
    spec = sl.slab_rfftn(x[z0 : z0 + nzl], world, all_to_all)
    want = np.fft.rfftn(x)  # [kz][ky][kx]
    e_fwd = np.abs(spec - want[:, y0 : y0 + nyl]).max() / np.abs(want).max()
    back = sl.slab_irfftn(spec, nx, world, all_to_all) / x.size
    e_inv = np.abs(back - x[z0 : z0 + nzl]).max()
    # sensor rows: every rank samples "its" points of a fake field (value == global linear index), rank 0 assembles
    cfg, arrays = synth.make_case(nx, ny, nz, nt=4, source="p_plane", shuffle_sensor=True, n_sensor=64)
    pos, loc = sl.index_partition(arrays["sensor_mask_index"], cfg, rank, world)
    field = (np.arange(nzl * ny * nx, dtype=np.int64) + z0 * ny * nx).astype(np.float64)
    rows = field[loc.astype(np.int64)][None, :]
    gathered = [None] * world
    dist.all_gather_object(gathered, (pos, rows))
    full = sl.assemble_rows(arrays["sensor_mask_index"].size, gathered)
    ok_rows = bool(np.array_equal(full[0], (arrays["sensor_mask_index"] - 1).astype(np.float64)))
    q.put((rank, float(e_fwd), float(e_inv), ok_rows))
    dist.destroy_process_group()


def _step_worker(rank, world, port, q):
    """The oracle's whole time loop with every 3-D transform running through the slab scheme (local x/y transforms, ONE
    all-to-all, local z transform, and back), sensors sampled by the rank that owns them and rows assembled by position."""
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sl = importlib.import_module("k-wave-fluid-cuda_b200.slab")
    sy = importlib.import_module("k-wave-fluid-cuda_b200.synth")
    from oracle import kspace_oracle as ko

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    def all_to_all(blocks):
        send = [torch.from_numpy(np.ascontiguousarray(b).view(np.float64).copy()) for b in blocks]
        recv = [torch.empty_like(s) for s in send]
        ops = []
        for peer in range(world):
            if peer == rank:
                recv[peer].copy_(send[peer])
                continue
            ops.append(dist.isend(send[peer], peer))
            ops.append(dist.irecv(recv[peer], peer))
        for o in ops:
            o.wait()
        return [r.numpy().view(np.complex128).reshape(blocks[0].shape) for r in recv]

    def gather(part, axis):
        parts = [None] * world
        dist.all_gather_object(parts, part)
        return np.concatenate(parts, axis=axis)

    nx, ny, nz, nt = 16, 8, 16, 12
    cfg, arrays = sy.make_case(nx, ny, nz, nt=nt, nonlinear=True, absorbing=True, source="p_many", source_mode=1, shuffle_sensor=True, n_sensor=40)
    z0, nzl = sl.slab_extent(nz, rank, world)
    y0, nyl = sl.slab_extent(ny, rank, world)

    class SlabOracle(ko.KSpaceOracle):
        def _fft(self, x):  # only this rank's z-slab goes in, only its ky range comes out of the exchange
            spec = sl.slab_rfftn(np.asarray(x, np.float64)[z0 : z0 + nzl], world, all_to_all)
            return gather(spec, 1).astype(self.cdt)

        def _ifft(self, xk):
            loc = sl.slab_irfftn(np.asarray(xk)[:, y0 : y0 + nyl, :], nx, world, all_to_all)
            return gather(loc, 0).astype(self.dt)

    o = SlabOracle(cfg, arrays, np.float64)
    pos, loc = sl.index_partition(arrays["sensor_mask_index"], cfg, rank, world)
    rows = []
    for _ in range(nt):
        o.step()
        rows.append(o.p[z0 : z0 + nzl].reshape(-1)[loc.astype(np.int64)].copy())  # this rank samples ITS points only
    parts = [None] * world
    dist.all_gather_object(parts, (pos, np.stack(rows)))
    full = sl.assemble_rows(arrays["sensor_mask_index"].size, parts)
    ref = ko.run(cfg, arrays, nt=nt, dtype=np.float64, record=("p_raw",))["p"]
    q.put((rank, float(np.abs(full - ref).max() / np.abs(ref).max())))
    dist.destroy_process_group()


def test_time_loop_through_slab_transforms_gloo():
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_step_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err in res:
        assert err < 1e-10, (rank, err)


@pytest.mark.parametrize("world,shape", [(2, (8, 12, 10)), (4, (16, 8, 16))])
def test_slab_transform_and_row_assembly_gloo(world, shape):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e_fwd, e_inv, ok_rows in res:
        assert e_fwd < 1e-12, (rank, e_fwd)
        assert e_inv < 1e-12, (rank, e_inv)
        assert ok_rows, rank


def test_exchange_layout_round_trip():
    a = np.arange(4 * 8 * 5).reshape(4, 8, 5).astype(np.complex128)
    for p in (1, 2, 4):
        b = slab.to_exchange_layout(a, p)
        assert b.shape == (p, 4, 8 // p, 5)
        # RowMap::off (csrc/fft_kernels.cuh): row (z, y) -> (y / nyl) * blk + (z * nyl + y % nyl) * nxp
        nyl, blk = 8 // p, 4 * (8 // p) * 5
        for z, y in ((0, 0), (3, 7), (2, 5)):
            off = (y // nyl) * blk + (z * nyl + y % nyl) * 5
            assert np.array_equal(b.reshape(-1)[off : off + 5], a[z, y])
        assert np.array_equal(slab.from_exchange_layout(b, p), a)


def test_index_partition_covers_list_in_order():
    cfg, arrays = synth.make_case(16, nt=4, source="p_many", shuffle_sensor=True, n_sensor=50)
    for name in ("sensor_mask_index", "p_source_index"):
        idx = arrays[name]
        for p in (1, 2, 4, 8):
            seen = np.zeros(idx.size, dtype=int)
            for r in range(p):
                pos, loc = slab.index_partition(idx, cfg, r, p)
                assert np.all(np.diff(pos.astype(np.int64)) > 0)  # list order kept
                z0, nzl = slab.slab_extent(16, r, p)
                assert np.array_equal(loc + np.uint64(z0 * 256), idx[pos.astype(np.int64)] - 1)
                assert loc.size == 0 or int(loc.max()) < nzl * 256
                seen[pos.astype(np.int64)] += 1
            assert np.all(seen == 1)


def test_cuboid_partition_matches_reference_ordering():
    cfg, arrays = synth.make_case(16, 16, 32, nt=4, source="p0", sensor="cuboid")
    corners = arrays["sensor_mask_corners"]
    nx, ny, nz = 16, 16, 32
    vol = np.arange(nx * ny * nz, dtype=np.int64).reshape(nz, ny, nx)
    # the undecomposed row: cuboids concatenated, x fastest inside each (CuboidOutputStream.cpp:265-338)
    full = np.concatenate([vol[c[2] - 1 : c[5], c[1] - 1 : c[4], c[0] - 1 : c[3]].reshape(-1) for c in corners.astype(np.int64)])
    for p in (1, 2, 4):
        parts = []
        for r in range(p):
            total, pos, loc = slab.cuboid_partition(corners, cfg, r, p)
            assert total == full.size
            z0, nzl = slab.slab_extent(nz, r, p)
            sl = vol[z0 : z0 + nzl]
            vals = [sl[c[2] : c[5] + 1, c[1] : c[4] + 1, c[0] : c[3] + 1].reshape(-1) for c in loc if c[0] >= 0]
            parts.append((pos, (np.concatenate(vals) if vals else np.empty(0, np.int64))[None, :]))
        assert np.array_equal(slab.assemble_rows(total, parts)[0], full)


def test_slice_arrays_accepts_grids_and_slabs():
    cfg, arrays = synth.make_case(16, nt=4, source="p0")
    a = slab.slice_arrays(cfg, arrays, 1, 2)
    assert a["c0"].shape == (8, 16, 16) and np.array_equal(a["c0"], arrays["c0"][8:])
    assert a["pml_z"].size == 16 and a["sensor_mask_index"] is not None
    b = slab.slice_arrays(cfg, a, 1, 2)  # slabs pass through
    assert np.array_equal(b["c0"], a["c0"])
    with pytest.raises(ValueError):
        slab.slice_arrays(cfg, {"c0": np.zeros(17, np.float32)}, 0, 2)
    with pytest.raises(ValueError):
        slab.slab_extent(10, 0, 4)
