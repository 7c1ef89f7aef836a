"""Worker of the slab-decomposed GPU tests (one process per GPU, spawned by tests/test_slab_gpu.py)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_rank(rank, world, nccl_id, shape, kwargs, nt, streams, start_index, q, compression=None):
    try:
        sys.path.insert(0, ROOT)
        kw = importlib.import_module("k-wave-fluid-cuda_b200")
        cfg, arrays = kw.synth.make_case(*shape, nt=nt, **kwargs)
        sim = kw.Simulation(cfg, arrays, streams=streams, start_index=start_index, raw_rows_capacity=nt, device=rank,
                            rank=rank, nranks=world, nccl_id=nccl_id, compression=compression)
        done = sim.run(nt)
        sim.finish()
        total, pos = sim.sensor_layout()
        out = {s: sim.fetch(s) for s in streams}
        out["p_final"] = sim.get_array("KW_P")
        out["ux_final"] = sim.get_array("KW_UX_SGX")
        out.update(done=done, total=total, pos=pos, slab=sim.local_slab(), comm_bytes=sim.comm_bytes(), comm_mode=sim.comm_mode(), launches=sim.launch_count())
        sim.close()
        q.put((rank, out))
    except Exception as e:  # surface the failure in the parent instead of a hang
        import traceback

        q.put((rank, {"error": f"{e}\n{traceback.format_exc()}"}))
