"""Peer-copy bandwidth between GPU0 and GPU1 (copy engines over NVLink): one copy vs the same bytes split over several
streams, one direction and both directions at once.  Guides how the all-to-all pushes are issued."""
import torch, time
assert torch.cuda.device_count() >= 2
n = 1 << 28  # 1 GiB of float32
a0 = torch.empty(n, dtype=torch.float32, device="cuda:0"); b1 = torch.empty(n, dtype=torch.float32, device="cuda:1")
a1 = torch.empty(n, dtype=torch.float32, device="cuda:1"); b0 = torch.empty(n, dtype=torch.float32, device="cuda:0")
def run(nstreams, both, chunk_mb=None, iters=5):
    s0 = [torch.cuda.Stream(device=0) for _ in range(nstreams)]
    s1 = [torch.cuda.Stream(device=1) for _ in range(nstreams)]
    per = n // nstreams
    def issue():
        for k in range(nstreams):
            with torch.cuda.stream(s0[k]):
                b1[k*per:(k+1)*per].copy_(a0[k*per:(k+1)*per], non_blocking=True)
            if both:
                with torch.cuda.stream(s1[k]):
                    b0[k*per:(k+1)*per].copy_(a1[k*per:(k+1)*per], non_blocking=True)
    issue(); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    t0 = time.perf_counter()
    for _ in range(iters): issue()
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    dt = (time.perf_counter() - t0) / iters
    return n * 4 / dt / 1e9
for both in (False, True):
    for ns in (1, 2, 4, 8):
        print(f"both_directions={both} streams={ns}: {run(ns, both):.0f} GB/s per direction", flush=True)
