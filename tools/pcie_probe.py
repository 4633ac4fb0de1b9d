"""PCIe ceilings for the host API: pinned H2D / D2H alone and together, then the host calls."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from lightweight_snappy_b200 import api, corpus

n = 1 << 30
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best

def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
for name, fn, gb in (("H2D", h2d, 1), ("D2H", d2h, 1), ("H2D+D2H", both, 2)):
    t = timed(fn)
    print(f"{name:8s} {n * gb / t / 1e9:6.1f} GB/s total ({t * 1e3:.1f} ms)")

data = corpus.make_corpus("mixed", n, device="cuda")
h_in = data.cpu().pin_memory()
comp = api.compress_host(h_in, 0) if hasattr(api, "compress_host") else None
h_comp = torch.from_numpy(comp.copy()).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
out_np = h_out.numpy()
for rep in range(3):
    if rep == 2:
        os.environ["SNAPPY_B200_TRACE"] = "1"
    t = time.perf_counter()
    got = api.decompress_host(h_comp, out_np)
    dt = time.perf_counter() - t
    print(f"decompress_host: {dt * 1e3:.1f} ms = {n / dt / 1e9:.1f} GB/s (stream {h_comp.numel() >> 20} MiB)", flush=True)
os.environ.pop("SNAPPY_B200_TRACE", None)
assert torch.equal(h_out, h_in)
c_out = torch.empty(api.max_compressed_bytes(n), dtype=torch.uint8).pin_memory()
for rep in range(3):
    t = time.perf_counter()
    c = api.compress_host(h_in, 0, c_out.numpy())
    dt = time.perf_counter() - t
    print(f"compress_host: {dt * 1e3:.1f} ms = {n / dt / 1e9:.1f} GB/s", flush=True)
