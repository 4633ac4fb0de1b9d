"""Sweep of the compress_host chunk size (env knob) on 1 GiB mixed."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from lightweight_snappy_b200 import api, corpus
n = 1 << 30
data = corpus.make_corpus(sys.argv[1] if len(sys.argv) > 1 else "mixed", n, device="cuda")
h_in = data.cpu().pin_memory()
c_out = torch.empty(api.max_compressed_bytes(n), dtype=torch.uint8).pin_memory()
ref = None
for chunk in (128, 64, 32, 96, 48, 128):
    os.environ["SNAPPY_B200_CHUNK_MIB"] = str(chunk)
    best = 1e9
    for rep in range(4):
        t = time.perf_counter()
        c = api.compress_host(h_in, 0, c_out.numpy())
        best = min(best, time.perf_counter() - t)
    if ref is None:
        ref = c.copy()
    assert c.size == ref.size and (c == ref).all()
    print(f"chunk {chunk:4d} MiB: {best * 1e3:.2f} ms = {n / best / 1e9:.1f} GB/s", flush=True)
