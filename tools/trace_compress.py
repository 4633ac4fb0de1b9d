import os, sys, time
sys.path.insert(0, ".")
import torch
from lightweight_snappy_b200 import api, corpus
n = 1 << 30
h_in = corpus.make_corpus("mixed", n, device="cuda").cpu().pin_memory()
c_out = torch.empty(api.max_compressed_bytes(n), dtype=torch.uint8).pin_memory()
for rep in range(3):
    if rep == 2: os.environ["SNAPPY_B200_TRACE"] = "1"
    t = time.perf_counter(); c = api.compress_host(h_in, 0, c_out.numpy()); dt = time.perf_counter() - t
    print(f"compress_host: {dt*1e3:.1f} ms", flush=True)
