"""End-to-end decompress_host / compress_host on 1 GiB of pinned host memory for the current environment knobs."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from lightweight_snappy_b200 import api, corpus
n = 1 << 30
data = corpus.make_corpus("mixed", n, device="cuda")
h_in = data.cpu().pin_memory()
comp = api.compress_host(h_in, 0)
h_comp = torch.from_numpy(comp.copy()).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
ts = []
for rep in range(6):
    t = time.perf_counter(); api.decompress_host(h_comp, h_out.numpy()); ts.append(time.perf_counter() - t)
assert torch.equal(h_out, h_in)
print({k: v for k, v in os.environ.items() if k.startswith("SNAPPY_B200_")}, f"decompress_host best {min(ts[1:])*1e3:.2f} ms median {sorted(ts[1:])[2]*1e3:.2f} ms", flush=True)
