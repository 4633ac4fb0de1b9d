"""Sweep of the decompress_host piece schedule (env knobs) on 1 GiB mixed."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from lightweight_snappy_b200 import api, corpus
n = int(float(sys.argv[2]) * (1 << 30)) if len(sys.argv) > 2 else 1 << 30
data = corpus.make_corpus(sys.argv[1] if len(sys.argv) > 1 else "mixed", n, device="cuda")
h_in = data.cpu().pin_memory()
comp = api.compress_host(h_in, 0)
h_comp = torch.from_numpy(comp.copy()).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
out_np = h_out.numpy()
api.decompress_host(h_comp, out_np)
for start, growth, piece in [(24, 2, 192), (16, 2, 128), (8, 2, 64), (16, 2, 96), (8, 2, 128), (16, 3, 144), (24, 2, 192), (16, 2, 128)]:
    os.environ["SNAPPY_B200_PIECE_START_MIB"] = str(start)
    os.environ["SNAPPY_B200_PIECE_GROWTH"] = str(growth)
    os.environ["SNAPPY_B200_PIECE_MIB"] = str(piece)
    best = 1e9
    for rep in range(4):
        t = time.perf_counter()
        api.decompress_host(h_comp, out_np)
        best = min(best, time.perf_counter() - t)
    print(f"start {start:3d} growth {growth} piece {piece}: {best * 1e3:.2f} ms = {n / best / 1e9:.1f} GB/s", flush=True)
assert torch.equal(h_out, h_in)
