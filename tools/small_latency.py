"""Steady-state latency of the host calls on small inputs."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from lightweight_snappy_b200 import api, corpus
for n in (4096, 65536, 1 << 20, 16 << 20):
    data = corpus.make_corpus("mixed", max(n, 1 << 20), device="cpu")[:n].numpy().copy()
    comp = api.compress_host(data, 0).copy()
    out = np.empty(n, dtype=np.uint8)
    for name, fn in (("compress", lambda: api.compress_host(data, 0)), ("decompress", lambda: api.decompress_host(comp, out))):
        fn(); fn()
        ts = []
        for _ in range(20):
            t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
        ts.sort()
        print(f"{n:9d} B {name:10s} median {ts[10] * 1e3:7.3f} ms  min {ts[0] * 1e3:7.3f} ms", flush=True)
