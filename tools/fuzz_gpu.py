"""Randomised parity run on the GPU: many inputs of random size and structure, both compressors
against the oracle byte for byte, and both decoders on the result.  (A longer-running sibling of
tests/test_gpu_parity.py::test_fuzz_against_oracle.)   usage: fuzz_gpu.py [n_cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import datasets, oracle_lib
from lightweight_snappy_b200 import api

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
oracle_lib.build()
orc = oracle_lib.Oracle()
done = 0
for case in range(n_cases):
    size = int(rng.choice([rng.integers(1, 64), rng.integers(1, 5000), rng.integers(60000, 70000), rng.integers(1, 300000)]))
    kind = rng.integers(0, 6)
    seed = int(rng.integers(0, 1 << 30))
    if kind == 0:
        spec = f"corpus:{['text', 'lowent', 'random', 'mixed'][rng.integers(0, 4)]}:{rng.integers(0, 50)}:{rng.integers(0, 1 << 19)}:{size}"
    elif kind == 1:
        spec = f"sym:{int(rng.choice([1, 2, 3, 4, 8, 16, 64, 256]))}:{seed}:{size}"
    elif kind == 2:
        spec = f"period:{int(rng.integers(1, 5000))}:{seed}:{size}"
    elif kind == 3:
        spec = f"rep:{int(rng.integers(0, 256))}:{size}"
    else:
        spec = f"lz:{seed}:{size}"
    try:
        data = datasets.gen(spec)
    except Exception:
        spec = f"sym:4:{seed}:{size}"
        data = datasets.gen(spec)
    for mode, fn in ((0, api.snappy_compress), (1, api.snappy_compress_bst)):
        want = orc.compress(data, mode)
        got = fn(data)
        assert got.size == want.size and np.array_equal(got, want), f"{spec} mode {mode}: stream differs"
        back = api.snappy_decompress(got)
        assert back.size == data.size and np.array_equal(back, data), f"{spec} mode {mode}: decode differs"
    done += 1
print(f"{done} cases x 2 modes: streams identical to the oracle, round trips exact")
