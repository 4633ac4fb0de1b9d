"""Named, deterministic test inputs.  A spec string fully determines the bytes, so the
golden fixtures only need to store specs and digests.

  hex:<hexbytes>                 literal bytes
  rep:<byte>:<count>             one byte value repeated
  corpus:<kind>:<segment>:<offset>:<size>   slice of a corpus segment (corpus.py)
  sym:<k>:<seed>:<size>          uniform random symbols from an alphabet of k bytes
  period:<p>:<seed>:<size>       a random p-byte pattern repeated
  lz:<seed>:<size>               LZ-synthetic: random runs and back-references
"""
from __future__ import annotations

import numpy as np
import torch

from lightweight_snappy_b200 import corpus


def _rand(seed: int, stream: int, n: int) -> np.ndarray:
    return corpus._draw(seed, 0, stream, n, torch.device("cpu")).numpy()


def gen(spec: str) -> np.ndarray:
    kind, _, rest = spec.partition(":")
    if kind == "hex":
        return np.frombuffer(bytes.fromhex(rest), dtype=np.uint8).copy()
    a = rest.split(":")
    if kind == "rep":
        return np.full(int(a[1]), int(a[0]), dtype=np.uint8)
    if kind == "corpus":
        ck, seg, off, size = a[0], int(a[1]), int(a[2]), int(a[3])
        out = np.empty(size, dtype=np.uint8)
        done = 0
        while done < size:
            s = corpus.make_segment(ck, seg).numpy()
            take = min(size - done, corpus.SEGMENT - off)
            out[done:done + take] = s[off:off + take]
            done += take
            seg += 1
            off = 0
        return out
    if kind == "sym":
        k, seed, size = int(a[0]), int(a[1]), int(a[2])
        alphabet = (_rand(seed, 11, k) % 256).astype(np.uint8)
        return alphabet[_rand(seed, 12, size) % k]
    if kind == "period":
        p, seed, size = int(a[0]), int(a[1]), int(a[2])
        pat = (_rand(seed, 13, p) % 256).astype(np.uint8)
        return np.resize(pat, size)
    if kind == "lz":
        seed, size = int(a[0]), int(a[1])
        r = _rand(seed, 14, 4 * (size // 4 + 16))
        out = np.empty(size + 300, dtype=np.uint8)
        n, j = 0, 0
        while n < size:
            t, ln, x = int(r[j] % 3), int(r[j + 1] % 40) + 1, int(r[j + 2])
            j += 3
            if t == 0 or n < 8:  # fresh random bytes
                out[n:n + ln] = (_rand(seed + x % 9973, 15, ln) % 256).astype(np.uint8)
            else:  # back-reference, possibly overlapping
                off = x % min(n, 70000) + 1
                for q in range(ln):
                    out[n + q] = out[n + q - off]
            n += ln
        return out[:size].copy()
    raise ValueError(spec)


def boundary_sizes() -> list[int]:
    s = [1, 2, 3, 4, 5, 14, 15, 16, 17, 18, 31, 32, 33, 63, 64, 65, 255, 256, 257, 511, 512, 513,
         1023, 1024, 1025, 2047, 2048, 2049, 4095, 4096, 4097, 8191, 8192, 8193, 65521, 65535, 65536,
         65537, 65551, 65552, 70000, 131071, 131072, 131073, 131087, 200000, 262147]
    return s
