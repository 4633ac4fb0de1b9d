"""Deterministic synthetic corpora for the Snappy hot path.

Workload generation only -- not part of the codec.  Every 1 MiB segment is a pure
function of (seed, segment index, class), built from a counter-based integer hash
(splitmix64 finaliser), so the same bytes come out of the CPU (tests, spot checks)
and of the GPU (bench.py generates the 1 GiB corpus directly in HBM; nothing has to
cross PCIe).  Only integer arithmetic, searchsorted and gathers are used, so torch
CPU and torch CUDA agree bit for bit.

Classes:
  text     8192-word vocabulary, word length 2..10, letters a-z, words drawn
           Zipf(s=1.05), joined by single spaces, '\n' after every 2**20-th word
  lowent   symbols {0,1,2,255} with p = {.70,.15,.10,.05}, each repeated 1..15 times
  random   uniform bytes
  mixed    segment i has class (text, lowent, random)[i % 3]
  lowent_random   segment i has class (lowent, random)[i % 2]   (BASELINE config 4)
"""
from __future__ import annotations

import functools

import torch

SEGMENT = 1 << 20
_M64 = (1 << 64) - 1


def _s64(x: int) -> int:
    """Python int -> two's-complement signed 64-bit value."""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


_GOLD = _s64(0x9E3779B97F4A7C15)
_C1 = _s64(0xBF58476D1CE4E5B9)
_C2 = _s64(0x94D049BB133111EB)


def _lsr(x: torch.Tensor, k: int) -> torch.Tensor:
    return (x >> k) & ((1 << (64 - k)) - 1)


def _mix(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors (wrapping arithmetic)."""
    z = x + _GOLD
    z = (z ^ _lsr(z, 30)) * _C1
    z = (z ^ _lsr(z, 27)) * _C2
    return z ^ _lsr(z, 31)


def _mix_int(x: int) -> int:
    z = (x + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def _key(seed: int, segment: int, stream: int) -> int:
    return _s64(_mix_int(_mix_int(_mix_int(seed) ^ (segment * 0x100000001B3)) ^ (stream * 0xD6E8FEB86659FD93)))


def _draw(seed: int, segment: int, stream: int, n: int, device) -> torch.Tensor:
    """n pseudo-random non-negative 63-bit integers."""
    ctr = torch.arange(n, dtype=torch.int64, device=device)
    return _lsr(_mix(ctr * _C2 + _key(seed, segment, stream)), 1)


def _iroot(n: int, k: int) -> int:
    """floor(n ** (1/k)) on Python ints (exact)."""
    if n < 2:
        return n
    x = 1 << -(-n.bit_length() // k)
    while True:
        y = ((k - 1) * x + n // x ** (k - 1)) // k
        if y >= x:
            return x
        x = y


VOCAB = 8192


@functools.lru_cache(maxsize=4)
def _vocabulary(seed: int):
    """(letters [VOCAB, 10] uint8, lengths [VOCAB] int64, cdf [VOCAB] int64) on CPU."""
    ctr = torch.arange(VOCAB * 11, dtype=torch.int64)
    h = _lsr(_mix(ctr * _C1 + _key(seed, -1, 7)), 1).view(VOCAB, 11)
    lengths = 2 + h[:, 0] % 9
    letters = (97 + h[:, 1:] % 26).to(torch.uint8)
    # Zipf(s=1.05) weights as exact integers: w_j = 2**70 / (j+1)**1.05
    weights = [(1 << 70) // _iroot(((j + 1) ** 21) << 400, 20) for j in range(VOCAB)]
    total = sum(weights)
    acc, cdf = 0, []
    for w in weights:
        acc += w
        cdf.append((acc << 40) // total)
    cdf[-1] = 1 << 40
    return letters, lengths, torch.tensor(cdf, dtype=torch.int64)


def _segment_text(seed: int, segment: int, device) -> torch.Tensor:
    letters, lengths, cdf = _vocabulary(seed)
    letters, lengths, cdf = letters.to(device), lengths.to(device), cdf.to(device)
    n_words = SEGMENT // 3 + 2  # shortest word + separator is 3 bytes
    u = _draw(seed, segment, 1, n_words, device) & ((1 << 40) - 1)
    word = torch.searchsorted(cdf, u, right=True).clamp_(max=VOCAB - 1)
    wlen = lengths[word]
    ends = torch.cumsum(wlen + 1, 0)
    pos = torch.arange(SEGMENT, dtype=torch.int64, device=device)
    wi = torch.searchsorted(ends, pos, right=True)
    within = pos - (ends[wi] - (wlen[wi] + 1))
    is_sep = within >= wlen[wi]
    ch = letters[word[wi], within.clamp(max=9)]
    gw = wi + segment * n_words
    sep = torch.where((gw & ((1 << 20) - 1)) == (1 << 20) - 1, 10, 32).to(torch.uint8)
    return torch.where(is_sep, sep, ch)


def _segment_lowent(seed: int, segment: int, device) -> torch.Tensor:
    n_runs = SEGMENT // 4  # expected 2 MiB of output; a shortfall is checked below
    h = _draw(seed, segment, 2, n_runs, device)
    p = h % 100
    sym = torch.where(p < 70, 0, torch.where(p < 85, 1, torch.where(p < 95, 2, 255))).to(torch.uint8)
    rlen = 1 + (h >> 8) % 15
    ends = torch.cumsum(rlen, 0)
    if int(ends[-1]) < SEGMENT:
        raise RuntimeError("low-entropy segment came up short")
    pos = torch.arange(SEGMENT, dtype=torch.int64, device=device)
    return sym[torch.searchsorted(ends, pos, right=True)]


def _segment_random(seed: int, segment: int, device) -> torch.Tensor:
    ctr = torch.arange(SEGMENT // 8, dtype=torch.int64, device=device)
    return _mix(ctr * _C2 + _key(seed, segment, 3)).view(torch.uint8)


_CLASSES = {"text": _segment_text, "lowent": _segment_lowent, "random": _segment_random}
_MIXES = {"mixed": ("text", "lowent", "random"), "lowent_random": ("lowent", "random")}
KINDS = tuple(_CLASSES) + tuple(_MIXES)


def segment_class(kind: str, segment: int) -> str:
    if kind in _CLASSES:
        return kind
    mix = _MIXES[kind]
    return mix[segment % len(mix)]


def make_segment(kind: str, segment: int, seed: int = 20261018, device="cpu") -> torch.Tensor:
    """One 1 MiB segment (uint8 tensor) of the given corpus."""
    return _CLASSES[segment_class(kind, segment)](seed, segment, torch.device(device))


def make_corpus(kind: str, n_bytes: int, seed: int = 20261018, device="cpu", first_segment: int = 0,
                out: torch.Tensor | None = None) -> torch.Tensor:
    """n_bytes of corpus `kind` starting at segment `first_segment` (truncated to size)."""
    device = torch.device(device)
    if out is None:
        out = torch.empty(n_bytes, dtype=torch.uint8, device=device)
    done = 0
    seg = first_segment
    while done < n_bytes:
        take = min(SEGMENT, n_bytes - done)
        out[done:done + take] = make_segment(kind, seg, seed, device)[:take]
        done += take
        seg += 1
    return out
