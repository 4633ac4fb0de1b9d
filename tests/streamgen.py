"""Random but valid raw-Snappy streams in this framing (elements never straddle a 64 KiB output
block), exercising element kinds the compressors never emit: copy-4, non-minimal literal length
encodings (including the 4-byte form), copy-2 with short lengths, offsets up to the block start,
self-overlapping copies with every small period, long literals in the middle of a block."""
from __future__ import annotations

import numpy as np

BLOCK = 65536


def _varint(n: int) -> bytes:
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def _literal(rng, data: bytes) -> bytes:
    m = len(data) - 1
    forms = []
    if m < 60:
        forms.append(bytes([m << 2]))
    if m < 1 << 8:
        forms.append(bytes([60 << 2, m]))
    if m < 1 << 16:
        forms.append(bytes([61 << 2]) + m.to_bytes(2, "little"))
    if m < 1 << 24:
        forms.append(bytes([62 << 2]) + m.to_bytes(3, "little"))
    forms.append(bytes([63 << 2]) + m.to_bytes(4, "little"))
    # mostly the minimal form, sometimes a wider one
    pick = 0 if rng.random() < 0.7 else int(rng.integers(0, len(forms)))
    return forms[pick] + data


def _copy(rng, length: int, off: int) -> bytes:
    forms = []
    if 4 <= length <= 11 and off < 2048:
        forms.append(bytes([((off >> 8) << 5) | ((length - 4) << 2) | 1, off & 0xFF]))
    if off < 65536:
        forms.append(bytes([((length - 1) << 2) | 2]) + off.to_bytes(2, "little"))
    forms.append(bytes([((length - 1) << 2) | 3]) + off.to_bytes(4, "little"))
    return forms[int(rng.integers(0, len(forms)))]


def make_crossing_stream(seed: int, total: int) -> tuple[np.ndarray, np.ndarray]:
    """Valid raw Snappy that is NOT framed in 64 KiB blocks: elements straddle the 64 KiB output
    boundaries, copies reach back further than 64 KiB (copy-4 with 32-bit offsets) and into earlier
    blocks.  The reference decoder resolves all of it in its whole-file buffer
    (src/snappy_decompression.c:253-280, :323-327).  Returns (stream, expected output)."""
    rng = np.random.default_rng(seed)
    out = bytearray()
    stream = bytearray(_varint(total))
    while len(out) < total:
        room = total - len(out)
        if len(out) == 0 or rng.random() < 0.3:
            n = int(min(room, rng.choice([1, 7, 60, 61, 300, 5000, 70000])))
            data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
            stream += _literal(rng, data)
            out += data
        else:
            n = int(min(room, rng.integers(1, 65)))
            r = rng.random()
            if r < 0.3:
                off = int(rng.integers(1, min(len(out), 70) + 1))
            elif r < 0.6:
                off = int(rng.integers(1, min(len(out), 65535) + 1))
            else:
                off = int(rng.integers(1, len(out) + 1))  # anywhere in the output so far: often > 64 KiB back
            stream += _copy(rng, n, off)
            start = len(out) - off
            for i in range(n):
                out.append(out[start + i])
    return np.frombuffer(bytes(stream), np.uint8).copy(), np.frombuffer(bytes(out), np.uint8).copy()


def make_stream(seed: int, total: int, style: str = "mixed") -> tuple[np.ndarray, np.ndarray]:
    """Returns (stream, expected output)."""
    rng = np.random.default_rng(seed)
    out = bytearray()
    stream = bytearray(_varint(total))
    while len(out) < total:
        block_start = len(out)
        block_len = min(BLOCK, total - block_start)
        produced = 0
        while produced < block_len:
            room = block_len - produced
            r = rng.random()
            if produced == 0 or r < (0.15 if style == "copies" else 0.4):
                if style == "bigliteral" and rng.random() < 0.3:
                    n = int(min(room, rng.integers(60, 70000)))
                else:
                    n = int(min(room, rng.choice([1, 2, 3, 5, 17, 59, 60, 61, 62, 63, 64, 65, 255, 256, 257, 1000])))
                data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
                stream += _literal(rng, data)
                out += data
                produced += n
            else:
                n = int(min(room, rng.integers(1, 65)))
                if rng.random() < 0.35:
                    off = int(rng.integers(1, min(produced, 70) + 1))  # short periods, often overlapping
                else:
                    off = int(rng.integers(1, produced + 1))
                stream += _copy(rng, n, off)
                start = len(out) - off
                for i in range(n):
                    out.append(out[start + i])
                produced += n
    return np.frombuffer(bytes(stream), np.uint8).copy(), np.frombuffer(bytes(out), np.uint8).copy()
