"""Multi-GPU sharding of the hot path: plain block partitioning.

Every 64 KiB block is compressed and decoded independently (reference:
src/snappy_compression.c:419-425 resets all state per block), so G ranks take contiguous block
ranges and never exchange data.  The only cross-rank datum is each partition's compressed byte
count (G integers): an exclusive scan of those places the partitions in one stream.  The first
rank's partition carries the varint preamble of the WHOLE input; the others emit bare blocks.
"""
from __future__ import annotations

BLOCK = 65536


def block_range(rank: int, world: int, n_blocks: int) -> tuple[int, int]:
    """Contiguous block range [lo, hi) of `rank`: [rank*n/world, (rank+1)*n/world)."""
    return rank * n_blocks // world, (rank + 1) * n_blocks // world


def byte_range(rank: int, world: int, n_bytes: int) -> tuple[int, int]:
    nb = (n_bytes + BLOCK - 1) // BLOCK
    lo, hi = block_range(rank, world, nb)
    return lo * BLOCK, min(hi * BLOCK, n_bytes)


def varint(n: int) -> bytes:
    """LEB128 preamble (reference: src/varint.c:12-20)."""
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def place_partitions(body_sizes: list[int], total_uncompressed: int) -> tuple[list[int], int]:
    """Stream offset of every rank's block bodies and the total stream length."""
    off = len(varint(total_uncompressed)) if total_uncompressed else 0
    offsets = []
    for s in body_sizes:
        offsets.append(off)
        off += s
    return offsets, off
