"""The N>1 path on CPU: world_size-2 gloo.  Each rank takes its block partition, "compresses"
it (the oracle stands in for the GPU here -- the host-side partition / placement logic is what
is under test), the ranks exchange only their compressed byte counts, and the assembled stream
must equal the single-process stream byte for byte."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_bytes, result_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import datasets
    import oracle_lib
    from lightweight_snappy_b200 import partition

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = oracle_lib.Oracle()
    data = datasets.gen(f"corpus:mixed:0:1040000:{n_bytes}")  # spans a class boundary
    lo, hi = partition.byte_range(rank, world, n_bytes)
    part = oracle.compress(data[lo:hi], 0)
    hdr = len(partition.varint(hi - lo))
    body = part[hdr:]  # a partition is a run of bare blocks
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([body.size], dtype=torch.int64))
    offsets, total = partition.place_partitions([int(s) for s in sizes], n_bytes)
    np.save(os.path.join(result_dir, f"part{rank}.npy"), body)
    dist.barrier()
    if rank == 0:
        stream = np.zeros(total, dtype=np.uint8)
        pre = partition.varint(n_bytes)
        stream[: len(pre)] = np.frombuffer(pre, np.uint8)
        for r in range(world):
            b = np.load(os.path.join(result_dir, f"part{r}.npy"))
            stream[offsets[r]: offsets[r] + b.size] = b
        whole = oracle.compress(data, 0)
        assert np.array_equal(stream, whole), "assembled stream differs from the single-process stream"
        assert np.array_equal(oracle.decompress(stream), data)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_bytes", [65536 * 5 + 777, 65536 * 2, 1000])
def test_two_rank_partition_matches_single_stream(tmp_path, n_bytes, oracle):
    port = 29500 + (os.getpid() + n_bytes) % 2000
    mp.spawn(_worker, args=(2, port, n_bytes, str(tmp_path)), nprocs=2, join=True)


def test_block_ranges_cover_everything():
    from lightweight_snappy_b200 import partition
    for nb in (0, 1, 7, 8, 16384, 16385):
        for world in (1, 2, 4, 8):
            edges = [partition.block_range(r, world, nb) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == nb
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            assert max(hi - lo for lo, hi in edges) - min(hi - lo for lo, hi in edges) <= 1
