"""GPU parity at BASELINE.json's full sizes, against the CPU oracle (not against another CUDA path).

The oracle is sharded over the host cores: every 64 MiB shard of the corpus is regenerated on the CPU
(the corpus is a pure function of seed and segment index), compressed by the oracle, and the shard's
block bodies are compared with the corresponding slice of the GPU stream through SHA-256 digests.
  configs[1]  1 GiB mixed, decompression: both decoders' output == original == oracle decoder output
  configs[2]  1 GiB mixed, hash compression: stream byte-identical to the oracle's
  configs[3]  1 GiB low-entropy + random, BST path: stream byte-identical to the oracle's `-b`
"""
import hashlib
import os

import numpy as np
import pytest

from lightweight_snappy_b200 import api, corpus

pytestmark = pytest.mark.gpu

GIB = 1 << 30
SHARD = 64 << 20
SEED = 20261018


def _shard_worker(args):
    kind, shard, mode = args
    import torch
    torch.set_num_threads(1)
    import oracle_lib
    from lightweight_snappy_b200 import corpus as cp
    data = cp.make_corpus(kind, SHARD, seed=SEED, first_segment=shard * (SHARD >> 20)).numpy()
    o = oracle_lib.Oracle()
    stream, sizes = o.compress(data, mode, with_sizes=True)
    hdr = len(o.varint_encode(data.size))
    back = o.decompress(stream, data.size)
    return {"shard": shard, "body_len": int(stream.size - hdr), "body_sha": hashlib.sha256(stream[hdr:].tobytes()).hexdigest(),
            "data_sha": hashlib.sha256(data.tobytes()).hexdigest(), "sizes_sum": int(sizes.sum()),
            "oracle_decodes": bool(back.size == data.size and np.array_equal(back, data))}


def _oracle_shards(kind: str, mode: int, n: int):
    import multiprocessing as mp
    import oracle_lib
    if not os.path.exists(oracle_lib.ORACLE_SO):
        oracle_lib.build(ref=False)
    ctx = mp.get_context("spawn")
    with ctx.Pool(min(os.cpu_count() or 1, n // SHARD)) as pool:
        return pool.map(_shard_worker, [(kind, s, mode) for s in range(n // SHARD)])


def _check_stream_against_oracle(kind: str, mode: int, n: int = GIB):
    import torch
    data = corpus.make_corpus(kind, n, seed=SEED, device="cuda")
    codec = api.DeviceCodec(n)
    codec.compress(data, mode)
    stream = codec.result_stream().clone()
    offs = codec.block_offsets[: api.block_count(n) + 1].cpu().numpy()
    host_stream = stream.cpu().numpy()
    shards = _oracle_shards(kind, mode, n)
    bps = SHARD // 65536  # blocks per shard
    for r in shards:
        s = r["shard"]
        lo, hi = int(offs[s * bps]), int(offs[(s + 1) * bps])
        assert r["oracle_decodes"], f"oracle round trip, shard {s}"
        assert hi - lo == r["body_len"], f"{kind} mode {mode}: shard {s} compressed size {hi - lo} vs oracle {r['body_len']}"
        assert hashlib.sha256(host_stream[lo:hi].tobytes()).hexdigest() == r["body_sha"], \
            f"{kind} mode {mode}: shard {s} bytes differ from the oracle's"
        got_data = data[s * SHARD:(s + 1) * SHARD].cpu().numpy()
        assert hashlib.sha256(got_data.tobytes()).hexdigest() == r["data_sha"], f"corpus shard {s}: CPU and GPU generators differ"
    total_c = sum(r["body_len"] for r in shards)
    hdr = int(offs[0])
    assert host_stream.size == hdr + total_c
    return data, codec, stream, hdr, total_c


def test_config2_config1_full_gib_mixed():
    import torch
    n = GIB
    data, codec, stream, hdr, total_c = _check_stream_against_oracle("mixed", api.MODE_HASH, n)
    side = codec.block_offsets[: api.block_count(n) + 1].clone()
    # configs[1]: the index-less decoder (K0 + block decode) and the indexed decoder, output == original
    out = torch.zeros(n, dtype=torch.uint8, device="cuda")
    k0 = torch.zeros_like(side)
    codec.decompress(stream, stream.numel(), hdr, n, out, k0)
    codec.check_status()
    assert torch.equal(out, data), "index-less decode of the 1 GiB stream"
    assert torch.equal(k0, side), "K0 block offsets vs the compressor's"
    out.zero_()
    codec.decompress_indexed(stream, side, n, out)
    codec.check_status()
    assert torch.equal(out, data), "indexed decode of the 1 GiB stream"


def test_config3_full_gib_lowent_random_bst():
    import torch
    n = GIB
    data, codec, stream, hdr, total_c = _check_stream_against_oracle("lowent_random", api.MODE_BST, n)
    ratio = n / stream.numel()
    assert 1.6 < ratio < 1.8, ratio
    out = torch.zeros(n, dtype=torch.uint8, device="cuda")
    codec.decompress(stream, stream.numel(), hdr, n, out)
    codec.check_status()
    assert torch.equal(out, data)
