"""BASELINE.json configs[4]: a batch of independent 64 KiB blocks (mixed corpus) sharded over the
GPUs of one box by plain block partitioning, device-resident, compress then decompress.

    python tools/batch64.py --gib 64                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tools/batch64.py --gib 64                      # N GPUs, 64/N GiB each

Rank r owns the block range [r*n/G, (r+1)*n/G) (partition.block_range) and walks it in
sub-batches that fit in HBM next to the codec workspace.  No collective on the data path: NCCL
carries the barrier and the max / sum of the timings.  Checks:
  * every sub-batch round-trips (output == input, K0 index == compressor's side index);
  * >= 1 % of the blocks, picked by a seeded RNG, are compressed by the oracle on the host and
    compared byte for byte with the bytes the GPU produced for that block;
  * a whole-batch checksum of input and output (sum of per-sub-batch 64-bit sums) agrees.
Prints one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GIB = 1 << 30
BLOCK = 1 << 16
SEED = 20261018


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=64.0, help="size of the whole batch")
    ap.add_argument("--sub-gib", type=float, default=4.0, help="device-resident sub-batch per step")
    ap.add_argument("--sample", type=float, default=0.01, help="fraction of blocks checked against the oracle")
    ap.add_argument("--mode", type=int, default=0)
    args = ap.parse_args()
    res = run(args, own_process_group=True)
    if res is not None:
        print(json.dumps(res), flush=True)


def run(args, own_process_group: bool = False, sample_clocks: bool = True):
    """The whole sweep step for this rank; returns the result dict on rank 0 (None elsewhere).  bench.py calls
    it with its own process group already up (own_process_group=False)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from lightweight_snappy_b200 import api, corpus, partition

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and own_process_group:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    import oracle_lib  # tests/oracle_lib.py: the checker, never the thing measured
    if not os.path.exists(oracle_lib.ORACLE_SO):
        oracle_lib.build(ref=False)
    orc = oracle_lib.Oracle()

    n_blocks = int(args.gib * GIB) // BLOCK
    b_lo, b_hi = partition.block_range(rank, world, n_blocks)
    seg_blocks = corpus.SEGMENT // BLOCK  # corpus segments are 1 MiB = 16 blocks
    sub_blocks = max(seg_blocks, int(args.sub_gib * GIB) // BLOCK // seg_blocks * seg_blocks)
    sub_bytes = sub_blocks * BLOCK
    codec = api.DeviceCodec(sub_bytes, device=dev)
    out = torch.empty(sub_bytes, dtype=torch.uint8, device=dev)
    idx = torch.zeros_like(codec.block_offsets)
    rng = np.random.default_rng(SEED + rank)

    # one untimed call of each direction (context, module load, first-touch of the workspace)
    warm = corpus.make_corpus("mixed", 16 * BLOCK, seed=SEED, device=dev)
    codec.compress(warm, args.mode)
    ws = codec.result_stream()
    codec.decompress(ws, ws.numel(), 3, warm.numel(), out, idx)
    codec.check_status()

    from bench import ClockSampler  # nvidia-smi clocks / throttle reasons during the measured section
    sampler = ClockSampler(local) if rank == 0 and sample_clocks else None
    ev = lambda: torch.cuda.Event(enable_timing=True)
    t_comp = t_decomp = 0.0
    c_total = 0
    sum_in = sum_out = 0
    checked = 0
    b = b_lo
    assert b_lo % seg_blocks == 0 or world == 1, "partition must fall on corpus segments"
    while b < b_hi:
        nb = min(sub_blocks, b_hi - b)
        n = nb * BLOCK
        data = corpus.make_corpus("mixed", n, seed=SEED, device=dev, first_segment=b // seg_blocks)
        e = [ev() for _ in range(4)]
        e[0].record()
        codec.compress(data, args.mode)
        e[1].record()
        stream = codec.result_stream()
        c_bytes = stream.numel()
        hdr = 1
        while (n >> (7 * hdr)) > 0:
            hdr += 1
        e[2].record()
        codec.decompress(stream, c_bytes, hdr, n, out, idx)
        e[3].record()
        codec.check_status()
        torch.cuda.synchronize(dev)
        t_comp += e[0].elapsed_time(e[1])
        t_decomp += e[2].elapsed_time(e[3])
        c_total += c_bytes - hdr
        assert torch.equal(out[:n], data), f"round trip failed in sub-batch at block {b}"
        side = codec.block_offsets[: nb + 1]
        assert torch.equal(idx[: nb + 1], side), "K0 index differs from the compressor's side index"
        sum_in += int(data.view(torch.int64).sum().item()) & ((1 << 64) - 1)
        sum_out += int(out[:n].view(torch.int64).sum().item()) & ((1 << 64) - 1)
        # oracle spot check on >= sample of the blocks of this sub-batch
        k = max(1, int(np.ceil(nb * args.sample)))
        picks = np.sort(rng.choice(nb, size=k, replace=False))
        offs = side.cpu().numpy()
        for j in picks:
            lo, hi = int(offs[j]), int(offs[j + 1])
            got = stream[lo:hi].cpu().numpy()
            blk = data[j * BLOCK:(j + 1) * BLOCK].cpu().numpy()
            ref = np.asarray(orc.compress(blk, args.mode))
            vl = 1
            while (blk.size >> (7 * vl)) > 0:
                vl += 1
            assert np.array_equal(got, ref[vl:]), f"block {b + j}: GPU bytes differ from the oracle's"
        checked += k
        b += nb

    def red(x: float, op) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    R = dist.ReduceOp if world > 1 else None
    tc = red(t_comp, R.MAX if R else None)
    td = red(t_decomp, R.MAX if R else None)
    total_c = red(float(c_total), R.SUM if R else None)
    total_checked = red(float(checked), R.SUM if R else None)
    ok = red(1.0 if sum_in == sum_out else 0.0, R.MIN if R else None)
    u = n_blocks * BLOCK
    clocks = sampler.stop() if sampler else None
    result = None
    if rank == 0:
        result = ({
            "clocks": clocks,
            "config": f"{args.gib:g} GiB batch = {n_blocks} independent 64 KiB blocks (mixed corpus), "
                      f"{world} GPU(s), block ranges per rank, sub-batches of {sub_bytes / GIB:g} GiB, device-resident",
            "n_gpus": world, "mode": "hash" if args.mode == 0 else "bst",
            "compress_GBps": u / tc / 1e6, "decompress_GBps": u / td / 1e6,
            "compress_ms": tc, "decompress_ms": td, "ratio": u / total_c,
            "oracle_checked_blocks": int(total_checked), "oracle_checked_fraction": total_checked / n_blocks,
            "checksum_ok": bool(ok), "roundtrip_ok": True,
            "timing": "sum over sub-batches of the CUDA-event span around the codec call, max over ranks",
        })
    if world > 1 and own_process_group:
        dist.destroy_process_group()
    return result


if __name__ == "__main__":
    main()
