"""Per-class timing of the two halves of index-less decompression (K0, block decode) and of compression.

    python tools/decode_bench.py [--mib 1024] [--kinds mixed,text,lowent,random] [--reps 5] [--mode 0]

Every number is a CUDA-event time on the launching stream after warm-up; the decoded bytes are compared
with the input before anything is timed.
"""
import argparse, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from lightweight_snappy_b200 import api, corpus

ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=1024)
ap.add_argument("--kinds", default="mixed,text,lowent,random")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--no-compress-timing", action="store_true")
a = ap.parse_args()
n = a.mib << 20
codec = api.DeviceCodec(n)
out = torch.empty(n, dtype=torch.uint8, device="cuda")
hdr = 1
while (n >> (7 * hdr)) > 0:
    hdr += 1


def ev_time(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


for kind in a.kinds.split(","):
    data = corpus.make_corpus(kind, n, device="cuda")
    codec.compress(data, a.mode)
    s = codec.result_stream().clone()
    c = s.numel()
    idx = torch.zeros_like(codec.block_offsets)
    codec.index(s, c, hdr, n, idx)
    out.zero_()
    codec.decode_segments(s, c, hdr, n, out, idx)
    codec.check_status()
    ok = torch.equal(out, data)
    t_k0 = ev_time(lambda: codec.index(s, c, hdr, n, idx), a.reps)
    t_dec = ev_time(lambda: codec.decode_segments(s, c, hdr, n, out, idx), a.reps)
    t_all = ev_time(lambda: codec.decompress(s, c, hdr, n, out, idx), a.reps)
    t_comp = float("nan") if a.no_compress_timing else ev_time(lambda: codec.compress(data, a.mode), max(2, a.reps // 2))
    print(f"{kind:14s} {a.mib} MiB mode {a.mode}: ok={ok} ratio {n / c:.3f} | K0 {t_k0:.3f} ms ({api.index_rounds()} rounds) | "
          f"decode {t_dec:.3f} ms = {(n + c) / t_dec / 1e6:.0f} GB/s (U+C) | decompress {t_all:.3f} ms = {n / t_all / 1e6:.0f} GB/s | "
          f"compress {t_comp:.3f} ms = {n / t_comp / 1e6:.0f} GB/s", flush=True)
