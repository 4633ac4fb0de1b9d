"""One compress + one index + one decode of N MiB of a corpus: the command ncu profiles."""
import argparse, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from lightweight_snappy_b200 import api, corpus

ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=64)
ap.add_argument("--kind", default="mixed")
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
n = a.mib << 20
data = corpus.make_corpus(a.kind, n, device="cuda")
codec = api.DeviceCodec(n)
out = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(a.reps):
    ec = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ec[0].record()
    codec.compress(data, a.mode)
    ec[1].record()
    s = codec.result_stream()
    t_comp = ec[0].elapsed_time(ec[1])
    hdr = 1
    while (n >> (7 * hdr)) > 0:
        hdr += 1
    idx = torch.zeros_like(codec.block_offsets)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    codec.index(codec.stream_buf, s.numel(), hdr, n, idx)
    ev[1].record()
    codec.decompress_indexed(codec.stream_buf, idx, n, out)
    ev[2].record()
    codec.check_status()
    assert torch.equal(out, data)
    out.zero_()
    ev[3].record()
    codec.decompress(codec.stream_buf, s.numel(), hdr, n, out, idx)
    ev[4].record()
    codec.check_status()
    assert torch.equal(out, data)
    print(f"{a.kind} {a.mib} MiB mode {a.mode}: compress {t_comp:.3f} ms, ratio {n / s.numel():.3f}, K0 rounds {api.index_rounds()}, "
          f"index {ev[0].elapsed_time(ev[1]):.3f} ms, window decode {ev[1].elapsed_time(ev[2]):.3f} ms, "
          f"index+segment decode {ev[3].elapsed_time(ev[4]):.3f} ms")
