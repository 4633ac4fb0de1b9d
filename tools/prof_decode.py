"""One compress + K0 + block decode of N MiB of a corpus: the command ncu profiles (decode kernels)."""
import argparse, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from lightweight_snappy_b200 import api, corpus

ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=96)
ap.add_argument("--kind", default="mixed")
a = ap.parse_args()
n = a.mib << 20
data = corpus.make_corpus(a.kind, n, device="cuda")
codec = api.DeviceCodec(n)
out = torch.empty(n, dtype=torch.uint8, device="cuda")
codec.compress(data, 0)
s = codec.result_stream().clone()
hdr = 1
while (n >> (7 * hdr)) > 0:
    hdr += 1
idx = torch.zeros_like(codec.block_offsets)
codec.index(s, s.numel(), hdr, n, idx)
codec.decode_segments(s, s.numel(), hdr, n, out, idx)
codec.check_status()
assert torch.equal(out, data)
print("ok", a.kind, a.mib)
