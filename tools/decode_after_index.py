import os, sys
sys.path.insert(0, ".")
import torch
from lightweight_snappy_b200 import api, corpus
n = 1 << 30
data = corpus.make_corpus("mixed", n, device="cuda")
codec = api.DeviceCodec(n)
out = torch.empty(n, dtype=torch.uint8, device="cuda")
codec.compress(data, 0)
s = codec.result_stream().clone(); c = s.numel()
idx = torch.zeros_like(codec.block_offsets)
hdr = 5
def once(with_index):
    if with_index:
        codec.index(s, c, hdr, n, idx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); codec.decode_segments(s, c, hdr, n, out, idx); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
codec.index(s, c, hdr, n, idx)
for w in (True, False, True, False):
    ts = sorted(once(w) for _ in range(7))
    print("index before each decode" if w else "decode only, repeated   ", [round(t, 3) for t in ts])
