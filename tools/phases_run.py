"""Debug: per-phase cycle counts of k_parse (needs the instrumented build)."""
import argparse, ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from lightweight_snappy_b200 import api, corpus
ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=1024)
ap.add_argument("--kind", default="text")
a = ap.parse_args()
n = a.mib << 20
data = corpus.make_corpus(a.kind, n, device="cuda")
codec = api.DeviceCodec(n)
L = api.lib()
buf = (ctypes.c_ulonglong * 16)()
codec.compress(data, 0)
L.sb200_dbg_phases(buf, 1)
codec.compress(data, 0)
L.sb200_dbg_phases(buf, 1)
v = list(buf)
nb = v[10]
print(f"{a.kind}: blocks {nb}, cycles/block {v[11] / nb:.0f}, wide hits/block {v[13] / nb:.1f}, with match_extend {v[12] / nb:.1f}")
names = {0: "loop top/narrow miss", 1: "key load", 2: "hash+lds", 3: "match", 4: "fetch+compare", 5: "ballot", 6: "commit+shfl", 7: "match_extend", 8: "record", 15: "rest (other paths)"}
for i, nm in names.items():
    print(f"  {nm:22s} cycles/block {v[i] / nb:10.0f}  per wide hit {v[i] / max(v[13], 1):7.0f}  share {v[i] / v[11] * 100:5.1f}%")
