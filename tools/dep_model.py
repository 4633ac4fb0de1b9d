"""Dependency structure of reference-compressed blocks (design aid for the decoder).

For one 64 KiB block of each corpus class: element counts, copy offsets, the data-flow depth of the
copy graph (element level), and the number of in-order multi-round-resolution rounds a warp would
need when it takes the elements of one 128-byte stream segment at a time.
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle_lib import Oracle
import lightweight_snappy_b200 as pkg  # noqa
from importlib import import_module
corpus = import_module("lightweight_snappy_b200").corpus if hasattr(import_module("lightweight_snappy_b200"), "corpus") else None


def parse(stream):
    """-> list of (kind, stream_pos, out_pos, length, offset)"""
    i = 0
    v = 0; sh = 0
    while True:
        b = stream[i]; i += 1
        v |= (b & 0x7f) << sh; sh += 7
        if b < 128: break
    out = 0; els = []
    n = len(stream)
    while i < n:
        t = stream[i]; k = t & 3; s = i
        if k == 0:
            l = t >> 2
            if l < 60: l += 1; i += 1
            else:
                nb = l - 59
                l = int.from_bytes(stream[i+1:i+1+nb], "little") + 1; i += 1 + nb
            els.append((0, s, out, l, 0)); i += l
        elif k == 1:
            l = ((t >> 2) & 7) + 4; off = ((t >> 5) << 8) | stream[i+1]; i += 2
            els.append((1, s, out, l, off))
        elif k == 2:
            l = (t >> 2) + 1; off = stream[i+1] | (stream[i+2] << 8); i += 3
            els.append((1, s, out, l, off))
        else:
            l = (t >> 2) + 1; off = int.from_bytes(stream[i+1:i+5], "little"); i += 5
            els.append((1, s, out, l, off))
        out += l
    return els


def analyse(name, data):
    o = Oracle()
    st = bytes(o.compress(data, 0))
    els = parse(st)
    n = len(data)
    ncopy = sum(1 for e in els if e[0]); nlit = len(els) - ncopy
    # data-flow depth at byte level, then per element
    lvl = np.zeros(n, dtype=np.int32)
    owner = np.zeros(n, dtype=np.int32)
    depth = []
    for ei, (k, s, op, l, off) in enumerate(els):
        owner[op:op+l] = ei
        if k == 0:
            lvl[op:op+l] = 0; depth.append(0)
        else:
            a = op - off; b = min(a + l, op)
            d = int(lvl[a:b].max()) + 1
            lvl[op:op+l] = d; depth.append(d)
    # MRR per 128-byte segment windows, in-order HWM, literals pre-placed
    seg_rounds = []; seg_n = []
    cur = None; members = []
    def flush(members):
        if not members: return
        pend = [m for m in members if els[m][0] == 1]
        seg_n.append(len(members))
        rounds = 0
        done_upto = None
        while pend:
            hwm = els[pend[0]][2]  # out start of first unfinished copy
            # elements after it that are literals are done already, but in-order rule: only below hwm is safe...
            nxt = []
            for m in pend:
                k, s, op, l, off = els[m]
                src_end = min(op - off + l, op)
                if src_end <= hwm or m == pend[0]:
                    pass
                else:
                    nxt.append(m)
            pend = nxt; rounds += 1
        seg_rounds.append(rounds)
    for ei, e in enumerate(els):
        sg = e[1] // 128
        if sg != cur:
            flush(members); members = []; cur = sg
        members.append(ei)
    flush(members)
    offs = np.array([e[4] for e in els if e[0]])
    lens = np.array([e[3] for e in els if e[0]])
    llens = np.array([e[3] for e in els if e[0] == 0])
    print(f"{name}: comp {len(st)} B, {nlit} literals (mean {llens.mean():.1f} B, max {llens.max()}), {ncopy} copies "
          f"(mean len {lens.mean() if ncopy else 0:.1f}); offset pct 10/50/90 = "
          f"{np.percentile(offs,[10,50,90]) if ncopy else 0}; overlapping {int((offs<lens).sum()) if ncopy else 0}")
    print(f"   data-flow depth max {max(depth)}, mean {np.mean(depth):.1f}; segments {len(seg_rounds)}, "
          f"elements/segment mean {np.mean(seg_n):.1f} max {max(seg_n)}; MRR rounds/segment mean {np.mean(seg_rounds):.2f} "
          f"max {max(seg_rounds)} total {sum(seg_rounds)}")


if __name__ == "__main__":
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "lightweight-snappy_b200"))
    import corpus as cp
    for kind, seg in (("text", 0), ("lowent", 1), ("random", 2)):
        d = cp.make_segment(kind, seg).numpy()[:65536].tobytes()
        analyse(kind, d)


def window_rounds(els, gw, wkeep):
    """Rounds of in-order multi-round resolution when the elements are taken gw at a time by one
    sub-warp group; literals and copies that reach further back than `wkeep` are ready at once."""
    total = 0; passes = 0
    i = 0
    n = len(els)
    while i < n:
        if els[i][0] == 0 and els[i][3] > 64:
            i += 1; passes += 1; total += 1; continue
        j = i
        while j < n and j - i < gw and not (els[j][0] == 0 and els[j][3] > 64):
            j += 1
        pend = [m for m in range(i, j) if els[m][0] == 1 and els[m][4] <= wkeep]
        rounds = 1  # round 0: literals, far copies and whatever is ready
        first = True
        while pend:
            hwm = els[pend[0]][2]
            nxt = [m for m in pend if m != pend[0] and min(els[m][2] - els[m][4] + els[m][3], els[m][2]) > hwm]
            if not first:
                rounds += 1
            first = False
            pend = nxt
        total += rounds; passes += 1
        i = j
    return passes, total


def report_windows():
    import corpus as cp
    o = Oracle()
    for kind, seg in (("text", 0), ("lowent", 1)):
        d = cp.make_segment(kind, seg).numpy()[:65536].tobytes()
        els = parse(bytes(o.compress(d, 0)))
        for gw in (8, 16, 32):
            p, r = window_rounds(els, gw, 3584)
            print(f"{kind}: group width {gw}: {p} passes, {r} rounds per block ({r / p:.2f} per pass); "
                  f"warp-level rounds per block when {32 // gw} blocks share a warp: {r * gw / 32:.0f}")


if __name__ == "__main__":
    report_windows()
