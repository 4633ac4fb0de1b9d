"""Python model of the K0 relaxation (csrc/index.cu), for debugging convergence offline."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import numpy as np
import datasets
from oracle_lib import Oracle

NO_DEAD_MARKS = int(os.environ.get('NDM', '1'))
SEG = 128; DEAD = 0x80; MARK = 0xff; NONE = (1 << 64) - 1; LOW = 1 << 63
RULE = int(os.environ.get('RULE', '1'))
C_MARK, C_PROP, C_DEAD = (0, 1 << 61, 2 << 61) if RULE else (0, 0, 1 << 63)

def decode_at(body, e):
    n = len(body); tag = body[e]; typ = tag & 3
    if typ == 0:
        m = tag >> 2; extra = m - 59 if m >= 60 else 0
    else:
        extra = 1 if typ == 1 else (2 if typ == 2 else 4)
    ok = e + 1 + extra <= n
    raw = 0
    if typ == 0 and extra and ok:
        for k in range(extra): raw |= body[e + 1 + k] << (8 * k)
    if typ == 0:
        ln = (raw if (tag >> 2) >= 60 else (tag >> 2)) + 1
        return 1 + extra + ln, ln, ok
    return 1 + extra, ((tag >> 2) & 7) + 4 if typ == 1 else (tag >> 2) + 1, ok

def walk(body, lo, hi, e, old, fresh):
    while e < hi:
        rel = e - lo
        if old is not None and rel in old: return e, True
        fresh.add(rel)
        size, _, _ = decode_at(body, e)
        e += size
    return e, False

def run(body, verbose=False):
    n = len(body); nseg = (n + SEG - 1) // SEG
    paths = []; exits = []; entry = [0] * nseg
    for t in range(nseg):
        lo = t * SEG; hi = min(lo + SEG, n); p = set()
        x, _ = walk(body, lo, hi, lo, None, p); paths.append(p); exits.append(x)
    rounds = 0
    while True:
        rounds += 1
        claim = [NONE] * nseg
        for t in range(nseg):
            dead = entry[t] & DEAD
            x = exits[t]; u = x // SEG
            if not dead:
                vmax = min(min(u, nseg), t + 1 + (65536 + 1024) // SEG)
                for v in range(t + 1, vmax): claim[v] = min(claim[v], C_MARK | (t << 8) | MARK)
                if u < nseg and x >= n: claim[u] = min(claim[u], C_MARK | (t << 8) | MARK)
            if u < nseg and x < n: claim[u] = min(claim[u], (C_DEAD if dead else C_PROP) | (t << 8) | (x - u * SEG))
        changed = 0; nch = 0
        for t in range(nseg):
            c = claim[t]; old = entry[t]; payload = c & 0xff
            if t == 0: ne = 0
            elif c == NONE or payload == MARK: ne = (old & 0x7f) | DEAD
            else: ne = payload
            if ne == old: continue
            entry[t] = ne; changed = 1; nch += 1
            if (ne & DEAD) or (ne & 0x7f) == (old & 0x7f): continue
            lo = t * SEG; hi = min(lo + SEG, n); fresh = set()
            x, merged = walk(body, lo, hi, lo + ne, paths[t], fresh)
            if merged: fresh |= paths[t]
            else: exits[t] = x
            paths[t] = fresh
        if verbose: print('round', rounds, 'changed segs', nch)
        if not changed or rounds > nseg + 2: break
    # truth
    true_entry = [DEAD] * nseg; e = 0
    while e < n:
        t = e // SEG
        if true_entry[t] == DEAD: true_entry[t] = e - t * SEG
        size, _, ok = decode_at(body, e); e += size
    bad = [t for t in range(nseg) if (entry[t] & DEAD) != (true_entry[t] & DEAD) or (not (entry[t] & DEAD) and entry[t] != true_entry[t])]
    return rounds, bad, nseg

if __name__ == '__main__':
    o = Oracle()
    specs = sys.argv[1:] or ['corpus:text:2:1:1', 'corpus:text:2:100:5000', 'corpus:mixed:0:900000:300000', 'corpus:random:7:0:140000', 'sym:5:70000:70000', 'lz:65536:65536']
    for spec in specs:
        data = datasets.gen(spec); s = o.compress(data, 0)
        hdr = len(o.varint_encode(data.size)); body = bytes(s[hdr:])
        r, bad, nseg = run(body, verbose='-v' in os.environ.get('K0V', ''))
        print(spec, 'nseg', nseg, 'rounds', r, 'bad', bad[:10])
