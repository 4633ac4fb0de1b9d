"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line.
usage: ncu_lines.py report.ncu-rep kernel_regex [min_pct]"""
import csv, subprocess, sys, collections, io
rep, kre = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
agg = collections.OrderedDict(); fname = "?"; hdr = None; seen_kernel = 0
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name":
        seen_kernel += 1
        if seen_kernel > 1: break   # first matching launch only
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
    d = dict(zip(hdr[2:], r[2:]))  # skip the two leading Line/Source columns (duplicate 'Source' key ok)
    key = (fname, int(r[0]), r[1].strip()[:100])
    a = agg.setdefault(key, collections.Counter())
    def num(v):
        try: return int(v)
        except ValueError: return 0
    a["inst"] += num(d["Instructions Executed"])
    a["samp"] += num(d["Warp Stall Sampling (All Samples)"])
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            a[k] += num(v)
ti = sum(a["inst"] for a in agg.values()) or 1; ts = sum(a["samp"] for a in agg.values()) or 1
print(f"total warp-instructions {ti}, stall samples {ts}")
for (f, ln, src), a in agg.items():
    pi, ps = 100 * a["inst"] / ti, 100 * a["samp"] / ts
    if pi < minpct and ps < minpct: continue
    top = ", ".join(f"{k[6:]}={100*v/ts:.1f}" for k, v in a.most_common(6) if k.startswith("stall_") and 100*v/ts >= 0.3)
    print(f"{pi:5.1f}%i {ps:5.1f}%s {f}:{ln:<4d} {src[:70]:70s} | {top}")
