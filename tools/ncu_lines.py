#!/usr/bin/env python3
"""Aggregate `ncu --page source --csv --print-source cuda,sass` output per CUDA source line.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:<k> > src.csv
    python tools/ncu_lines.py src.csv [top_n]
"""
import collections
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
cur_file, hdr = None, None
agg = collections.defaultdict(lambda: collections.Counter())
src_text = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
        continue
    key = (cur_file, int(r[0]))
    src_text[key] = r[1].strip()
    d = agg[key]
    for name, val in zip(hdr, r):
        if name == "# Samples" or name.startswith("stall_") and "Not Issued" not in name or name in (
                "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal"):
            try:
                d[name] += float(val)
            except ValueError:
                pass
tot = sum(d["# Samples"] for d in agg.values()) or 1
inst = sum(d["Instructions Executed"] for d in agg.values()) or 1
print(f"total samples {tot:.0f}, warp instructions {inst:.0f}")
for key, d in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    stalls = sorted(((v, k) for k, v in d.items() if k.startswith("stall_")), reverse=True)[:3]
    st = " ".join(f"{k[6:]}={100 * v / max(1, d['# Samples']):.0f}%" for v, k in stalls)
    conf = ""
    if d["L1 Wavefronts Shared"]:
        conf = f" smem wf {d['L1 Wavefronts Shared']:.0f}/{d['L1 Wavefronts Shared Ideal']:.0f}"
    print(f"{100 * d['# Samples'] / tot:5.1f}% inst {100 * d['Instructions Executed'] / inst:4.1f}%  {key[0]}:{key[1]:<4d} "
          f"{src_text[key][:70]:70s} | {st}{conf}")
