#!/usr/bin/env python3
"""Source-line hot spots of one kernel from an .ncu-rep captured with --import-source on (and -lineinfo):

    python tools/ncu_hotspots.py <report.ncu-rep> <kernel-name regex> [top_n] > profiles/<name>_source_hotspots.txt

Per CUDA source line: share of executed warp instructions, share of stall samples, shared-memory wavefronts (actual / ideal)."""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    data, fname, hdr = [], "", None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[2] != "-":          # rows with an address are the SASS lines under a source line
            continue
        col = {h: i for i, h in enumerate(hdr)}
        try:
            data.append((fname, r[0], r[1].strip(), int(r[col["# Samples"]]), int(r[col["Instructions Executed"]]),
                         int(r[col["L1 Wavefronts Shared"]] or 0), int(r[col["L1 Wavefronts Shared Ideal"]] or 0)))
        except (ValueError, KeyError):
            pass
    ts, te = sum(d[3] for d in data) or 1, sum(d[4] for d in data) or 1
    print(f"# {rep}: {kern}; {te} warp instructions, {ts} stall samples")
    for d in sorted(data, key=lambda x: -x[4])[:top]:
        wf = f"  smem wf {d[5]}/{d[6]}" if d[5] else ""
        print(f"{100 * d[4] / te:5.1f}% inst {100 * d[3] / ts:5.1f}% samples  {d[0]}:{d[1]:<4s} {d[2][:100]}{wf}")


if __name__ == "__main__":
    main()
