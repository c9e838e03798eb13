#!/usr/bin/env python
"""Per-CUDA-line summary of an ncu report (needs -lineinfo and --import-source on):
    python profiles/ncu_lines.py gpurun_out/prof.ncu-rep [top]
Prints stall-reason totals, and the hottest source lines by warp-stall samples with their share of
executed instructions. Reads `ncu --page source --print-source cuda,sass --csv`."""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    lines = collections.OrderedDict()
    stalls = collections.Counter()
    cur = None
    for r in rows:
        if "# Samples" in r and "Source" in r:
            hdr = r
            ci, ie, ws = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
            sc = [(i, x) for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
            continue
        if hdr is None or len(r) <= ws:
            continue
        # cuda,sass view: a CUDA line row (first column = line number) is followed by its SASS rows (address)
        key = r[0]
        try:
            ins = float(r[ie] or 0)
            smp = float(r[ws] or 0)
        except ValueError:
            continue
        if not key.startswith("0x") and key.strip().isdigit():
            cur = (int(key), r[ci].strip())
            lines.setdefault(cur, [0.0, 0.0])
            lines[cur][0] += ins
            lines[cur][1] += smp
            for i, x in sc:
                try:
                    stalls[x] += float(r[i] or 0)
                except ValueError:
                    pass
    ti = sum(v[0] for v in lines.values()) or 1
    ts = sum(v[1] for v in lines.values()) or 1
    print(f"instructions {ti:.3e}  samples {ts:.0f}")
    tot = sum(stalls.values()) or 1
    print("stalls: " + ", ".join(f"{k[6:]} {v / tot * 100:.1f}%" for k, v in stalls.most_common(8)))
    for (ln, src), (ins, smp) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{ln:5d} inst {ins / ti * 100:5.2f}%  samples {smp / ts * 100:5.2f}%  {src[:110]}")


if __name__ == "__main__":
    main()
