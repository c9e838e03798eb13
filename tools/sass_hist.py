#!/usr/bin/env python
"""Opcode histogram of an ncu report's SASS page, weighted by executed warp instructions:
    python tools/sass_hist.py report.ncu-rep [--lines]
Prints the share of each opcode class in executed instructions and in stall samples, and (with --lines)
per CUDA source line: executed instructions split into fp64 / other."""
import collections, csv, subprocess, sys

def load(rep, what):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", what, "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    for i, r in enumerate(rows):
        if "# Samples" in r and "Source" in r:
            return r, rows[i + 1:]
    raise SystemExit("no source page")

def opclass(op):
    op = op.split(".")[0]
    if op in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"): return "fp64:" + op
    return op

def main():
    rep = sys.argv[1]
    hdr, rows = load(rep, "sass")
    ci, ie, ws = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ins = collections.Counter(); smp = collections.Counter()
    for r in rows:
        if len(r) <= ws: continue
        toks = r[ci].split()
        if not toks: continue
        op = toks[1] if toks[0].startswith("@") else toks[0]
        try: a, b = float(r[ie] or 0), float(r[ws] or 0)
        except ValueError: continue
        ins[opclass(op)] += a; smp[opclass(op)] += b
    ti, ts = sum(ins.values()), sum(smp.values())
    f64 = sum(v for k, v in ins.items() if k.startswith("fp64"))
    print(f"executed warp instructions {ti:.4e}, fp64 share {f64 / ti * 100:.1f}%")
    for k, v in ins.most_common(40):
        print(f"  {k:14s} inst {v / ti * 100:5.2f}%  samples {smp[k] / ts * 100:5.2f}%")

if __name__ == "__main__":
    main()
