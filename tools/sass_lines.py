#!/usr/bin/env python
"""Per CUDA source line: executed warp instructions split into fp64 / other, from an ncu report
(needs -lineinfo, --import-source on):  python tools/sass_lines.py report.ncu-rep [min_pct]"""
import collections, csv, subprocess, sys

def main():
    rep = sys.argv[1]; minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None; cur = None
    tot = collections.OrderedDict()
    for r in rows:
        if "# Samples" in r and "Source" in r:
            hdr = r; ci, ie, ws = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"); continue
        if hdr is None or len(r) <= ws: continue
        key = r[0]
        if not key.startswith("0x") and key.strip().isdigit():
            cur = (int(key), r[ci].strip()); tot.setdefault(cur, [0.0, 0.0, 0.0, collections.Counter()]); continue
        if cur is None or not key.startswith("0x"): continue
        toks = r[ci].split()
        if not toks: continue
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        try: a, b = float(r[ie] or 0), float(r[ws] or 0)
        except ValueError: continue
        t = tot[cur]
        if op in ("DFMA", "DMUL", "DADD", "DSETP"): t[0] += a
        else: t[1] += a; t[3][op] += a
        t[2] += b
    ti = sum(v[0] + v[1] for v in tot.values()); ts = sum(v[2] for v in tot.values())
    print(f"total {ti:.4e} instr")
    for (ln, src), (f, o, s, ops) in sorted(tot.items(), key=lambda kv: kv[0][0]):
        if (f + o) / ti * 100 < minp: continue
        top = ",".join(f"{k}:{v / ti * 100:.2f}" for k, v in ops.most_common(4))
        print(f"{ln:5d} f64 {f / ti * 100:5.2f}% other {o / ti * 100:5.2f}% smp {s / ts * 100:5.2f}% | {src[:70]} | {top}")

if __name__ == "__main__":
    main()
