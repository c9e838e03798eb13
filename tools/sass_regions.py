#!/usr/bin/env python
"""Joins the per-instruction counters of an ncu report with nvdisasm -gi line/inline info and sums
executed warp instructions per code region (line ranges of one source file).
    python tools/sass_regions.py report.ncu-rep function.sass file.cuh name:lo-hi [name:lo-hi ...]
function.sass = the function's section of `nvdisasm -gi -c <cubin>`."""
import collections, csv, re, subprocess, sys

F64 = ("DFMA", "DMUL", "DADD", "DSETP")

def main():
    rep, sass, fname = sys.argv[1:4]
    regions = []
    for a in sys.argv[4:]:
        n, r = a.split(":"); lo, hi = r.split("-"); regions.append((n, int(lo), int(hi)))
    # nvdisasm: annotation lines precede each instruction
    loc = {}; chain = []
    pat = re.compile(r'//## File "([^"]+)", line (\d+)')
    ins = re.compile(r'^\s*/\*([0-9a-f]+)\*/\s+(.*?);')
    for ln in open(sass):
        m = pat.search(ln)
        if m:
            chain.append((m.group(1), int(m.group(2)))); continue
        m = ins.match(ln)
        if m:
            if chain: cur = chain
            loc[int(m.group(1), 16)] = cur
            chain = []
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if "# Samples" in r][0]
    hdr = rows[hi]; ci, ie, ws = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = None
    tot = collections.defaultdict(lambda: [0.0, 0.0, 0.0, collections.Counter()])
    lines = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= ws or not r[0].startswith("0x"): continue
        addr = int(r[0], 16)
        if base is None: base = addr
        toks = r[ci].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        a, b = float(r[ie] or 0), float(r[ws] or 0)
        ch = loc.get(addr - base, [("?", 0)])
        own = [l for f, l in ch if f.endswith(fname)]
        line = own[0] if own else 0
        reg = "other"
        for n, lo, hi2 in regions:
            if any(lo <= l <= hi2 for l in own): reg = n; break
        t = tot[reg]
        if op in F64: t[0] += a
        else: t[1] += a; t[3][op] += a
        t[2] += b
        lines[(reg, line)][0 if op in F64 else 1] += a; lines[(reg, line)][2] += b
    ti = sum(v[0] + v[1] for v in tot.values()); ts = sum(v[2] for v in tot.values())
    print(f"total {ti:.4e} warp instructions")
    for reg, (f, o, s, ops) in sorted(tot.items(), key=lambda kv: -(kv[1][0] + kv[1][1])):
        top = " ".join(f"{k}:{v / ti * 100:.2f}" for k, v in ops.most_common(8))
        print(f"{reg:12s} f64 {f / ti * 100:5.2f}% other {o / ti * 100:5.2f}% samples {s / ts * 100:5.2f}% | {top}")
    print("-- lines >= 0.4% --")
    for (reg, line), (f, o, s) in sorted(lines.items(), key=lambda kv: -(kv[1][0] + kv[1][1])):
        if (f + o) / ti * 100 >= 0.4:
            print(f"{reg:12s} line {line:5d} f64 {f / ti * 100:5.2f}% other {o / ti * 100:5.2f}% samples {s / ts * 100:5.2f}%")

if __name__ == "__main__":
    main()
