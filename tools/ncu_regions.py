"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by file:line ranges and by
opcode. usage: ncu_regions.py dump.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; fpath = None
per_line = collections.Counter(); per_line_s = collections.Counter(); src = {}
ops = collections.Counter(); ops_by_line = collections.defaultdict(collections.Counter)
cur = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if "# Samples" in r: hdr = r; ie = hdr.index("Instructions Executed"); ws = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= ie: continue
    k = r[0].strip()
    if k.isdigit():
        cur = (fpath, int(k)); src[cur] = r[1].strip(); continue
    if k == "" and r[2].startswith("0x") or k.startswith("0x"):
        pass
    # sass row: Address in col 2, Source(sass) col 3
    try: ins = float(r[ie] or 0); smp = float(r[ws] or 0)
    except ValueError: continue
    sass = r[3].strip() if len(r) > 3 else ""
    if not sass: continue
    toks = sass.split()
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0]
    per_line[cur] += ins; per_line_s[cur] += smp
    ops[op] += ins; ops_by_line[cur][op] += ins
tot = sum(per_line.values()); ts = sum(per_line_s.values())
print("total warp instr %.4e samples %d" % (tot, ts))
print("opcodes:", ", ".join("%s %.1f%%" % (k, v / tot * 100) for k, v in ops.most_common(25)))
for (f, l), v in sorted(per_line.items(), key=lambda kv: -kv[1])[:70]:
    top = ",".join("%s:%.0f%%" % (k, x / v * 100) for k, x in ops_by_line[(f, l)].most_common(4))
    print("%-22s %4d inst %5.2f%% smp %5.2f%%  [%s] %s" % (f, l, v / tot * 100, per_line_s[(f, l)] / ts * 100, top, src.get((f, l), "")[:70]))
if len(sys.argv) > 2:
    regs = [tuple(map(int, a.split("-"))) for a in sys.argv[2:]]
    for lo, hi in regs:
        v = sum(x for (f, l), x in per_line.items() if f == "kernels_blo_site.cuh" and lo <= l <= hi)
        vs = sum(x for (f, l), x in per_line_s.items() if f == "kernels_blo_site.cuh" and lo <= l <= hi)
        print("lines %d-%d: inst %.2f%% samples %.2f%%" % (lo, hi, v / tot * 100, vs / ts * 100))
    v = sum(x for (f, l), x in per_line.items() if f != "kernels_blo_site.cuh")
    print("other files: %.2f%%" % (v / tot * 100))
    by = collections.Counter()
    for (f, l), x in per_line.items():
        if f != "kernels_blo_site.cuh": by[f] += x
    print({k: "%.2f%%" % (v / tot * 100) for k, v in by.items()})
