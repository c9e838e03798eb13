#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show which hardware paths the built library uses:
tcgen05 MMA (UTCIMMA), tcgen05 commit/barrier (UTCBAR), tensor-memory loads/stores (LDTM/STTM), TMA bulk copies
(UBLKCP), tensor-map TMA (UTMALDG: none - every staged tile is contiguous, bulk copies suffice), mbarrier
(SYNCS), fp64 tensor-core MMA (DMMA), fp64 arithmetic (DFMA/DMUL/DADD), local-memory spills (LDL/STL).
    python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "epa-ng_b200", "libepa_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ins = re.compile(r"^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)")
kern = collections.OrderedDict()
cur = None
for ln in out.splitlines():
    if "Function :" in ln:
        cur = ln.split("Function :")[1].strip(); kern[cur] = collections.Counter(); continue
    m = ins.match(ln)
    if m and cur: kern[cur][m.group(1)] += 1; kern[cur]["_n"] += 1
names = subprocess.run(["c++filt"], input="\n".join(kern), capture_output=True, text=True).stdout.splitlines()
cols = ["UTCIMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "DMMA", "DFMA", "DMUL", "DADD", "MUFU", "LDL", "STL"]
print("# cuobjdump -sass epa-ng_b200/libepa_b200.so (sm_100a): static instruction counts per kernel")
print("# kernel | instructions | " + " | ".join(cols))
rows = []
for (k, c), nm in zip(kern.items(), names):
    nm = re.sub(r"^void ", "", nm); nm = re.sub(r"\(.*$", "", nm); nm = nm.replace("epa::", "").replace("(anonymous namespace)::", "")
    nm = nm.replace("(int)", "").replace("(bool)", "")
    rows.append(f"{nm} | {c['_n']} | " + " | ".join(str(c[x]) for x in cols))
print("\n".join(sorted(rows)))
tot = collections.Counter()
for c in kern.values(): tot.update(c)
print("# total | %d | " % tot["_n"] + " | ".join(str(tot[x]) for x in cols))
