#!/bin/bash
# tools/regions.sh <lib.so> <mangled-substring> <report.ncu-rep>: region breakdown of the lane = site kernel
lib=$1; sub=$2; rep=$3
d=$(mktemp -d); (cd $d && cuobjdump -xelf all $OLDPWD/$lib > /dev/null && nvdisasm -gi -c epa_b200.sm_100a.cubin > all.sass 2>/dev/null)
python - "$d/all.sass" "$sub" "$d/fn.sass" <<'PY'
import sys
src, sub, dst = sys.argv[1:4]
out = []; p = False
for ln in open(src):
    if ln.lstrip().startswith(".section"):
        if p: break
        if ".text." in ln and sub in ln: p = True
    if p: out.append(ln)
open(dst, "w").writelines(out)
PY
python tools/sass_regions.py $rep $d/fn.sass kernels_blo_site.cuh $(python - <<'PY'
import re
src = open("epa-ng_b200/csrc/kernels_blo_site.cuh").read().split("\n")
names = {"site_derivatives": "deriv", "site_newton": "newton", "site_pass_tip": "pass_tip", "site_pass_first": "pass_first",
         "site_pass_distal": "pass_distal", "site_pmatrix": "pmatrix", "site_tipvec": "tipvec", "site_next_item": "next_item",
         "site_store_row": "store_row", "blo_site_kernel": "main"}
starts = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:__device__ __forceinline__ \w[\w ]*|blo_site_kernel)\b.*?(\w+)\(", l)
    for k in names:
        if re.search(r"\b" + k + r"\(", l) and not l.startswith(" "): starts.append((i, names[k]))
starts.sort()
out = []
for j, (ln, nm) in enumerate(starts):
    end = starts[j + 1][0] - 1 if j + 1 < len(starts) else len(src)
    out.append(f"{nm}:{ln}-{end}")
print(" ".join(out))
PY
)
rm -rf $d
