#!/bin/bash
# ncu --set full of one kernel of a bench.py run with the full library: tools/capture_cfg.sh <kernel regex> <tag> <bench args...>
k=$1; tag=$2; shift; shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/${tag} python bench.py --steps 1 --warmup 1 --no-cpu --no-files "$@" > gpurun_out/${tag}.log 2>&1
python profiles/ncu_summary.py gpurun_out/${tag}.ncu-rep > gpurun_out/${tag}.txt 2>&1
python profiles/ncu_lines.py gpurun_out/${tag}.ncu-rep 40 >> gpurun_out/${tag}.txt 2>&1
python tools/sass_hist.py gpurun_out/${tag}.ncu-rep >> gpurun_out/${tag}.txt 2>&1
