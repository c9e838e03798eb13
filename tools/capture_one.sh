#!/bin/bash
# One ncu --set full capture of one kernel of a short bench run; the report travels back in gpurun_out/.
#   gpurun --timeout 600 -- 'bash tools/capture_one.sh blo_site_kernel r2a'
k=$1; tag=${2:-r2}; shift; shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/${tag}_${k} python bench.py --steps 1 --warmup 1 --no-cpu --queries 262144 "$@" > gpurun_out/${tag}_ncu_${k}.log 2>&1
python profiles/ncu_summary.py gpurun_out/${tag}_${k}.ncu-rep > gpurun_out/${tag}_ncu_${k}.txt 2>&1
python profiles/ncu_lines.py gpurun_out/${tag}_${k}.ncu-rep 40 >> gpurun_out/${tag}_ncu_${k}.txt 2>&1
du -sh gpurun_out/*
