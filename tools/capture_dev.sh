#!/bin/bash
# ncu --set full of one kernel with the developer library; report comes back in gpurun_out/
k=$1; tag=$2; shift; shift
export EPA_B200_LIB=$PWD/epa-ng_b200/libepa_dev.so
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/${tag}_${k} python bench.py --steps 1 --warmup 1 --no-cpu --queries 262144 "$@" > gpurun_out/${tag}_ncu_${k}.log 2>&1
python profiles/ncu_summary.py gpurun_out/${tag}_${k}.ncu-rep > gpurun_out/${tag}_ncu_${k}.txt 2>&1
python profiles/ncu_lines.py gpurun_out/${tag}_${k}.ncu-rep 40 >> gpurun_out/${tag}_ncu_${k}.txt 2>&1
