#!/bin/bash
# Round-end capture on the GPU box: bench lines, ncu launch list, ncu --set full of the main kernels.
# The .ncu-rep files stay in /tmp; their summaries (profiles/ncu_summary.py + profiles/ncu_lines.py) travel back.
#   gpurun --timeout 1500 -- 'bash tools/capture_profiles.sh r2'
tag=${1:-r2}
GTR='GTR{0.676278/2.012275/0.478487/0.753965/2.406436/1.0}+FU{0.245629/0.235012/0.253054/0.266305}+G4{1.078763}'
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench.err
python bench.py --model "$GTR" --no-files --ref-queries 8000 > gpurun_out/${tag}_bench_n1_general_gtr.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-files > gpurun_out/${tag}_launch_run.log 2>&1
cap() {  # kernel regex, short name, extra bench args...
  k=$1; short=$2; shift; shift
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o /tmp/${tag}_${short} python bench.py --steps 1 --warmup 1 --no-cpu --no-files "$@" > gpurun_out/${tag}_ncu_${short}.log 2>&1
  python profiles/ncu_summary.py /tmp/${tag}_${short}.ncu-rep > gpurun_out/${tag}_ncu_${short}.txt 2>&1
  python profiles/ncu_lines.py /tmp/${tag}_${short}.ncu-rep 25 >> gpurun_out/${tag}_ncu_${short}.txt 2>&1
  python tools/sass_hist.py /tmp/${tag}_${short}.ncu-rep >> gpurun_out/${tag}_ncu_${short}.txt 2>&1
}
cap blo_site_kernel blo_site --queries 262144
cap blo_site_kernel blo_site_general_gtr --queries 262144 --model "$GTR"
cap lookup_build_site_kernel lookup_build_site --queries 262144
cap preplace_mma_kernel preplace_mma --queries 262144
cap select_count_kernel select_count --queries 262144
cap blo_generic_kernel blo_generic --config cfg4
cap lookup_build_kernel lookup_build_aa --config cfg4
cap "blo_site_kernel" blo_site_per_rate --config cfg5 --queries 512
du -sh gpurun_out
tail -c 300 gpurun_out/${tag}_bench_n1.json
