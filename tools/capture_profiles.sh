#!/bin/bash
# Round-end capture on the GPU box: bench lines, ncu launch list, ncu --set full of the four main
# kernels. The .ncu-rep files stay in /tmp (gpurun_out/ is limited to 64 MiB); their summaries
# (profiles/ncu_summary.py + profiles/ncu_lines.py) are what travels back.
#   gpurun --timeout 900 -- 'bash tools/capture_profiles.sh r1e'
tag=${1:-r1e}
skip_bench=$2
mkdir -p gpurun_out
if [ -z "$skip_bench" ]; then
  python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench.err
  python bench.py --impl reference > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench.err
fi
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${tag}_launch_run.log 2>&1
for k in blo_site_kernel lookup_build_site_kernel preplace_mma_kernel select_count_kernel; do
  short=${k%_kernel}
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o /tmp/${tag}_${short} python bench.py --steps 1 --warmup 1 --no-cpu --queries 262144 > gpurun_out/${tag}_ncu_${short}.log 2>&1
  python profiles/ncu_summary.py /tmp/${tag}_${short}.ncu-rep > gpurun_out/${tag}_ncu_${short}.txt 2>&1
  python profiles/ncu_lines.py /tmp/${tag}_${short}.ncu-rep 25 >> gpurun_out/${tag}_ncu_${short}.txt 2>&1
done
du -sh gpurun_out
tail -c 300 gpurun_out/${tag}_bench_n1.json
