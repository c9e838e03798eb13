#!/bin/bash
tag=r2
cd /root/repo
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_models.py -x -q -m gpu 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:lookup_build_site_kernel -c 1 -f -o /tmp/${tag}_lk python bench.py --steps 1 --warmup 1 --no-cpu --no-files --queries 262144 > gpurun_out/${tag}_ncu_lookup_build_site.log 2>&1
python profiles/ncu_summary.py /tmp/${tag}_lk.ncu-rep > gpurun_out/${tag}_ncu_lookup_build_site.txt 2>&1
python profiles/ncu_lines.py /tmp/${tag}_lk.ncu-rep 25 >> gpurun_out/${tag}_ncu_lookup_build_site.txt 2>&1
python tools/sass_hist.py /tmp/${tag}_lk.ncu-rep >> gpurun_out/${tag}_ncu_lookup_build_site.txt 2>&1
python bench.py --no-files --ref-queries 4000 2>/dev/null | tail -1 > gpurun_out/bench_lk.json
python -c "
import json; d=json.load(open('gpurun_out/bench_lk.json')); print(d['value'], d['e2e']['value'], d['kernels']['lookup_build'], d['parity_vs_reference'])"
