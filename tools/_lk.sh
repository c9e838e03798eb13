ncu --set full --clock-control none --import-source on -k regex:lookup_build_site_kernel -c 1 -f -o /tmp/lk python bench.py --steps 1 --warmup 1 --no-cpu --queries 131072 > gpurun_out/lk_ncu.log 2>&1
python profiles/ncu_summary.py /tmp/lk.ncu-rep > gpurun_out/lk_ncu.txt 2>&1
python profiles/ncu_lines.py /tmp/lk.ncu-rep 25 >> gpurun_out/lk_ncu.txt 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu --queries 131072 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['kernels']['lookup_build'])"
