#!/bin/bash
# amino-acid thorough kernel, developer A/B over libepa_dev_w*.so variants
for w in 8 16; do
  EPA_B200_LIB=$PWD/epa-ng_b200/libepa_dev_w$w.so python bench.py --config cfg4 --steps 2 --warmup 2 --ref-queries 500 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('warps $w', round(d['value']), 'thorough', round(d['kernels']['thorough']['ms_per_step'],1), d['parity_vs_reference'])"
done
