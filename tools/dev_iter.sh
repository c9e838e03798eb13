#!/bin/bash
# Developer iteration on the GPU box with the EPA_DEV_MIN library: a few thorough-path parity tests,
# then a short bench (thorough ms, parity against the reference binary on 3000 queries).
export EPA_B200_LIB=$PWD/epa-ng_b200/libepa_dev.so
python -m pytest tests/test_gpu_parity.py -x -q -k "thorough_all_pairs or thorough_candidates or synth64_placements or dense_chunk_placements or mid_length or edge_cases" 2>&1 | tail -5
python bench.py --steps 2 --warmup 1 --queries 262144 --ref-queries 3000 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); k=d['kernels']
print('thorough ms', round(k['thorough']['ms_per_step'],3), 'pre', round(k['preplace']['ms_per_step'],3), 'sel', round(k['select']['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'same', d['resident_equals_e2e'])
print('parity', d.get('parity_vs_reference'))"
