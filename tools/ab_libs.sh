#!/bin/bash
# A/B of developer builds on the GPU box: tools/ab_libs.sh libA.so libB.so ... (paths relative to epa-ng_b200/)
# prints the thorough time of a 262144-query cfg2 run, twice per library, alternating
GTR='GTR{0.676278/2.012275/0.478487/0.753965/2.406436/1.0}+FU{0.245629/0.235012/0.253054/0.266305}+G4{1.078763}'
for rep in 1 2; do
  for lib in "$@"; do
    for model in default general; do
      extra=""; [ $model = general ] && extra="--model $GTR"
      EPA_B200_LIB=$PWD/epa-ng_b200/$lib python bench.py --steps 3 --warmup 2 --queries 262144 --no-files --no-cpu $extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); k=d['kernels']
print('$lib $model thorough ms', round(k['thorough']['ms_per_step'],3), 'value', round(d['value']))"
    done
  done
done
