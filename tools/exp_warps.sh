for w in 4 6 8 10 12; do
  EPA_B200_SITE_WARPS=$w python bench.py --steps 2 --warmup 1 --no-cpu --queries 262144 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('warps $w', d['kernels']['thorough'], d['value'])"
done
EPA_B200_NO_TMEM=1 python bench.py --steps 2 --warmup 1 --no-cpu --queries 262144 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('no_tmem', d['kernels']['thorough'], d['value'])"
