#!/bin/bash
# bench.py on the other BASELINE.json configurations (one GPU), lines into gpurun_out/<tag>_<cfg>.json
tag=${1:-r2}
mkdir -p gpurun_out
for c in cfg4 cfg5; do
  python bench.py --config $c --steps 2 --warmup 3 > gpurun_out/${tag}_bench_${c}.json 2> gpurun_out/${tag}_bench_${c}.err || tail -5 gpurun_out/${tag}_bench_${c}.err
done
python bench.py --config cfg3 --steps 2 --warmup 3 --ref-queries 8000 > gpurun_out/${tag}_bench_cfg3.json 2> gpurun_out/${tag}_bench_cfg3.err || tail -5 gpurun_out/${tag}_bench_cfg3.err
python - <<PY
import json
for c in ("cfg4", "cfg5", "cfg3"):
    try:
        d = json.load(open("gpurun_out/${tag}_bench_%s.json" % c))
        print(c, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 1), d["config"].get("bounded", ""),
              "cpu", d.get("cpu_baseline", {}).get("value"), "parity", d.get("parity_vs_reference"))
        print("   kernels", {k: round(v.get("ms_per_step", v.get("ms", 0)), 2) for k, v in d["kernels"].items()})
    except Exception as e:
        print(c, "failed", e)
PY
