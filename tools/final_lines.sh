#!/bin/bash
# Final bench lines of the round on one GPU -> gpurun_out/<tag>_*.json (copied to profiles/ afterwards)
tag=${1:-r2}
GTR='GTR{0.676278/2.012275/0.478487/0.753965/2.406436/1.0}+FU{0.245629/0.235012/0.253054/0.266305}+G4{1.078763}'
mkdir -p gpurun_out
python bench.py 2>gpurun_out/${tag}_final.err | tail -1 > gpurun_out/${tag}_bench_n1.json
python bench.py --impl reference 2>>gpurun_out/${tag}_final.err | tail -1 > gpurun_out/${tag}_bench_reference_arm.json
python bench.py --model "$GTR" --no-files --ref-queries 8000 2>>gpurun_out/${tag}_final.err | tail -1 > gpurun_out/${tag}_bench_n1_general_gtr.json
python bench.py --config cfg3 --ref-queries 8000 --steps 2 2>>gpurun_out/${tag}_final.err | tail -1 > gpurun_out/${tag}_bench_cfg3_n1.json
python bench.py --config cfg4 2>>gpurun_out/${tag}_final.err | tail -1 > gpurun_out/${tag}_bench_cfg4_n1.json
python bench.py --config cfg5 --steps 2 2>>gpurun_out/${tag}_final.err | tail -1 > gpurun_out/${tag}_bench_cfg5_n1_bounded.json
python - <<PY
import json
for f in ("bench_n1", "bench_reference_arm", "bench_n1_general_gtr", "bench_cfg3_n1", "bench_cfg4_n1", "bench_cfg5_n1_bounded"):
    try:
        d = json.load(open("gpurun_out/${tag}_%s.json" % f))
        print(f, round(d["value"]), round(d.get("e2e", {}).get("value", 0)), round(d.get("ms_per_step", 0), 1),
              {k: round(v.get("ms_per_step", v.get("ms", 0)), 2) for k, v in d.get("kernels", {}).items()},
              d.get("parity_vs_reference"), (d.get("cpu_baseline") or {}).get("value"), d.get("clocks"))
        if "e2e_files" in d:
            print("   files", {k: round(v["value"]) for k, v in d["e2e_files"].items() if isinstance(v, dict) and "value" in v})
        if "roofline" in d: print("   roofline", {k: v for k, v in d["roofline"].items() if k in ("frac", "fp64", "achieved")})
    except Exception as e:
        print(f, "FAILED", e)
PY
