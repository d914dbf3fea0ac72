#!/bin/bash
# every BASELINE.json workload through bench.py on one GPU (one JSON line each)
out=gpurun_out; mkdir -p $out
run() { tag=$1; shift; timeout 900 python bench.py "$@" > $out/r2i_$tag.json 2> $out/r2i_$tag.err; echo "$tag rc=$?"
  python - $out/r2i_$tag.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
    print("  ", d["config"]["workload"], d["config"]["mode"], "T", d["config"]["T"], round(d["value"], 1), "frames/s", round(d["ms_per_step"], 1), "ms/step",
          round(d["tflops_algorithmic"], 1), "TFLOP/s alg | e2e", round(d["e2e"]["value"], 1), "| c1", round(r.get("contraction1_us_per_launch", 0), 1), "c2", round(r.get("us_per_launch", 0), 1), "us | obj", d["objective"], "|", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("   no result", e)
PY
}
run batch256 --workload batch_256utt_20k --steps 1 --warmup 1 --no-cpu-baseline
run stacked_3xtf32 --workload context_stacked_50k --steps 2 --warmup 1 --no-cpu-baseline
run stacked_tf32 --workload context_stacked_50k --steps 2 --warmup 1 --no-cpu-baseline --mode tf32
run stacked_bf16 --workload context_stacked_50k --steps 2 --warmup 1 --no-cpu-baseline --mode bf16
run large200k --workload large_dictionary_200k --steps 1 --warmup 1 --no-cpu-baseline
run refdefault --workload reference_default --steps 3 --warmup 1
run tf32 --mode tf32 --steps 5 --warmup 3 --no-cpu-baseline --no-extras
