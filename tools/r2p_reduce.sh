#!/bin/bash
# reduction kernel with one box load per block; ratio straight from TMEM on the batch workload (A/B)
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $out/r2p_pytest.log
tools/ab_bench.sh "default:EVC_X=1" "default2:EVC_X=1"
for v in direct separate; do
  extra="EVC_X=1"; [ $v = separate ] && extra="EVC_NO_FUSED_REDUCE=1"
  env $extra timeout 600 python bench.py --workload batch_256utt_20k --steps 1 --warmup 1 --iterations 100 --no-cpu-baseline --no-extras > $out/r2p_batch_$v.json 2> $out/r2p_batch_$v.err; echo "batch $v rc=$?"
  python - $out/r2p_batch_$v.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d["roofline"]
    print(round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms", r["class_ms_launches"], "obj", d["objective"])
except Exception as e:
    print("no result", e)
PY
done
