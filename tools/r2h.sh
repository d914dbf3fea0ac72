#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/r2h_pytest.log 2>&1; prc=$?; echo "pytest rc=$prc"; tail -n 4 $out/r2h_pytest.log; grep -E "^FAILED" $out/r2h_pytest.log | head
for m in 3xtf32 tf32 bf16; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --mode $m > $out/r2h_bench_$m.json 2> $out/r2h_bench_$m.err; echo "bench $m rc=$?"
  python - $out/r2h_bench_$m.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d["roofline"]; c = r["class_ms_launches"]
    print(d["config"]["mode"], round(d["value"]), "frames/s c1", round(r["contraction1_us_per_launch"], 1), "us c2", round(r["us_per_launch"], 1), "us reduce", round(c["reduce_ratio"][0] / c["reduce_ratio"][1] * 1e3, 1), "us obj", d["objective"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("no result", e)
PY
done
timeout 300 python bench.py --steps 3 --warmup 1 --workload reference_default --no-cpu-baseline > $out/r2h_bench_refdefault.json 2> $out/r2h_bench_refdefault.err; echo "refdefault rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2h_bench_refdefault.json").read().strip().splitlines()[-1])
    print("reference_default e2e", round(d["e2e"]["value"]), "frames/s", d["e2e"]["ms_each_step"])
except Exception as e:
    print("no result", e)
PY
