#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q -s > $out/r2c_pytest.log 2>&1; prc=$?; echo "pytest rc=$prc"
grep -E "headline\[|config1-full\[|passed|failed|Error|error" $out/r2c_pytest.log | head -40
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/r2c_bench.json 2> $out/r2c_bench.err; brc=$?; echo "bench rc=$brc"; tail -n 5 $out/r2c_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c_bench.json").read().strip().splitlines()[-1]); r = d["roofline"]
    print("3xtf32", round(d["value"]), "frames/s e2e", round(d["e2e"]["value"]), "copy_ms", d["e2e"]["copy_ms"], "c1", round(r["contraction1_us_per_launch"], 1), "c2", round(r["us_per_launch"], 1), r["class_ms_launches"], "obj", d["objective"], d["clocks"])
    print("extra", json.dumps(d.get("extra")))
except Exception as e:
    print("no result", e)
PY
for m in tf32 bf16; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --mode $m > $out/r2c_bench_$m.json 2> $out/r2c_bench_$m.err; echo "bench $m rc=$?"
  python - $out/r2c_bench_$m.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d["roofline"]
    print(d["config"]["mode"], round(d["value"]), "frames/s c1", round(r["contraction1_us_per_launch"], 1), "us c2", round(r["us_per_launch"], 1), "us obj", d["objective"])
except Exception as e:
    print("no result", e)
PY
done
timeout 300 python tests/manual/accuracy_modes.py > $out/r2c_accuracy.log 2>&1; tail -n 20 $out/r2c_accuracy.log
if [ $prc -eq 0 ] && [ $brc -eq 0 ]; then
  cmd="python bench.py --steps 1 --warmup 1 --iterations 20 --no-cpu-baseline --no-extras"
  $cmd > $out/r2c_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/r2c_launches.csv $cmd > $out/r2c_ncu1.log 2>&1
  echo "ncu launches rc=$?"
  $cmd > $out/r2c_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel|reduce_partials" -s 30 -c 6 -o $out/r2c_prof $cmd > $out/r2c_ncu2.log 2>&1
  echo "ncu full rc=$?"
fi
