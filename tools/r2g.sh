#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/r2g_pytest.log 2>&1; prc=$?; echo "pytest rc=$prc"; tail -n 4 $out/r2g_pytest.log
tools/ab_bench.sh "keepl2:EVC_X=0" "nokeep:EVC_NO_KEEP_L2=1" 2>&1 | tee $out/r2g_ab.log
timeout 900 python bench.py --steps 5 --warmup 3 > $out/r2g_bench.json 2> $out/r2g_bench.err; brc=$?; echo "bench rc=$brc"; tail -n 3 $out/r2g_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/r2g_bench_reference.json 2> $out/r2g_bench_reference.err; echo "reference arm rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2g_bench.json", "gpurun_out/r2g_bench_reference.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, round(d["value"], 1), "frames/s e2e", round(d["e2e"]["value"], 1), d["e2e"].get("ms_each_step"), r.get("class_ms_launches"), d.get("clocks"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "no result", e)
PY
if [ $prc -eq 0 ] && [ $brc -eq 0 ]; then
  cmd="python bench.py --steps 1 --warmup 1 --iterations 20 --no-cpu-baseline --no-extras"
  $cmd > $out/r2g_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/r2g_launches.csv $cmd > $out/r2g_ncu1.log 2>&1
  echo "ncu launches rc=$?"
  $cmd > $out/r2g_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel|reduce_partials" -s 30 -c 6 -o $out/r2g_prof $cmd > $out/r2g_ncu2.log 2>&1
  echo "ncu full rc=$?"
fi
