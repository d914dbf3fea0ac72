#!/bin/bash
# One GPU-box visit: diagnostics, parity tests, bench in every mode and (only when all of that exited 0) the ncu
# launch list + one full capture of the tensor-core kernels.  usage: tools/gpu_round.sh <tag> [noncu]
tag="${1:-run}"; out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.log 2>&1
timeout 900 python tests/manual/gpu_diag.py > $out/${tag}_diag.log 2>&1; echo "diag rc=$?"
tail -n 25 $out/${tag}_diag.log
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; prc=$?; echo "pytest rc=$prc"
tail -n 15 $out/${tag}_pytest.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err; brc=$?; echo "bench rc=$brc"
tail -c 1500 $out/${tag}_bench.json; tail -n 5 $out/${tag}_bench.err
for m in tf32 bf16; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mode $m > $out/${tag}_bench_$m.json 2> $out/${tag}_bench_$m.err; echo "bench $m rc=$?"
  python - $out/${tag}_bench_$m.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d["roofline"]
    print(d["config"]["mode"], round(d["value"]), "frames/s c1", round(r["contraction1_us_per_launch"], 1), "us c2", round(r["us_per_launch"], 1), "us obj", d["objective"])
except Exception as e:
    print("no result", e)
PY
done
if [ "$2" != "noncu" ] && [ $prc -eq 0 ] && [ $brc -eq 0 ]; then
  cmd="python bench.py --steps 1 --warmup 1 --iterations 20 --no-cpu-baseline"
  $cmd > $out/${tag}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches.csv $cmd > $out/${tag}_ncu1.log 2>&1
  echo "ncu launches rc=$?"
  $cmd > $out/${tag}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 20 -c 4 -o $out/${tag}_prof $cmd > $out/${tag}_ncu2.log 2>&1
  echo "ncu full rc=$?"
fi
