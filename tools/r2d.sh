#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q -s > $out/r2d_pytest.log 2>&1; prc=$?; echo "pytest rc=$prc"
grep -E "griffin-lim|headline\[|config1-full\[|passed|failed|Error|error|assert" $out/r2d_pytest.log | head -40
timeout 300 python tests/manual/griffin_lim_timing.py 2>&1 | tail -2 | tee $out/r2d_griffin.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/r2d_bench.json 2> $out/r2d_bench.err; brc=$?; echo "bench rc=$brc"; tail -n 5 $out/r2d_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2d_bench.json").read().strip().splitlines()[-1]); r = d["roofline"]
    print("3xtf32", round(d["value"]), "frames/s e2e", round(d["e2e"]["value"]), "copy_ms", d["e2e"]["copy_ms"], "c1", round(r["contraction1_us_per_launch"], 1), "c2", round(r["us_per_launch"], 1), r["class_ms_launches"], "obj", d["objective"], d["clocks"])
    print("extra", json.dumps(d.get("extra")))
except Exception as e:
    print("no result", e)
PY
timeout 300 python bench.py --steps 3 --warmup 1 --workload reference_default > $out/r2d_bench_refdefault.json 2> $out/r2d_bench_refdefault.err; echo "refdefault rc=$?"; tail -c 1200 $out/r2d_bench_refdefault.json; tail -n 3 $out/r2d_bench_refdefault.err
