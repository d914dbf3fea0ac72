#!/bin/bash
out=gpurun_out; mkdir -p $out; : > $out/r2k_ab.log
for mt in 1 2; do
  tag="c2mt$mt"
  EVC_C2_MTILES=$mt timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_parity_round2_gpu.py tests/test_properties_gpu.py -q -x -k "3xtf32" > $out/r2k_pytest_$tag.log 2>&1
  prc=$?
  EVC_C2_MTILES=$mt timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $out/r2k_bench_$tag.json 2> $out/r2k_bench_$tag.err
  brc=$?
  python - "$tag" "$prc" "$brc" <<'PY' | tee -a gpurun_out/r2k_ab.log
import json, sys
tag, prc, brc = sys.argv[1:4]
try:
    d = json.loads(open(f"gpurun_out/r2k_bench_{tag}.json").read().strip().splitlines()[-1]); r = d["roofline"]
    print(f"{tag:8s} pytest rc={prc} bench rc={brc} {d['value']:8.0f} frames/s {d['ms_per_step']:7.2f} ms  c1 {r['contraction1_us_per_launch']:6.1f} us  c2 {r['us_per_launch']:6.1f} us obj {d['objective']:.9f} sm {d['clocks']['sm_mhz']}")
except Exception as e:
    print(tag, "pytest rc=", prc, "bench rc=", brc, "no result:", e)
PY
  tail -n 3 $out/r2k_pytest_$tag.log | head -2
done
