#!/bin/bash
# A/B of the cluster-multicast variants: EVC_C1_PAIRS x EVC_C2_PAIRS (CTA pairs per cluster of each contraction)
out=gpurun_out; mkdir -p $out; : > $out/r2f_ab.log
for cfg in "1 1" "1 2" "1 4" "2 1" "4 1" "2 2" "2 4" "4 4"; do
  set -- $cfg; c1=$1; c2=$2; tag="c1p${c1}_c2p${c2}"
  EVC_C1_PAIRS=$c1 EVC_C2_PAIRS=$c2 timeout 240 python -m pytest tests/test_parity_gpu.py tests/test_parity_round2_gpu.py -q -x -k "3xtf32" > $out/r2f_pytest_$tag.log 2>&1
  prc=$?
  EVC_C1_PAIRS=$c1 EVC_C2_PAIRS=$c2 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $out/r2f_bench_$tag.json 2> $out/r2f_bench_$tag.err
  brc=$?
  python - "$tag" "$prc" "$brc" <<'PY' | tee -a gpurun_out/r2f_ab.log
import json, sys
tag, prc, brc = sys.argv[1:4]
try:
    d = json.loads(open(f"gpurun_out/r2f_bench_{tag}.json").read().strip().splitlines()[-1]); r = d["roofline"]
    print(f"{tag:12s} pytest rc={prc} bench rc={brc} {d['value']:8.0f} frames/s {d['ms_per_step']:7.2f} ms  c1 {r['contraction1_us_per_launch']:6.1f} us  c2 {r['us_per_launch']:6.1f} us  reduce {r['class_ms_launches']['reduce_ratio'][0] / r['class_ms_launches']['reduce_ratio'][1] * 1e3:5.1f} us obj {d['objective']:.9f} sm {d['clocks']['sm_mhz']}")
except Exception as e:
    print(tag, "pytest rc=", prc, "bench rc=", brc, "no result:", e)
PY
  tail -n 3 $out/r2f_pytest_$tag.log | head -2
done
