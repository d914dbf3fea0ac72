#!/bin/bash
# A/B runs of bench.py under environment variants; one JSON line per variant into gpurun_out/ab_<tag>.json.
# usage: tools/ab_bench.sh "<tag>:<ENV=VAL ...>" ...
mkdir -p gpurun_out
for spec in "$@"; do
  tag="${spec%%:*}"; envs="${spec#*:}"
  env $envs timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err || echo "variant $tag failed"
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{tag}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f"{tag:14s} {d['value']:8.0f} frames/s  {d['ms_per_step']:7.2f} ms  c1 {r['contraction1_us_per_launch']:6.1f} us  c2 {r['us_per_launch']:6.1f} us  "
          f"obj {d['objective']:.6f}  sm {d['clocks']['sm_mhz']}")
except Exception as e:
    print(tag, "no result:", e, open(f"gpurun_out/ab_{tag}.err").read()[-300:])
PY
done
