#!/bin/bash
# Bottleneck decomposition of the two GEMM kernels: rerun the bench with parts of the pipeline disabled
# (EVC_DEBUG_FLAGS: 1 skip hi/lo split, 2 skip MMA issue, 4 skip TMA loads, 8 skip epilogue memory ops).
# Results are numerically meaningless when flags != 0; only the per-kernel times are read.
MODES=${MODES:-"3xtf32 tf32"}
FLAGS=${FLAGS:-"0 1 2 4 8 3 5 6 9 12 13 14 15"}
for mode in $MODES; do
for f in $FLAGS; do
  EVC_DEBUG_FLAGS=$f python bench.py --steps 2 --warmup 2 --iterations 60 --no-cpu-baseline --mode $mode 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('mode $mode flags %2d: contraction1 %7.1f us  contraction2 %7.1f us  step %7.2f ms' % ($f, r['contraction1_us_per_launch'], r['us_per_launch'], d['ms_per_step']))"
done; done
