#!/bin/bash
# Bottleneck decomposition of the two GEMM kernels: rerun the bench on the INSTRUMENTED build of the library
# (build_variants/libevc_b200_instr.so, nvcc -DEVC_INSTRUMENT; the default build has no such switches) with parts
# of the pipeline disabled.  EVC_DEBUG_FLAGS bits: 1 no plane split (contraction 1), 2 no MMA issue, 4 no operand TMA,
# 8 no update arithmetic, 16 no H chunk loads/stores, 32 no leftover-row partials, 64 no TMEM loads.
# Results are numerically meaningless when flags != 0; only the per-kernel times are read.
MODES=${MODES:-"3xtf32 bf16"}
FLAGS=${FLAGS:-"0 2 4 8 16 32 6 24 22 30 94"}
LIB=${LIB:-build_variants/libevc_b200_instr.so}
for mode in $MODES; do
for f in $FLAGS; do
  EVC_LIB_PATH=$LIB EVC_DEBUG_FLAGS=$f timeout 120 python bench.py --steps 2 --warmup 2 --iterations 60 --no-cpu-baseline --mode $mode 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('mode $mode flags %3d: contraction1 %7.1f us  contraction2 %7.1f us  step %7.2f ms' % ($f, r['contraction1_us_per_launch'], r['us_per_launch'], d['ms_per_step']))"
done; done
