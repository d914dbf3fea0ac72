#!/bin/bash
# per-role cycles with parts of contraction 2 disabled (instrumented build): 188 = MMA only, 184 = MMA + operand TMA, 128 = everything
out=gpurun_out; mkdir -p $out
for f in ${FLAGSETS:-188 184 152 128}; do
  EVC_NO_FUSED_REDUCE=1 EVC_LIB_PATH=build_variants/libevc_b200_instr.so EVC_DEBUG_FLAGS=$f timeout 200 python bench.py --steps 1 --warmup 1 --iterations 4 \
    --no-cpu-baseline --no-extras > $out/r2s_clk_$f.log 2>&1
  echo "== flags $f rc=$?"
  grep '^clk cta 0' $out/r2s_clk_$f.log | grep -v "wait_raw [1-9]" | tail -n 12 | grep "mma\|producer\|epilogue warp 2\|hstorer" | tail -n 4
done
