#!/bin/bash
# per-role wait cycles of the two contractions (instrumented build, EVC_DEBUG_FLAGS=128): where the CTAs wait
out=gpurun_out; mkdir -p $out
for v in ${VARIANTS:-fused separate}; do
  extra="EVC_FUSED_REDUCE=1"; [ $v = separate ] && extra="EVC_NO_FUSED_REDUCE=1"
  env $extra EVC_LIB_PATH=build_variants/libevc_b200_instr.so EVC_DEBUG_FLAGS=128 timeout 200 python bench.py --steps 1 --warmup 1 --iterations 4 \
    --no-cpu-baseline --no-extras > $out/r2m_clk_$v.log 2>&1
  echo "== $v rc=$? lines $(grep -c '^clk' $out/r2m_clk_$v.log)"
  grep '^clk cta [01] ' $out/r2m_clk_$v.log | tail -n 28
done
