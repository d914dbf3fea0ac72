#!/bin/bash
# round 2, in-kernel split-K sum: bit identity against the separate pass, A/B bench, per-phase cycles
out=gpurun_out; mkdir -p $out
EVC_FUSED_REDUCE=1 timeout 600 python tests/manual/fused_reduce_ab.py > $out/r2l_fused.txt 2> $out/r2l_fused.err; echo "fused rc=$?"; tail -n 3 $out/r2l_fused.err
EVC_NO_FUSED_REDUCE=1 timeout 600 python tests/manual/fused_reduce_ab.py > $out/r2l_separate.txt 2> $out/r2l_separate.err; echo "separate rc=$?"
if cmp -s $out/r2l_fused.txt $out/r2l_separate.txt; then echo "BIT-IDENTICAL ($(wc -l < $out/r2l_fused.txt) lines)"; else echo "DIFFERENT"; diff $out/r2l_fused.txt $out/r2l_separate.txt | head -20; fi
tools/ab_bench.sh "fused:EVC_FUSED_REDUCE=1" "separate:EVC_NO_FUSED_REDUCE=1" "fused2:EVC_FUSED_REDUCE=1" "separate2:EVC_NO_FUSED_REDUCE=1"
VARIANTS=fused tools/r2m_clk.sh | grep "fused reduce" | tail -4
