#!/bin/bash
# leftover rows carried by contraction 1's split warps: parity tests, bench, bit identity of the reduce variants
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 $out/r2o_pytest.log
tools/ab_bench.sh "default:EVC_X=1" "fused:EVC_FUSED_REDUCE=1" "default2:EVC_X=1"
