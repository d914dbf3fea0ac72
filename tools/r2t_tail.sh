#!/bin/bash
# contraction 2 tail: 20 tail tiles as 60 pieces of 96 / 96 / 64 frames (default) vs whole tiles (EVC_NO_HALF_TILES=1)
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $out/r2t_pytest.log
tools/ab_bench.sh "thirds:EVC_X=1" "whole:EVC_NO_HALF_TILES=1" "thirds2:EVC_X=1" "whole2:EVC_NO_HALF_TILES=1"
timeout 300 python tests/manual/dtw_timing.py > $out/r2t_dtw.log 2>&1; echo "dtw rc=$?"; tail -n 3 $out/r2t_dtw.log
