#!/bin/bash
# ring depth A/B: default (5 operand stages + 4 H chunk buffers in contraction 2), 6 + 2, 4 + 4
out=gpurun_out; mkdir -p $out
tools/ab_bench.sh "default:EVC_X=1" "hb2:EVC_LIB_PATH=build_variants/libevc_b200_hb2.so" "st4:EVC_LIB_PATH=build_variants/libevc_b200_st4.so" "default2:EVC_X=1" "hb2b:EVC_LIB_PATH=build_variants/libevc_b200_hb2.so"
