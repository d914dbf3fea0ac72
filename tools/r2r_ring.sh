#!/bin/bash
# contraction 2, shared memory split between operand stages and H chunk buffers: 5 + 4 (default), 4 + 6, 3 + 8
out=gpurun_out; mkdir -p $out
tools/ab_bench.sh "default:EVC_X=1" "s4h6:EVC_LIB_PATH=build_variants/libevc_b200_s4h6.so" "s3h8:EVC_LIB_PATH=build_variants/libevc_b200_s3h8.so" "default2:EVC_X=1" "s4h6b:EVC_LIB_PATH=build_variants/libevc_b200_s4h6.so"
