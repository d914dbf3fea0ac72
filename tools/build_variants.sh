#!/bin/bash
# Compile-time variants of libevc_b200.so for A/B runs (EVC_LIB_PATH selects one at run time; still no fallback).
# usage: tools/build_variants.sh name:"-DFLAG=VAL ..." ...
mkdir -p build_variants
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $defs \
       -o build_variants/libevc_b200_$name.so exemplars_vc_b200/csrc/evc_api.cu -ldl && echo "built $name ($defs)"
done
