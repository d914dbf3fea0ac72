#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/r2b_pytest.log 2>&1; prc=$?; echo "pytest rc=$prc"; tail -n 8 $out/r2b_pytest.log
tools/flag_sweep.sh > $out/r2b_sweep.log 2>&1; cat $out/r2b_sweep.log
tools/ab_bench.sh "base:EVC_X=0" "h6s4:EVC_LIB_PATH=build_variants/libevc_b200_h6s4.so" "h8s3:EVC_LIB_PATH=build_variants/libevc_b200_h8s3.so" "nopdl:EVC_NO_PDL=1" 2>&1 | tee $out/r2b_ab.log
cmd="python bench.py --steps 1 --warmup 1 --iterations 20 --no-cpu-baseline"
$cmd > $out/r2b_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 20 -c 4 -o $out/r2b_prof $cmd > $out/r2b_ncu2.log 2>&1
echo "ncu full rc=$?"; tail -n 3 $out/r2b_ncu2.log
