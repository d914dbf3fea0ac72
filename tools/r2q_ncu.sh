#!/bin/bash
# ncu of the final round-2 kernels: launch list, then one full capture (with source) of the three hot kernels
out=gpurun_out; mkdir -p $out
cmd="python bench.py --steps 1 --warmup 1 --iterations 20 --no-cpu-baseline --no-extras"
$cmd > $out/r2q_plain.log 2>&1 || { echo "plain run failed"; tail -n 5 $out/r2q_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/r2q_launches.csv $cmd > $out/r2q_ncu1.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel|reduce_partials" -s 30 -c 6 -o $out/r2q_prof $cmd > $out/r2q_ncu2.log 2>&1
echo "ncu full rc=$?"; ls -la $out/r2q_prof.ncu-rep
