#!/bin/bash
# contraction 2: updated activations stored straight from registers (EVC_C2_DIRECT_STORE=1) vs the staged TMA store
out=gpurun_out; mkdir -p $out
tools/ab_bench.sh "staged:EVC_NO_FUSED_REDUCE=1" "direct:EVC_NO_FUSED_REDUCE=1 EVC_C2_DIRECT_STORE=1" "staged2:EVC_NO_FUSED_REDUCE=1" "direct2:EVC_NO_FUSED_REDUCE=1 EVC_C2_DIRECT_STORE=1"
EVC_C2_DIRECT_STORE=1 timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_properties_gpu.py -m gpu -x -q > $out/r2n_pytest.log 2>&1; echo "pytest(direct) rc=$?"; tail -n 3 $out/r2n_pytest.log
EVC_C2_DIRECT_STORE=1 EVC_NO_FUSED_REDUCE=1 EVC_LIB_PATH=build_variants/libevc_b200_instr.so EVC_DEBUG_FLAGS=128 timeout 200 python bench.py --steps 1 --warmup 1 --iterations 4 \
    --no-cpu-baseline --no-extras > $out/r2n_clk_direct.log 2>&1
grep '^clk cta 0' $out/r2n_clk_direct.log | tail -n 12 | grep -v "wait_raw [1-9]" 
