#!/bin/bash
# multi-GPU visit: sharded tests, bench at N ranks with the exemplar-sharded extra line, reference arm under torchrun
N=${1:-2}; out=gpurun_out; mkdir -p $out
nvidia-smi -L | head -9
[ "$2" = "notests" ] || { timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s > $out/r2e_pytest_n$N.log 2>&1; echo "pytest rc=$?"; tail -n 6 $out/r2e_pytest_n$N.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 > $out/r2e_bench_n$N.json 2> $out/r2e_bench_n$N.err; echo "bench N=$N rc=$?"; tail -n 4 $out/r2e_bench_n$N.err
python - $out/r2e_bench_n$N.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("N", d["n_gpus"], round(d["value"]), "frames/s e2e", round(d["e2e"]["value"]), d["e2e"].get("ms_each_step"), d["clocks"])
    print("extra", json.dumps(d.get("extra")))
except Exception as e:
    print("no result", e)
PY
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-extras --no-p2p --workload large_dictionary_200k --iterations 50 > $out/r2e_bench200k_nccl_n$N.json 2> $out/r2e_bench200k_nccl_n$N.err; echo "200k nccl rc=$?"
python - $out/r2e_bench200k_nccl_n$N.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("200k nccl: ms/step", round(d["ms_per_step"], 2), d["config"]["all_reduce"], d["roofline"]["class_ms_launches"])
except Exception as e:
    print("no result", e)
PY
timeout 600 $TR bench.py --impl reference --gpus $N --steps 1 --warmup 0 --ref-sample-iters 2 > $out/r2e_ref_n$N.json 2> $out/r2e_ref_n$N.err; echo "ref arm rc=$?"
python - $out/r2e_ref_n$N.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("reference arm under torchrun:", round(d["value"], 1), "frames/s blas_threads", d["blas_threads"], "env", d["omp_num_threads_env"], "cores", d["cpu_baseline"]["cores"])
except Exception as e:
    print("no result", e)
PY
