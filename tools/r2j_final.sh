#!/bin/bash
# final confirmation on N GPUs: smoke, GPU tests (N=1 only), bench both arms the way the driver launches them
N=${1:-1}; out=gpurun_out; mkdir -p $out
if [ "$N" = "1" ]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/r2j_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 5 $out/r2j_smoke.log
  timeout 1200 python -m pytest tests -m gpu -q > $out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $out/r2j_pytest.log
  timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $out/r2j_ref_n1.json 2> $out/r2j_ref_n1.err; echo "reference rc=$?"
  timeout 900 python bench.py --gpus 1 > $out/r2j_bench_n1.json 2> $out/r2j_bench_n1.err; echo "bench rc=$?"; tail -n 3 $out/r2j_bench_n1.err
else
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
  timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 > $out/r2j_bench_n$N.json 2> $out/r2j_bench_n$N.err; echo "bench N=$N rc=$?"; tail -n 3 $out/r2j_bench_n$N.err
fi
python - $N <<'PY'
import json, sys
N = sys.argv[1]
for f in (f"gpurun_out/r2j_ref_n{N}.json", f"gpurun_out/r2j_bench_n{N}.json"):
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["e2e"].get("ms_each_step"), "frac", r.get("frac"), "exec", r.get("executed_frac"),
              "c1/c2 us", r.get("contraction1_us_per_launch"), r.get("us_per_launch"), d.get("clocks"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
        if d.get("extra"): print("   extra", json.dumps(d["extra"]["exemplar_sharded"]))
    except Exception as e:
        print(f, "no result", e)
PY
