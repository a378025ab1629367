#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/tests_gpu.log 2>&1; echo "tests_gpu rc=$?"; tail -n 12 gpurun_out/tests_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_new.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_new.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("roofline"))
ks=d.get("kernels") or {}
for k,v in sorted(ks.items(), key=lambda kv:-kv[1].get("share_of_step",0))[:22]:
    print(f"   {k:40s} n={v['launches_per_step']:5.0f} avg={v['avg_ms']*1e3:8.1f}us share={v['share_of_step']*100:5.1f}% frac={v.get('frac')}")
PY
