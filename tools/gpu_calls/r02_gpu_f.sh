#!/usr/bin/env bash
# relu_dropout with one Philox-7 call per 16 elements, K4 backward on padded operands: tests + bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_glue_gpu.py tests/test_head_gpu.py tests/test_model_gpu.py -m gpu -q > gpurun_out/tests_f.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/tests_f.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_f.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d.get("gpu_launches"))
for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1].get("share_of_step", 0)):
    if "relu" in k or "head" in k: print(f"  {k:34s} n {v['launches_per_step']:5.0f} avg {v['avg_ms']*1e3:8.1f} us share {v.get('share_of_step', 0):.3f} frac {v.get('frac', 0):.3f}")
PY
