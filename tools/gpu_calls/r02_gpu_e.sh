#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_head_gpu.py -m gpu -q > gpurun_out/tests_head.log 2>&1; echo "tests_head rc=$?"; tail -n 4 gpurun_out/tests_head.log
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_model_gpu.py -m gpu -q -s > gpurun_out/tests_train.log 2>&1; echo "tests_train rc=$?"; grep -E "passed|failed|\[bf16|Error|assert" gpurun_out/tests_train.log | tail -n 8 | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_k4dl.json 2> gpurun_out/bench_k4dl.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_k4dl.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d.get("gpu_launches"))
for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1].get("share_of_step", 0)):
    if "head" in k: print(f"  {k:34s} n {v['launches_per_step']:5.0f} avg {v['avg_ms']*1e3:8.1f} us share {v.get('share_of_step', 0):.3f}")
PY
