#!/usr/bin/env bash
# Philox4x32-7 in the shared dropout helper: every test that touches dropout + the bench line
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_glue_gpu.py tests/test_embed_gpu.py tests/test_model_gpu.py tests/test_variants_gpu.py tests/test_train_gpu.py -m gpu -q > gpurun_out/tests_h.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/tests_h.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_h.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1))
for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1].get("share_of_step", 0)):
    if "add_" in k or "relu" in k: print(f"  {k:26s} n {v['launches_per_step']:4.0f} avg {v['avg_ms']*1e3:7.1f} us frac {v.get('frac', 0):.3f}")
PY
