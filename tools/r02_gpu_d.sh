#!/usr/bin/env bash
# whole GPU suite + smoke + the default bench line (with cpu_baseline) + the reference arm
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x > gpurun_out/tests_gpu.log 2>&1; echo "tests_gpu rc=$?"; grep -E "passed|failed|\[fp32|\[bf16|\[30|Error|assert" gpurun_out/tests_gpu.log | tail -n 25
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}, "e2e", d["e2e"], "cpu", d.get("cpu_baseline"), "roofline", d["roofline"])
PY
