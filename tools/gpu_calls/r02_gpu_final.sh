#!/usr/bin/env bash
# final state of the round: whole GPU suite, smoke, default bench line (+ cpu_baseline), reference arm
# (the ncu capture of every kernel family was taken by an earlier run of this script: profiles/r02_ncu_*)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/tests_gpu.log 2>&1; echo "tests_gpu rc=$?"; grep -E "passed|failed|Error|assert " gpurun_out/tests_gpu.log | tail -n 6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"; tail -c 400 gpurun_out/bench_reference.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}, "e2e", d["e2e"]["value"], "cpu", d.get("cpu_baseline", {}).get("value"), "roofline", {k: d["roofline"][k] for k in ("kernel", "frac", "traffic")})
PY
