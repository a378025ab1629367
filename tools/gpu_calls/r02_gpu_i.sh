#!/usr/bin/env bash
# A/B: remainder of a key row as (64, 32) tiles (product) or as one 128-wide masked tile
set -u
mkdir -p gpurun_out
for V in 0 1; do
  export PVQA_FWD_ONE_REM_TILE=$V
  timeout 600 python -m pytest tests/test_attn_gpu.py -m gpu -q > gpurun_out/tests_i$V.log 2>&1; echo "attn tests (one_rem_tile=$V) rc=$?"; tail -n 2 gpurun_out/tests_i$V.log | head -n 1
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_i$V.json 2> gpurun_out/bench_i$V.err; echo "bench rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_i$V.json").read().strip().splitlines()[-1])
print("one_rem_tile=$V", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3))
for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1].get("share_of_step", 0)):
    if k.startswith("attn_fwd"): print(f"  {k:26s} n {v['launches_per_step']:4.0f} avg {v['avg_ms']*1e3:7.1f} us frac {v.get('frac', 0):.3f}")
PY
done
