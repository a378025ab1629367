#!/usr/bin/env bash
# A/B: add_dropout_ln_bwd at 2 (product) vs 3 resident CTAs per SM; add_dropout_ln_fwd now at 4 CTAs per SM
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_glue_gpu.py -m gpu -q > gpurun_out/tests_g.log 2>&1; echo "tests (product lib) rc=$?"; tail -n 2 gpurun_out/tests_g.log
PVQA_LIB_PATH=$PWD/tools/_variants/libpvqa_lnbwd3.so timeout 300 python -m pytest tests/test_glue_gpu.py -m gpu -q > gpurun_out/tests_g3.log 2>&1; echo "tests (variant lib) rc=$?"; tail -n 2 gpurun_out/tests_g3.log
for V in product variant product variant; do
  if [ $V = variant ]; then export PVQA_LIB_PATH=$PWD/tools/_variants/libpvqa_lnbwd3.so; else unset PVQA_LIB_PATH; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g_$V.json 2> gpurun_out/bench_g_$V.err; echo "bench $V rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_g_$V.json").read().strip().splitlines()[-1])
print("$V", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3))
for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1].get("share_of_step", 0)):
    if "add_" in k or "rms" in k: print(f"  {k:26s} n {v['launches_per_step']:4.0f} avg {v['avg_ms']*1e3:7.1f} us frac {v.get('frac', 0):.3f}")
PY
done
