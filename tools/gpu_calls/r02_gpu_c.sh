#!/usr/bin/env bash
# fp32 gradient diagnostic at base dims, head + training tests, ncu evidence exported as CSV (the .ncu-rep stays on the box:
# gpurun_out/ is capped at 64 MiB)
set -u
mkdir -p gpurun_out
timeout 600 python tools/diag_base_grads.py > gpurun_out/diag_base.log 2>&1; echo "diag rc=$?"; tail -n 60 gpurun_out/diag_base.log
timeout 300 python -m pytest tests/test_head_gpu.py -m gpu -q > gpurun_out/tests_head.log 2>&1; echo "tests_head rc=$?"; tail -n 5 gpurun_out/tests_head.log
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q -s > gpurun_out/tests_train.log 2>&1; echo "tests_train rc=$?"; grep -E "passed|failed|gradient|graphed|Error|assert" gpurun_out/tests_train.log | tail -n 30
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_k4tc.json 2> gpurun_out/bench_k4tc.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_k4tc.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d.get("gpu_launches"))
for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1].get("share_of_step", 0)):
    print(f"  {k:34s} n {v['launches_per_step']:5.0f} avg {v['avg_ms']*1e3:8.1f} us share {v.get('share_of_step', 0):.3f} frac {v.get('frac', 0):.3f}")
PY
KREGEX='regex:^(attn_|add_dropout|cast_rows|col_sum|embed_|phoneme_head|relu_dropout|residual_dropout|rms_norm|vocab_ce)'
timeout 300 python tools/ncu_one_layer.py > gpurun_out/ncu_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$KREGEX" -c 120 -f -o /tmp/r02_kernels python tools/ncu_one_layer.py > gpurun_out/ncu_run.log 2>&1; echo "ncu rc=$?"; tail -n 2 gpurun_out/ncu_run.log
ncu -i /tmp/r02_kernels.ncu-rep --page raw --csv > gpurun_out/r02_kernels_raw.csv 2> gpurun_out/ncu_export.err; echo "export rc=$?"
python tools/ncu_summary.py < gpurun_out/r02_kernels_raw.csv > gpurun_out/r02_ncu_kernels.txt 2>> gpurun_out/ncu_export.err
ls -la gpurun_out/ | head -30; du -sh gpurun_out
