#!/usr/bin/env bash
# final state of the round: whole GPU suite, smoke, default bench line (+ cpu_baseline), reference arm, ncu evidence
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
KREGEX='regex:^(attn_|add_dropout|cast_rows|col_sum|embed_|phoneme_head|relu_dropout|residual_dropout|rms_norm|vocab_ce)'
timeout 300 python tools/ncu_one_layer.py > gpurun_out/ncu_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$KREGEX" -c 120 -f -o /tmp/r02_kernels python tools/ncu_one_layer.py > gpurun_out/ncu_run.log 2>&1; echo "ncu rc=$?"; tail -n 2 gpurun_out/ncu_run.log
ncu -i /tmp/r02_kernels.ncu-rep --page raw --csv > gpurun_out/r02_kernels_raw.csv 2> gpurun_out/ncu_export.err; echo "export rc=$?"
python tools/ncu_summary.py < gpurun_out/r02_kernels_raw.csv > gpurun_out/r02_ncu_kernels.txt 2>> gpurun_out/ncu_export.err
ls -la gpurun_out | head -20
