#!/usr/bin/env bash
# fused K4 (tcgen05) parity, training tests, eager phonosal arm, ncu evidence for every kernel family
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_head_gpu.py -m gpu -q -x > gpurun_out/tests_head.log 2>&1; echo "tests_head rc=$?"; tail -n 15 gpurun_out/tests_head.log
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_parallel_gpu.py -m gpu -q -s > gpurun_out/tests_train.log 2>&1; echo "tests_train rc=$?"; grep -E "passed|failed|gradient|Error|assert" gpurun_out/tests_train.log | tail -n 30
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_k4tc.json 2> gpurun_out/bench_k4tc.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_k4tc.json
timeout 600 python bench.py --impl eager-gpu --workload phonosal --batch 32 --steps 4 --warmup 2 > gpurun_out/eager_phonosal.json 2> gpurun_out/eager_phonosal.err; echo "eager phonosal rc=$?"; tail -c 600 gpurun_out/eager_phonosal.json
KREGEX='regex:^(attn_|add_dropout|cast_rows|col_sum|embed_|phoneme_head|relu_dropout|residual_dropout|rms_norm|vocab_ce)'
timeout 300 python tools/ncu_one_layer.py > gpurun_out/ncu_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$KREGEX" -c 120 -f -o gpurun_out/r02_kernels python tools/ncu_one_layer.py > gpurun_out/ncu_run.log 2>&1; echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_run.log
ls -la gpurun_out/*.ncu-rep 2>/dev/null
