#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_train_gpu.py tests/test_variants_gpu.py -m gpu -q -s > gpurun_out/tests_gpu.log 2>&1; echo "tests_gpu rc=$?"; grep -E "passed|failed|\[fp32|\[bf16|\[30|Error|assert" gpurun_out/tests_gpu.log | tail -n 25
