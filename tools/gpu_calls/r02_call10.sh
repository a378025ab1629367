#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/attn_probe.log
timeout 300 python tools/attn_probe.py all > gpurun_out/probe_all.out 2>&1; echo "probe rc=$?"
grep -E "done|median" gpurun_out/attn_probe.log | cut -c1-200
timeout 900 python -m pytest tests/test_attn_gpu.py -m gpu -q -x > gpurun_out/tests_attn.log 2>&1; echo "tests_attn rc=$?"; tail -n 5 gpurun_out/tests_attn.log
timeout 600 bash -c "python tools/attn_trace.py build && python tools/attn_trace.py 64" > gpurun_out/attn_trace.log 2>&1; echo "trace rc=$?"
grep -A25 "== fwd enc_self B=64 p=0.1" gpurun_out/attn_trace.log
