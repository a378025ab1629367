#!/usr/bin/env bash
# Round-2 second GPU call: first device run of the rewritten attention kernels.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/attn_probe.log
timeout 300 python tools/attn_probe.py parity > gpurun_out/probe_parity.out 2>&1; echo "parity rc=$?"
tail -n 40 gpurun_out/attn_probe.log
timeout 300 python tools/attn_probe.py timing > gpurun_out/probe_timing.out 2>&1; echo "timing rc=$?"
tail -n 12 gpurun_out/attn_probe.log
timeout 900 python -m pytest tests/test_attn_gpu.py -m gpu -q -x > gpurun_out/tests_attn.log 2>&1; echo "tests_attn rc=$?"; tail -n 15 gpurun_out/tests_attn.log
