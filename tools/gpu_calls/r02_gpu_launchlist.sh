#!/usr/bin/env bash
# ncu launch list (gpu__time_duration only) of ONE eager training step of the 12-layer bench model, HEAD kernels
set -u
mkdir -p gpurun_out
timeout 300 python tools/one_step.py 2 > gpurun_out/one_step_plain.log 2>&1; echo "plain rc=$?"; tail -n 1 gpurun_out/one_step_plain.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_one_step.csv python tools/one_step.py 2 > gpurun_out/one_step_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launch_table.py gpurun_out/r02_launches_one_step.csv 60 > gpurun_out/r02_launch_table.txt; head -n 30 gpurun_out/r02_launch_table.txt
ls -la gpurun_out/r02_launches_one_step.csv
